/* TEST INFRASTRUCTURE (oracle) — not part of the product path.
 *
 * Our own mixed-radix complex FFT behind the two third-party call surfaces the
 * reference's speedy.c uses (neither library is vendored by the reference nor
 * present in this environment):
 *   - kissfft, float:  kiss_fft_alloc / kiss_fft        (speedy.c:269, 449)
 *   - FFTW3, double:   fftw_plan_dft_1d / fftw_execute  (speedy.c:274-277, 467)
 *
 * Algorithm: Stockham autosort, decimation in frequency, arbitrary factor list.
 * For a stage of radix R on sub-transforms of length n (stride s = N/n):
 *     y[q + s(Rp + t)] = W_n^{pt} * sum_r x[q + s(p + r n/R)] W_R^{rt}
 * Radix 2 and 4 are the multiplication-free butterflies written out; every
 * other prime factor (3, 5, and 11 for the 22.05 kHz frame, N = 660) uses the
 * O(R^2) form with roots taken from the length-N table.  All roots are computed
 * in double and rounded once to the working type.
 *
 * The FFT sizes the reference needs are 2*(int)(1.5*rate/100) (speedy.c:213-214):
 * 480, 660, 720, 1440.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "shim/fftw3.h"
#include "shim/kiss_fft.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define ORACLE_MAX_FACTORS 32
#define ORACLE_MAX_RADIX 2048

static int factorize(int n, int* factors) {
  int count = 0;
  while (n % 4 == 0) { factors[count++] = 4; n /= 4; }
  while (n % 2 == 0) { factors[count++] = 2; n /= 2; }
  for (int p = 3; n > 1; p += 2) {
    while (n % p == 0) { factors[count++] = p; n /= p; }
  }
  return count;
}

#define DEFINE_FFT(SUFFIX, REAL)                                               \
  typedef struct { REAL r, i; } cpx_##SUFFIX;                                  \
  typedef struct {                                                             \
    int n;                                                                     \
    int inverse;                                                               \
    int nfactors;                                                              \
    int factors[ORACLE_MAX_FACTORS];                                           \
    cpx_##SUFFIX* roots; /* W_N^k, k = 0..N-1 (conjugated when inverse) */     \
    cpx_##SUFFIX* work0;                                                       \
    cpx_##SUFFIX* work1;                                                       \
  } plan_##SUFFIX;                                                             \
                                                                               \
  static size_t plan_bytes_##SUFFIX(int n) {                                   \
    return sizeof(plan_##SUFFIX) + 3 * (size_t)n * sizeof(cpx_##SUFFIX);       \
  }                                                                            \
                                                                               \
  static void plan_init_##SUFFIX(plan_##SUFFIX* p, int n, int inverse) {       \
    p->n = n;                                                                  \
    p->inverse = inverse;                                                      \
    p->nfactors = factorize(n, p->factors);                                    \
    p->roots = (cpx_##SUFFIX*)(p + 1);                                         \
    p->work0 = p->roots + n;                                                   \
    p->work1 = p->work0 + n;                                                   \
    for (int k = 0; k < n; k++) {                                              \
      double phase = -2.0 * M_PI * (double)k / (double)n;                      \
      if (inverse) phase = -phase;                                             \
      p->roots[k].r = (REAL)cos(phase);                                        \
      p->roots[k].i = (REAL)sin(phase);                                        \
    }                                                                          \
  }                                                                            \
                                                                               \
  static inline cpx_##SUFFIX cmul_##SUFFIX(cpx_##SUFFIX a, cpx_##SUFFIX b) {   \
    cpx_##SUFFIX c;                                                            \
    c.r = a.r * b.r - a.i * b.i;                                               \
    c.i = a.r * b.i + a.i * b.r;                                               \
    return c;                                                                  \
  }                                                                            \
                                                                               \
  /* One Stockham stage: x (length n*s, n = current transform length) -> y. */ \
  static void stage_##SUFFIX(const plan_##SUFFIX* pl, int radix, int n, int s, \
                             const cpx_##SUFFIX* x, cpx_##SUFFIX* y) {         \
    const int N = pl->n;                                                       \
    const int m = n / radix;                                                   \
    const cpx_##SUFFIX* w = pl->roots;                                         \
    /* sign of the transform: forward multiplies by -i where W_4 appears */    \
    const REAL sg = pl->inverse ? (REAL)-1 : (REAL)1;                          \
    for (int p = 0; p < m; p++) {                                              \
      /* W_n^{pt} = W_N^{p t s} */                                             \
      const int step = (int)(((long)p * s) % N);                               \
      for (int q = 0; q < s; q++) {                                            \
        const cpx_##SUFFIX* xi = x + q + s * p;                                \
        cpx_##SUFFIX* yo = y + q + s * radix * p;                              \
        const int xs = s * m;                                                  \
        cpx_##SUFFIX o[ORACLE_MAX_RADIX];                                      \
        o[0].r = 0; o[0].i = 0;                                                \
        if (radix == 2) {                                                      \
          cpx_##SUFFIX a = xi[0], b = xi[xs];                                  \
          o[0].r = a.r + b.r; o[0].i = a.i + b.i;                              \
          o[1].r = a.r - b.r; o[1].i = a.i - b.i;                              \
        } else if (radix == 4) {                                               \
          cpx_##SUFFIX a = xi[0], b = xi[xs], c = xi[2 * xs], d = xi[3 * xs];  \
          cpx_##SUFFIX apc = {a.r + c.r, a.i + c.i};                           \
          cpx_##SUFFIX amc = {a.r - c.r, a.i - c.i};                           \
          cpx_##SUFFIX bpd = {b.r + d.r, b.i + d.i};                           \
          /* -i*(b-d) forward, +i*(b-d) inverse */                             \
          cpx_##SUFFIX jbmd = {sg * (b.i - d.i), -sg * (b.r - d.r)};           \
          o[0].r = apc.r + bpd.r;  o[0].i = apc.i + bpd.i;                     \
          o[1].r = amc.r + jbmd.r; o[1].i = amc.i + jbmd.i;                    \
          o[2].r = apc.r - bpd.r;  o[2].i = apc.i - bpd.i;                     \
          o[3].r = amc.r - jbmd.r; o[3].i = amc.i - jbmd.i;                    \
        } else {                                                               \
          cpx_##SUFFIX a[ORACLE_MAX_RADIX]; /* 44.1 kHz: N = 1322 = 2 * 661 */   \
          for (int r = 0; r < radix; r++) a[r] = xi[r * xs];                   \
          for (int t = 0; t < radix; t++) {                                    \
            cpx_##SUFFIX acc = a[0];                                           \
            for (int r = 1; r < radix; r++) {                                  \
              /* W_R^{rt} = W_N^{(N/R) * (rt mod R)} */                        \
              int e = (r * t) % radix;                                         \
              if (e == 0) {                                                    \
                acc.r += a[r].r; acc.i += a[r].i;                              \
              } else {                                                         \
                cpx_##SUFFIX term = cmul_##SUFFIX(a[r], w[(N / radix) * e]);   \
                acc.r += term.r; acc.i += term.i;                              \
              }                                                                \
            }                                                                  \
            o[t] = acc;                                                        \
          }                                                                    \
        }                                                                      \
        yo[0] = o[0];                                                          \
        if (p == 0) {                                                          \
          for (int t = 1; t < radix; t++) yo[s * t] = o[t];                    \
        } else {                                                               \
          int e = step;                                                        \
          for (int t = 1; t < radix; t++) {                                    \
            yo[s * t] = cmul_##SUFFIX(o[t], w[e]);                             \
            e += step; if (e >= N) e -= N;                                     \
          }                                                                    \
        }                                                                      \
      }                                                                        \
    }                                                                          \
  }                                                                            \
                                                                               \
  static void execute_##SUFFIX(plan_##SUFFIX* pl, const cpx_##SUFFIX* in,      \
                               cpx_##SUFFIX* out) {                            \
    const int N = pl->n;                                                       \
    cpx_##SUFFIX* src = pl->work0;                                             \
    cpx_##SUFFIX* dst = pl->work1;                                             \
    memcpy(src, in, (size_t)N * sizeof(cpx_##SUFFIX));                         \
    int n = N, s = 1;                                                          \
    for (int f = 0; f < pl->nfactors; f++) {                                   \
      int radix = pl->factors[f];                                              \
      stage_##SUFFIX(pl, radix, n, s, src, dst);                               \
      n /= radix;                                                              \
      s *= radix;                                                              \
      cpx_##SUFFIX* tmp = src; src = dst; dst = tmp;                           \
    }                                                                          \
    memcpy(out, src, (size_t)N * sizeof(cpx_##SUFFIX));                        \
  }

DEFINE_FFT(f32, float)
DEFINE_FFT(f64, double)

/* ---- kissfft-compatible surface (float) ------------------------------- */

struct kiss_fft_state {
  plan_f32 plan;
};

kiss_fft_cfg kiss_fft_alloc(int nfft, int inverse_fft, void* mem,
                            size_t* lenmem) {
  size_t need = plan_bytes_f32(nfft);
  kiss_fft_cfg cfg = NULL;
  if (lenmem == NULL) {
    cfg = (kiss_fft_cfg)malloc(need);
  } else {
    if (mem != NULL && *lenmem >= need) cfg = (kiss_fft_cfg)mem;
    *lenmem = need;
  }
  if (cfg) plan_init_f32(&cfg->plan, nfft, inverse_fft);
  return cfg;
}

void kiss_fft(kiss_fft_cfg cfg, const kiss_fft_cpx* fin, kiss_fft_cpx* fout) {
  execute_f32(&cfg->plan, (const cpx_f32*)fin, (cpx_f32*)fout);
}

void kiss_fft_cleanup(void) {}

/* ---- FFTW3-compatible surface (double) -------------------------------- */

struct fftw_plan_s {
  fftw_complex* in;
  fftw_complex* out;
  plan_f64* plan;
};

void* fftw_malloc(size_t n) { return malloc(n); }
void fftw_free(void* p) { free(p); }

fftw_plan fftw_plan_dft_1d(int n, fftw_complex* in, fftw_complex* out, int sign,
                           unsigned flags) {
  (void)flags;
  fftw_plan p = (fftw_plan)malloc(sizeof(struct fftw_plan_s));
  if (!p) return NULL;
  p->plan = (plan_f64*)malloc(plan_bytes_f64(n));
  if (!p->plan) {
    free(p);
    return NULL;
  }
  plan_init_f64(p->plan, n, sign == FFTW_BACKWARD);
  p->in = in;
  p->out = out;
  return p;
}

void fftw_execute(const fftw_plan p) {
  /* `double complex` is layout-compatible with {double re, im}. */
  execute_f64(p->plan, (const cpx_f64*)p->in, (cpx_f64*)p->out);
}

void fftw_destroy_plan(fftw_plan p) {
  if (p) {
    free(p->plan);
    free(p);
  }
}
