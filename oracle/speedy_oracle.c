/* TEST INFRASTRUCTURE (oracle) — not part of the product path.
 *
 * CPU restatement of the reference's nonlinear speed-up path in the
 * whole-stream ("batch") form the CUDA kernels use: every analysis frame, every
 * tension/speed value and the Sonic feed schedule are written as closed-form
 * functions of the frame index instead of the reference's incremental ring
 * buffers.  Each function cites the reference code it follows.  Floating-point
 * expressions keep the reference's C types and evaluation order (float vs
 * double promotion), so with the same FFT behind it this file reproduces the
 * compiled reference bit for bit (tests/test_oracle.py checks that against
 * oracle/_ref).
 *
 * Pinned by: the known-answer tests of /root/reference/speedy_test.cc
 * re-expressed in tests/test_oracle.py, and equality with the unmodified
 * speedy.c + soniclib.c compiled into oracle/_ref.  The Sonic stage
 * (oracle/sonic_oracle.c) and the FFT (oracle/fft_oracle.c) are our own
 * restatements of absent third-party code; see their headers.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use
 * this file.
 */
#include "speedy_oracle.h"

#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "shim/fftw3.h"
#include "shim/kiss_fft.h"
#define SONIC_INTERNAL 1
#include "shim/sonic.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* speedy.c:90 */
#define FRAME_RATE_HZ 100.0
/* speedy.c:92 */
#define MINIMUM_SPEED 0.01

void oracle_geometry(int rate, int match_matlab, oracle_geom* g) {
  /* speedy.c:213-214 */
  g->window = (int)(1.5 * rate / (float)FRAME_RATE_HZ);
  g->fft = 2 * g->window;
  /* speedy.c:335-338: double quotient truncated by the int return type */
  g->step = (int)(rate / FRAME_RATE_HZ);
  /* soniclib.c:406-411 */
  g->partial = g->window - (g->window / g->step) * g->step;
  /* speedy.h:136-146 */
  g->future = match_matlab ? 8 : 12;
  g->past = match_matlab ? 12 : 8;
  /* upstream Sonic: SONIC_MAX_PITCH 400, SONIC_MIN_PITCH 65, AMDF at 4 kHz */
  g->min_period = rate / SONIC_MAX_PITCH;
  g->max_period = rate / SONIC_MIN_PITCH;
  g->max_required = 2 * g->max_period;
  g->skip = rate > SONIC_AMDF_FREQ ? rate / SONIC_AMDF_FREQ : 1;
}

/* Number of analysis frames the shim has sent to Speedy after `total` sample
 * frames were written.  soniclib.c:440-444: window k (samples k*S .. k*S+W-1)
 * is sent when sample number k*S + W + 1 arrives (location == partial + 1 in
 * buffer k + W/S). */
int oracle_frames_analyzed(const oracle_geom* g, long total) {
  if (total < g->window + 1) return 0;
  return (int)((total - g->window - 1) / g->step) + 1;
}

/* Number of tension/speed values computed after `frames` analysis frames.
 * soniclib.c:317 + speedy.c:755: tension r is ready once r + Future <= at_time,
 * where at_time = k + 1 for window k (soniclib.c:296), one per send. */
int oracle_tensions_ready(const oracle_geom* g, int frames) {
  int n = frames - g->future + 1;
  return n > 0 ? n : 0;
}

/* soniclib.c:262-287: mono = (sum over channels) / channels, C integer
 * division (truncates toward zero), stored to short. */
static void downmix(const short* in, long n, int channels, short* mono) {
  for (long i = 0; i < n; i++) {
    int sum = 0;
    for (int c = 0; c < channels; c++) sum += in[i * channels + c];
    mono[i] = (short)(sum / channels);
  }
}

/* speedy.c:256-258: Hamming window, double arithmetic, stored float. */
void oracle_hamming(int window, float* w) {
  for (int i = 0; i < window; i++) {
    w[i] = 0.54 - 0.46 * cos(2 * M_PI * i / (window - 1.0));
  }
}

typedef struct {
  int n;
  int use_double;
  kiss_fft_cfg kiss;
  kiss_fft_cpx* kin;
  kiss_fft_cpx* kout;
  fftw_plan fftw;
  fftw_complex* din;
  fftw_complex* dout;
} spec_plan;

static int spec_plan_init(spec_plan* p, int n, int use_double) {
  memset(p, 0, sizeof(*p));
  p->n = n;
  p->use_double = use_double;
  if (use_double) {
    p->din = (fftw_complex*)fftw_malloc(sizeof(fftw_complex) * n);
    p->dout = (fftw_complex*)fftw_malloc(sizeof(fftw_complex) * n);
    if (!p->din || !p->dout) return 0;
    p->fftw = fftw_plan_dft_1d(n, p->din, p->dout, FFTW_FORWARD, FFTW_ESTIMATE);
    return p->fftw != NULL;
  }
  p->kin = (kiss_fft_cpx*)malloc(sizeof(kiss_fft_cpx) * n);
  p->kout = (kiss_fft_cpx*)malloc(sizeof(kiss_fft_cpx) * n);
  p->kiss = kiss_fft_alloc(n, 0, NULL, NULL);
  return p->kin && p->kout && p->kiss;
}

static void spec_plan_free(spec_plan* p) {
  if (p->fftw) fftw_destroy_plan(p->fftw);
  fftw_free(p->din);
  fftw_free(p->dout);
  free(p->kin);
  free(p->kout);
  free(p->kiss);
}

/* speedy.c:438-454 (float, kissfft) and :458-473 (double, FFTW): window,
 * zero-pad to N, complex FFT, magnitude of all N bins. */
static void spectrogram_frame(spec_plan* p, const float* input,
                              const float* window, int w, float* spec) {
  int n = p->n;
  if (p->use_double) {
    for (int i = 0; i < w; i++) p->din[i] = CMPLX(input[i] * window[i], 0);
    for (int i = w; i < n; i++) p->din[i] = CMPLX(0, 0);
    fftw_execute(p->fftw);
    for (int i = 0; i < n; i++) spec[i] = cabs(p->dout[i]);
  } else {
    for (int i = 0; i < w; i++) {
      p->kin[i].r = input[i] * window[i];
      p->kin[i].i = 0.0;
    }
    for (int i = w; i < n; i++) {
      p->kin[i].r = 0.0;
      p->kin[i].i = 0.0;
    }
    kiss_fft(p->kiss, p->kin, p->kout);
    for (int i = 0; i < n; i++) {
      kiss_fft_cpx c = p->kout[i];
      spec[i] = sqrt(c.r * c.r + c.i * c.i); /* speedy.c:434-436 */
    }
  }
}

/* speedy.c:628-647.  Returns the frame energy (DC skipped in the sum, not in
 * the normalised output). */
static float normalize_by_energy(const float* spec, float* normalized,
                                 int length) {
  float signal_energy = 0.0;
  for (int i = 1; i < length; i++) signal_energy += spec[i] * spec[i];
  const float eps = 2.2204e-16;
  float inverse_norm = 1.0 / (sqrt(signal_energy) + eps);
  for (int i = 0; i < length; i++) normalized[i] = spec[i] * inverse_norm;
  return signal_energy;
}

/* speedy.c:590-610 with the ring replaced by direct indexing: comp[a] is the
 * compressed energy stored at at_time a (0 for a <= 0, which the ring holds
 * because nothing is ever written there before it is read). */
static float hysteresis(const float* comp, int n_comp, int at, int future,
                        int past) {
  float past_max = 0.0, future_max = 0.0;
  for (int i = 0; i <= future; i++) {
    int a = at + i;
    float value = (a >= 1 && a <= n_comp) ? comp[a] : 0.0f;
    value *= (future - i) / (float)future;
    if (value > future_max) future_max = value;
  }
  for (int i = 0; i <= past; i++) {
    int a = at - i;
    float value = (a >= 1 && a <= n_comp) ? comp[a] : 0.0f;
    value *= (past - i) / (float)past;
    if (value > past_max) past_max = value;
  }
  return (past_max + future_max) / 2.0;
}

int oracle_analyze(const oracle_cfg* cfg, const short* in, long n_frames,
                   oracle_taps* taps) {
  oracle_geom g;
  oracle_geometry(cfg->rate, cfg->match_matlab, &g);
  const int W = g.window, N = g.fft, S = g.step, F = g.future, B = g.past;
  const int half = N / 2;
  const int nA = oracle_frames_analyzed(&g, n_frames);
  const int nT = oracle_tensions_ready(&g, nA);
  taps->n_analysis = nA;
  taps->n_tension = nT;
  if (taps->max_frames < nA) return -1;

  short* mono = (short*)malloc(sizeof(short) * (size_t)(n_frames > 0 ? n_frames : 1));
  float* window = (float*)malloc(sizeof(float) * W);
  float* input = (float*)malloc(sizeof(float) * W);
  /* spectrogram rows are indexed by at_time; rows 0 and "-1" are all zero:
   * the history ring is zero-initialised (speedy.c:242-248) and at_time starts
   * at 1 (soniclib.c:296). */
  float* spec = (float*)calloc((size_t)(nA + 2) * N, sizeof(float));
  float* E = (float*)calloc((size_t)nA + 2, sizeof(float));
  float* lp = (float*)calloc((size_t)nA + 2, sizeof(float));
  float* local = (float*)calloc((size_t)nA + 2, sizeof(float));
  float* comp = (float*)calloc((size_t)nA + 2, sizeof(float));
  float* norm = (float*)malloc(sizeof(float) * N);
  float* norm_last = (float*)malloc(sizeof(float) * N);
  spec_plan plan;
  if (!mono || !window || !input || !spec || !E || !lp || !local || !comp ||
      !norm || !norm_last || !spec_plan_init(&plan, N, cfg->fft_double)) {
    return -2;
  }
#define SPEC(a) (spec + (size_t)((a) + 1) * N) /* a >= -1 */

  downmix(in, n_frames, cfg->channels, mono);
  oracle_hamming(W, window);

  /* speedy.c:263-267, 287-292 */
  const float mean_spectrogram_energy = 2.14204;
  const float mean_emphasis_weighted_local_difference = 123.837;
  const float mean_emphasis_weighted_lpf = 123.979;
  const float mean_relative_spectral_difference = 0.971975;
  const float max_energy_hysteresis = 1.41421;
  const float alpha = exp(-1.0 / (float)FRAME_RATE_HZ); /* speedy.c:67 */
  float energy_lp_state = mean_spectrogram_energy;
  float diff_lp_state = mean_emphasis_weighted_local_difference;
  float preemph_state = 0.0;

  /* ---- AddData time: one pass over the analysis frames (speedy.c:553-565) */
  for (int k = 0; k < nA; k++) {
    const int at = k + 1; /* soniclib.c:295-296 */
    const short* x = mono + (size_t)k * S;
    for (int i = 0; i < W; i++) input[i] = x[i] / 32768.0; /* speedy.c:558 */
    /* speedy.c:416-425: the state carried in is the last sample of the
     * previous (overlapping) window. */
    for (int i = 0; i < W; i++) {
      float last_sample = input[i];
      input[i] = 1.0 * input[i] - 0.97 * preemph_state;
      preemph_state = last_sample;
    }
    float* s = SPEC(at);
    spectrogram_frame(&plan, input, window, W, s);
    if (taps->spectrogram) {
      memcpy(taps->spectrogram + (size_t)k * N, s, sizeof(float) * N);
    }
    /* speedy.c:510-523 */
    float e = 0.0;
    for (int i = 1; i < half; i++) e += s[i] * s[i];
    energy_lp_state = (1 - alpha) * e + alpha * energy_lp_state; /* :73-76 */
    E[at] = e;
    lp[at] = energy_lp_state;
    local[at] = e / energy_lp_state;
    comp[at] = sqrt(local[at] > 2 ? 2.0 : local[at]);
  }

  /* ---- ComputeTension time (speedy.c:752-766, 664-729) and speed
   * (speedy.c:768-788, soniclib.c:339-345) */
  float current_duration = 0.0, desired_duration = 0.0;
  const float Rg = cfg->speed;
  for (int r = 0; r < nT; r++) {
    float f[ORACLE_FEATURES];
    memset(f, 0, sizeof(f));
    const int a_now = r + F; /* at_time of the frame added just before */
    /* features written at AddData time keep the newest frame's values */
    f[1] = lp[a_now];
    f[2] = local[a_now];
    f[3] = comp[a_now];
    f[12] = a_now;

    const float* cur = SPEC(r);
    const float* last = SPEC(r - 1);
    float hyst = hysteresis(comp, nA, r, F, B);
    f[4] = hyst;
    f[0] = normalize_by_energy(cur, norm, half);
    normalize_by_energy(last, norm_last, half);
    if (taps->normalized) {
      /* soniclib.c:303-310 exposes N floats; only the first N/2 are defined */
      memcpy(taps->normalized + (size_t)r * half, norm, sizeof(float) * half);
    }
    f[14] = 0.04 * max_energy_hysteresis;
    f[5] = f[0] <= f[14];
    f[13] = r;
    /* speedy.c:685-703: skip_frame_count starts at 1 and is re-armed by every
     * low-energy frame, then consumed at once; frame 0 is always low energy
     * (all-zero spectrum), so "skipped" == "low energy". */
    if (f[5]) {
      f[5] = 1;
      f[6] = f[7] = f[9] = f[10] = 0;
      diff_lp_state = (1 - alpha) * 0.0f + alpha * diff_lp_state;
      f[8] = diff_lp_state;
    } else {
      float bin_threshold = 0;
      for (int i = 1; i < half; i++) bin_threshold = fmax(bin_threshold, cur[i]);
      bin_threshold /= 100.0;
      float lsd = 0.0;
      const float eps = 2.2204e-16;
      for (int i = 1; i < half; i++) {
        if (cur[i] > bin_threshold && last[i] > bin_threshold) {
          lsd += fabs(log((norm[i] + eps) / (norm_last[i] + eps)));
        }
      }
      f[6] = lsd;
      f[7] = lsd * hyst;
      diff_lp_state = (1 - alpha) * f[7] + alpha * diff_lp_state;
      f[8] = diff_lp_state;
      f[9] = f[7] / (f[8] + 0.01 * mean_emphasis_weighted_lpf);
      f[10] = fmin(f[9], 4 * mean_relative_spectral_difference);
    }
    /* speedy.c:754-762 */
    float a = 1 / 2.0, b = 1 / 4.0, M_E_ = 0.7, M_S = 1.0;
    float tension = a * (hyst - M_E_) + b * (f[10] - M_S);
    f[11] = tension;

    /* speedy.c:768-788 */
    float requested_speed;
    if (Rg > 1.0) {
      requested_speed = fmax(1, Rg + (1 - Rg) * tension);
    } else {
      requested_speed = fmax(MINIMUM_SPEED, fmin(1, Rg - (1 - Rg) * tension));
    }
    if (cfg->feedback > 0) {
      float excess_duration = current_duration - desired_duration;
      requested_speed += fmax(MINIMUM_SPEED, cfg->feedback * excess_duration);
    }
    float frame_duration = 1.0 / FRAME_RATE_HZ;
    current_duration += frame_duration / requested_speed;
    desired_duration += frame_duration / Rg;
    /* soniclib.c:343-345 */
    float new_rate =
        requested_speed * cfg->nonlinear + Rg * (1 - cfg->nonlinear);

    if (taps->features) {
      memcpy(taps->features + (size_t)r * ORACLE_FEATURES, f, sizeof(f));
    }
    if (taps->tension) taps->tension[r] = tension;
    if (taps->speed) taps->speed[r] = new_rate;
  }
  if (taps->energy) {
    for (int k = 0; k < nA; k++) taps->energy[k] = E[k + 1];
  }
#undef SPEC
  spec_plan_free(&plan);
  free(mono); free(window); free(input); free(spec); free(E); free(lp);
  free(local); free(comp); free(norm); free(norm_last);
  return 0;
}

long oracle_resynthesize(const oracle_cfg* cfg, const short* in, long n_frames,
                         const float* speeds, int n_speeds, int flush,
                         short* out, long out_cap) {
  oracle_geom g;
  oracle_geometry(cfg->rate, cfg->match_matlab, &g);
  const int S = g.step, C = cfg->channels;
  sonicStream s = sonicIntCreateStream(cfg->rate, C);
  if (!s) return -2;
  sonicIntSetSpeed(s, cfg->speed); /* soniclib.c:177-183 */
  long produced = 0;
  short* tmp = (short*)malloc(sizeof(short) * (size_t)C * 4096);
#define DRAIN()                                                              \
  for (;;) {                                                                 \
    int got = sonicIntReadShortFromStream(s, tmp, 4096);                     \
    if (got <= 0) break;                                                     \
    long room = out_cap - produced;                                          \
    long take = got < room ? got : (room > 0 ? room : 0);                    \
    memcpy(out + produced * C, tmp, sizeof(short) * (size_t)take * C);       \
    produced += got; /* keeps counting past the capacity */                 \
  }
  if (cfg->nonlinear == 0) {
    /* soniclib.c:397-399: the shim is bypassed entirely */
    sonicIntWriteShortToStream(s, in, (int)n_frames);
    DRAIN();
  } else {
    /* soniclib.c:354, 369-371: one delayed 10 ms buffer per speed value */
    for (int r = 0; r < n_speeds; r++) {
      sonicIntSetSpeed(s, speeds[r]);
      sonicIntWriteShortToStream(s, in + (size_t)r * S * C, S);
      DRAIN();
    }
    if (flush) {
      /* soniclib.c:538-550: remaining complete buffers at the last speed; the
       * partial buffer being filled is dropped. */
      long write_index = n_frames / S;
      for (long r = n_speeds; r < write_index; r++) {
        sonicIntWriteShortToStream(s, in + (size_t)r * S * C, S);
        DRAIN();
      }
    }
  }
  if (flush) {
    sonicIntFlushStream(s); /* soniclib.c:551 */
    DRAIN();
  }
#undef DRAIN
  free(tmp);
  sonicIntDestroyStream(s);
  return produced;
}

long oracle_process(const oracle_cfg* cfg, const short* in, long n_frames,
                    const float* speed_override, short* out, long out_cap,
                    oracle_taps* taps) {
  oracle_taps local_taps;
  float* own_speed = NULL;
  oracle_geom g;
  oracle_geometry(cfg->rate, cfg->match_matlab, &g);
  int nA = oracle_frames_analyzed(&g, n_frames);
  int nT = oracle_tensions_ready(&g, nA);
  if (cfg->nonlinear == 0) {
    return oracle_resynthesize(cfg, in, n_frames, NULL, 0, 1, out, out_cap);
  }
  if (!taps) {
    memset(&local_taps, 0, sizeof(local_taps));
    taps = &local_taps;
    taps->max_frames = nA;
  }
  if (!taps->speed) {
    own_speed = (float*)malloc(sizeof(float) * (size_t)(nT > 0 ? nT : 1));
    taps->speed = own_speed;
  }
  int rc = oracle_analyze(cfg, in, n_frames, taps);
  long produced = rc;
  if (rc == 0) {
    produced = oracle_resynthesize(cfg, in, n_frames,
                                   speed_override ? speed_override : taps->speed,
                                   nT, 1, out, out_cap);
  }
  if (own_speed) {
    free(own_speed);
    taps->speed = NULL;
  }
  return produced;
}

/* ---- many streams over OS threads (CPU baseline, kind "port") ---------- */
#include <pthread.h>

typedef struct {
  const oracle_cfg* cfg;
  const short* in;
  long n_frames;
  int n_streams;
  short* out;
  long out_cap;
  long* out_counts;
  int tid, n_threads;
} batch_job;

static void* batch_worker(void* arg) {
  batch_job* j = (batch_job*)arg;
  const int C = j->cfg->channels;
  for (int i = j->tid; i < j->n_streams; i += j->n_threads) {
    j->out_counts[i] =
        oracle_process(j->cfg, j->in + (size_t)i * j->n_frames * C, j->n_frames,
                       NULL, j->out + (size_t)i * j->out_cap * C, j->out_cap, NULL);
  }
  return NULL;
}

int oracle_process_batch(const oracle_cfg* cfg, const short* in, long n_frames,
                         int n_streams, short* out, long out_cap,
                         long* out_counts, int n_threads) {
  if (n_threads < 1) n_threads = 1;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * n_threads);
  batch_job* jobs = (batch_job*)malloc(sizeof(batch_job) * n_threads);
  for (int t = 0; t < n_threads; t++) {
    batch_job j = {cfg, in, n_frames, n_streams, out, out_cap, out_counts, t, n_threads};
    jobs[t] = j;
    pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
  }
  for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
  free(th);
  free(jobs);
  return 0;
}

/* ---- synthetic input (same integer generator as the CUDA side) --------- */
#include "../speedy_b200/csrc/synth.h"

void oracle_synth_fill(short* dst, unsigned long long first_id, int n_streams,
                       int rate, int channels, long n_frames) {
  for (int s = 0; s < n_streams; s++) {
    short* p = dst + (size_t)s * n_frames * channels;
    for (long n = 0; n < n_frames; n++) {
      for (int c = 0; c < channels; c++) {
        p[n * channels + c] = synth_sample(first_id + s, rate, channels, c, n);
      }
    }
  }
}
