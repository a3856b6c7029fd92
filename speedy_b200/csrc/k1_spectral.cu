// K1 — per-frame windowed spectrogram, frame energy and raw spectral difference.
//
// Replaces, for every analysis window k of every stream:
//   speedyAddDataShort          speedy.c:553-565   int16 -> float (/32768)
//   speedyPreemphasisFilter     speedy.c:416-425   y = x - 0.97*state (double)
//   speedySpectrogram           speedy.c:438-454   Hamming, zero-pad, FFT, |X|
//   speedyComputeLocalEnergy    speedy.c:513-516   E = sum_{i=1}^{N/2-1} |X_i|^2
//   speedyNormalizeByEnergy     speedy.c:628-647   (both frames)
//   speedyComputeSpectralDiff.  speedy.c:705-719   thresholded sum |log ratio|
// The low-energy gate, both low-pass recurrences, hysteresis, tension and speed
// are K2 (k2_tension.cu): they need values from up to Future frames ahead.
//
// The local spectral difference of at_time a depends only on the spectra of
// windows a-1 and a-2, so it is computed here, eagerly, instead of Future
// frames later as the reference does; the spectrogram never goes to HBM
// (8 bytes per 10 ms frame do: energy and difference).
//
// 16 kHz fast path (N = 480): one warp owns a run of consecutive windows of one
// stream.  The zero-padded real 480-point transform is computed as two
// 120-point complex FFTs (even/odd output bins of the packed half-length
// sequence z[m] = v[2m] + i v[2m+1]) followed by the real-FFT split; each
// 120-point FFT is radix-8 x radix-15 (the 15 as a 3x5 prime-factor butterfly),
// entirely in registers with one shared-memory transpose.  Two windows are in
// flight per warp so that both butterfly passes fill the warp: 2 x 30 radix-8
// butterflies, then 32 radix-15 butterflies.
//
// Any other rate: k1_spectral_mixed (even windows whose half factors into radices up
// to 13: N = 660, 720, 960, 1440 ...: the same packed decomposition, the transforms
// as a Stockham FFT in shared memory, one CTA per run), k1_spectral_bluestein (odd or
// prime windows, 44.1 kHz: chirp-z on the same machinery).

#include <stdlib.h>

#include "kernels.cuh"

namespace speedy {

namespace {

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cscale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }
// a * (-i)
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }

__device__ __forceinline__ void dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
  float2 t0 = cadd(a0, a2), t1 = csub(a0, a2);
  float2 t2 = cadd(a1, a3), t3 = mul_mi(csub(a1, a3));
  a0 = cadd(t0, t2);
  a1 = cadd(t1, t3);
  a2 = csub(t0, t2);
  a3 = csub(t1, t3);
}

// In-place forward 8-point DFT, natural order in and out.
__device__ __forceinline__ void dft8(float2 (&a)[8]) {
  float2 e0 = a[0], e1 = a[2], e2 = a[4], e3 = a[6];
  float2 o0 = a[1], o1 = a[3], o2 = a[5], o3 = a[7];
  dft4(e0, e1, e2, e3);
  dft4(o0, o1, o2, o3);
  const float h = 0.70710678118654752440f;
  // W8^1 = h(1 - i), W8^2 = -i, W8^3 = h(-1 - i)
  float2 w1 = make_float2(h * (o1.x + o1.y), h * (o1.y - o1.x));
  float2 w2 = mul_mi(o2);
  float2 w3 = make_float2(h * (o3.y - o3.x), -h * (o3.x + o3.y));
  a[0] = cadd(e0, o0);  a[4] = csub(e0, o0);
  a[1] = cadd(e1, w1);  a[5] = csub(e1, w1);
  a[2] = cadd(e2, w2);  a[6] = csub(e2, w2);
  a[3] = cadd(e3, w3);  a[7] = csub(e3, w3);
}

__device__ __forceinline__ void dft3(float2& a0, float2& a1, float2& a2) {
  const float s = 0.86602540378443864676f;  // sin(2pi/3)
  float2 t = cadd(a1, a2);
  float2 m = make_float2(a0.x - 0.5f * t.x, a0.y - 0.5f * t.y);
  float2 d = cscale(csub(a1, a2), s);
  float2 mi = mul_mi(d);  // -i * s * (a1 - a2)
  a0 = cadd(a0, t);
  a1 = cadd(m, mi);
  a2 = csub(m, mi);
}

__device__ __forceinline__ void dft5(float2& a0, float2& a1, float2& a2, float2& a3, float2& a4) {
  const float c1 = 0.30901699437494742410f;   // cos(2pi/5)
  const float c2 = -0.80901699437494742410f;  // cos(4pi/5)
  const float s1 = 0.95105651629515357212f;   // sin(2pi/5)
  const float s2 = 0.58778525229247312917f;   // sin(4pi/5)
  float2 t1 = cadd(a1, a4), t2 = cadd(a2, a3);
  float2 t3 = csub(a1, a4), t4 = csub(a2, a3);
  float2 m1 = make_float2(a0.x + c1 * t1.x + c2 * t2.x, a0.y + c1 * t1.y + c2 * t2.y);
  float2 m2 = make_float2(a0.x + c2 * t1.x + c1 * t2.x, a0.y + c2 * t1.y + c1 * t2.y);
  float2 u1 = mul_mi(make_float2(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y));
  float2 u2 = mul_mi(make_float2(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y));
  a0 = make_float2(a0.x + t1.x + t2.x, a0.y + t1.y + t2.y);
  a1 = cadd(m1, u1);
  a4 = csub(m1, u1);
  a2 = cadd(m2, u2);
  a3 = csub(m2, u2);
}

// Forward 15-point DFT as a 3 x 5 prime-factor (Good-Thomas) butterfly:
// n = (5 n1 + 3 n2) mod 15, k = (10 k1 + 6 k2) mod 15, no inner twiddles.
__device__ __forceinline__ void dft15(const float2 (&in)[15], float2 (&out)[15]) {
  float2 t[3][5];
#pragma unroll
  for (int n2 = 0; n2 < 5; n2++) {
    float2 x0 = in[(3 * n2) % 15], x1 = in[(5 + 3 * n2) % 15], x2 = in[(10 + 3 * n2) % 15];
    dft3(x0, x1, x2);
    t[0][n2] = x0; t[1][n2] = x1; t[2][n2] = x2;
  }
#pragma unroll
  for (int k1 = 0; k1 < 3; k1++) {
    dft5(t[k1][0], t[k1][1], t[k1][2], t[k1][3], t[k1][4]);
#pragma unroll
    for (int k2 = 0; k2 < 5; k2++) out[(10 * k1 + 6 * k2) % 15] = t[k1][k2];
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// max over the warp of non-negative floats (their bit patterns order like unsigned integers)
__device__ __forceinline__ float warp_max(float v) {
  return __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(v)));
}

// ---------------------------------------------------------------------------
// 16 kHz fast path
// ---------------------------------------------------------------------------
constexpr int W16 = 240;            // window
constexpr int S16 = 160;            // step
constexpr int P16 = 80;             // window - step
constexpr int H16 = 240;            // N/2 bins kept
constexpr int kRun = K1_RUN;        // new windows per warp (plus one halo)
constexpr int kSampN = (kRun + 1) * S16 + P16;
static_assert(W16 == 30 * 8 && W16 == S16 + P16, "pass 0 gives eight samples to each of 30 lanes");
constexpr float kPreHi = 0.97f;                          // speedy.c:422
constexpr float kPreLo = (float)(0.97 - (double)0.97f);  // remainder of the double constant

struct WarpSmem480 {
  float2 z[2][240];      // per slot: windowed input (aliased), then transposes
  float lmag[3][240];    // log2 |X|^2 of the previous window, slot A, slot B
  short samp[kSampN + 8];
};

}  // namespace

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k1_spectral_480(K1Params p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* s_win = reinterpret_cast<float*>(smem_raw);                 // [240]
  float2* s_tw480 = reinterpret_cast<float2*>(s_win + 240);         // [240]
  WarpSmem480* s_warp = reinterpret_cast<WarpSmem480*>(s_tw480 + 240);

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 240; i += WARPS * 32) {
    s_win[i] = p.window[i] * 3.0517578125e-05f;  // Hamming / 32768: exact scaling
    s_tw480[i] = p.tw_n[i];  // W_480^k
  }
  __syncthreads();

  const long long item = (long long)blockIdx.x * WARPS + warp;
  const int s = (int)(item / p.runs_per_stream);
  const int run = (int)(item % p.runs_per_stream);
  if (s >= p.n_streams) return;

  const Geometry& g = p.g;
  const Range rg = write_range(p.st.total, p.counts, p.frames, p.done, s);
  const long long t_old = rg.t_old, t_new = rg.t_new;
  const int kA = frames_analyzed(g, t_old);          // scratch rows count from here
  const int kD = frames_analyzed(g, rg.t_done);      // first window of this launch
  const int kB = frames_analyzed(g, t_new);
  const int k0 = kD + run * kRun;  // first new window of this run
  if (k0 >= kB) return;
  const int k1 = min(k0 + kRun, kB);

  Source src;
  src.channels = g.channels;
  src.hist = p.hist + (size_t)s * p.hist_stride;
  src.in = p.in ? p.in + (size_t)s * p.in_stride_frames * g.channels : nullptr;
  src.hist_base = p.st.hist_base[s];
  src.t_old = t_old;
  src.t_new = t_new;

  WarpSmem480& ws = s_warp[warp];

  // Stage the run's samples (mono down-mix) once: windows k0-1 .. k1-1.
  const long long base = (long long)(k0 - 1) * S16;
  const int need = (k1 - k0 + 1) * S16 + P16;
  stage_mono<32, short>(src, base, need, t_new, ws.samp, nullptr, lane);
  __syncwarp();

  // Per-lane constants.
  // pass 1: lane = q*16 + m2 (m2 < 15) handles the radix-8 butterfly of residue m2 of FFT q:
  // one FFT per half-warp, so that the 8-byte stores of a half-warp stay inside one FFT's
  // rows (no bank conflict between the two)
  const int q1 = lane >> 4, m2 = lane & 15;
  const bool p1_live = m2 < 15;
  float2 tw[8];  // W_240^{m2 (2 j1 + q)}
#pragma unroll
  for (int j1 = 0; j1 < 8; j1++) tw[j1] = p.tw_half[(m2 * (2 * j1 + q1)) % 240];  // (m2 = 15: unused)
  // pass 2: lane = slot*16 + q*8 + j1 handles one radix-15 butterfly
  const int slot2 = lane >> 4, q2 = (lane >> 3) & 1, j1_2 = lane & 7;

  // real-FFT split: bins t = 1 + lane + 32 j, j < 4; in the last step lane 23 lands on
  // t = 120, lane 24 takes t = 0, lanes 25-31 idle
  int split_t[4];
  float2 split_w[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    int t = 1 + lane + 32 * j;
    if (t == 121) t = 0;
    else if (t > 121) t = -1;
    split_t[j] = t;
    split_w[j] = s_tw480[t < 0 ? 0 : t];
  }

  int ip = 0, ia = 1, ib = 2;  // rotating indices into ws.mag
  float linv_prev = 0.0f;

  // windows are processed in pairs (kk, kk+1), starting at the halo k0-1
  for (int kk = k0 - 1; kk < k1; kk += 2) {
    // ---- pass 0: int16 -> float, pre-emphasis, Hamming ---------------------
    // lane l < 30 takes samples 4l .. 4l+3 and 120 + 4l .. 120 + 4l+3: 8-byte loads of
    // samples, 16-byte loads of the window table and 16-byte stores, all at consecutive
    // addresses across the warp (no bank conflicts)
#pragma unroll
    for (int slot = 0; slot < 2; slot++) {
      const int k = kk + slot;
      const int o = (k - (k0 - 1)) * S16;
      float4* v4 = reinterpret_cast<float4*>(ws.z[slot]) + lane;
      const bool live = (k >= 0) && (k < k1);
      if (lane < 30) {
        float4 o0 = make_float4(0.0f, 0.0f, 0.0f, 0.0f), o1 = o0;
        if (live) {
          const short* sp = ws.samp + o + 4 * lane;
          const int2 ra = *reinterpret_cast<const int2*>(sp);
          const int2 rb = *reinterpret_cast<const int2*>(sp + 120);
          // state entering sample 0 is the last sample of the previous window,
          // i.e. sample P-1 of this one (speedy.c:416-425; window k-1 ends at
          // k*S + P - 1); 0 before the first window.
          const float xm = (float)(lane > 0 ? sp[-1] : (k >= 1 ? ws.samp[o + P16 - 1] : 0));
          const float xn = (float)sp[119];
          const float x0 = (float)(short)(ra.x & 0xffff), x1 = (float)(ra.x >> 16);
          const float x2 = (float)(short)(ra.y & 0xffff), x3 = (float)(ra.y >> 16);
          const float x4 = (float)(short)(rb.x & 0xffff), x5 = (float)(rb.x >> 16);
          const float x6 = (float)(short)(rb.y & 0xffff), x7 = (float)(rb.y >> 16);
          const float4 w0 = *reinterpret_cast<const float4*>(s_win + 4 * lane);
          const float4 w1 = *reinterpret_cast<const float4*>(s_win + 120 + 4 * lane);
          // y = x - 0.97 * state (speedy.c:422, evaluated there in double): 0.97 is
          // split into a float and its remainder so the constant carries no error;
          // the /32768 of speedy.c:558 is folded into the window table.
          o0.x = __fmul_rn(__fmaf_rn(-kPreLo, xm, __fmaf_rn(-kPreHi, xm, x0)), w0.x);
          o0.y = __fmul_rn(__fmaf_rn(-kPreLo, x0, __fmaf_rn(-kPreHi, x0, x1)), w0.y);
          o0.z = __fmul_rn(__fmaf_rn(-kPreLo, x1, __fmaf_rn(-kPreHi, x1, x2)), w0.z);
          o0.w = __fmul_rn(__fmaf_rn(-kPreLo, x2, __fmaf_rn(-kPreHi, x2, x3)), w0.w);
          o1.x = __fmul_rn(__fmaf_rn(-kPreLo, xn, __fmaf_rn(-kPreHi, xn, x4)), w1.x);
          o1.y = __fmul_rn(__fmaf_rn(-kPreLo, x4, __fmaf_rn(-kPreHi, x4, x5)), w1.y);
          o1.z = __fmul_rn(__fmaf_rn(-kPreLo, x5, __fmaf_rn(-kPreHi, x5, x6)), w1.z);
          o1.w = __fmul_rn(__fmaf_rn(-kPreLo, x6, __fmaf_rn(-kPreHi, x6, x7)), w1.w);
        }
        v4[0] = o0;
        v4[30] = o1;
      }
    }
    __syncwarp();

    // ---- pass 1: radix-8 butterflies, 30 lanes, one slot at a time --------
#pragma unroll
    for (int slot = 0; slot < 2; slot++) {
      float2 a[8];
      if (p1_live) {
        const float2* zin = ws.z[slot];
#pragma unroll
        for (int m1 = 0; m1 < 8; m1++) a[m1] = zin[15 * m1 + m2];  // z[m] = (v[2m], v[2m+1])
      }
      __syncwarp();
      if (p1_live) {
        if (q1) {
          // odd output bins: z[m] * W_240^m = z[m] * W_16^{m1} * W_240^{m2};
          // W_240^{m2} is folded into tw[], W_16^{m1} applied here.
          const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f;
          const float h = 0.70710678118654752440f;
          a[1] = cmul(a[1], make_float2(c1, -s1));
          a[2] = cmul(a[2], make_float2(h, -h));
          a[3] = cmul(a[3], make_float2(s1, -c1));
          a[4] = mul_mi(a[4]);
          a[5] = cmul(a[5], make_float2(-s1, -c1));
          a[6] = cmul(a[6], make_float2(-h, -h));
          a[7] = cmul(a[7], make_float2(-c1, -s1));
        }
        dft8(a);
        float2* zout = ws.z[slot];
#pragma unroll
        for (int j1 = 0; j1 < 8; j1++) zout[(q1 * 8 + j1) * 15 + m2] = cmul(a[j1], tw[j1]);
      }
    }
    __syncwarp();

    // ---- pass 2: radix-15 butterflies, 32 lanes ----------------------------
    {
      float2 b[15], c[15];
      const float2* zin = ws.z[slot2] + (q2 * 8 + j1_2) * 15;
#pragma unroll
      for (int i = 0; i < 15; i++) b[i] = zin[i];
      dft15(b, c);
      __syncwarp();
      // Z[kbin], kbin = 2 (j1 + 8 j2) + q : even bins from FFT 0, odd from FFT 1
      float2* zout = ws.z[slot2];
#pragma unroll
      for (int j2 = 0; j2 < 15; j2++) zout[2 * (j1_2 + 8 * j2) + q2] = c[j2];
    }
    __syncwarp();

    // ---- real-FFT split, power, energy -------------------------------------
    // The features need |X|^2 (energy) and log|X| (difference), never |X| itself:
    // keep p = re^2 + im^2 and log2 p; the square root is taken only for the tap.
    float e_slot[2], pmax_slot[2];
#pragma unroll
    for (int slot = 0; slot < 2; slot++) {
      const int k = kk + slot;
      const float2* Z = ws.z[slot];
      float* lmag = ws.lmag[slot == 0 ? ia : ib];
      float* tap = nullptr;
      if (p.tap_spec && k >= k0 && k < k1) tap = p.tap_spec + ((size_t)s * p.tap_stride + (k - kA)) * 480;
      float e = 0.0f, mx = 0.0f;
      // Real-FFT split, two bins per step.  With S = (Z[t] + conj Z[240-t]) / 2,
      // D = (Z[t] - conj Z[240-t]) / 2 and P = W_480^t D:
      //   X[t] = S - i P,   X[240-t] = conj(S) - i conj(P)      (W_480^(240-t) = -conj W_480^t)
      // The halves are left out: everything below is 2 X, its square 4 |X|^2 (exact
      // scalings), which only shifts every log2 by the same 2 and is undone once for the
      // energy.  t = 120 pairs with itself and t = 0 with Z[240] = Z[0] (its partner is
      // bin N/2, kept for the tap only): lanes 23 and 24 of the last step take them.
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int t = split_t[j];
        if (t < 0) continue;
        const int u = t == 0 ? 0 : 240 - t;
        const float2 zk = Z[t], zm = Z[u];
        const float sx = zk.x + zm.x, sy = zk.y - zm.y;
        const float dx = zk.x - zm.x, dy = zk.y + zm.y;
        const float2 w = split_w[j];
        const float px = w.x * dx - w.y * dy, py = w.x * dy + w.y * dx;
        const float re0 = sx + py, im0 = sy - px;    // 2 X[t]
        const float re1 = sx - py, im1 = sy + px;    // 2 X[240-t] up to the sign of im
        // speedy.c:434-436 squares and sums in float; so does this
        const float q0 = __fadd_rn(__fmul_rn(re0, re0), __fmul_rn(im0, im0));
        const float q1 = __fadd_rn(__fmul_rn(re1, re1), __fmul_rn(im1, im1));
        lmag[t] = __log2f(q0);
        if (t != 0) lmag[240 - t] = __log2f(q1);
        const float c0 = t >= 1 ? q0 : 0.0f;               // bin 0 is not part of the energy
        const float c1 = (t >= 1 && t < 120) ? q1 : 0.0f;  // bin N/2 neither; bin 120 counts once
        e += c0 + c1;
        mx = fmaxf(mx, fmaxf(c0, c1));
        if (tap) {
          const float m0 = 0.5f * __fsqrt_rn(q0), m1 = 0.5f * __fsqrt_rn(q1);
          tap[t] = m0;
          if (t != 0) tap[480 - t] = m0;
          tap[240 - t] = m1;
          tap[240 + t] = m1;
        }
      }
      e_slot[slot] = 0.25f * warp_sum(e);
      pmax_slot[slot] = warp_max(mx);  // 4 max |X|^2: the same shift as the stored log2 values
    }
    __syncwarp();

    // ---- spectral difference against the previous window, outputs ---------
    // speedy.c:705-719 in the log2 domain: with n_i = |X_i| / (sqrt(E) + eps),
    //   log(n_c / n_l) = ln2 * (0.5 (lp_c - lp_l) + (linv_c - linv_l)),
    //   lp = log2 |X|^2, linv = -log2(sqrt(E) + eps);
    // |X_i| > max|X| / 100 (speedy.c:709, 714)  <=>  lp_i > log2(max p) - log2(1e4).
    float linv_slot[2];
#pragma unroll
    for (int slot = 0; slot < 2; slot++) {
      const int k = kk + slot;
      linv_slot[slot] = -__log2f(__fsqrt_rn(e_slot[slot]) + 2.2204e-16f);
      const float* lc = ws.lmag[slot == 0 ? ia : ib];
      const float* ll = ws.lmag[slot == 0 ? ip : ia];
      const float linv_last = slot == 0 ? linv_prev : linv_slot[0];
      if (k >= k0 && k < k1) {
        const float thr = __log2f(pmax_slot[slot]) - 13.287712379549449f;  // log2(1e4)
        const float d2 = 2.0f * (linv_slot[slot] - linv_last);
        float acc = 0.0f;
        for (int j = lane; j < H16 / 4; j += 32) {  // bins 4j .. 4j+3, bin 0 excluded
          const float4 c = *reinterpret_cast<const float4*>(lc + 4 * j);
          const float4 l = *reinterpret_cast<const float4*>(ll + 4 * j);
          if (j > 0 && c.x > thr && l.x > thr) acc += fabsf((c.x - l.x) + d2);
          if (c.y > thr && l.y > thr) acc += fabsf((c.y - l.y) + d2);
          if (c.z > thr && l.z > thr) acc += fabsf((c.z - l.z) + d2);
          if (c.w > thr && l.w > thr) acc += fabsf((c.w - l.w) + d2);
        }
        const float lsd = warp_sum(acc) * 0.34657359027997264f;  // ln2 / 2
        const int j = k - kA;
        if (lane == 0) p.feat[(size_t)s * p.feat_stride + j] = make_float2(e_slot[slot], lsd);
      }
    }
    linv_prev = linv_slot[1];
    // rotate: slot B becomes "previous"
    int t = ip; ip = ib; ib = ia; ia = t;
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------
// Mixed-radix path: any even window W whose half M = W/2 factors into radices
// <= 16 (22.05 kHz: M = 165 = 5*3*11, 24 kHz: 180, 32 kHz: 240, 48 kHz: 360 ...).
// One CTA per run of windows, two windows in flight.  The same decomposition as
// the 16 kHz kernel: the zero-padded real N = 2W transform is the real-FFT split
// of the W-point complex transform of z[m] = v[2m] + i v[2m+1] (m < M, zero
// beyond), whose even / odd output bins are the M-point transforms of z[m] and
// of z[m] W_W^m.  The four M-point transforms of a window pair run together as
// an autosort (Stockham) FFT in shared memory, one thread per radix-R butterfly.
// ---------------------------------------------------------------------------
namespace {

template <int R>
__device__ __forceinline__ void small_dft(float2 (&a)[R], const float2* twM, int M) {
  if (R == 2) {
    const float2 t = a[1];
    a[1] = csub(a[0], t);
    a[0] = cadd(a[0], t);
  } else if (R == 3) {
    dft3(a[0], a[1], a[2]);
  } else if (R == 4) {
    dft4(a[0], a[1], a[2], a[3]);
  } else if (R == 5) {
    dft5(a[0], a[1], a[2], a[3], a[4]);
  }
}

// One radix-R butterfly of the stage (n = R m, stride st): inputs x[q + st (pp + r m)],
// outputs y[q + st (R pp + t)] = W_n^{pp t} sum_r x_r W_R^{r t};  W_n^{pp t} = W_M^{st pp t}.
template <int R>
__device__ __forceinline__ void butterfly(const float2* x, float2* y, const float2* twM, int M, int m, int st,
                                          int pp, int q) {
  float2 a[R];
#pragma unroll
  for (int r = 0; r < R; r++) a[r] = x[q + st * (pp + r * m)];
  if constexpr (R == 8) dft8(a);
  else small_dft<R>(a, twM, M);
  const int tstep = st * pp;
  y[q + st * (R * pp)] = a[0];
#pragma unroll
  for (int t = 1; t < R; t++) y[q + st * (R * pp + t)] = cmul(a[t], twM[tstep * t]);
}

// Odd prime radix R (7, 11, 13) in registers: with S_r = a_r + a_{R-r}, D_r = a_r - a_{R-r},
//   y_t, y_{R-t} = (a_0 + sum_r S_r cos(2 pi r t / R))  -/+  i sum_r D_r sin(2 pi r t / R),
// r, t = 1 .. (R-1)/2, which halves the multiplications of the direct form.
template <int R>
__device__ __forceinline__ void butterfly_prime(const float2* x, float2* y, const float2* twM, int M, int m, int st,
                                                int pp, int q) {
  constexpr int H = (R - 1) / 2;
  float2 a[R];
#pragma unroll
  for (int r = 0; r < R; r++) a[r] = x[q + st * (pp + r * m)];
  float c[H + 1], sn[H + 1];
  const int wr = M / R;
#pragma unroll
  for (int j = 1; j <= H; j++) {
    const float2 w = twM[wr * j];  // (cos, -sin)(2 pi j / R)
    c[j] = w.x;
    sn[j] = -w.y;
  }
  float2 S[H + 1], D[H + 1];
  float2 y0 = a[0];
#pragma unroll
  for (int r = 1; r <= H; r++) {
    S[r] = cadd(a[r], a[R - r]);
    D[r] = csub(a[r], a[R - r]);
    y0 = cadd(y0, S[r]);
  }
  const int tstep = st * pp;
  y[q + st * (R * pp)] = y0;
#pragma unroll
  for (int t = 1; t <= H; t++) {
    float2 P = a[0], Q = make_float2(0.0f, 0.0f);
#pragma unroll
    for (int r = 1; r <= H; r++) {
      const int e = (r * t) % R;
      const float cc = e <= H ? c[e] : c[R - e];
      const float ss = e <= H ? sn[e] : -sn[R - e];
      P.x = fmaf(S[r].x, cc, P.x);
      P.y = fmaf(S[r].y, cc, P.y);
      Q.x = fmaf(D[r].x, ss, Q.x);
      Q.y = fmaf(D[r].y, ss, Q.y);
    }
    const float2 yt = make_float2(P.x + Q.y, P.y - Q.x);   // P - i Q
    const float2 yr = make_float2(P.x - Q.y, P.y + Q.x);   // P + i Q
    y[q + st * (R * pp + t)] = cmul(yt, twM[tstep * t]);
    y[q + st * (R * pp + R - t)] = cmul(yr, twM[tstep * (R - t)]);
  }
}

// Any other (odd prime) radix up to 16: direct O(R^2) butterfly.
__device__ __forceinline__ void butterfly_any(int R, const float2* x, float2* y, const float2* twM, int M, int m,
                                              int st, int pp, int q) {
  float2 a[16];
  for (int r = 0; r < R; r++) a[r] = x[q + st * (pp + r * m)];
  const int wr = M / R;
  const int tstep = st * pp;
  for (int t = 0; t < R; t++) {
    float2 acc = a[0];
    int e = 0;
    for (int r = 1; r < R; r++) {
      e += t;
      if (e >= R) e -= R;
      acc = cadd(acc, cmul(a[r], twM[wr * e]));
    }
    y[q + st * (R * pp + t)] = cmul(acc, twM[tstep * t]);
  }
}

// nT transforms of length M (x[t * M + i]) as one autosort FFT, radices from the plan in
// p; ping-pongs between x and y and returns the buffer that holds the result.  Ends
// with a block barrier.
template <int THREADS>
__device__ __forceinline__ float2* stockham(float2* x, float2* y, int nT, int M, const K1Params& p,
                                            const float2* s_twM, int tid) {
  int st = 1;
  for (int f = 0; f < p.n_factors; f++) {
    const int R = p.factors[f];
    const int m = p.plan_m[f];      // sub-transform length after this stage: n / R
    const int per = p.plan_per[f];  // butterflies per transform = M / R = m * st
    // floor(a / b) == int((a + 0.5) / b) in float for these small integers
    const float rcp_per = __frcp_rn((float)per), rcp_st = __frcp_rn((float)st);
    for (int b = tid; b < nT * per; b += THREADS) {
      const int which = (int)(((float)b + 0.5f) * rcp_per);
      const int idx = b - which * per;
      const int pp = (int)(((float)idx + 0.5f) * rcp_st), q = idx - pp * st;
      const float2* xi = x + which * M;
      float2* yo = y + which * M;
      switch (R) {
        case 8: butterfly<8>(xi, yo, s_twM, M, m, st, pp, q); break;
        case 5: butterfly<5>(xi, yo, s_twM, M, m, st, pp, q); break;
        case 4: butterfly<4>(xi, yo, s_twM, M, m, st, pp, q); break;
        case 3: butterfly<3>(xi, yo, s_twM, M, m, st, pp, q); break;
        case 2: butterfly<2>(xi, yo, s_twM, M, m, st, pp, q); break;
        case 7: butterfly_prime<7>(xi, yo, s_twM, M, m, st, pp, q); break;
        case 11: butterfly_prime<11>(xi, yo, s_twM, M, m, st, pp, q); break;
        case 13: butterfly_prime<13>(xi, yo, s_twM, M, m, st, pp, q); break;
        default: butterfly_any(R, xi, yo, s_twM, M, m, st, pp, q); break;
      }
    }
    __syncthreads();
    float2* tmp = x; x = y; y = tmp;
    st *= R;
  }
  return x;
}

}  // namespace

template <int THREADS>
__global__ void __launch_bounds__(THREADS) k1_spectral_mixed(K1Params p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const Geometry& g = p.g;
  const int W = g.window, S = g.step, P = g.partial, N = g.fft, M = W / 2, HW = W / 2;
  constexpr int NWARP = THREADS / 32;
  float2* X = reinterpret_cast<float2*>(smem_raw);     // [4][M]
  float2* Y = X + 4 * M;                                // [4][M]
  float2* s_twM = Y + 4 * M;                            // [M]   W_M^k
  float2* s_twW = s_twM + M;                            // [M]   W_W^m
  float2* s_twN = s_twW + M;                            // [HW]  W_N^t
  float* s_win = reinterpret_cast<float*>(s_twN + HW);  // [W]   Hamming / 32768
  float* lmag = s_win + W;                              // [3][W]
  float* red = lmag + 3 * W;                            // [NWARP][4]
  short* samp = reinterpret_cast<short*>(red + NWARP * 4);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < M; i += THREADS) {
    s_twM[i] = p.tw_n[4 * i];
    s_twW[i] = p.tw_half[i];
  }
  for (int i = tid; i < HW; i += THREADS) s_twN[i] = p.tw_n[i];
  for (int i = tid; i < W; i += THREADS) s_win[i] = p.window[i] * 3.0517578125e-05f;

  const long long item = blockIdx.x;
  const int s = (int)(item / p.runs_per_stream);
  const int run = (int)(item % p.runs_per_stream);
  if (s >= p.n_streams) return;
  const Range rg = write_range(p.st.total, p.counts, p.frames, p.done, s);
  const long long t_old = rg.t_old, t_new = rg.t_new;
  const int kA = frames_analyzed(g, t_old);      // scratch rows count from here
  const int kD = frames_analyzed(g, rg.t_done);  // first window of this launch
  const int kB = frames_analyzed(g, t_new);
  const int k0 = kD + run * kRun;
  if (k0 >= kB) return;
  const int k1 = min(k0 + kRun, kB);

  Source src;
  src.channels = g.channels;
  src.hist = p.hist + (size_t)s * p.hist_stride;
  src.in = p.in ? p.in + (size_t)s * p.in_stride_frames * g.channels : nullptr;
  src.hist_base = p.st.hist_base[s];
  src.t_old = t_old;
  src.t_new = t_new;

  // Stage the run's samples (mono down-mix) once: windows k0-1 .. k1-1.  With disjoint windows
  // (P == 0, the white-box frame step) the pre-emphasis state of the first one lies before it.
  const int lead = P == 0 ? 8 : 0;
  const long long base = (long long)(k0 - 1) * S - lead;
  const int need = (k1 - k0 + 1) * S + P + lead;
  stage_mono<THREADS, short>(src, base, need, t_new, samp, nullptr, tid);
  __syncthreads();

  int ip = 0, ia = 1, ib = 2;  // rotating rows of lmag: previous window, slot A, slot B
  float linv_prev = 0.0f;

  for (int kk = k0 - 1; kk < k1; kk += 2) {
    // ---- pass 0: int16 -> float, pre-emphasis, Hamming, packing ------------
    for (int i = tid; i < 2 * M; i += THREADS) {
      const int slot = i >= M ? 1 : 0;
      const int m = i - slot * M;
      const int k = kk + slot;
      float2 z = make_float2(0.0f, 0.0f);
      if (k >= 0 && k < k1) {
        const int o = lead + (k - (k0 - 1)) * S;
        const int n = 2 * m;
        // state entering sample 0 is the last sample of the previous window, i.e.
        // sample P-1 of this one (speedy.c:416-425); 0 before the first window
        const float xm = (float)((n > 0) ? samp[o + n - 1] : (k >= 1 ? samp[o + P - 1] : 0));
        const float x0 = (float)samp[o + n];
        const float x1 = (float)samp[o + n + 1];
        const float y0 = __fmaf_rn(-kPreLo, xm, __fmaf_rn(-kPreHi, xm, x0));
        const float y1 = __fmaf_rn(-kPreLo, x0, __fmaf_rn(-kPreHi, x0, x1));
        z = make_float2(__fmul_rn(y0, s_win[n]), __fmul_rn(y1, s_win[n + 1]));
      }
      X[(2 * slot) * M + m] = z;
      X[(2 * slot + 1) * M + m] = cmul(z, s_twW[m]);
    }
    __syncthreads();

    // ---- the four M-point FFTs ----------------------------------------------
    float2* x = stockham<THREADS>(X, Y, 4, M, p, s_twM, tid);
    // Z[t] of slot: x[(2 slot + (t & 1)) M + (t >> 1)]

    // ---- real-FFT split, power, energy (as in the 16 kHz kernel) -------------
    float e0 = 0.0f, e1 = 0.0f, mx0 = 0.0f, mx1 = 0.0f;
    for (int i = tid; i < 2 * (HW - 1); i += THREADS) {
      const int slot = i >= HW - 1 ? 1 : 0;
      const int t = 1 + i - slot * (HW - 1);
      const int k = kk + slot;
      const float2* Z = x + 2 * slot * M;
      const int u = W - t;
      const float2 zk = Z[(t & 1) * M + (t >> 1)], zm = Z[(u & 1) * M + (u >> 1)];
      const float sx = 0.5f * (zk.x + zm.x), sy = 0.5f * (zk.y - zm.y);
      const float dx = 0.5f * (zk.x - zm.x), dy = 0.5f * (zk.y + zm.y);
      const float2 w = s_twN[t];
      const float px = w.x * dx - w.y * dy, py = w.x * dy + w.y * dx;
      const float re0 = sx + py, im0 = sy - px;  // X[t]
      const float re1 = sx - py, im1 = sy + px;  // X[W-t] up to the sign of im
      const float p0 = __fadd_rn(__fmul_rn(re0, re0), __fmul_rn(im0, im0));
      const float p1 = __fadd_rn(__fmul_rn(re1, re1), __fmul_rn(im1, im1));
      float* lm = lmag + (slot == 0 ? ia : ib) * W;
      lm[t] = __log2f(p0);
      lm[u] = __log2f(p1);
      if (slot == 0) { e0 += p0 + p1; mx0 = fmaxf(mx0, fmaxf(p0, p1)); }
      else { e1 += p0 + p1; mx1 = fmaxf(mx1, fmaxf(p0, p1)); }
      if (p.tap_spec && k >= k0 && k < k1) {
        float* tap = p.tap_spec + ((size_t)s * p.tap_stride + (k - kA)) * N;
        const float m0 = __fsqrt_rn(p0), m1 = __fsqrt_rn(p1);
        tap[t] = m0;
        tap[N - t] = m0;
        tap[u] = m1;
        tap[N - u] = m1;
      }
    }
    if (tid == 0 || tid == 32) {
      // bins 0, N/4 and N/2: X[0] = Re Z0 + Im Z0, X[N/2] = Re Z0 - Im Z0,
      // X[N/4] = conj(Z[W/2]) (W_N^(N/4) = -i)
      const int slot = tid == 0 ? 0 : 1;
      const int k = kk + slot;
      const float2* Z = x + 2 * slot * M;
      const float2 z0 = Z[0], zq = Z[(HW & 1) * M + (HW >> 1)];
      const float x0 = z0.x + z0.y;
      const float p0 = __fmul_rn(x0, x0);
      const float pq = __fadd_rn(__fmul_rn(zq.x, zq.x), __fmul_rn(zq.y, zq.y));
      float* lm = lmag + (slot == 0 ? ia : ib) * W;
      lm[0] = __log2f(p0);
      lm[HW] = __log2f(pq);
      if (slot == 0) { e0 += pq; mx0 = fmaxf(mx0, pq); }
      else { e1 += pq; mx1 = fmaxf(mx1, pq); }
      if (p.tap_spec && k >= k0 && k < k1) {
        float* tap = p.tap_spec + ((size_t)s * p.tap_stride + (k - kA)) * N;
        tap[0] = fabsf(x0);
        tap[W] = fabsf(z0.x - z0.y);
        const float mq = __fsqrt_rn(pq);
        tap[HW] = mq;
        tap[N - HW] = mq;
      }
    }
    e0 = warp_sum(e0);
    e1 = warp_sum(e1);
    mx0 = warp_max(mx0);
    mx1 = warp_max(mx1);
    if (lane == 0) {
      red[warp * 4 + 0] = e0;
      red[warp * 4 + 1] = e1;
      red[warp * 4 + 2] = mx0;
      red[warp * 4 + 3] = mx1;
    }
    __syncthreads();
    float e_slot[2] = {0.0f, 0.0f}, pmax_slot[2] = {0.0f, 0.0f};
#pragma unroll
    for (int w = 0; w < NWARP; w++) {
      e_slot[0] += red[w * 4 + 0];
      e_slot[1] += red[w * 4 + 1];
      pmax_slot[0] = fmaxf(pmax_slot[0], red[w * 4 + 2]);
      pmax_slot[1] = fmaxf(pmax_slot[1], red[w * 4 + 3]);
    }
    __syncthreads();  // red is reused below

    // ---- spectral difference against the previous window (log2 domain) -------
    float linv_slot[2];
    linv_slot[0] = -__log2f(__fsqrt_rn(e_slot[0]) + 2.2204e-16f);
    linv_slot[1] = -__log2f(__fsqrt_rn(e_slot[1]) + 2.2204e-16f);
    const float thr0 = __log2f(pmax_slot[0]) - 13.287712379549449f;  // log2(1e4)
    const float thr1 = __log2f(pmax_slot[1]) - 13.287712379549449f;
    const float d20 = 2.0f * (linv_slot[0] - linv_prev);
    const float d21 = 2.0f * (linv_slot[1] - linv_slot[0]);
    float acc0 = 0.0f, acc1 = 0.0f;
    {
      // bins 1 .. W-1 of both slots, four per 16-byte load (W is a multiple of 4 here,
      // otherwise one at a time)
      const float* ca = lmag + ia * W;
      const float* cb = lmag + ib * W;
      const float* la = lmag + ip * W;
      if ((W & 3) == 0) {
        const int nv = W >> 2;
        for (int i = tid; i < 2 * nv; i += THREADS) {
          const int slot = i >= nv ? 1 : 0;
          const int j = i - slot * nv;
          const float4 c = *reinterpret_cast<const float4*>((slot ? cb : ca) + 4 * j);
          const float4 l = *reinterpret_cast<const float4*>((slot ? ca : la) + 4 * j);
          const float thr = slot ? thr1 : thr0, d2 = slot ? d21 : d20;
          float a = 0.0f;
          if (j > 0 && c.x > thr && l.x > thr) a += fabsf((c.x - l.x) + d2);
          if (c.y > thr && l.y > thr) a += fabsf((c.y - l.y) + d2);
          if (c.z > thr && l.z > thr) a += fabsf((c.z - l.z) + d2);
          if (c.w > thr && l.w > thr) a += fabsf((c.w - l.w) + d2);
          if (slot) acc1 += a; else acc0 += a;
        }
      } else {
        for (int i = tid; i < 2 * (W - 1); i += THREADS) {
          const int slot = i >= W - 1 ? 1 : 0;
          const int bin = 1 + i - slot * (W - 1);
          const float c = (slot ? cb : ca)[bin], l = (slot ? ca : la)[bin];
          if (slot == 0) {
            if (c > thr0 && l > thr0) acc0 += fabsf((c - l) + d20);
          } else {
            if (c > thr1 && l > thr1) acc1 += fabsf((c - l) + d21);
          }
        }
      }
    }
    acc0 = warp_sum(acc0);
    acc1 = warp_sum(acc1);
    if (lane == 0) {
      red[warp * 4 + 0] = acc0;
      red[warp * 4 + 1] = acc1;
    }
    __syncthreads();
    if (tid < 2) {
      const int slot = tid;
      const int k = kk + slot;
      if (k >= k0 && k < k1) {
        float t = 0.0f;
        for (int w = 0; w < NWARP; w++) t += red[w * 4 + slot];
        p.feat[(size_t)s * p.feat_stride + (k - kA)] = make_float2(e_slot[slot], t * 0.34657359027997264f);  // ln2 / 2
      }
    }
    linv_prev = linv_slot[1];
    int t = ip; ip = ib; ib = ia; ia = t;  // slot B becomes "previous"
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// Chirp-z (Bluestein) path: windows the mixed-radix kernel cannot factor, e.g.
// 44.1 kHz where W = 661 is prime.  With c[n] = e^{-i pi n^2 / N},
//   X[k] = c[k] sum_n (v[n] c[n]) conj(c)[k - n],
// a convolution carried out with two L-point FFTs (L >= N, 7-smooth) per window
// on the same Stockham machinery; only |X[k]| is needed, so the final chirp
// multiplication and both conjugations of the inverse transform drop out.
// ---------------------------------------------------------------------------
template <int THREADS>
__global__ void __launch_bounds__(THREADS) k1_spectral_bluestein(K1Params p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const Geometry& g = p.g;
  const int W = g.window, S = g.step, P = g.partial, N = g.fft, L = p.bl_L, H = N / 2;
  const int HP = (H + 3) & ~3;  // row pitch of lmag
  constexpr int NWARP = THREADS / 32;
  float2* X = reinterpret_cast<float2*>(smem_raw);  // [2][L]
  float2* Y = X + 2 * L;                             // [2][L]
  float2* s_tw = Y + 2 * L;                          // [L]  W_L^k
  float2* s_B = s_tw + L;                            // [L]  FFT of the chirp filter / L
  float2* s_ch = s_B + L;                            // [W]  chirp
  float* s_win = reinterpret_cast<float*>(s_ch + W + (W & 1));  // [W]  Hamming / 32768
  float* lmag = s_win + ((W + 3) & ~3);              // [3][HP]
  float* red = lmag + 3 * HP;                        // [NWARP][4]
  short* samp = reinterpret_cast<short*>(red + NWARP * 4);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < L; i += THREADS) {
    s_tw[i] = p.bl_tw[i];
    s_B[i] = p.bl_B[i];
  }
  for (int i = tid; i < W; i += THREADS) {
    s_ch[i] = p.bl_chirp[i];
    s_win[i] = p.window[i] * 3.0517578125e-05f;
  }

  const long long item = blockIdx.x;
  const int s = (int)(item / p.runs_per_stream);
  const int run = (int)(item % p.runs_per_stream);
  if (s >= p.n_streams) return;
  const Range rg = write_range(p.st.total, p.counts, p.frames, p.done, s);
  const long long t_old = rg.t_old, t_new = rg.t_new;
  const int kA = frames_analyzed(g, t_old);      // scratch rows count from here
  const int kD = frames_analyzed(g, rg.t_done);  // first window of this launch
  const int kB = frames_analyzed(g, t_new);
  const int k0 = kD + run * kRun;
  if (k0 >= kB) return;
  const int k1 = min(k0 + kRun, kB);

  Source src;
  src.channels = g.channels;
  src.hist = p.hist + (size_t)s * p.hist_stride;
  src.in = p.in ? p.in + (size_t)s * p.in_stride_frames * g.channels : nullptr;
  src.hist_base = p.st.hist_base[s];
  src.t_old = t_old;
  src.t_new = t_new;

  const int lead = P == 0 ? 8 : 0;  // (disjoint windows: the first one's pre-emphasis state lies before it)
  const long long base = (long long)(k0 - 1) * S - lead;
  const int need = (k1 - k0 + 1) * S + P + lead;
  stage_mono<THREADS, short>(src, base, need, t_new, samp, nullptr, tid);
  __syncthreads();

  int ip = 0, ia = 1, ib = 2;  // rotating rows of lmag: previous window, slot A, slot B
  float linv_prev = 0.0f;

  for (int kk = k0 - 1; kk < k1; kk += 2) {
    // ---- pass 0: int16 -> float, pre-emphasis, Hamming, chirp ---------------
    for (int i = tid; i < 2 * L; i += THREADS) {
      const int slot = i >= L ? 1 : 0;
      const int n = i - slot * L;
      const int k = kk + slot;
      float2 a = make_float2(0.0f, 0.0f);
      if (n < W && k >= 0 && k < k1) {
        const int o = lead + (k - (k0 - 1)) * S;
        const float xm = (float)((n > 0) ? samp[o + n - 1] : (k >= 1 ? samp[o + P - 1] : 0));
        const float x0 = (float)samp[o + n];
        const float v = __fmul_rn(__fmaf_rn(-kPreLo, xm, __fmaf_rn(-kPreHi, xm, x0)), s_win[n]);
        a = cscale(s_ch[n], v);
      }
      X[i] = a;
    }
    __syncthreads();

    // ---- convolution with the chirp filter: FFT, multiply, FFT again --------
    float2* x = stockham<THREADS>(X, Y, 2, L, p, s_tw, tid);
    for (int i = tid; i < 2 * L; i += THREADS) {
      const int j = i >= L ? i - L : i;
      const float2 t = cmul(x[i], s_B[j]);
      x[i] = make_float2(t.x, -t.y);  // conj: the second forward transform then inverts
    }
    __syncthreads();
    x = stockham<THREADS>(x, x == X ? Y : X, 2, L, p, s_tw, tid);
    // |X[k]| of slot = |x[slot * L + k]|, k = 0 .. H

    // ---- power, energy -------------------------------------------------------
    float e0 = 0.0f, e1 = 0.0f, mx0 = 0.0f, mx1 = 0.0f;
    for (int i = tid; i < 2 * (H + 1); i += THREADS) {
      const int slot = i >= H + 1 ? 1 : 0;
      const int t = i - slot * (H + 1);
      const int k = kk + slot;
      const float2 z = x[slot * L + t];
      const float pw = __fadd_rn(__fmul_rn(z.x, z.x), __fmul_rn(z.y, z.y));
      if (t < H) {
        lmag[(slot == 0 ? ia : ib) * HP + t] = __log2f(pw);
        if (t >= 1) {
          if (slot == 0) { e0 += pw; mx0 = fmaxf(mx0, pw); }
          else { e1 += pw; mx1 = fmaxf(mx1, pw); }
        }
      }
      if (p.tap_spec && k >= k0 && k < k1) {
        float* tap = p.tap_spec + ((size_t)s * p.tap_stride + (k - kA)) * N;
        const float m = __fsqrt_rn(pw);
        tap[t] = m;
        if (t >= 1 && t < H) tap[N - t] = m;
      }
    }
    e0 = warp_sum(e0);
    e1 = warp_sum(e1);
    mx0 = warp_max(mx0);
    mx1 = warp_max(mx1);
    if (lane == 0) {
      red[warp * 4 + 0] = e0;
      red[warp * 4 + 1] = e1;
      red[warp * 4 + 2] = mx0;
      red[warp * 4 + 3] = mx1;
    }
    __syncthreads();
    float e_slot[2] = {0.0f, 0.0f}, pmax_slot[2] = {0.0f, 0.0f};
#pragma unroll
    for (int w = 0; w < NWARP; w++) {
      e_slot[0] += red[w * 4 + 0];
      e_slot[1] += red[w * 4 + 1];
      pmax_slot[0] = fmaxf(pmax_slot[0], red[w * 4 + 2]);
      pmax_slot[1] = fmaxf(pmax_slot[1], red[w * 4 + 3]);
    }
    __syncthreads();  // red is reused below

    // ---- spectral difference against the previous window (log2 domain) -------
    float linv_slot[2];
    linv_slot[0] = -__log2f(__fsqrt_rn(e_slot[0]) + 2.2204e-16f);
    linv_slot[1] = -__log2f(__fsqrt_rn(e_slot[1]) + 2.2204e-16f);
    const float thr0 = __log2f(pmax_slot[0]) - 13.287712379549449f;  // log2(1e4)
    const float thr1 = __log2f(pmax_slot[1]) - 13.287712379549449f;
    const float d20 = 2.0f * (linv_slot[0] - linv_prev);
    const float d21 = 2.0f * (linv_slot[1] - linv_slot[0]);
    float acc0 = 0.0f, acc1 = 0.0f;
    for (int i = tid; i < 2 * (H - 1); i += THREADS) {
      const int slot = i >= H - 1 ? 1 : 0;
      const int bin = 1 + i - slot * (H - 1);
      const float c = lmag[(slot == 0 ? ia : ib) * HP + bin];
      const float l = lmag[(slot == 0 ? ip : ia) * HP + bin];
      if (slot == 0) {
        if (c > thr0 && l > thr0) acc0 += fabsf((c - l) + d20);
      } else {
        if (c > thr1 && l > thr1) acc1 += fabsf((c - l) + d21);
      }
    }
    acc0 = warp_sum(acc0);
    acc1 = warp_sum(acc1);
    if (lane == 0) {
      red[warp * 4 + 0] = acc0;
      red[warp * 4 + 1] = acc1;
    }
    __syncthreads();
    if (tid < 2) {
      const int slot = tid;
      const int k = kk + slot;
      if (k >= k0 && k < k1) {
        float t = 0.0f;
        for (int w = 0; w < NWARP; w++) t += red[w * 4 + slot];
        p.feat[(size_t)s * p.feat_stride + (k - kA)] = make_float2(e_slot[slot], t * 0.34657359027997264f);  // ln2 / 2
      }
    }
    linv_prev = linv_slot[1];
    int t = ip; ip = ib; ib = ia; ia = t;  // slot B becomes "previous"
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// Host-side launch
// ---------------------------------------------------------------------------
size_t k1_smem_bytes_480(int warps) {
  return 240 * sizeof(float) + 240 * sizeof(float2) + (size_t)warps * sizeof(WarpSmem480);
}

// Does launch_k1 take the tensor-core kernel for this geometry?  (It occupies an SM's shared memory
// alone, so the analysis no longer shares SMs with the resynthesis kernel: the write pipeline in
// batch.cu cuts long writes into fewer prefixes then.)
bool k1_uses_dft16(const Geometry& g) {
  const char* e = getenv("SPEEDY_K1_TC");
  K1Params q;
  q.g = g;
  return (e ? atoi(e) : 1) != 0 && k1_dft16_supported(q);
}

cudaError_t launch_k1(const K1Params& p, cudaStream_t stream) {
  if (p.max_new_frames <= 0) return cudaSuccess;
  // 16 kHz mono: the tensor-core shape (k1_dft16.cu) for every launch, long or short, so that a
  // stream's windows always go through the same arithmetic however its input is chunked;
  // SPEEDY_K1_TC=0 selects the FFT kernel below instead
  {
    if (k1_uses_dft16(p.g)) return launch_k1_dft16(p, stream);
  }
  K1Params q = p;
  q.runs_per_stream = (p.max_new_frames + kRun - 1) / kRun;
  const long long items = (long long)q.runs_per_stream * p.n_streams;
  if (p.g.fft == 480 && p.g.step == S16 && p.g.partial == P16) {
    constexpr int WARPS = 4;
    const size_t smem = k1_smem_bytes_480(WARPS);
    static SmemOptIn opt_480;
    if (cudaError_t e = opt_480.ensure(k1_spectral_480<WARPS>, smem)) return e;
    const unsigned blocks = (unsigned)((items + WARPS - 1) / WARPS);
    k1_spectral_480<WARPS><<<blocks, WARPS * 32, smem, stream>>>(q);
  } else {
    constexpr int THREADS = 128;
    const int N = p.g.fft, W = p.g.window, M = W / 2;
    // radix plan for the M-point transforms: 8, 4, 5, 3, 2, then odd primes up to 16
    int plan[kMaxFactors], np = 0, rest = M;
    bool mixed = (W % 2 == 0) && M >= 4;
    if (mixed) {
      const int pref[] = {8, 4, 5, 3, 2, 7, 11, 13};
      for (int r : pref) {
        while (rest % r == 0 && np < kMaxFactors) {
          plan[np++] = r;
          rest /= r;
        }
      }
      mixed = rest == 1;
    }
    if (!mixed && p.bl_L > 0) {
      // chirp-z: the plan is for the L-point transforms
      const int L = p.bl_L;
      const int pref[] = {8, 4, 5, 3, 2, 7};
      np = 0;
      rest = L;
      for (int r : pref) {
        while (rest % r == 0 && np < kMaxFactors) {
          plan[np++] = r;
          rest /= r;
        }
      }
      if (rest != 1) return cudaErrorInvalidValue;
      q.n_factors = np;
      for (int i = 0, n = L; i < np; i++) {
        q.factors[i] = plan[i];
        n /= plan[i];
        q.plan_m[i] = n;
        q.plan_per[i] = L / plan[i];
      }
      const int HP = (N / 2 + 3) & ~3;
      const size_t smem = (size_t)(6 * L + W + (W & 1)) * sizeof(float2) + (size_t)(((W + 3) & ~3) + 3 * HP) * sizeof(float) +
                          (THREADS / 32) * 4 * sizeof(float) +
                          (size_t)((kRun + 1) * p.g.step + p.g.partial + 24) * sizeof(short);
      static SmemOptIn opt_b;
      if (cudaError_t e = opt_b.ensure(k1_spectral_bluestein<THREADS>, smem)) return e;
      k1_spectral_bluestein<THREADS><<<(unsigned)items, THREADS, smem, stream>>>(q);
    } else if (mixed) {
      q.n_factors = np;
      for (int i = 0, n = M; i < np; i++) {
        q.factors[i] = plan[i];
        n /= plan[i];
        q.plan_m[i] = n;
        q.plan_per[i] = M / plan[i];
      }
      const size_t smem = (size_t)(8 * M + 2 * M + W / 2) * sizeof(float2) + (size_t)(4 * W) * sizeof(float) +
                          (THREADS / 32) * 4 * sizeof(float) +
                          (size_t)((kRun + 1) * p.g.step + p.g.partial + 24) * sizeof(short);
      static SmemOptIn opt_m;
      if (cudaError_t e = opt_m.ensure(k1_spectral_mixed<THREADS>, smem)) return e;
      k1_spectral_mixed<THREADS><<<(unsigned)items, THREADS, smem, stream>>>(q);
    } else {
      return cudaErrorInvalidValue;  // (speedyBatchCreate plans a chirp-z length for every window the mixed-radix kernel cannot take)
    }
  }
  count_launch();
  return cudaGetLastError();
}

}  // namespace speedy
