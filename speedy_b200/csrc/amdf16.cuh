// 16 kHz pitch-period search (Sonic's findPitchPeriod, oracle/sonic_oracle.c:169-257) with
// the geometry as compile-time constants: step 160, lags 40 .. 246, AMDF decimation by four,
// coarse lags 10 .. 61 on 123 decimated values, refinement over +-16 lags at the full rate.
// Shared by the pipelined kernel's chain warp (k4_splice.cu) and the one-warp kernel
// (k4_sonic.cu); `Ctx` supplies the warp's shared-memory arrays and lane assignment:
//   int lane, cGi, cSub, cG (coarse: group, lane in group, lanes of the group), fg, fGi0 (fine);
//   int prevPeriod, prevMinDiff;
//   int* win() / ds() (32-bit mono window, decimated copy), float* rcp() (1 / lag table),
//   unsigned* part() (kPartWords per-lag partial sums, 7 rows of kPartStride; zero at start).
#pragma once

#include <cuda_runtime.h>

namespace speedy {
namespace amdf16 {

constexpr unsigned kFull = 0xffffffffu;
// Row stride of the per-lag partial sums (four coarse rows, three fine rows of up to 64 lags).
// 76 words = 19 sixteen-byte slots: the 16-byte stores of a quarter-warp (lanes enumerate
// lag group x lane-in-group) then fall into distinct bank groups (4 - 5 wavefronts per store,
// 12 with a stride of 64).
constexpr int kPartStride = 76;
constexpr int kPartWords = 7 * kPartStride;

#ifdef K4_TIMING
// developer build: cycles of the search's phases (0 decimate, 1 coarse blocks, 2 coarse pick,
// 3 fine blocks, 4 fine pick), accumulated by lane 0 of the stream that has k.timing set
static __device__ unsigned long long g_amdf_cycles[8];
#define AT_BEGIN() long long _at0 = clock64()
#define AT_MARK(slot) do { const long long _t1 = clock64(); if (k.lane == 0 && k.timing) atomicAdd(&g_amdf_cycles[slot], (unsigned long long)(_t1 - _at0)); _at0 = _t1; } while (0)
#else
#define AT_BEGIN() do {} while (0)
#define AT_MARK(slot) do {} while (0)
#endif

// The 16 |a - b| terms of one aligned block of four samples for the four lags pg .. pg+3
// (pg a multiple of four), every sample valid for every lag.
__device__ __forceinline__ void sad16(const int4& av, const int4& b0, const int4& b1, unsigned (&d)[4]) {
  d[0] = __sad(av.x, b0.x, d[0]); d[1] = __sad(av.x, b0.y, d[1]);
  d[2] = __sad(av.x, b0.z, d[2]); d[3] = __sad(av.x, b0.w, d[3]);
  d[0] = __sad(av.y, b0.y, d[0]); d[1] = __sad(av.y, b0.z, d[1]);
  d[2] = __sad(av.y, b0.w, d[2]); d[3] = __sad(av.y, b1.x, d[3]);
  d[0] = __sad(av.z, b0.z, d[0]); d[1] = __sad(av.z, b0.w, d[1]);
  d[2] = __sad(av.z, b1.x, d[2]); d[3] = __sad(av.z, b1.y, d[3]);
  d[0] = __sad(av.w, b0.w, d[0]); d[1] = __sad(av.w, b1.x, d[1]);
  d[2] = __sad(av.w, b1.y, d[2]); d[3] = __sad(av.w, b1.z, d[3]);
}

#define LD4(p) (*reinterpret_cast<const int4*>(p))

// Fully valid blocks j, j + step, ... < jend of one lag group (3 LDS.128 + 16 VABSDIFF each).
// Two register sets in turn: the loads of the next block are in flight while this one's
// differences issue (a lone warp has nobody to hide the shared-memory latency behind).
__device__ __forceinline__ void blocks_run(const int* base, int pg, int j, int step, int jend, unsigned (&d)[4]) {
  if (j >= jend) return;
  const int* pa = base + 4 * j;
  const int stride = 4 * step;
  int4 a0 = LD4(pa), b00 = LD4(pa + pg), b01 = LD4(pa + pg + 4);
  int4 a1, b10, b11;
#pragma unroll 1
  for (;;) {
    j += step;
    if (j >= jend) {
      sad16(a0, b00, b01, d);
      break;
    }
    pa += stride;
    a1 = LD4(pa); b10 = LD4(pa + pg); b11 = LD4(pa + pg + 4);
    sad16(a0, b00, b01, d);
    j += step;
    if (j >= jend) {
      sad16(a1, b10, b11, d);
      break;
    }
    pa += stride;
    a0 = LD4(pa); b00 = LD4(pa + pg); b01 = LD4(pa + pg + 4);
    sad16(a1, b10, b11, d);
  }
}

// The ragged start of a lag group's range: the block that holds the first sample when the
// range does not start on a block boundary (hd = 1 .. 3 samples in): sample m counts for
// every lag iff m >= hd.
__device__ __forceinline__ void block_head(const int* base, int pg, int hd, unsigned (&d)[4]) {
  const int4 a = LD4(base), b0 = LD4(base + pg), b1 = LD4(base + pg + 4);
  d[0] = __sad(a.w, b0.w, d[0]); d[1] = __sad(a.w, b1.x, d[1]);
  d[2] = __sad(a.w, b1.y, d[2]); d[3] = __sad(a.w, b1.z, d[3]);
  if (hd <= 2) {
    d[0] = __sad(a.z, b0.z, d[0]); d[1] = __sad(a.z, b0.w, d[1]);
    d[2] = __sad(a.z, b1.x, d[2]); d[3] = __sad(a.z, b1.y, d[3]);
  }
  if (hd <= 1) {
    d[0] = __sad(a.y, b0.y, d[0]); d[1] = __sad(a.y, b0.z, d[1]);
    d[2] = __sad(a.y, b0.w, d[2]); d[3] = __sad(a.y, b1.x, d[3]);
  }
}

// The ragged end: `pa` is the first block that is not fully valid for all four lags.  Its
// sample m counts for lag pg + l iff m < c + l, c = hd (pg is a multiple of four): the c
// samples every lag still has, then l more for lag pg + l.  Six terms are unconditional
// (m < l), the other twelve hang off three predicates.
__device__ __forceinline__ void block_tail(const int* pa, int pg, int c, unsigned (&d)[4]) {
  const int4 a0 = LD4(pa), a1 = LD4(pa + 4);
  const int4 b0 = LD4(pa + pg), b1 = LD4(pa + pg + 4), b2 = LD4(pa + pg + 8);
  d[1] = __sad(a0.x, b0.y, d[1]);
  d[2] = __sad(a0.x, b0.z, d[2]); d[2] = __sad(a0.y, b0.w, d[2]);
  d[3] = __sad(a0.x, b0.w, d[3]); d[3] = __sad(a0.y, b1.x, d[3]); d[3] = __sad(a0.z, b1.y, d[3]);
  if (c > 0) {
    d[0] = __sad(a0.x, b0.x, d[0]); d[1] = __sad(a0.y, b0.z, d[1]);
    d[2] = __sad(a0.z, b1.x, d[2]); d[3] = __sad(a0.w, b1.z, d[3]);
  }
  if (c > 1) {
    d[0] = __sad(a0.y, b0.y, d[0]); d[1] = __sad(a0.z, b0.w, d[1]);
    d[2] = __sad(a0.w, b1.y, d[2]); d[3] = __sad(a1.x, b1.w, d[3]);
  }
  if (c > 2) {
    d[0] = __sad(a0.z, b0.z, d[0]); d[1] = __sad(a0.w, b1.x, d[1]);
    d[2] = __sad(a1.x, b1.z, d[2]); d[3] = __sad(a1.y, b2.x, d[3]);
  }
}

// The same for the one-lag-per-lane form: lane i holds the sums of lags base + i (sa) and
// base + 32 + i (sb); candidates are ballots.
static __device__ __noinline__ void resolve_exact2(unsigned sa, unsigned sb, unsigned bal_a, unsigned bal_b, int base,
                                            int want_min, unsigned* rd, int* rp) {
  unsigned bd = 0u;
  int bp = want_min ? 0 : 255;
#pragma unroll 1
  for (int half = 0; half < 2; half++) {
    unsigned bal = half ? bal_b : bal_a;
    while (bal) {
      const int src = __ffs(bal) - 1;
      bal &= bal - 1;
      const unsigned cd = __shfl_sync(kFull, half ? sb : sa, src);
      const int cp = base + 32 * half + src;
      const unsigned long long lhs = (unsigned long long)cd * (unsigned)bp;
      const unsigned long long rhs = (unsigned long long)bd * (unsigned)cp;
      if (want_min ? (bp == 0 || lhs < rhs) : (lhs > rhs)) {
        bd = cd;
        bp = cp;
      }
    }
  }
  *rd = bd;
  *rp = bp;
}

// floor(a / b) for a < 2^27, 0 < b < 2^11 (quotient < 2^16) from the float reciprocal of b:
// the estimate is off by at most one either way, one fix-up per side.
__device__ __forceinline__ int udiv_small(unsigned a, int b, float rcp_b) {
  const int q = (int)(__uint2float_rn(a) * rcp_b);
  const int rem = (int)a - q * b;
  return q + (rem >= b ? 1 : 0) - (rem < 0 ? 1 : 0);
}


// AMDF over lags lo..hi on a[i] = arr[off + i] (oracle/sonic_oracle.c:185-209): lane `sub`
// of the `G` adjacent lanes of a lag group (lags pg .. pg+3, pg a multiple of four) takes
// every G-th fully valid block; lane 0 of the group also the ragged start, lane 1 (or 0 when
// alone) the ragged end.  The partial sums go through shared memory, one row per lane of a
// group, and come back one lag per lane (lag base + lane, and base + 32 + lane): no shuffle
// tree.  Rows a group has no lane for stay zero (coarse pass: the lane assignment is
// static); lags outside [lo, hi] are masked.
//
// Arg-min / arg-max of diff / lag.  The C scan compares by cross-multiplication with strict
// inequalities, so ties go to the smaller lag.  Float keys diff * (1 / lag) (relative error
// < 2e-7): the lags within 2e-6 of the warp-wide extremum are a superset of the true
// extremum; nearly always that is one lag, otherwise the short list is resolved exactly.
// FINE: also the per-sample difference at the best lag and what the previous-period rule
// needs of the worst one (the coarse pass needs neither).
template <class Ctx, int ROWS, bool FINE>
__device__ __forceinline__ int search(const Ctx& k, const int* arr, int off, int lo, int hi, int pg, bool live, int sub,
                                      int G, int* minDiff, int* maxDiff) {
  constexpr bool WANT_DIFFS = FINE;
  unsigned d[4] = {0u, 0u, 0u, 0u};
  AT_BEGIN();
  if (live) {
    const int B0 = off & ~3, hd = off & 3;
    const int* base = arr + B0;
    const int jf1 = (hd + pg) >> 2;  // blocks jf0 .. jf1-1 are fully valid
    blocks_run(base, pg, (hd ? 1 : 0) + sub, G, jf1, d);
    if (sub == 0 && hd) block_head(base, pg, hd, d);
    if (sub == (G > 1 ? 1 : 0)) block_tail(base + 4 * jf1, pg, hd, d);
  }
  {
    // 16 kHz: the partial sums go through shared memory, one row per lane of a group, and come
    // back one lag per lane (lag base + lane, and base + 32 + lane): no shuffle tree, no
    // four-lags-per-leader selects.  Rows a group has no lane for stay zero (coarse pass: the
    // lane assignment is static); lags outside [lo, hi] are masked below.
    unsigned* part = k.part() + (WANT_DIFFS ? 4 * kPartStride : 0);
    const int base = lo & ~3;
    if (live) *reinterpret_cast<uint4*>(part + sub * kPartStride + (pg - base)) = make_uint4(d[0], d[1], d[2], d[3]);
    __syncwarp();
    AT_MARK(FINE ? 3 : 1);
    unsigned sa = 0u, sb = 0u;
#pragma unroll
    for (int r = 0; r < ROWS; r++) {
      sa += part[r * kPartStride + k.lane];
      sb += part[r * kPartStride + 32 + k.lane];
    }
    __syncwarp();  // (the next search of this kind rewrites the rows)
    const int la = base + k.lane, lb = la + 32;
    const bool va = la >= lo && la <= hi, vb = lb <= hi;
    const float ka = __uint2float_rn(sa) * k.rcp()[va ? la : 0], kb = __uint2float_rn(sb) * k.rcp()[vb ? lb : 0];
    const float big = 3.0e38f;
    const float emin = __uint_as_float(__reduce_min_sync(kFull, __float_as_uint(fminf(va ? ka : big, vb ? kb : big))));
    const float tmin = emin * 1.000002f;
    const unsigned bal_a = __ballot_sync(kFull, va && ka <= tmin), bal_b = __ballot_sync(kFull, vb && kb <= tmin);
    unsigned wal_a = 0u, wal_b = 0u;
    float emax = 1.0f;
    if (WANT_DIFFS) {
      emax = __uint_as_float(__reduce_max_sync(kFull, __float_as_uint(fmaxf(va ? ka : 0.f, vb ? kb : 0.f))));
      const float tmax = emax * 0.999998f;
      wal_a = __ballot_sync(kFull, va && ka >= tmax);
      wal_b = __ballot_sync(kFull, vb && kb >= tmax);
    }
    unsigned best_diff, worst_diff = 0u;
    int best, worst = 255;
    const bool one_min = __popc(bal_a) + __popc(bal_b) == 1;
    const bool one_max = !WANT_DIFFS || (__popc(wal_a) + __popc(wal_b) == 1 && emax > 0.f);
    if (one_min && one_max) {
      const int src = __ffs(bal_a | bal_b) - 1;
      best = base + src + (bal_a ? 0 : 32);
      best_diff = __shfl_sync(kFull, bal_a ? sa : sb, src);
      if (WANT_DIFFS) {
        const int srw = __ffs(wal_a | wal_b) - 1;
        worst = base + srw + (wal_a ? 0 : 32);
        worst_diff = __shfl_sync(kFull, wal_a ? sa : sb, srw);
      }
    } else {
      resolve_exact2(sa, sb, bal_a, bal_b, base, 1, &best_diff, &best);
      if (WANT_DIFFS) resolve_exact2(sa, sb, wal_a, wal_b, base, 0, &worst_diff, &worst);
    }
    if (WANT_DIFFS) {
      const int md = udiv_small(best_diff, best, k.rcp()[best]);
      *minDiff = md;
      *maxDiff = worst_diff >= (unsigned)(3 * md + 1) * (unsigned)worst ? 3 * md + 1 : 0;
    }
#ifdef K4_TIMING
    _at0 += (best & 0);  // (the pick's result is needed before the clock is read)
#endif
    AT_MARK(FINE ? 4 : 2);
    return best;
  }
}

// Upstream downSampleInput: four frames per value, C integer division as
// (v + (v < 0 ? 3 : 0)) >> 2.  Lane l loads the aligned vectors l, l + 32, l + 64, l + 96 of
// the span (consecutive 16-byte slots across the warp: no bank conflicts) and splits each at
// the span's offset inside a vector; a value is the upper part of its vector plus the lower
// part of the next one, which the neighbouring lane holds.
//
// Hook: work of the caller that does not depend on this search (the previous splice's
// overlap-add and copy-through in k4_chain16.cu).  load() issues its shared-memory loads
// before this function's own, finish() does its arithmetic and global stores under their
// latency.  Neither may write shared memory.
struct NoHook {
  __device__ __forceinline__ void load() {}
  __device__ __forceinline__ void finish() {}
};

template <class Ctx, class Hook>
__device__ __forceinline__ void decimate(const Ctx& k, int off, Hook& hook) {
  AT_BEGIN();
  {
    const int r = off & 3;
    const int4* p = reinterpret_cast<const int4*>(k.win() + (off & ~3)) + k.lane;
    hook.load();
    const int4 x0 = p[0], x1 = p[32], x2 = p[64], x3 = p[96];
    hook.finish();
    int lo[4], hi[4];
    const int4 xs[4] = {x0, x1, x2, x3};
    // (r is the same for the whole warp: one branch instead of a chain of selects per vector)
    if (r == 0) {
#pragma unroll
      for (int q = 0; q < 4; q++) { lo[q] = 0; hi[q] = (xs[q].x + xs[q].y) + (xs[q].z + xs[q].w); }
    } else if (r == 1) {
#pragma unroll
      for (int q = 0; q < 4; q++) { lo[q] = xs[q].x; hi[q] = xs[q].y + xs[q].z + xs[q].w; }
    } else if (r == 2) {
#pragma unroll
      for (int q = 0; q < 4; q++) { lo[q] = xs[q].x + xs[q].y; hi[q] = xs[q].z + xs[q].w; }
    } else {
#pragma unroll
      for (int q = 0; q < 4; q++) { lo[q] = xs[q].x + xs[q].y + xs[q].z; hi[q] = xs[q].w; }
    }
    const int nb = (k.lane + 1) & 31;
    int t[4];
#pragma unroll
    for (int q = 0; q < 4; q++) t[q] = __shfl_sync(kFull, lo[q], nb);
    const bool last = k.lane == 31;  // its neighbour is lane 0 of the next round of vectors
    __syncwarp();  // every lane is done reading the previous decimated copy
#pragma unroll
    for (int q = 0; q < 4; q++) {
      // (value 127 would need vector 128: it only ever meets masked samples)
      const int s = hi[q] + (last ? (q < 3 ? t[q < 3 ? q + 1 : 3] : 0) : t[q]);
      k.ds()[32 * q + k.lane] = (s + ((s >> 31) & 3)) >> 2;
    }
  }
  __syncwarp();
  AT_MARK(0);
}

// findPitchPeriod at window offset `off`: coarse pass on the decimated copy (a static lane
// assignment with more lanes for the longer lags: group q has q + 1 blocks), refinement at
// the full rate (three lanes per lag group), previous-period rule.
template <class Ctx, class Hook>
__device__ __forceinline__ int find_pitch_period(Ctx& k, int off, Hook& hook) {
  int minDiff = 0, maxDiff = 0;
  decimate(k, off, hook);
  int period = 4 * search<Ctx, 4, false>(k, k.ds(), 0, 10, 61, 4 * (2 + k.cGi), k.cGi >= 0, k.cSub, k.cG, nullptr, nullptr);
  int lo = period - 16, hi = period + 16;
  if (lo < 40) lo = 40;
  if (hi > 246) hi = 246;
  const int g0 = lo >> 2;
  period = search<Ctx, 3, true>(k, k.win(), off, lo, hi, 4 * (g0 + k.fGi0), k.fGi0 < (hi >> 2) - g0 + 1, k.fg, 3, &minDiff,
                                &maxDiff);
  // previousPeriodBetter(preferNew = 1), oracle/sonic_oracle.c:213-223
  const bool keep_prev =
      minDiff != 0 && k.prevPeriod != 0 && !(maxDiff > minDiff * 3) && !(minDiff * 2 <= k.prevMinDiff * 3);
  const int result = keep_prev ? k.prevPeriod : period;
  k.prevMinDiff = minDiff;
  k.prevPeriod = period;
  return result;
}

template <class Ctx>
__device__ __forceinline__ int find_pitch_period(Ctx& k, int off) {
  NoHook none;
  return find_pitch_period(k, off, none);
}

}  // namespace amdf16
}  // namespace speedy
