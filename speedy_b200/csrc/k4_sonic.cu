// K4 — Sonic time-scale modification: AMDF pitch-period search and
// pitch-synchronous overlap-add, one CTA of NW warps per stream (NW = 1, 2 or 4).
//
// Replaces what the reference does through upstream Sonic
// (soniclib.c:354, 369-370, 398, 547, 551 -> sonicIntSetSpeed,
// sonicIntWriteShortToStream, sonicIntFlushStream; algorithm restated in
// oracle/sonic_oracle.c and SURVEY.md Appendix A): processStreamInput,
// changeSpeed, findPitchPeriod (down-sampled coarse search + full-rate
// refinement), prevPeriodBetter, skipPitchPeriod / insertPitchPeriod, overlapAdd,
// copy-through of unmodified input, and flush.
//
// Everything here is integer arithmetic on int16 samples except the handful of
// float expressions that size a splice (period / (speed - 1) ...), which are
// written with explicit IEEE _rn intrinsics (and the file is built with
// --fmad=false) so they round exactly as the C code does.  Integer sums are
// associative, so splitting the AMDF sums across lanes cannot change a result:
// given identical per-frame speeds the output is bit-exact.
//
// Warps per stream.  The splice cursor is strictly sequential (the next position
// depends on the period just found), so a stream is a chain of ~80 pitch iterations
// per second of audio and parallelism comes from the streams.  With thousands of
// streams per GPU one warp per stream is the efficient shape: no block barrier, the
// uniform bookkeeping executed once.  With few streams (the 1024-stream benchmark
// leaves 7 per SM) the chain's latency is what is measured; 2 or 4 warps can share
// the parallel loops of a stream (window refill, decimation, AMDF blocks, overlap-
// add): all warps run the same uniform control flow on replicated scalar state and
// meet at three barriers per pitch iteration.  Measured at 1024 streams that does not
// pay (K4 18.5 ms with one warp, 23.1 / 33.2 ms with two / four: the replicated scalar
// code and the barriers cost more than the split loops save), so one warp per stream
// is the default everywhere and the wider shapes stay selectable per batch
// (threads_per_stream; the single-stream drop-in API uses four).
//
// Variants.  CH = 1 is a mono specialisation (the multi-channel paths drop out of the
// instruction stream: the chain is sensitive to instruction-cache misses).  Short
// launches (10 ms streaming writes, flush) take HOSTMAP = true: the lane assignment of
// the two searches is computed by the launcher, because the kernel's own set-up would
// be half of such a launch's instructions.  The shared window is sized by the launcher
// so that one wave of streams stays resident (k4_buf_frames).
//
// Sonic's input FIFO is never materialised: the stream keeps two absolute
// cursors (head = first unconsumed frame, fed = one past the last frame handed
// to Sonic) and the warp slides a shared-memory window over the caller's buffer.
// The output cursor is the per-stream pending count in the output buffer.
//
// AMDF layout.  The window is kept in shared memory as 32-bit mono samples, so
// that one LDS.128 fetches four operands ready for VABSDIFF (|a-b|+c in one
// instruction).  Lags are processed in groups of four consecutive lags starting
// at a multiple of four: for an aligned block of four samples a[blk..blk+3] the
// operands of lags 4k..4k+3 are the seven values b[blk+4k .. blk+4k+6], i.e. two
// more aligned LDS.128: 3 loads + 16 VABSDIFF per 16 differences.  The first and
// last block of a lag group are the same code under a per-element mask.  A group's
// blocks are spread over a few adjacent lanes, which keep their partial sums in
// registers and combine them with shuffles; no shared-memory accumulators.
#include <stdlib.h>

#include "amdf16.cuh"
#include "kernels.cuh"

namespace speedy {

#ifdef K4_TIMING
__device__ unsigned long long g_k4_cycles[16];
#endif

namespace {


#ifdef K4_TIMING
#define T_BEGIN() const long long _t0 = clock64()
#define T_END(slot) do { if (vl == 0 && timing) atomicAdd(&g_k4_cycles[slot], (unsigned long long)(clock64() - _t0)); } while (0)
#else
#define T_BEGIN() do {} while (0)
#define T_END(slot) do {} while (0)
#endif

constexpr int kPad = 32;  // over-read slack behind the window and the decimated copy
constexpr unsigned kFull = 0xffffffffu;
constexpr int kMaxGroups = 64;  // lag groups of four per search (>= 2 * skip + 2)
// per-lag totals: [2][4 * kMaxGroups], or the 16 kHz search's padded rows (amdf16.cuh)
constexpr int kSumsWords = ((2 * 4 * kMaxGroups > amdf16::kPartWords ? 2 * 4 * kMaxGroups : amdf16::kPartWords) + 3) & ~3;

template <int NW, int CH, bool K16 = false>
struct Sonic {
  static constexpr int VL = 32 * NW;  // lanes cooperating on one stream
  // geometry
  int Crt, S, minP, maxP, maxReq, skip;  // Crt: channel count when not a template constant
  long long cap;
  // shared memory (this warp's slice)
  int* w32;                 // mono window, 32-bit [bufN + kPad]
  short* buf;               // interleaved raw window (C > 1 only) [bufN * C]
  int* ds32;                // decimated mono [maxReq / skip + kPad]
  int bufN;
  // window state (warp-uniform)
  long long bufStart;
  int bufLen;
  // source
  Source src;
  long long zero_from;  // frames >= this read as silence (flush padding)
  // stream state (warp-uniform)
  long long head, fed, outTotal;
  int prevPeriod, prevMinDiff, remCopy, outCount, status;
  short* out;
  int lane, warp, vl;
  unsigned* sums;  // [2][4 * kMaxGroups] per-lag totals handed from the group leaders to every warp
  int parity;
  // fine-pass lane mapping: fG sub-lanes per lag group
  int fG, fPerRound, fGi0, fg;
  // coarse-pass lane mapping: group cGi (-1 = idle), sub-lane cSub of cG, largest cG
  int cGi, cSub, cG, cMaxG;
  unsigned dec_magic;
  const unsigned* magic_tab;  // ceil(2^(32+s) / n) for n = 2 .. maxP, or null (computed on demand)
  float* rcp16;               // K16: 1 / lag table of the shared 16 kHz pitch search (amdf16.cuh)
  __device__ __forceinline__ int* win() const { return w32; }
  __device__ __forceinline__ int* ds() const { return ds32; }
  __device__ __forceinline__ float* rcp() const { return rcp16; }
  __device__ __forceinline__ unsigned* part() const { return sums; }
  bool fold_all;  // few streams per SM (latency bound): every lane runs the fold rather than branching around it
  bool timing;

  // channel count: a compile-time constant in the mono specialisation (CH = 1)
  __device__ __forceinline__ int nch() const { return CH ? CH : Crt; }

  // Make [start, start + count) resident in the shared window (count <= bufN - 8).
  __device__ __forceinline__ void sync() {
    if (NW == 1) __syncwarp(); else __syncthreads();
  }

  __device__ __forceinline__ void ensure(long long start, int count) {
    if (start >= bufStart && start + count <= bufStart + bufLen) return;
    T_BEGIN();
    sync();  // every lane is done with the old window
    bufStart = start & ~7LL;  // keeps the 16-byte loads of the refill aligned
    bufLen = bufN;
    if (NW == 1 && CH == 1 && fold_all) stage_mono<VL, int, 16>(src, bufStart, bufN, zero_from, w32, nullptr, vl);
    else stage_mono<VL, int>(src, bufStart, bufN, zero_from, w32, nch() > 1 ? buf : nullptr, vl);
    sync();
    T_END(0);
  }

  __device__ __forceinline__ void advance_out(int n) {
    outTotal += n;
    if (outCount + n > cap) {
      status |= 1;  // SPEEDY_STATUS_OUTPUT_OVERFLOW
      outCount = (int)cap;
    } else {
      outCount += n;
    }
  }

  // Append n frames starting at absolute frame `from` to the output.
  __device__ __forceinline__ void emit_copy(long long from, int n, int out_offset_frames) {
    const int o0 = (int)(from - bufStart);
    const int total = n * nch();
    const long long base = (long long)(outCount + out_offset_frames) * nch();
    const long long room = cap * nch() - base;
    short* o = out + base;
    T_BEGIN();
    if (nch() == 1 && NW == 1 && fold_all) {
      // one warp, few streams per SM (latency bound): up to 128 frames as four independent slots (all loads first, then the stores), the
      // rest -- rare -- in a loop; a lone warp otherwise waits out a shared-memory round trip per 32 frames
      const int* src = w32 + o0 + vl;
      int v[4];
#pragma unroll
      for (int q = 0; q < 4; q++) v[q] = vl + 32 * q < total ? src[32 * q] : 0;
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int i = vl + 32 * q;
        if (i < total && i < room) o[i] = (short)v[q];
      }
      for (int i = 128 + vl; i < total; i += 32) {
        if (i < room) o[i] = (short)w32[o0 + i];
      }
    } else if (nch() == 1) {
      for (int i = vl; i < total; i += VL) {
        if (i < room) o[i] = (short)w32[o0 + i];
      }
    } else if (nch() == 2) {
      // stereo: a frame is one aligned 32-bit word on both sides (room is a whole number of frames)
      const int* p32 = reinterpret_cast<const int*>(buf) + o0;
      int* o32 = reinterpret_cast<int*>(o);
      for (int t = vl; t < n; t += VL) {
        if (2 * t < room) o32[t] = p32[t];
      }
    } else {
      const short* p = buf + (size_t)o0 * nch();
      for (int i = vl; i < total; i += VL) {
        if (i < room) o[i] = p[i];
      }
    }
    T_END(7);
  }

  // ceil(2^(32+shift) / n) for 2^shift < n <= 2^(shift+1): floor from the correctly rounded
  // double quotient (the true one is at least 1/n away from the integers it does not hit)
  static __device__ __forceinline__ unsigned division_magic(int n, int shift) {
    const double qd = __ddiv_rn((double)(1ULL << (32 + shift)), (double)n);
    return (unsigned)(unsigned long long)qd + ((n & (n - 1)) ? 1u : 0u);
  }

  // out[t] = (down[t]*(n-t) + up[t]*t) / n per channel, C integer arithmetic.
  // down/up are absolute frames inside the window.
  __device__ __forceinline__ void overlap_add(int n, long long down, long long up, int out_offset_frames) {
    const int d0 = (int)(down - bufStart), u0 = (int)(up - bufStart);
    const int total = n * nch();
    const long long base = (long long)(outCount + out_offset_frames) * nch();
    const long long room = cap * nch() - base;
    short* o = out + base;
    T_BEGIN();
    // trunc(|num| / n) == umulhi(|num|, magic) >> shift for |num| < 2^25, 1 < n < 2^11,
    // 2^shift < n <= 2^(shift+1), magic = ceil(2^(32+shift) / n)
    unsigned magic = 0u;
    int shift = 0;
    if (n > 1) {
      shift = 31 - __clz(n - 1);
      magic = (magic_tab && n <= maxP) ? magic_tab[n] : division_magic(n, shift);
    }
    if (nch() == 1) {
      if (n == 1) {
        if (vl == 0 && 0 < room) o[0] = (short)w32[d0];
      } else if (NW == 1 && fold_all) {
        // (as in emit_copy: four independent slots, loads first; the many-stream shapes keep the loop:
        // 8192 x 30 s measured 39.2 ms with the slots against 38.9 with it)
        const int* dp = w32 + d0 + vl;
        const int* up = w32 + u0 + vl;
        int a[4], b[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const bool in = vl + 32 * q < total;
          a[q] = in ? dp[32 * q] : 0;
          b[q] = in ? up[32 * q] : 0;
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const int t = vl + 32 * q;
          const int num = a[q] * (n - t) + b[q] * t;
          const int qq = (int)(__umulhi((unsigned)abs(num), magic) >> shift);
          if (t < total && t < room) o[t] = (short)(num < 0 ? -qq : qq);
        }
        for (int t = 128 + vl; t < total; t += 32) {
          const int num = w32[d0 + t] * (n - t) + w32[u0 + t] * t;
          const int qq = (int)(__umulhi((unsigned)abs(num), magic) >> shift);
          if (t < room) o[t] = (short)(num < 0 ? -qq : qq);
        }
      } else {
        for (int t = vl; t < total; t += VL) {
          const int num = w32[d0 + t] * (n - t) + w32[u0 + t] * t;
          const int q = (int)(__umulhi((unsigned)abs(num), magic) >> shift);
          if (t < room) o[t] = (short)(num < 0 ? -q : q);
        }
      }
    } else if (nch() == 2) {
      // stereo: one frame (two samples) per lane and pass, 32-bit loads and stores
      const int* dp32 = reinterpret_cast<const int*>(buf) + d0;
      const int* up32 = reinterpret_cast<const int*>(buf) + u0;
      int* o32 = reinterpret_cast<int*>(o);
      for (int t = vl; t < n; t += VL) {
        const int dw = dp32[t], uw = up32[t];
        const int num_l = (int)(short)(dw & 0xffff) * (n - t) + (int)(short)(uw & 0xffff) * t;
        const int num_r = (dw >> 16) * (n - t) + (uw >> 16) * t;
        const int ql = n == 1 ? abs(num_l) : (int)(__umulhi((unsigned)abs(num_l), magic) >> shift);
        const int qr = n == 1 ? abs(num_r) : (int)(__umulhi((unsigned)abs(num_r), magic) >> shift);
        const int l = num_l < 0 ? -ql : ql, r = num_r < 0 ? -qr : qr;
        if (2 * t < room) o32[t] = (l & 0xffff) | (r << 16);
      }
    } else {
      const short* dp = buf + (size_t)d0 * nch();
      const short* up_ = buf + (size_t)u0 * nch();
      for (int i = vl; i < total; i += VL) {
        const int t = nch() == 2 ? (i >> 1) : i / nch();  // (stereo: no division per sample)
        const int num = (int)dp[i] * (n - t) + (int)up_[i] * t;
        const int q = n == 1 ? abs(num) : (int)(__umulhi((unsigned)abs(num), magic) >> shift);
        if (i < room) o[i] = (short)(num < 0 ? -q : q);
      }
    }
    T_END(6);
  }

  // Upstream downSampleInput: sum `skip` frames x C channels, C integer division
  // (truncating).  |sum| < 2^21 and the divisor is small, so the quotient is exact
  // as (|sum| * ceil(2^32 / divisor)) >> 32.
  __device__ __forceinline__ void decimate(int off) {
    T_BEGIN();
    sync();  // every lane is done reading the previous decimated copy
    const int count = maxReq / skip;
    const int per = nch() * skip;
    if (nch() == 1 && (skip & 3) == 0) {
      const int r = off & 3;
      const int nmid = (skip >> 2) - 1;
      const int* base = w32 + (off & ~3);
#pragma unroll 2
      for (int i = vl; i < count; i += VL) {
        const int4* p = reinterpret_cast<const int4*>(base + i * skip);
        const int4 x = p[0];
        const int4 z = p[nmid + 1];
        int v = x.w + (r == 0 ? x.x : z.x) + (r <= 1 ? x.y : z.y) + (r <= 2 ? x.z : z.z);
#pragma unroll 1
        for (int m = 1; m <= nmid; m++) {
          const int4 t = p[m];
          v += (t.x + t.y) + (t.z + t.w);
        }
        const int qa = (int)__umulhi((unsigned)abs(v), dec_magic);
        ds32[i] = v < 0 ? -qa : qa;
      }
    } else if (nch() == 2 && (skip & 3) == 0) {
      // stereo: one 32-bit word per frame (left | right << 16), the same aligned walk;
      // a dp2a with unit weights adds both halves of a word to the sum in one instruction
      const int r = off & 3;
      const int nmid = (skip >> 2) - 1;
      const int* base = reinterpret_cast<const int*>(buf) + (off & ~3);
#pragma unroll 2
      for (int i = vl; i < count; i += VL) {
        const int4* p = reinterpret_cast<const int4*>(base + i * skip);
        const int4 x = p[0];
        const int4 z = p[nmid + 1];
        int v = __dp2a_lo(x.w, 0x0101, 0);
        v = __dp2a_lo(r == 0 ? x.x : z.x, 0x0101, v);
        v = __dp2a_lo(r <= 1 ? x.y : z.y, 0x0101, v);
        v = __dp2a_lo(r <= 2 ? x.z : z.z, 0x0101, v);
#pragma unroll 1
        for (int m = 1; m <= nmid; m++) {
          const int4 t = p[m];
          v = __dp2a_lo(t.x, 0x0101, v);
          v = __dp2a_lo(t.y, 0x0101, v);
          v = __dp2a_lo(t.z, 0x0101, v);
          v = __dp2a_lo(t.w, 0x0101, v);
        }
        const int qa = (int)__umulhi((unsigned)abs(v), dec_magic);
        ds32[i] = v < 0 ? -qa : qa;
      }
    } else {
#pragma unroll 1
      for (int i = vl; i < count; i += VL) {
        int v = 0;
        if (nch() == 1) {
          const int* q = w32 + off + i * skip;
#pragma unroll 1
          for (int j = 0; j < skip; j++) v += q[j];
        } else {
          const short* q = buf + ((size_t)off + (size_t)i * skip) * nch();
          int j = 0;
#pragma unroll 1
          for (; j + 4 <= per; j += 4) v += (q[j] + q[j + 1]) + (q[j + 2] + q[j + 3]);
#pragma unroll 1
          for (; j < per; j++) v += q[j];
        }
        const int qa = (int)__umulhi((unsigned)abs(v), dec_magic);
        ds32[i] = v < 0 ? -qa : qa;
      }
    }
    sync();
    T_END(1);
  }

  // |a - b| sums of one aligned block of four samples for the four lags
  // pg .. pg+3 (pg a multiple of four), every sample valid for every lag.
  __device__ __forceinline__ void block_full(const int* arr, int blk, int pg, unsigned (&d)[4]) {
    const int4 av = *reinterpret_cast<const int4*>(arr + blk);
    const int4 b0 = *reinterpret_cast<const int4*>(arr + blk + pg);
    const int4 b1 = *reinterpret_cast<const int4*>(arr + blk + pg + 4);
    d[0] = __sad(av.x, b0.x, d[0]); d[0] = __sad(av.y, b0.y, d[0]);
    d[0] = __sad(av.z, b0.z, d[0]); d[0] = __sad(av.w, b0.w, d[0]);
    d[1] = __sad(av.x, b0.y, d[1]); d[1] = __sad(av.y, b0.z, d[1]);
    d[1] = __sad(av.z, b0.w, d[1]); d[1] = __sad(av.w, b1.x, d[1]);
    d[2] = __sad(av.x, b0.z, d[2]); d[2] = __sad(av.y, b0.w, d[2]);
    d[2] = __sad(av.z, b1.x, d[2]); d[2] = __sad(av.w, b1.y, d[2]);
    d[3] = __sad(av.x, b0.w, d[3]); d[3] = __sad(av.y, b1.x, d[3]);
    d[3] = __sad(av.z, b1.y, d[3]); d[3] = __sad(av.w, b1.z, d[3]);
  }

  // The same for a block at the edge of the range: sample x = blk + m contributes to
  // lag pg + l iff off <= x < off + pg + l.
  __device__ __forceinline__ void block_edge(const int* arr, int blk, int pg, int off, unsigned (&d)[4]) {
    const int4 av = *reinterpret_cast<const int4*>(arr + blk);
    const int4 b0 = *reinterpret_cast<const int4*>(arr + blk + pg);
    const int4 b1 = *reinterpret_cast<const int4*>(arr + blk + pg + 4);
    const int a[4] = {av.x, av.y, av.z, av.w};
    const int b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    const int first = off - blk;      // m >= first
    const int last = off + pg - blk;  // m < last + l
#pragma unroll
    for (int l = 0; l < 4; l++) {
#pragma unroll
      for (int m = 0; m < 4; m++) {
        if (m >= first && m < last + l) d[l] = __sad(a[m], b[m + l], d[l]);
      }
    }
  }

  // Exact arg-min and arg-max of diff/period over the warp's candidates (one
  // (diff, period) pair per lane for each, period 0 = none).  The C scan compares by
  // cross-multiplication with strict inequalities, so ties go to the smaller lag.
  // A float quotient picks the lanes within 2e-6 of the extremum (a superset of the
  // true extremum: its relative error is below 4e-7); almost always that is one
  // lane, otherwise the short list is resolved exactly.
  __device__ __forceinline__ void pick2(unsigned bd, int bp, unsigned wd, int wp, unsigned* min_diff,
                                        int* min_period, unsigned* max_diff, int* max_period) {
    const float fb = bp ? __fdividef((float)bd, (float)bp) : 3.0e38f;
    const float fw = wp ? __fdividef((float)wd, (float)wp) : 0.0f;
    const float emin = __uint_as_float(__reduce_min_sync(kFull, __float_as_uint(fb)));
    const float emax = __uint_as_float(__reduce_max_sync(kFull, __float_as_uint(fw)));
    unsigned ballot_b = __ballot_sync(kFull, bp != 0 && fb <= emin * 1.000002f);
    unsigned ballot_w = __ballot_sync(kFull, wp != 0 && fw >= emax * 0.999998f);
    int sb = __ffs(ballot_b) - 1, sw = __ffs(ballot_w) - 1;
    unsigned rbd = __shfl_sync(kFull, bd, sb);
    int rbp = __shfl_sync(kFull, bp, sb);
    unsigned rwd = __shfl_sync(kFull, wd, sw);
    int rwp = __shfl_sync(kFull, wp, sw);
    ballot_b &= ballot_b - 1;
    ballot_w &= ballot_w - 1;
    while (ballot_b) {  // rare: several candidates within rounding of each other
      sb = __ffs(ballot_b) - 1;
      ballot_b &= ballot_b - 1;
      const unsigned cd = __shfl_sync(kFull, bd, sb);
      const int cp = __shfl_sync(kFull, bp, sb);
      const unsigned long long l = (unsigned long long)cd * (unsigned)rbp;
      const unsigned long long r = (unsigned long long)rbd * (unsigned)cp;
      if (l < r || (l == r && cp < rbp)) { rbd = cd; rbp = cp; }
    }
    while (ballot_w) {
      sw = __ffs(ballot_w) - 1;
      ballot_w &= ballot_w - 1;
      const unsigned cd = __shfl_sync(kFull, wd, sw);
      const int cp = __shfl_sync(kFull, wp, sw);
      const unsigned long long l = (unsigned long long)cd * (unsigned)rwp;
      const unsigned long long r = (unsigned long long)rwd * (unsigned)cp;
      if (l > r || (l == r && cp < rwp)) { rwd = cd; rwp = cp; }
    }
    *min_diff = rbd; *min_period = rbp; *max_diff = rwd; *max_period = rwp;
  }

  // floor(a / b) for a < 2^27, 0 < b < 2^11 (quotient < 2^16): float estimate + fix-up.
  // The estimate carries three roundings of 2^-24 each, under 0.02 at that size of
  // quotient, so its truncation is off by at most one either way.
  static __device__ __forceinline__ int udiv_small(unsigned a, int b) {
    const int q = (int)((float)a * __frcp_rn((float)b));
    const int rem = (int)a - q * b;
    return q + (rem >= b ? 1 : 0) - (rem < 0 ? 1 : 0);
  }

  // Fold the four lag sums a lane holds (lags pg .. pg+3) into its running best /
  // worst candidates; lags ascend, so strict comparisons reproduce the C scan.
  __device__ __forceinline__ void fold(const unsigned (&d)[4], int pg, int lo, int hi, unsigned& bd, int& bp,
                                       unsigned& wd, int& wp) {
    // branch-free: the splice chain waits for this, and only a few lanes hold sums
#pragma unroll
    for (int l = 0; l < 4; l++) {
      const int p = pg + l;
      const bool in = p >= lo && p <= hi;
      const bool first = bp == 0;
      const bool lt = (unsigned long long)d[l] * (unsigned)bp < (unsigned long long)bd * (unsigned)p;
      const bool gt = (unsigned long long)d[l] * (unsigned)wp > (unsigned long long)wd * (unsigned)p;
      const bool tb = in && (first || lt), tw = in && (first || gt);
      bd = tb ? d[l] : bd;
      bp = tb ? p : bp;
      wd = tw ? d[l] : wd;
      wp = tw ? p : wp;
    }
  }

  // Sum the partial lag sums of the `G` adjacent lanes of a group (all in one warp)
  // into its first lane.
  __device__ __forceinline__ void combine(unsigned (&d)[4], int sub, int G, int maxG) {
    for (int delta = 1; delta < maxG; delta <<= 1) {
#pragma unroll
      for (int l = 0; l < 4; l++) {
        const unsigned t = __shfl_down_sync(kFull, d[l], delta);
        if (sub + delta < G) d[l] += t;
      }
    }
  }

  // AMDF over lags lo..hi on a[i] = arr[off + i].  Returns the best lag; *minDiff /
  // *maxDiff are the per-sample differences at the best and worst lag.
  __device__ __forceinline__ int search(const int* arr, int off, int lo, int hi, bool coarse, int* minDiff,
                                        int* maxDiff) {
    const int g0 = lo >> 2;  // first lag group (lags 4*g0 .. 4*g0+3)
    const int ngroups = (hi >> 2) - g0 + 1;
    unsigned* tot = sums + parity * (4 * kMaxGroups);
    parity ^= 1;
    unsigned bd = 0, wd = 0;
    int bp = 0, wp = 0;
    T_BEGIN();
    if (coarse) {
      // off == 0 and the lag range is fixed: a static lane assignment with more
      // lanes for the longer lags; group q has q fully valid blocks (j < q) and one
      // partially valid block (j == q: sample m counts for lag 4q + l while m < l)
      unsigned d[4] = {0u, 0u, 0u, 0u}, e[4] = {0u, 0u, 0u, 0u};
      const int q = g0 + cGi;
      const int pg = 4 * q;
      if (cGi >= 0) {
        int j = cSub;
        for (; j + cG < q; j += 2 * cG) {
          block_full(arr, 4 * j, pg, d);
          block_full(arr, 4 * (j + cG), pg, e);
        }
        if (j < q) block_full(arr, 4 * j, pg, d);
        if (cSub == cG - 1) block_edge(arr, pg, pg, 0, e);
#pragma unroll
        for (int l = 0; l < 4; l++) d[l] += e[l];
      }
      combine(d, cSub, cG, cMaxG);
      if (NW == 1) {
        // every lane runs the (branch-free) fold; only a group's first lane has a range
        const bool leader = cGi >= 0 && cSub == 0;
        if (fold_all) fold(d, pg, leader ? lo : 1, leader ? hi : 0, bd, bp, wd, wp);
        else if (leader) fold(d, pg, lo, hi, bd, bp, wd, wp);
      } else if (cGi >= 0 && cSub == 0) {
        *reinterpret_cast<uint4*>(tot + 4 * cGi) = make_uint4(d[0], d[1], d[2], d[3]);
      }
    } else {
      // blocks of group gi: j = 0 .. nblk-1 at B0 + 4j; fully valid for jf0 <= j < jf1,
      // the head block (j = 0 when off is unaligned) and one or two tail blocks are
      // handled by the group's sub-lanes under a per-element mask
      const int B0 = off & ~3;
      const int jf0 = (off & 3) ? 1 : 0;
      for (int gbase = 0; gbase < ngroups; gbase += fPerRound) {
        const int gi = fGi0 + gbase;
        const bool live = gi < ngroups;  // also false for idle lanes
        const int pg = 4 * (g0 + gi);
        unsigned d[4] = {0u, 0u, 0u, 0u}, e[4] = {0u, 0u, 0u, 0u};
        if (live) {
          const int jf1 = (off - B0 + pg) >> 2;
          const int nblk = ((((off + pg + 2) & ~3) + 4) - B0) >> 2;
          int j = jf0 + fg;
          for (; j + fG < jf1; j += 2 * fG) {
            block_full(arr, B0 + 4 * j, pg, d);
            block_full(arr, B0 + 4 * (j + fG), pg, e);
          }
          if (j < jf1) block_full(arr, B0 + 4 * j, pg, d);
          for (int which = fg; which < 3; which += fG) {
            const int je = which == 0 ? 0 : jf1 + which - 1;
            if (which == 0 ? jf0 == 1 : je < nblk) block_edge(arr, B0 + 4 * je, pg, off, e);
          }
#pragma unroll
          for (int l = 0; l < 4; l++) d[l] += e[l];
        }
        combine(d, fg, fG, fG);
        if (NW == 1) {
          // (one warp: rounds ascend, so the running candidates see ascending lags)
          const bool leader = live && fg == 0;
          if (fold_all) fold(d, pg, leader ? lo : 1, leader ? hi : 0, bd, bp, wd, wp);
          else if (leader) fold(d, pg, lo, hi, bd, bp, wd, wp);
        } else if (live && fg == 0) {
          *reinterpret_cast<uint4*>(tot + 4 * gi) = make_uint4(d[0], d[1], d[2], d[3]);
        }
      }
    }
    T_END(coarse ? 2 : 4);
#ifdef K4_TIMING
    const long long _t1 = clock64();
#endif
    if (NW > 1) {
      // every warp folds all groups (ascending lags per lane) and picks: the result
      // is identical in all warps, no broadcast needed.  The totals are double-
      // buffered, so the next search may start writing while a slower warp reads.
      sync();
      for (int gi = lane; gi < ngroups; gi += 32) {
        const uint4 t4 = *reinterpret_cast<const uint4*>(tot + 4 * gi);
        const unsigned d[4] = {t4.x, t4.y, t4.z, t4.w};
        fold(d, 4 * (g0 + gi), lo, hi, bd, bp, wd, wp);
      }
    }
    unsigned best_diff, worst_diff;
    int best, worst;
    pick2(bd, bp, wd, wp, &best_diff, &best, &worst_diff, &worst);
    // the C scan starts from (maxDiff = 0, worstPeriod = 255) and only replaces
    // it with a strictly larger ratio
    if (worst_diff == 0u) worst = 255;
    *minDiff = udiv_small(best_diff, best);
    *maxDiff = udiv_small(worst_diff, worst);
#ifdef K4_TIMING
    if (vl == 0 && timing) atomicAdd(&g_k4_cycles[coarse ? 3 : 5], (unsigned long long)(clock64() - _t1 + (*minDiff & 0)));
#endif
    return best;
  }

  __device__ __forceinline__ int find_pitch_period(long long pos) {
    const int off = (int)(pos - bufStart);
    if (K16) return amdf16::find_pitch_period(*this, off);  // 16 kHz mono, one warp: compile-time geometry
    int minDiff = 0, maxDiff = 0, period = 0;
    const int* arr = w32;
    int aoff = off;
    int lo = minP, hi = maxP, stages = 1;
    if (!(nch() == 1 && skip == 1)) {
      decimate(off);
      arr = ds32;
      aoff = 0;
      lo = minP / skip;
      hi = maxP / skip;
      stages = skip != 1 ? 2 : 1;
    }
    for (int stage = 0; stage < stages; stage++) {
      period = search(arr, aoff, lo, hi, stage == 0 && stages == 2, &minDiff, &maxDiff);
      if (stage == 0 && stages == 2) {
        // refine around the coarse estimate at the full rate (mono window)
        period *= skip;
        lo = period - (skip << 2);
        hi = period + (skip << 2);
        if (lo < minP) lo = minP;
        if (hi > maxP) hi = maxP;
        arr = w32;
        aoff = off;
      }
    }
    // prevPeriodBetter(preferNew = 1)
    const bool keep_prev = minDiff != 0 && prevPeriod != 0 && !(maxDiff > minDiff * 3) && !(minDiff * 2 <= prevMinDiff * 3);
    const int result = keep_prev ? prevPeriod : period;
    prevMinDiff = minDiff;
    prevPeriod = period;
    return result;
  }

  // processStreamInput with the speed that is current now.
  __device__ __forceinline__ void process(float speed) {
    const long long numInput = fed - head;
    // upstream: speed > 1.00001 || speed < 0.99999 with speed promoted to double; the same
    // test on the float itself: 0x3F800054 is the smallest float above 1.00001, 0x3F7FFF58
    // the largest below 0.99999
    if (speed >= __uint_as_float(0x3F800054u) || speed <= __uint_as_float(0x3F7FFF58u)) {
      if (numInput < maxReq) return;
      long long position = 0;
      do {
        int newSamples;
        const long long pos = head + position;
        if (remCopy > 0) {
          newSamples = remCopy < maxReq ? remCopy : maxReq;
          ensure(pos, newSamples);
          emit_copy(pos, newSamples, 0);
          advance_out(newSamples);
          remCopy -= newSamples;
          position += newSamples;
        } else {
          ensure(pos, maxReq);
#ifdef K4_TIMING
          if (vl == 0 && timing) atomicAdd(&g_k4_cycles[9], 1ULL);
#endif
          const int period = find_pitch_period(pos);
          if (speed > 1.0f) {
            if (speed >= 2.0f) {
              newSamples = (int)(long long)__fdiv_rn((float)period, __fsub_rn(speed, 1.0f));
            } else {
              newSamples = period;
              remCopy = (int)__fdiv_rn(__fmul_rn((float)period, __fsub_rn(2.0f, speed)),
                                       __fsub_rn(speed, 1.0f));
            }
            overlap_add(newSamples, pos, pos + period, 0);
            advance_out(newSamples);
            position += period + newSamples;
          } else {
            if (speed < 0.5f) {
              newSamples = (int)(long long)__fdiv_rn(__fmul_rn((float)period, speed),
                                                     __fsub_rn(1.0f, speed));
            } else {
              newSamples = period;
              remCopy = (int)__fdiv_rn(
                  __fmul_rn((float)period, __fsub_rn(__fmul_rn(2.0f, speed), 1.0f)),
                  __fsub_rn(1.0f, speed));
            }
            // the period itself, then the cross-fade back into it
            emit_copy(pos, period, 0);
            overlap_add(newSamples, pos + period, pos, period);
            advance_out(period + newSamples);
            position += newSamples;
          }
        }
        if (newSamples == 0) return;  // nothing produced: the input is not consumed
      } while (position + maxReq <= numInput);
      head += position;
    } else {
      // speed == 1: copy the whole FIFO through
      long long left = numInput;
      while (left > 0) {
        int n = left < bufN - 8 ? (int)left : bufN - 8;
        ensure(head, n);
        emit_copy(head, n, 0);
        advance_out(n);
        head += n;
        left -= n;
      }
    }
  }
};

__host__ __device__ inline size_t k4_stream_smem(const Geometry& g, int buf_frames) {
  size_t b = (size_t)(buf_frames + kPad) * sizeof(int);
  b += (size_t)((g.max_required / g.skip + kPad + 3) & ~3) * sizeof(int);
  b += (size_t)kSumsWords * sizeof(unsigned);
  b += (size_t)((g.max_period + 4) & ~3) * sizeof(unsigned);  // overlap-add division constants
  b += (size_t)((g.max_period + 8) & ~3) * sizeof(float);     // reciprocals of the lags (16 kHz search)
  if (g.channels > 1) b += (size_t)buf_frames * g.channels * sizeof(short) + 16;  // + one vector of over-read
  return (b + 15) & ~(size_t)15;
}

}  // namespace

// MINB: resident CTAs per SM the register allocation is held to (1 = unconstrained).
// Few streams: registers are free, latency is what counts.  Many streams: 16 resident
// warps per SM hide the chain's latency, worth a tighter allocation.
template <int NW, int MINB, int CH, bool HOSTMAP, bool K16 = false>
__global__ void __launch_bounds__(NW * 32, MINB) k4_sonic(K4Params p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int s = blockIdx.x;
  if (s >= p.n_streams) return;
  if (p.flush && p.flush_mask && p.flush_mask[s] == 0) return;  // a flush of some streams only
  const Geometry& g = p.g;

  Sonic<NW, CH, K16> k;
  k.lane = threadIdx.x & 31;
  k.warp = threadIdx.x >> 5;
  k.vl = threadIdx.x;
  k.parity = 0;
#ifdef K4_TIMING
  k.timing = (s == 0) && !p.flush;
  const long long t_kernel = clock64();
#else
  k.timing = false;
#endif
  k.Crt = g.channels;
  k.S = g.step;
  k.minP = g.min_period;
  k.maxP = g.max_period;
  k.maxReq = g.max_required;
  k.skip = g.skip;
  k.cap = p.out_capacity;
  k.bufN = p.buf_frames;
  // carve-up (every piece a multiple of 16 bytes)
  k.w32 = reinterpret_cast<int*>(smem_raw);
  k.ds32 = k.w32 + k.bufN + kPad;
  k.sums = reinterpret_cast<unsigned*>(k.ds32 + ((k.maxReq / k.skip + kPad + 3) & ~3));
  unsigned* magic_tab = k.sums + kSumsWords;  // [(maxP + 4) & ~3]
  k.rcp16 = reinterpret_cast<float*>(magic_tab + ((k.maxP + 4) & ~3));
  k.buf = reinterpret_cast<short*>(k.rcp16 + ((k.maxP + 8) & ~3));
  if (K16) {
    for (int n = k.vl; n < ((k.maxP + 8) & ~3); n += Sonic<NW, CH>::VL) k.rcp16[n] = n ? __frcp_rn((float)n) : 0.f;
    for (int n = k.vl; n < kSumsWords; n += Sonic<NW, CH>::VL) k.sums[n] = 0u;  // rows no lane writes stay zero
  }
  // long launches: the overlap-add's division constants once, off the splice chain (a
  // double division per pitch iteration otherwise); short ones compute the few they need
  k.magic_tab = HOSTMAP ? nullptr : magic_tab;
  k.fold_all = MINB == 1;
  if (!HOSTMAP) {
    for (int n = 2 + k.vl; n <= k.maxP; n += Sonic<NW, CH>::VL) {
      magic_tab[n] = Sonic<NW, CH>::division_magic(n, 31 - __clz(n - 1));
    }
  }
  k.bufStart = 0;
  k.bufLen = 0;
  k.dec_magic = (unsigned)((0x100000000ULL + (unsigned)(g.channels * k.skip) - 1) / (unsigned)(g.channels * k.skip));
  if (HOSTMAP) {
    // short launches (10 ms streaming writes, flush): the lane mappings come from the
    // launcher (k4_lane_map), the kernel's own set-up would be half of its instructions
    const unsigned m = p.lane_map[threadIdx.x];
    k.cGi = (int)(m & 0xffu) - 1;
    k.cSub = (int)((m >> 8) & 0xffu);
    k.cG = (int)((m >> 16) & 0xffu);
    k.cMaxG = p.c_max_g;
    k.fG = p.f_g;
    k.fPerRound = p.f_per_round;
    __builtin_assume(k.cG >= 1 && k.cG <= 32 && k.cSub >= 0 && k.cSub < 32 && k.cMaxG >= 1 && k.cMaxG <= 32);
    __builtin_assume(k.fG >= 1 && k.fG <= 32 && k.fPerRound >= 1);
    const int gpw = 32 / k.fG;  // groups per warp
    const int slot = k.lane / k.fG;
    k.fg = k.lane - slot * k.fG;
    k.fGi0 = slot < gpw ? k.warp * gpw + slot : (1 << 30);  // idle lanes never match
  } else {
    // lane mappings (lags rounded out to groups of four)
    const int c_lo = k.minP / k.skip, c_hi = k.maxP / k.skip;
    const int qlo = c_lo >> 2, qhi = c_hi >> 2;
    // coarse: group q has q + 1 blocks.  Groups are dealt to the warps largest first
    // (least-loaded warp takes the next one); inside a warp every group gets
    // ceil((q + 1) / T) adjacent lanes with the smallest T that fits 32 lanes.
    int load[NW];
    int owner[kMaxGroups];
#pragma unroll
    for (int w = 0; w < NW; w++) load[w] = 0;
    for (int q = qhi; q >= qlo; q--) {
      int best_w = 0;
#pragma unroll
      for (int w = 1; w < NW; w++) {
        if (load[w] < load[best_w]) best_w = w;
      }
      owner[q - qlo] = best_w;
      load[best_w] += q + 1;
    }
    int T = 1;
    for (;; T++) {
      int sum = 0;
      for (int q = qlo; q <= qhi; q++) {
        if (owner[q - qlo] == k.warp) sum += (q + T) / T;
      }
      if (sum <= 32) break;
    }
    k.cGi = -1;
    k.cSub = 0;
    k.cG = 1;
    k.cMaxG = 1;
    int first = 0;
    for (int q = qlo; q <= qhi; q++) {
      if (owner[q - qlo] != k.warp) continue;
      const int n = (q + T) / T;
      if (k.lane >= first && k.lane < first + n) {
        k.cGi = q - qlo;
        k.cSub = k.lane - first;
        k.cG = n;
      }
      if (n > k.cMaxG) k.cMaxG = n;
      first += n;
    }
    // fine (and single-stage) pass: fG adjacent lanes per group, groups never
    // straddle a warp: the largest fG with NW * (32 / fG) >= the group count
    const int fine_groups = k.skip != 1 ? 2 * k.skip + 2 : qhi - qlo + 1;
    k.fG = 32;
    while (k.fG > 1 && NW * (32 / k.fG) < fine_groups) k.fG--;
    const int gpw = 32 / k.fG;  // groups per warp
    k.fPerRound = NW * gpw;
    const int slot = k.lane / k.fG;
    k.fg = k.lane - slot * k.fG;
    k.fGi0 = slot < gpw ? k.warp * gpw + slot : (1 << 30);  // idle lanes never match
  }
  for (int i = k.vl; i < kPad; i += Sonic<NW, CH>::VL) {  // the over-read pads
    k.w32[k.bufN + i] = 0;
    k.ds32[k.maxReq / k.skip + i] = 0;
  }
  k.sync();

  const Range rg = write_range(p.st.total, p.counts, p.frames, p.done, s);
  const long long t_old = rg.t_old;
  const long long t_new = p.flush ? t_old : rg.t_new;
  const long long t_done = p.flush ? t_old : rg.t_done;
  k.src.channels = CH ? CH : g.channels;
  k.src.hist = p.hist + (size_t)s * p.hist_stride;
  k.src.in = p.in ? p.in + (size_t)s * p.in_stride_frames * g.channels : nullptr;
  k.src.hist_base = p.st.hist_base[s];
  k.src.t_old = t_old;
  k.src.t_new = t_new;
  k.zero_from = 1LL << 56;  // "never": still safe to multiply by the channel count

  k.head = p.st.sonic_head[s];
  k.fed = p.st.sonic_fed[s];
  k.outTotal = p.st.out_total[s];
  k.outCount = p.st.out_count[s];
  k.prevPeriod = p.st.prev_period[s];
  k.prevMinDiff = p.st.prev_min_diff[s];
  k.remCopy = p.st.remaining_copy[s];
  k.status = 0;
  k.out = p.out + (size_t)s * p.out_capacity * g.channels;
  float speed = p.st.sonic_speed[s];
  const bool nonlinear = p.st.nonlinear[s] != 0.0f;

  // Feed events, in the reference's order, through ONE process() call site:
  //   write, nonlinear (soniclib.c:354, 369-371): one 10 ms buffer per new speed
  //   write, linear    (soniclib.c:397-399): the whole write at the global speed
  //   flush, nonlinear (soniclib.c:538-550): the complete delayed buffers at the
  //                    last speed (the partial buffer being filled is dropped)
  //   flush, both      upstream sonicFlushStream: expected length, 2*maxRequired
  //                    frames of silence, process, trim
  long long ev = 0, ev_end = 0;
  const float* sp = p.speeds ? p.speeds + (size_t)s * p.speeds_stride : nullptr;
  int rA = 0;
  if (!p.flush) {
    if (nonlinear) {
      rA = tensions_ready(g, frames_analyzed(g, t_old));  // the speeds rows count from here
      ev = tensions_ready(g, frames_analyzed(g, t_done));
      ev_end = tensions_ready(g, frames_analyzed(g, t_new));
    } else {
      ev = 0;
      ev_end = t_new > t_done ? 1 : 0;
    }
  } else if (nonlinear) {
    ev = k.fed / k.S;
    ev_end = t_old / k.S;
    if (ev_end < ev) ev_end = ev;
  }
  const long long n_events = (ev_end - ev) + (p.flush ? 1 : 0);
  const bool per_frame_speed = nonlinear && !p.flush;
  long long expected = 0;
  float speed_batch = 0.0f;  // 32 speeds at a time, one per lane
  for (long long i = 0; i < n_events; i++, ev++) {
    const bool final_flush = p.flush && i == n_events - 1;
    if (final_flush) {
      const long long remaining = k.fed - k.head;
      expected = k.outTotal +
                 (int)__fadd_rn(__fdiv_rn(__fdiv_rn((float)(int)remaining, speed), 1.0f), 0.5f);
      k.zero_from = k.fed;
      k.bufLen = 0;  // the window may hold real samples past the padding point
      k.fed += 2 * k.maxReq;
    } else if (nonlinear) {
      if (per_frame_speed) {
        const int j = (int)(ev - rA);
        if ((j & 31) == 0 || i == 0) {
          const long long idx = (long long)(j & ~31) + k.lane;
          speed_batch = idx < ev_end - rA ? sp[idx] : 0.0f;  // rows past this launch are not ready
        }
        speed = __shfl_sync(kFull, speed_batch, j & 31);
      }
      k.fed = (ev + 1) * k.S;
    } else {
      k.fed = t_new;
    }
    k.process(speed);
    if (final_flush) {
      if (k.outTotal > expected) {
        long long excess = k.outTotal - expected;
        k.outTotal = expected;
        k.outCount = k.outCount > excess ? (int)(k.outCount - excess) : 0;
      }
      // the padding is not input: the stream carries on from the real end of the data
      // (upstream sonicFlushStream leaves numInputSamples = 0 and stays usable)
      k.fed -= 2 * k.maxReq;
      k.head = k.fed;
      k.remCopy = 0;
      k.status |= 2;  // SPEEDY_STATUS_FLUSHED
    }
  }

#ifdef K4_TIMING
  if (k.vl == 0 && k.timing) {
    atomicAdd(&g_k4_cycles[8], (unsigned long long)(clock64() - t_kernel));
    atomicAdd(&g_k4_cycles[10], (unsigned long long)n_events);
  }
#endif
  if (k.vl == 0) {
    p.st.sonic_head[s] = k.head;
    p.st.sonic_fed[s] = k.fed;
    p.st.out_total[s] = k.outTotal;
    p.st.out_count[s] = k.outCount;
    p.st.prev_period[s] = k.prevPeriod;
    p.st.prev_min_diff[s] = k.prevMinDiff;
    p.st.remaining_copy[s] = k.remCopy;
    p.st.sonic_speed[s] = speed;
    if (k.status) atomicOr(&p.st.status[s], k.status);
  }
}

// Lane assignment of the two AMDF searches for NW warps per stream (lags rounded out to
// groups of four).
//   coarse: group q has q + 1 blocks.  Groups are dealt to the warps largest first
//   (least-loaded warp takes the next one); inside a warp every group gets
//   ceil((q + 1) / T) adjacent lanes with the smallest T that fits 32 lanes.
//   fine (and single-stage) pass: fG adjacent lanes per group, groups never straddle a
//   warp: the largest fG with NW * (32 / fG) >= the group count.
void k4_lane_map(K4Params& p, int NW) {
  const Geometry& g = p.g;
  const int c_lo = g.min_period / g.skip, c_hi = g.max_period / g.skip;
  const int qlo = c_lo >> 2, qhi = c_hi >> 2;
  int load[4] = {0, 0, 0, 0};
  int owner[kMaxGroups];
  for (int q = qhi; q >= qlo; q--) {
    int best_w = 0;
    for (int w = 1; w < NW; w++) {
      if (load[w] < load[best_w]) best_w = w;
    }
    owner[q - qlo] = best_w;
    load[best_w] += q + 1;
  }
  p.c_max_g = 1;
  for (int i = 0; i < 128; i++) p.lane_map[i] = 1u << 16;  // idle: no group, cG = 1
  for (int w = 0; w < NW; w++) {
    int T = 1;
    for (;; T++) {
      int sum = 0;
      for (int q = qlo; q <= qhi; q++) {
        if (owner[q - qlo] == w) sum += (q + T) / T;
      }
      if (sum <= 32) break;
    }
    int first = 0;
    for (int q = qlo; q <= qhi; q++) {
      if (owner[q - qlo] != w) continue;
      const int n = (q + T) / T;
      for (int l = first; l < first + n; l++) {
        p.lane_map[w * 32 + l] = (unsigned)(q - qlo + 1) | ((unsigned)(l - first) << 8) | ((unsigned)n << 16);
      }
      first += n;
    }
  }
  // cMaxG is per warp in the kernel's combine(); the largest over the warps is a valid bound
  for (int i = 0; i < 32 * NW; i++) {
    const int n = (int)((p.lane_map[i] >> 16) & 0xffu);
    if (n > p.c_max_g) p.c_max_g = n;
  }
  const int fine_groups = g.skip != 1 ? 2 * g.skip + 2 : qhi - qlo + 1;
  int fG = 32;
  while (fG > 1 && NW * (32 / fG) < fine_groups) fG--;
  p.f_g = fG;
  p.f_per_round = NW * (32 / fG);
}

static int k4_buf_frames(const Geometry& g, int n_streams) {
  // window: several search spans; smaller when many streams share an SM
  int n = (n_streams >= 148 * 12 ? 4 : 8) * g.max_required;
  const int floor_n = n_streams >= 148 * 12 ? 2048 : 4096;
  if (n < floor_n) n = floor_n;
  // ... but never so large that the streams of one wave stop fitting on their SM: the
  // splice chain is latency bound, a second wave of CTAs would double the run time
  // (48 kHz stereo: 8 search spans are 95 KB per stream)
  int ctas = (n_streams + 147) / 148;
  if (ctas > 20) ctas = 20;
  const long long budget = (227LL * 1024) / ctas - 1024;  // per CTA, 1 KB reserved by the system
  const long long per_frame = sizeof(int) + (g.channels > 1 ? g.channels * sizeof(short) : 0);
  const long long fit = ((long long)budget - (long long)k4_stream_smem(g, 0)) / per_frame;
  if (n > fit) n = (int)fit;
  const int least = g.max_required + g.max_required / 2;  // one search span plus room to slide
  if (n < least) n = least;
  if (const char* e = getenv("SPEEDY_K4_BUF")) n = atoi(e) > 2 * g.max_required ? atoi(e) : n;
  return n & ~63;
}

template <int NW, int MINB, int CH, bool HOSTMAP, bool K16 = false>
static cudaError_t launch_k4_c(K4Params& p, cudaStream_t stream) {
  const size_t smem = k4_stream_smem(p.g, p.buf_frames);
  static SmemOptIn opt;  // (one per instantiation)
  if (cudaError_t e = opt.ensure(k4_sonic<NW, MINB, CH, HOSTMAP, K16>, smem)) return e;
  k4_sonic<NW, MINB, CH, HOSTMAP, K16><<<p.n_streams, NW * 32, smem, stream>>>(p);
  count_launch();
  return cudaGetLastError();
}

// 16 kHz mono: the pitch search with compile-time geometry (amdf16.cuh)
static bool k4_is_16k_mono(const Geometry& g) {
  return g.channels == 1 && g.rate == 16000 && g.step == 160 && g.min_period == 40 && g.max_period == 246 && g.skip == 4;
}

template <int NW, int MINB>
static cudaError_t launch_k4_t(K4Params& p, cudaStream_t stream) {
  // short launches take the variant with host-computed lane mappings
  const bool short_launch = p.flush || p.frames - p.done <= p.g.rate;
  if (short_launch) {
    k4_lane_map(p, NW);
    if (NW == 1 && k4_is_16k_mono(p.g)) return launch_k4_c<NW, MINB, 1, true, NW == 1>(p, stream);
    return p.g.channels == 1 ? launch_k4_c<NW, MINB, 1, true>(p, stream) : launch_k4_c<NW, MINB, 0, true>(p, stream);
  }
  if (NW == 1 && k4_is_16k_mono(p.g)) return launch_k4_c<NW, MINB, 1, false, NW == 1>(p, stream);
  return p.g.channels == 1 ? launch_k4_c<NW, MINB, 1, false>(p, stream) : launch_k4_c<NW, MINB, 0, false>(p, stream);
}

cudaError_t launch_k4(const K4Params& p0, cudaStream_t stream) {
  // The pipelined shape (k4_splice.cu: chain / filler / output warps, TMA bulk loads) is
  // selected with SPEEDY_K4_PIPELINE=1 (optionally SPEEDY_K4_SPLICE_MIN=<frames>: launches
  // shorter than that stay here).  Measured (profiles/README.md): with the same 16 kHz pitch
  // search (amdf16.cuh) in both, the three-warp CTA is held to 80 registers by the per-
  // scheduler register file at 7 streams per SM and its helper warps compete with the
  // analysis kernels for issue slots, so the one-warp shape is the faster step; flushes,
  // short launches and multi-channel streams always run here.
  {
    const int pipeline = getenv("SPEEDY_K4_PIPELINE") ? atoi(getenv("SPEEDY_K4_PIPELINE")) : 0;
    const long long min_frames = getenv("SPEEDY_K4_SPLICE_MIN") ? atoll(getenv("SPEEDY_K4_SPLICE_MIN")) : -1;
    const bool short_launch = p0.flush || p0.frames - p0.done <= (min_frames >= 0 ? min_frames : (long long)p0.g.rate);
    if (pipeline && !short_launch && p0.threads_per_stream <= 32 && k4_splice_supported(p0)) {
      return launch_k4_splice(p0, stream);
    }
  }
  // The chain shape (k4_chain16.cu: fixed overlapping windows prefetched by TMA bulk copies,
  // output deferred into the next search) is selected with SPEEDY_K4_CHAIN=1 for 16 kHz mono
  // writes of at least a second of audio.  Measured at 1024 x 60 s (profiles/README.md): 11.9 ms
  // against 12.4 ms for this kernel when each runs alone, but 19.0 against 18.1 ms for the whole
  // step, where the analysis kernels share the SMs: its extra (predicated) instructions cost more
  // there than the latency it hides, so it is not the default.
  {
    const int chain = getenv("SPEEDY_K4_CHAIN") ? atoi(getenv("SPEEDY_K4_CHAIN")) : 0;
    const bool short_launch = p0.flush || p0.frames - p0.done <= (long long)p0.g.rate;
    const int max_streams = getenv("SPEEDY_K4_CHAIN_MAX") ? atoi(getenv("SPEEDY_K4_CHAIN_MAX")) : 148 * 8;
    if (chain && !short_launch && p0.threads_per_stream <= 32 && p0.n_streams <= max_streams && k4_chain16_supported(p0)) {
      return launch_k4_chain16(p0, stream);
    }
  }
  K4Params p = p0;
  p.buf_frames = k4_buf_frames(p.g, p.n_streams);
  // threads per stream: with few streams per SM the serial splice chain is latency
  // bound and extra warps shorten it; with many streams one warp each is the most
  // work-efficient shape
  int t = p.threads_per_stream;
  if (const char* e = getenv("SPEEDY_K4_THREADS")) t = atoi(e);
  if (t == 0) t = 32;  // measured: extra warps do not shorten the chain enough to pay for their barriers
  if (t <= 32) {
    if (p.n_streams > 148 * 16) return launch_k4_t<1, 20>(p, stream);
    return p.n_streams > 148 * 12 ? launch_k4_t<1, 16>(p, stream) : launch_k4_t<1, 1>(p, stream);
  }
  if (t <= 64) return launch_k4_t<2, 7>(p, stream);
  return launch_k4_t<4, 7>(p, stream);
}

}  // namespace speedy

#ifdef K4_TIMING
// developer build only (SPEEDY_K4_TIMING=1): read / reset the per-phase cycle counters
extern "C" void speedyDebugK4Cycles(unsigned long long* out, int reset) {
  if (out) {
    cudaMemcpyFromSymbol(out, speedy::g_k4_cycles, sizeof(unsigned long long) * 16);
    unsigned long long a[8];  // the 16 kHz search's own phases (amdf16.cuh)
    cudaMemcpyFromSymbol(a, speedy::amdf16::g_amdf_cycles, sizeof(a));
    for (int i = 0; i < 5; i++) out[1 + i] += a[i];
  }
  if (reset) {
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(speedy::g_k4_cycles, z, sizeof(z));
    cudaMemcpyToSymbol(speedy::amdf16::g_amdf_cycles, z, sizeof(unsigned long long) * 8);
  }
}
#endif
