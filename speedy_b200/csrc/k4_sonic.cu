// K4 — Sonic time-scale modification: AMDF pitch-period search and
// pitch-synchronous overlap-add, one warp per stream.
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
// Why one warp per stream.  The splice cursor is strictly sequential (the next
// position depends on the period just found), so a stream is a chain of ~80 pitch
// iterations per second of audio; parallelism comes from the streams.  A warp
// needs no block barrier, no cross-warp reduction and executes the uniform
// bookkeeping once instead of once per warp; the profile of the earlier
// multi-warp version was dominated by exactly that overhead.
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
// last block of a lag group are the same code under a per-element mask.
#include <stdlib.h>

#include "kernels.cuh"

namespace speedy {

namespace {

constexpr int kPad = 16;  // over-read slack behind the window and the decimated copy
constexpr unsigned kFull = 0xffffffffu;

struct Sonic {
  // geometry
  int C, S, minP, maxP, maxReq, skip;
  long long cap;
  // shared memory (this warp's slice)
  int* w32;                 // mono window, 32-bit [bufN + kPad]
  short* buf;               // interleaved raw window (C > 1 only) [bufN * C]
  int* ds32;                // decimated mono [maxReq / skip + kPad]
  unsigned* acc;            // per-lag AMDF sums
  unsigned short* items;    // coarse pass: work item -> (group << 8 | block)
  int n_items;
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
  int lane;
  // fine-pass lane mapping: G sub-lanes per lag group
  int fG, fPerRound, fGi0, fg;
  unsigned dec_magic;

  // Make [start, start + count) resident in the shared window (count <= bufN - 8).
  __device__ __forceinline__ void ensure(long long start, int count) {
    if (start >= bufStart && start + count <= bufStart + bufLen) return;
    __syncwarp();  // every lane is done with the old window
    bufStart = start & ~7LL;  // keeps the 16-byte loads of the refill aligned
    bufLen = bufN;
    stage_mono<32, int>(src, bufStart, bufN, zero_from, w32, C > 1 ? buf : nullptr, lane);
    __syncwarp();
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
    const int total = n * C;
    const long long base = (long long)(outCount + out_offset_frames) * C;
    const long long room = cap * C - base;
    short* o = out + base;
    if (C == 1) {
      for (int i = lane; i < total; i += 32) {
        if (i < room) o[i] = (short)w32[o0 + i];
      }
    } else {
      const short* p = buf + (size_t)o0 * C;
      for (int i = lane; i < total; i += 32) {
        if (i < room) o[i] = p[i];
      }
    }
  }

  // out[t] = (down[t]*(n-t) + up[t]*t) / n per channel, C integer arithmetic.
  // down/up are absolute frames inside the window.
  __device__ __forceinline__ void overlap_add(int n, long long down, long long up, int out_offset_frames) {
    const int d0 = (int)(down - bufStart), u0 = (int)(up - bufStart);
    const int total = n * C;
    const long long base = (long long)(outCount + out_offset_frames) * C;
    const long long room = cap * C - base;
    short* o = out + base;
    if (C == 1) {
      for (int t = lane; t < total; t += 32) {
        int v = (w32[d0 + t] * (n - t) + w32[u0 + t] * t) / n;
        if (t < room) o[t] = (short)v;
      }
    } else {
      const short* dp = buf + (size_t)d0 * C;
      const short* up_ = buf + (size_t)u0 * C;
      for (int i = lane; i < total; i += 32) {
        int t = i / C;
        int v = ((int)dp[i] * (n - t) + (int)up_[i] * t) / n;
        if (i < room) o[i] = (short)v;
      }
    }
  }

  // Upstream downSampleInput: sum `skip` frames x C channels, C integer division
  // (truncating).  |sum| < 2^21 and the divisor is small, so the quotient is exact
  // as (|sum| * ceil(2^32 / divisor)) >> 32.
  __device__ __forceinline__ void decimate(int off) {
    const int count = maxReq / skip;
    const int per = C * skip;
    for (int i = lane; i < count; i += 32) {
      int v = 0;
      if (C == 1) {
        const int* q = w32 + off + i * skip;
        for (int j = 0; j < skip; j++) v += q[j];
      } else {
        const short* q = buf + ((size_t)off + (size_t)i * skip) * C;
        for (int j = 0; j < per; j++) v += q[j];
      }
      const int qa = (int)__umulhi((unsigned)abs(v), dec_magic);
      ds32[i] = v < 0 ? -qa : qa;
    }
    __syncwarp();
  }

  // |a - b| sums of one aligned block of four samples for the four lags
  // pg .. pg+3 (pg a multiple of four).  Sample x = blk + m contributes to lag
  // pg + l iff off <= x < off + pg + l.
  __device__ __forceinline__ void block_sads(const int* arr, int blk, int pg, int off, unsigned (&d)[4]) {
    const int4 av = *reinterpret_cast<const int4*>(arr + blk);
    const int4 b0 = *reinterpret_cast<const int4*>(arr + blk + pg);
    const int4 b1 = *reinterpret_cast<const int4*>(arr + blk + pg + 4);
    if (blk >= off && blk + 3 < off + pg) {
      d[0] = __sad(av.x, b0.x, d[0]); d[0] = __sad(av.y, b0.y, d[0]);
      d[0] = __sad(av.z, b0.z, d[0]); d[0] = __sad(av.w, b0.w, d[0]);
      d[1] = __sad(av.x, b0.y, d[1]); d[1] = __sad(av.y, b0.z, d[1]);
      d[1] = __sad(av.z, b0.w, d[1]); d[1] = __sad(av.w, b1.x, d[1]);
      d[2] = __sad(av.x, b0.z, d[2]); d[2] = __sad(av.y, b0.w, d[2]);
      d[2] = __sad(av.z, b1.x, d[2]); d[2] = __sad(av.w, b1.y, d[2]);
      d[3] = __sad(av.x, b0.w, d[3]); d[3] = __sad(av.y, b1.x, d[3]);
      d[3] = __sad(av.z, b1.y, d[3]); d[3] = __sad(av.w, b1.z, d[3]);
    } else {
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
  }

  // Exact arg-min / arg-max of diff/period over the warp's candidates (one per
  // lane, period 0 = none).  The C scan compares by cross-multiplication with
  // strict inequalities, so ties go to the smaller lag.  A float quotient picks
  // the lanes within 2e-6 of the extremum (a superset of the true extremum, its
  // relative error is below 4e-7); almost always that is one lane, otherwise the
  // short list is resolved exactly.
  template <bool kMin>
  __device__ __forceinline__ void pick(unsigned diff, int period, unsigned* out_diff, int* out_period) {
    const float f = period ? __fdividef((float)diff, (float)period) : (kMin ? 3.0e38f : 0.0f);
    const unsigned key = __float_as_uint(f);
    const float ext = __uint_as_float(kMin ? __reduce_min_sync(kFull, key) : __reduce_max_sync(kFull, key));
    const bool near = period != 0 && (kMin ? f <= ext * 1.000002f : f >= ext * 0.999998f);
    unsigned ballot = __ballot_sync(kFull, near);
    int src_lane = __ffs(ballot) - 1;
    unsigned bd = __shfl_sync(kFull, diff, src_lane);
    int bp = __shfl_sync(kFull, period, src_lane);
    ballot &= ballot - 1;
    while (ballot) {  // rare: several candidates within rounding of each other
      src_lane = __ffs(ballot) - 1;
      ballot &= ballot - 1;
      const unsigned cd = __shfl_sync(kFull, diff, src_lane);
      const int cp = __shfl_sync(kFull, period, src_lane);
      const unsigned long long l = (unsigned long long)cd * (unsigned)bp;
      const unsigned long long r = (unsigned long long)bd * (unsigned)cp;
      const bool better = kMin ? (l < r || (l == r && cp < bp)) : (l > r || (l == r && cp < bp));
      if (better) {
        bd = cd;
        bp = cp;
      }
    }
    *out_diff = bd;
    *out_period = bp;
  }

  // AMDF over lags lo..hi on a[i] = arr[off + i].  Returns the best lag; *minDiff /
  // *maxDiff are the per-sample differences at the best and worst lag.
  __device__ __forceinline__ int search(const int* arr, int off, int lo, int hi, bool coarse, int* minDiff,
                                        int* maxDiff) {
    const int g0 = lo >> 2;  // first lag group (lags 4*g0 .. 4*g0+3)
    const int ngroups = (hi >> 2) - g0 + 1;
    const int nlag = ngroups * 4;
    for (int li = lane; li < nlag; li += 32) acc[li] = 0;
    __syncwarp();
    if (coarse) {
      // off == 0 and the lag range is fixed: a precomputed flat list of blocks
      for (int w = lane; w < n_items; w += 32) {
        const int it = items[w];
        const int gi = it >> 8, j = it & 255;
        unsigned d[4] = {0u, 0u, 0u, 0u};
        block_sads(arr, 4 * j, 4 * (g0 + gi), 0, d);
        atomicAdd(&acc[4 * gi + 0], d[0]);
        atomicAdd(&acc[4 * gi + 1], d[1]);
        atomicAdd(&acc[4 * gi + 2], d[2]);
        atomicAdd(&acc[4 * gi + 3], d[3]);
      }
    } else {
      const int B0 = off & ~3;
      for (int gi = fGi0; gi < ngroups; gi += fPerRound) {
        const int pg = 4 * (g0 + gi);
        const int nblk = ((((off + pg + 2) & ~3) + 4) - B0) >> 2;
        unsigned d[4] = {0u, 0u, 0u, 0u};
        for (int j = fg; j < nblk; j += fG) block_sads(arr, B0 + 4 * j, pg, off, d);
        atomicAdd(&acc[4 * gi + 0], d[0]);
        atomicAdd(&acc[4 * gi + 1], d[1]);
        atomicAdd(&acc[4 * gi + 2], d[2]);
        atomicAdd(&acc[4 * gi + 3], d[3]);
      }
    }
    __syncwarp();
    // ---- best / worst lag ----------------------------------------------------
    unsigned bd = 0, wd = 0;
    int bp = 0, wp = 0;
    for (int li = lane; li < nlag; li += 32) {
      const int p = 4 * g0 + li;
      if (p >= lo && p <= hi) {
        const unsigned d = acc[li];
        if (bp == 0) {
          bd = wd = d;
          bp = wp = p;
        } else {
          // p is larger than the lags this lane already holds: strict comparisons
          if ((unsigned long long)d * (unsigned)bp < (unsigned long long)bd * (unsigned)p) { bd = d; bp = p; }
          if ((unsigned long long)d * (unsigned)wp > (unsigned long long)wd * (unsigned)p) { wd = d; wp = p; }
        }
      }
    }
    unsigned best_diff, worst_diff;
    int best, worst;
    pick<true>(bd, bp, &best_diff, &best);
    pick<false>(wd, wp, &worst_diff, &worst);
    __syncwarp();  // acc[] is rewritten by the next search
    // the C scan starts from (maxDiff = 0, worstPeriod = 255) and only replaces
    // it with a strictly larger ratio
    if (worst_diff == 0u) worst = 255;
    *minDiff = (int)(best_diff / (unsigned)best);
    *maxDiff = (int)(worst_diff / (unsigned)worst);
    return best;
  }

  __device__ __forceinline__ int find_pitch_period(long long pos) {
    const int off = (int)(pos - bufStart);
    int minDiff = 0, maxDiff = 0, period = 0;
    const int* arr = w32;
    int aoff = off;
    int lo = minP, hi = maxP, stages = 1;
    if (!(C == 1 && skip == 1)) {
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
    int result = period;
    if (minDiff != 0 && prevPeriod != 0) {
      if (!(maxDiff > minDiff * 3) && !(minDiff * 2 <= prevMinDiff * 3)) result = prevPeriod;
    }
    prevMinDiff = minDiff;
    prevPeriod = period;
    return result;
  }

  // processStreamInput with the speed that is current now.
  __device__ __forceinline__ void process(float speed) {
    const long long numInput = fed - head;
    if ((double)speed > 1.00001 || (double)speed < 0.99999) {
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

__host__ __device__ inline int k4_acc_entries(const Geometry& g) {
  int coarse = g.max_period / g.skip - g.min_period / g.skip + 1;
  int fine = g.skip != 1 ? 8 * g.skip + 1 : 0;
  int n = (coarse > fine ? coarse : fine) + 8;
  return (n + 3) & ~3;
}

// coarse work items: groups q = lo/4 .. hi/4, blocks 0 .. q (the last one masked)
__host__ __device__ inline int k4_coarse_items(const Geometry& g) {
  const int lo = g.min_period / g.skip, hi = g.max_period / g.skip;
  int n = 0;
  for (int q = lo >> 2; q <= (hi >> 2); q++) n += q + 1;
  return n;
}

__host__ __device__ inline size_t k4_warp_smem(const Geometry& g, int buf_frames) {
  size_t b = (size_t)(buf_frames + kPad) * sizeof(int);
  b += (size_t)((g.max_required / g.skip + kPad + 3) & ~3) * sizeof(int);
  b += (size_t)k4_acc_entries(g) * sizeof(unsigned);
  b += (size_t)((k4_coarse_items(g) + 7) & ~7) * sizeof(unsigned short);
  if (g.channels > 1) b += (size_t)buf_frames * g.channels * sizeof(short);
  return (b + 15) & ~(size_t)15;
}

}  // namespace

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k4_sonic(K4Params p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5;
  const int s = blockIdx.x * WARPS + warp;
  if (s >= p.n_streams) return;
  const Geometry& g = p.g;

  Sonic k;
  k.lane = threadIdx.x & 31;
  k.C = g.channels;
  k.S = g.step;
  k.minP = g.min_period;
  k.maxP = g.max_period;
  k.maxReq = g.max_required;
  k.skip = g.skip;
  k.cap = p.out_capacity;
  k.bufN = p.buf_frames;
  // carve-up of this warp's slice (every piece a multiple of 16 bytes)
  unsigned char* base = smem_raw + (size_t)warp * k4_warp_smem(g, p.buf_frames);
  k.w32 = reinterpret_cast<int*>(base);
  k.ds32 = k.w32 + k.bufN + kPad;
  k.acc = reinterpret_cast<unsigned*>(k.ds32 + ((k.maxReq / k.skip + kPad + 3) & ~3));
  k.items = reinterpret_cast<unsigned short*>(k.acc + k4_acc_entries(g));
  k.n_items = k4_coarse_items(g);
  k.buf = reinterpret_cast<short*>(k.items + ((k.n_items + 7) & ~7));
  k.bufStart = 0;
  k.bufLen = 0;
  k.dec_magic = (unsigned)((0x100000000ULL + (unsigned)(k.C * k.skip) - 1) / (unsigned)(k.C * k.skip));
  {
    // coarse item list and fine-pass lane mapping (lags rounded out to fours)
    const int c_lo = k.minP / k.skip, c_hi = k.maxP / k.skip;
    if (k.lane == 0) {
      int w = 0;
      for (int q = c_lo >> 2; q <= (c_hi >> 2); q++) {
        for (int j = 0; j <= q; j++) k.items[w++] = (unsigned short)(((q - (c_lo >> 2)) << 8) | j);
      }
    }
    const int fine_groups = k.skip != 1 ? 2 * k.skip + 2 : (c_hi >> 2) - (c_lo >> 2) + 1;
    k.fG = 32 / fine_groups;
    if (k.fG < 1) k.fG = 1;
    k.fPerRound = 32 / k.fG;
    k.fGi0 = k.lane / k.fG;
    k.fg = k.lane - k.fGi0 * k.fG;
    if (k.lane >= k.fPerRound * k.fG) k.fGi0 = 1 << 30;  // idle lane
  }
  for (int i = k.lane; i < kPad; i += 32) {  // the over-read pads
    k.w32[k.bufN + i] = 0;
    k.ds32[k.maxReq / k.skip + i] = 0;
  }
  __syncwarp();

  const long long t_old = p.st.total[s];
  const long long t_new = p.flush ? t_old : t_old + (p.counts ? p.counts[s] : p.frames);
  k.src.channels = g.channels;
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
      rA = tensions_ready(g, frames_analyzed(g, t_old));
      ev = rA;
      ev_end = tensions_ready(g, frames_analyzed(g, t_new));
    } else {
      ev = 0;
      ev_end = t_new > t_old ? 1 : 0;
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
        if ((j & 31) == 0) {
          const long long idx = (long long)j + k.lane;
          speed_batch = idx < ev_end - rA ? sp[idx] : 0.0f;
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
      k.head = k.fed;
      k.remCopy = 0;
      k.status |= 2;  // SPEEDY_STATUS_FLUSHED
    }
  }

  if (k.lane == 0) {
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

static int k4_buf_frames(const Geometry& g, int n_streams) {
  // window: several search spans; smaller when many streams share an SM
  int n = (n_streams >= 148 * 12 ? 4 : 8) * g.max_required;
  const int floor_n = n_streams >= 148 * 12 ? 2048 : 4096;
  if (n < floor_n) n = floor_n;
  if (const char* e = getenv("SPEEDY_K4_BUF")) n = atoi(e) > 2 * g.max_required ? atoi(e) : n;
  return (n + 63) & ~63;
}

template <int WARPS>
static cudaError_t launch_k4_t(K4Params& p, cudaStream_t stream) {
  const size_t smem = (size_t)WARPS * k4_warp_smem(p.g, p.buf_frames);
  static size_t attr = 0;
  if (smem > attr) {
    cudaError_t e = cudaFuncSetAttribute(k4_sonic<WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr = smem;
  }
  const int blocks = (p.n_streams + WARPS - 1) / WARPS;
  k4_sonic<WARPS><<<blocks, WARPS * 32, smem, stream>>>(p);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_k4(const K4Params& p0, cudaStream_t stream) {
  K4Params p = p0;
  p.buf_frames = k4_buf_frames(p.g, p.n_streams);
  // one warp per stream; CTAs of one warp keep the grid fine-grained
  return launch_k4_t<1>(p, stream);
}

}  // namespace speedy
