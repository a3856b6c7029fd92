// K4 — Sonic time-scale modification: AMDF pitch-period search and
// pitch-synchronous overlap-add, one CTA per stream.
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
// associative, so splitting the AMDF sums across threads cannot change a result:
// given identical per-frame speeds the output is bit-exact.
//
// Sonic's input FIFO is never materialised: the stream keeps two absolute
// cursors (head = first unconsumed frame, fed = one past the last frame handed
// to Sonic) and the CTA slides a shared-memory window over the caller's buffer.
// The output cursor is the per-stream pending count in the output buffer.
//
// AMDF layout.  The window is kept in shared memory as 32-bit mono samples, so
// that one LDS.128 fetches four operands ready for VABSDIFF (|a-b|+c in one
// instruction).  Lags are processed in groups of four consecutive lags starting
// at a multiple of four: for an aligned block of four samples a[blk..blk+3] the
// operands of lags 4k..4k+3 are the seven values b[blk+4k .. blk+4k+6], i.e. two
// more aligned LDS.128.  3 loads + 16 VABSDIFF per 16 differences.  The few
// samples before the first / after the last aligned block of each lag (at most
// nine) are summed by one thread per lag, which also initialises the per-lag
// accumulator the block sums are added to.
#include <stdlib.h>

#include "kernels.cuh"

namespace speedy {

namespace {

template <int THREADS>
struct Sonic {
  static constexpr int NW = THREADS / 32;

  // geometry
  int C, S, minP, maxP, maxReq, skip;
  long long cap;
  // shared memory
  int* w32;        // mono window, 32-bit [bufN + 8]
  short* buf;      // interleaved raw window (C > 1 only) [bufN * C]
  int* ds32;       // decimated mono [maxReq / skip + 8]
  unsigned* acc;   // per-lag AMDF sums
  unsigned long long* red;  // [2 * NW]
  int bufN;
  // window state (uniform across the CTA)
  long long bufStart;
  int bufLen;
  // source
  Source src;
  long long zero_from;  // frames >= this read as silence (flush padding)
  // stream state (uniform)
  long long head, fed, outTotal;
  int prevPeriod, prevMinDiff, remCopy, outCount, status;
  short* out;
  int tid;

  // Make [start, start + count) resident in the shared window (count <= bufN - 8).
  __device__ void ensure(long long start, int count) {
    if (start >= bufStart && start + count <= bufStart + bufLen) return;
    __syncthreads();  // everyone is done with the old window
    bufStart = start & ~7LL;  // keeps the 16-byte loads of the refill aligned
    bufLen = bufN;
    stage_mono<THREADS, int>(src, bufStart, bufN, zero_from, w32, C > 1 ? buf : nullptr, tid);
    __syncthreads();
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
  __device__ void emit_copy(long long from, int n, int out_offset_frames) {
    const int o0 = (int)(from - bufStart);
    const int total = n * C;
    const long long base = (long long)(outCount + out_offset_frames) * C;
    const long long room = cap * C - base;
    short* o = out + base;
    if (C == 1) {
      for (int i = tid; i < total; i += THREADS) {
        if (i < room) o[i] = (short)w32[o0 + i];
      }
    } else {
      const short* p = buf + (size_t)o0 * C;
      for (int i = tid; i < total; i += THREADS) {
        if (i < room) o[i] = p[i];
      }
    }
  }

  // out[t] = (down[t]*(n-t) + up[t]*t) / n per channel, C integer arithmetic.
  // down/up are absolute frames inside the window.
  __device__ void overlap_add(int n, long long down, long long up, int out_offset_frames) {
    const int d0 = (int)(down - bufStart), u0 = (int)(up - bufStart);
    const int total = n * C;
    const long long base = (long long)(outCount + out_offset_frames) * C;
    const long long room = cap * C - base;
    short* o = out + base;
    if (C == 1) {
      for (int t = tid; t < total; t += THREADS) {
        int v = (w32[d0 + t] * (n - t) + w32[u0 + t] * t) / n;
        if (t < room) o[t] = (short)v;
      }
    } else {
      const short* dp = buf + (size_t)d0 * C;
      const short* up_ = buf + (size_t)u0 * C;
      for (int i = tid; i < total; i += THREADS) {
        int t = i / C;
        int v = ((int)dp[i] * (n - t) + (int)up_[i] * t) / n;
        if (i < room) o[i] = (short)v;
      }
    }
  }

  // Upstream downSampleInput: sum `skip` frames x C channels, C integer division
  // (truncating).  |sum| < 2^21 and the divisor is small, so the quotient is exact
  // as (|sum| * ceil(2^32 / divisor)) >> 32.
  __device__ void decimate(int off) {
    const int count = maxReq / skip;
    const int per = C * skip;
    const unsigned magic = (unsigned)((0x100000000ULL + per - 1) / per);
    for (int i = tid; i < count; i += THREADS) {
      int v = 0;
      if (C == 1) {
        const int* q = w32 + off + i * skip;
        for (int j = 0; j < skip; j++) v += q[j];
      } else {
        const short* q = buf + ((size_t)off + (size_t)i * skip) * C;
        for (int j = 0; j < per; j++) v += q[j];
      }
      const int qa = (int)__umulhi((unsigned)abs(v), magic);
      ds32[i] = v < 0 ? -qa : qa;
    }
    __syncthreads();
  }

  // Thread mapping of the aligned-block pass for up to `max_groups` lag groups:
  // G sub-lanes per group, fixed per search stage so that no thread divides.
  struct Map {
    int G, per_round, gi0, g;
    __device__ void init(int max_groups, int tid) {
      G = THREADS / max_groups;
      if (G < 1) G = 1;
      per_round = THREADS / G;
      gi0 = tid / G;
      g = tid - gi0 * G;
      if (tid >= per_round * G) gi0 = 1 << 30;  // idle thread
    }
  };
  Map map_coarse, map_fine;

  // AMDF over lags lo..hi on a[i] = arr[off + i].  Returns the best lag; *minDiff /
  // *maxDiff are the per-sample differences at the best and worst lag.
  __device__ __forceinline__ int search(const int* arr, int off, int lo, int hi, const Map& map,
                                        int* minDiff, int* maxDiff) {
    const int g0 = lo >> 2;                 // first lag group (lags 4*g0 .. 4*g0+3)
    const int ngroups = (hi >> 2) - g0 + 1;
    const int nlag = ngroups * 4;
    const int A0 = (off + 3) & ~3;          // first aligned block inside the range
    // ---- edge samples, one thread per lag; initialises acc[] -------------------
    for (int li = tid; li < nlag; li += THREADS) {
      const int p = 4 * g0 + li;
      const int pg = p & ~3;
      int A1 = (off + pg) & ~3;             // end of the group's aligned blocks
      if (A1 < A0) A1 = A0;
      unsigned d = 0;
      if (p >= lo && p <= hi) {
        const int* a = arr + off;
        const int head_n = min(A0 - off, p);
        for (int i = 0; i < head_n; i++) d = __sad(a[i], a[i + p], d);
        for (int i = max(A1 - off, head_n); i < p; i++) d = __sad(a[i], a[i + p], d);
      }
      acc[li] = d;
    }
    __syncthreads();
    // ---- aligned blocks: 4 lags x 4 samples per step ----------------------------
    for (int gi = map.gi0; gi < ngroups; gi += map.per_round) {
      const int pg = 4 * (g0 + gi);
      const int A1 = (off + pg) & ~3;
      const int nb = (A1 - A0) >> 2;
      unsigned d0 = 0, d1 = 0, d2 = 0, d3 = 0;
      for (int j = map.g; j < nb; j += map.G) {
        const int blk = A0 + 4 * j;
        const int4 av = *reinterpret_cast<const int4*>(arr + blk);
        const int4 b0 = *reinterpret_cast<const int4*>(arr + blk + pg);
        const int4 b1 = *reinterpret_cast<const int4*>(arr + blk + pg + 4);
        d0 = __sad(av.x, b0.x, d0); d0 = __sad(av.y, b0.y, d0);
        d0 = __sad(av.z, b0.z, d0); d0 = __sad(av.w, b0.w, d0);
        d1 = __sad(av.x, b0.y, d1); d1 = __sad(av.y, b0.z, d1);
        d1 = __sad(av.z, b0.w, d1); d1 = __sad(av.w, b1.x, d1);
        d2 = __sad(av.x, b0.z, d2); d2 = __sad(av.y, b0.w, d2);
        d2 = __sad(av.z, b1.x, d2); d2 = __sad(av.w, b1.y, d2);
        d3 = __sad(av.x, b0.w, d3); d3 = __sad(av.y, b1.x, d3);
        d3 = __sad(av.z, b1.y, d3); d3 = __sad(av.w, b1.z, d3);
      }
      if (map.g < nb) {
        atomicAdd(&acc[4 * gi + 0], d0);
        atomicAdd(&acc[4 * gi + 1], d1);
        atomicAdd(&acc[4 * gi + 2], d2);
        atomicAdd(&acc[4 * gi + 3], d3);
      }
    }
    __syncthreads();
    // ---- best / worst lag ---------------------------------------------------------
    // The C scan keeps the lag with the smallest (largest) diff/period, comparing by
    // cross-multiplication with strict inequalities, i.e. ties go to the smaller lag.
    // key = floor(diff * 2^23 / period) orders the ratios exactly: two different
    // ratios of integers with periods < 2^11 differ by at least 2^-22, so their keys
    // differ by at least 1 (the numerator diff * 2^23 < 2^50 is exact in a double and
    // the correctly rounded quotient is monotone).  Packing the lag below the key
    // turns both selections into one 64-bit min / max.
    unsigned long long kmin = ~0ULL, kmax = 0ULL;
    for (int li = tid; li < nlag; li += THREADS) {
      const int p = 4 * g0 + li;
      if (p >= lo && p <= hi) {
        const double q = (double)((unsigned long long)acc[li] << 23) / (double)p;
        const unsigned long long key = (unsigned long long)q;
        const unsigned long long lo_key = (key << 11) | (unsigned)p;
        const unsigned long long hi_key = (key << 11) | (unsigned)(2047 - p);
        kmin = lo_key < kmin ? lo_key : kmin;
        kmax = hi_key > kmax ? hi_key : kmax;
      }
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
      const unsigned long long a = __shfl_xor_sync(0xffffffffu, kmin, m);
      const unsigned long long b = __shfl_xor_sync(0xffffffffu, kmax, m);
      kmin = a < kmin ? a : kmin;
      kmax = b > kmax ? b : kmax;
    }
    if (NW > 1) {
      if ((tid & 31) == 0) {
        red[tid >> 5] = kmin;
        red[NW + (tid >> 5)] = kmax;
      }
      __syncthreads();
      kmin = red[0];
      kmax = red[NW];
#pragma unroll
      for (int w = 1; w < NW; w++) {
        kmin = red[w] < kmin ? red[w] : kmin;
        kmax = red[NW + w] > kmax ? red[NW + w] : kmax;
      }
    }
    const int best = (int)(kmin & 2047);
    int worst = 2047 - (int)(kmax & 2047);
    const unsigned best_diff = acc[best - 4 * g0];
    unsigned worst_diff = acc[worst - 4 * g0];
    __syncthreads();  // acc[] and red[] may be rewritten by the next search
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
      period = search(arr, aoff, lo, hi, stage == 0 && stages == 2 ? map_coarse : map_fine, &minDiff, &maxDiff);
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

static __host__ __device__ inline int k4_acc_entries(const Geometry& g) {
  int coarse = g.max_period / g.skip - g.min_period / g.skip + 1;
  int fine = g.skip != 1 ? 8 * g.skip + 1 : 0;
  int n = (coarse > fine ? coarse : fine) + 8;
  return (n + 3) & ~3;
}

}  // namespace

template <int THREADS>
__global__ void __launch_bounds__(THREADS) k4_sonic(K4Params p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int s = blockIdx.x;
  if (s >= p.n_streams) return;
  const Geometry& g = p.g;

  Sonic<THREADS> k;
  k.tid = threadIdx.x;
  k.C = g.channels;
  k.S = g.step;
  k.minP = g.min_period;
  k.maxP = g.max_period;
  k.maxReq = g.max_required;
  k.skip = g.skip;
  k.cap = p.out_capacity;
  k.bufN = p.buf_frames;
  // carve-up (every piece a multiple of 16 bytes)
  k.w32 = reinterpret_cast<int*>(smem_raw);
  k.ds32 = k.w32 + k.bufN + 8;
  k.acc = reinterpret_cast<unsigned*>(k.ds32 + ((k.maxReq / k.skip + 8 + 3) & ~3));
  k.red = reinterpret_cast<unsigned long long*>(k.acc + k4_acc_entries(g));
  k.buf = reinterpret_cast<short*>(k.red + 2 * Sonic<THREADS>::NW);
  k.bufStart = 0;
  k.bufLen = 0;
  {
    // lag groups per search stage (lags rounded out to multiples of four)
    const bool two_stage = k.skip != 1;
    const int c_lo = k.minP / k.skip, c_hi = k.maxP / k.skip;
    const int full_groups = (c_hi >> 2) - (c_lo >> 2) + 1;
    k.map_coarse.init(full_groups, threadIdx.x);
    k.map_fine.init(two_stage ? 2 * k.skip + 2 : full_groups, threadIdx.x);
  }
  for (int i = threadIdx.x; i < 8; i += THREADS) {  // the over-read pads
    k.w32[k.bufN + i] = 0;
    k.ds32[k.maxReq / k.skip + i] = 0;
  }

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
  long long expected = 0;
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
      if (!p.flush) speed = sp[ev - rA];
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

  if (threadIdx.x == 0) {
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

static int k4_buf_frames(const Geometry& g) {
  // window: several search spans, multiple of 64 frames
  int n = 8 * g.max_required;
  if (n < 4096) n = 4096;
  if (const char* e = getenv("SPEEDY_K4_BUF")) n = atoi(e) > 2 * g.max_required ? atoi(e) : n;
  return (n + 63) & ~63;
}

static size_t k4_smem(const Geometry& g, int buf_frames, int nw) {
  size_t b = (size_t)(buf_frames + 8) * sizeof(int);
  b += (size_t)((g.max_required / g.skip + 8 + 3) & ~3) * sizeof(int);
  b += (size_t)k4_acc_entries(g) * sizeof(unsigned);
  b += (size_t)2 * nw * sizeof(unsigned long long);
  if (g.channels > 1) b += (size_t)buf_frames * g.channels * sizeof(short);
  return (b + 15) & ~(size_t)15;
}

template <int THREADS>
static cudaError_t launch_k4_t(K4Params& p, cudaStream_t stream) {
  const size_t smem = k4_smem(p.g, p.buf_frames, THREADS / 32);
  static size_t attr = 0;
  if (smem > attr) {
    cudaError_t e = cudaFuncSetAttribute(k4_sonic<THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return e;
    attr = smem;
  }
  k4_sonic<THREADS><<<p.n_streams, THREADS, smem, stream>>>(p);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_k4(const K4Params& p0, cudaStream_t stream) {
  K4Params p = p0;
  p.buf_frames = k4_buf_frames(p.g);
  int t = p.threads_per_stream;
  if (t == 0) t = p.n_streams >= 148 * 16 ? 32 : 64;
  if (t <= 32) return launch_k4_t<32>(p, stream);
  if (t <= 64) return launch_k4_t<64>(p, stream);
  return launch_k4_t<128>(p, stream);
}

}  // namespace speedy
