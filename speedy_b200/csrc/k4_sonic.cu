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
#include "kernels.cuh"

namespace speedy {

namespace {

struct Cand {           // one AMDF candidate: summed |difference| at a lag
  unsigned diff;
  int period;           // 0 = no candidate
};

// "a is a better minimum than b": smaller diff/period, ties to the smaller lag
// (the scan order of the C loop with strict inequalities).
__device__ __forceinline__ bool better_min(const Cand& a, const Cand& b) {
  if (b.period == 0) return a.period != 0;
  if (a.period == 0) return false;
  unsigned long long l = (unsigned long long)a.diff * (unsigned)b.period;
  unsigned long long r = (unsigned long long)b.diff * (unsigned)a.period;
  return l < r || (l == r && a.period < b.period);
}
__device__ __forceinline__ bool better_max(const Cand& a, const Cand& b) {
  if (b.period == 0) return a.period != 0;
  if (a.period == 0) return false;
  unsigned long long l = (unsigned long long)a.diff * (unsigned)b.period;
  unsigned long long r = (unsigned long long)b.diff * (unsigned)a.period;
  return l > r || (l == r && a.period < b.period);
}

__device__ __forceinline__ Cand shfl_xor_cand(const Cand& c, int mask) {
  Cand o;
  o.diff = __shfl_xor_sync(0xffffffffu, c.diff, mask);
  o.period = __shfl_xor_sync(0xffffffffu, c.period, mask);
  return o;
}

template <int THREADS>
struct Sonic {
  static constexpr int NW = THREADS / 32;

  // geometry
  int C, S, minP, maxP, maxReq, skip;
  long long cap;
  // shared memory
  short* buf;     // window of interleaved frames [bufN * C]
  short* ds;      // decimated mono [maxReq / skip]
  short* mono;    // full-rate mono (C > 1) [maxReq]
  Cand* red;      // [2 * NW]
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

  __device__ __forceinline__ int sample(long long frame, int c) const {
    if (frame >= zero_from) return 0;
    return src.raw(frame, c);
  }

  // Make [start, start + count) resident in the shared window.
  __device__ void ensure(long long start, int count) {
    if (start >= bufStart && start + count <= bufStart + bufLen) return;
    __syncthreads();  // everyone is done with the old window
    bufStart = start;
    bufLen = bufN;
    const int total = bufN * C;
    const long long first = start * C;
    const long long lim_zero = zero_from * C;
    const long long lim_data = src.t_new * C;
    for (int i = tid; i < total; i += THREADS) {
      long long e = first + i;
      short v = 0;
      if (e < lim_zero && e < lim_data) {
        long long frame = e / C;
        v = (short)src.raw(frame, (int)(e - frame * C));
      }
      buf[i] = v;
    }
    __syncthreads();
  }

  __device__ __forceinline__ const short* at(long long frame) const {
    return buf + (frame - bufStart) * C;
  }

  // Append n frames starting at window frame `from` to the output.
  __device__ void emit_copy(long long from, int n) {
    ensure(from, n);
    const short* p = at(from);
    const int total = n * C;
    const long long room = (cap - outCount) * C;
    short* o = out + (long long)outCount * C;
    for (int i = tid; i < total; i += THREADS) {
      if (i < room) o[i] = p[i];
    }
    advance_out(n);
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

  // out[t] = (down[t]*(n-t) + up[t]*t) / n per channel, C integer arithmetic.
  __device__ void overlap_add(int n, const short* down, const short* up, int out_offset_frames) {
    const int total = n * C;
    const long long base = (long long)(outCount + out_offset_frames) * C;
    const long long room = cap * C - base;
    short* o = out + base;
    for (int i = tid; i < total; i += THREADS) {
      int t = i / C;
      int v = ((int)down[i] * (n - t) + (int)up[i] * t) / n;
      if (i < room) o[i] = (short)v;
    }
  }

  // Sum over channels and `sk` consecutive frames, C integer division.
  __device__ void decimate(const short* x, int sk, short* dst, int count) {
    const int per = C * sk;
    for (int i = tid; i < count; i += THREADS) {
      const short* q = x + (size_t)i * per;
      int v = 0;
      for (int j = 0; j < per; j++) v += q[j];
      dst[i] = (short)(v / per);
    }
    __syncthreads();
  }

  // AMDF over lags lo..hi on `a`.  Returns the best lag; *minDiff / *maxDiff are
  // the per-sample differences at the best and worst lag.
  __device__ __forceinline__ int search(const short* a, int lo, int hi, int* minDiff, int* maxDiff) {
    const int nl = hi - lo + 1;
    // G sub-lanes cooperate on one lag (power of two, groups of adjacent lanes)
    int G = 1;
    while (G < 32 && 2 * G * nl <= THREADS) G *= 2;
    const int g = tid & (G - 1);
    const int slots = THREADS / G;  // lags evaluated concurrently
    Cand best = {0u, 0}, worst = {0u, 0};
    const int rounds = (nl + slots - 1) / slots;
    for (int round = 0; round < rounds; round++) {
      // every thread runs every round: the shuffles below need the whole warp
      const int li = tid / G + round * slots;
      const bool live = li < nl;
      const int p = lo + (live ? li : 0);
      unsigned d = 0;
      if (live) {
        const short* b = a + p;
        int i = g;
        for (; i + 3 * G < p; i += 4 * G) {
          d += __sad((int)a[i], (int)b[i], 0u) + __sad((int)a[i + G], (int)b[i + G], 0u) +
               __sad((int)a[i + 2 * G], (int)b[i + 2 * G], 0u) +
               __sad((int)a[i + 3 * G], (int)b[i + 3 * G], 0u);
        }
        for (; i < p; i += G) d += __sad((int)a[i], (int)b[i], 0u);
      }
      for (int m = 1; m < G; m <<= 1) d += __shfl_xor_sync(0xffffffffu, d, m);
      if (live) {
        Cand c = {d, p};
        if (better_min(c, best)) best = c;
        if (better_max(c, worst)) worst = c;
      }
    }
    // CTA-wide reduction
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
      Cand ob = shfl_xor_cand(best, m), ow = shfl_xor_cand(worst, m);
      if (better_min(ob, best)) best = ob;
      if (better_max(ow, worst)) worst = ow;
    }
    if (NW > 1) {
      __syncthreads();
      if ((tid & 31) == 0) {
        red[tid >> 5] = best;
        red[NW + (tid >> 5)] = worst;
      }
      __syncthreads();
      best = red[0];
      worst = red[NW];
#pragma unroll
      for (int w = 1; w < NW; w++) {
        if (better_min(red[w], best)) best = red[w];
        if (better_max(red[NW + w], worst)) worst = red[NW + w];
      }
    }
    // the C scan starts from (maxDiff = 0, worstPeriod = 255) and only replaces
    // it with a strictly larger ratio
    if (worst.diff == 0u) {
      worst.diff = 0u;
      worst.period = 255;
    }
    *minDiff = (int)(best.diff / (unsigned)best.period);
    *maxDiff = (int)(worst.diff / (unsigned)worst.period);
    return best.period;
  }

  __device__ __forceinline__ int find_pitch_period(long long pos) {
    const short* x = at(pos);
    int minDiff = 0, maxDiff = 0, period = 0;
    const short* arr = x;
    int lo = minP, hi = maxP, stages = 1;
    __syncthreads();  // previous readers of ds/mono are done
    if (!(C == 1 && skip == 1)) {
      decimate(x, skip, ds, maxReq / skip);
      arr = ds;
      lo = minP / skip;
      hi = maxP / skip;
      stages = skip != 1 ? 2 : 1;
    }
    for (int stage = 0; stage < stages; stage++) {
      period = search(arr, lo, hi, &minDiff, &maxDiff);
      if (stage == 0 && stages == 2) {
        // refine around the coarse estimate at the full rate
        period *= skip;
        lo = period - (skip << 2);
        hi = period + (skip << 2);
        if (lo < minP) lo = minP;
        if (hi > maxP) hi = maxP;
        if (C == 1) {
          arr = x;
        } else {
          decimate(x, 1, mono, maxReq);
          arr = mono;
        }
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
        if (remCopy > 0) {
          newSamples = remCopy < maxReq ? remCopy : maxReq;
          emit_copy(head + position, newSamples);
          remCopy -= newSamples;
          position += newSamples;
        } else {
          const long long pos = head + position;
          ensure(pos, maxReq);
          const int period = find_pitch_period(pos);
          const short* x = at(pos);
          if (speed > 1.0f) {
            if (speed >= 2.0f) {
              newSamples = (int)(long long)__fdiv_rn((float)period, __fsub_rn(speed, 1.0f));
            } else {
              newSamples = period;
              remCopy = (int)__fdiv_rn(__fmul_rn((float)period, __fsub_rn(2.0f, speed)),
                                       __fsub_rn(speed, 1.0f));
            }
            overlap_add(newSamples, x, x + (size_t)period * C, 0);
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
            {
              const int total = period * C;
              const long long room = (cap - outCount) * C;
              short* o = out + (long long)outCount * C;
              for (int i = tid; i < total; i += THREADS) {
                if (i < room) o[i] = x[i];
              }
            }
            overlap_add(newSamples, x + (size_t)period * C, x, period);
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
        int n = left < bufN ? (int)left : bufN;
        emit_copy(head, n);
        head += n;
        left -= n;
      }
    }
  }
};

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
  k.buf = reinterpret_cast<short*>(smem_raw);
  k.ds = k.buf + (size_t)k.bufN * k.C;
  k.mono = k.ds + ((k.maxReq / k.skip + 7) & ~7);
  k.red = reinterpret_cast<Cand*>(k.mono + (k.C > 1 ? ((k.maxReq + 7) & ~7) : 8));
  k.bufStart = 0;
  k.bufLen = 0;

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
    // the history must still hold everything from head on (tail kernel checks)
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
  // window: a few search spans, multiple of 64 frames
  int n = 4 * g.max_required;
  if (n < 2048) n = 2048;
  return (n + 63) & ~63;
}

static size_t k4_smem(const Geometry& g, int buf_frames, int nw) {
  size_t b = (size_t)buf_frames * g.channels * sizeof(short);
  b += (size_t)((g.max_required / g.skip + 7) & ~7) * sizeof(short);
  b += (size_t)(g.channels > 1 ? ((g.max_required + 7) & ~7) : 8) * sizeof(short);
  b += (size_t)2 * nw * sizeof(Cand);
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
  if (t == 0) t = p.n_streams >= 148 * 24 ? 32 : (p.n_streams >= 148 * 8 ? 64 : 128);
  if (t <= 32) return launch_k4_t<32>(p, stream);
  if (t <= 64) return launch_k4_t<64>(p, stream);
  return launch_k4_t<128>(p, stream);
}

}  // namespace speedy
