// K4, 16 kHz mono chain shape — Sonic time-scale modification with one warp per stream,
// built around the latency of the splice cursor (the next position depends on the period
// just found, so a stream is one long dependent chain and a launch lasts as long as the
// slowest chain):
//
//   * the stream's input is cut into overlapping windows at fixed positions (window k =
//     relative frames [k * G, k * G + N), G = N - 512 >= N - (maxRequired + over-read)), so
//     every pitch search lies inside exactly one window and the NEXT window is known long
//     before the cursor reaches it: one lane fetches it with a 1-D TMA bulk copy
//     (cp.async.bulk, completion on an mbarrier) into a raw int16 buffer while the chain
//     works in the current one; entering a window costs one shared-to-shared widening pass
//     instead of a round trip to HBM;
//   * the overlap-add / copy-through of splice i does not depend on search i + 1, so it is
//     deferred and issued inside the decimation step of the next search: its shared-memory
//     loads, multiplies and global stores run under the latency of the search's own loads;
//   * cursors are 32-bit frames relative to the launch's base; the feed events of the
//     reference's shim are walked as in k4_plan.cuh.
//
// Replaces what the reference does through upstream Sonic (soniclib.c:354, 369-370, 398 ->
// sonicIntSetSpeed, sonicIntWriteShortToStream; algorithm restated in
// oracle/sonic_oracle.c:169-355 and SURVEY.md Appendix A).  Results are bit-identical to
// k4_sonic.cu (which keeps flushes, short launches, other rates and multi-channel streams):
// the pitch search is the same code (amdf16.cuh), integer sums are associative, and the float
// expressions are the same _rn intrinsics (the file is built with --fmad=false).
#include <stdlib.h>

#include "amdf16.cuh"
#include "k4_plan.cuh"
#include "kernels.cuh"

namespace speedy {

void k4_lane_map(K4Params& p, int NW);  // k4_sonic.cu

namespace {

using amdf16::kFull;
constexpr int kOverlap = 512;  // >= maxRequired (492) + what the AMDF reads past a span (16)
constexpr int kMaxReq = 492, kMaxP = 246, kStep = 160;

extern __shared__ __align__(128) unsigned char chain_smem[];

// ---- mbarrier / bulk-copy primitives (PTX) ---------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// global -> shared bulk copy (TMA, 1-D), completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_load(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// Shared-memory carve-up for a window of N frames (byte offsets, multiples of 16).
template <int N>
struct Lay {
  static constexpr int bar = 0;
  static constexpr int win = 16;                              // int[N + 32]
  static constexpr int raw = win + (N + 32) * 4;              // short[N]: the prefetched next window
  static constexpr int ds = raw + N * 2;                      // int[128 + 32]
  static constexpr int part = ds + (128 + 32) * 4;            // unsigned[kPartWords]
  static constexpr int rcp = part + amdf16::kPartWords * 4;   // float[256]
  static constexpr int magic = rcp + 256 * 4;                 // unsigned[248]
  static constexpr int total = magic + 248 * 4;
};

// An output operation that has been sized but not yet carried out.
//   kind 1: skipPitchPeriod's cross-fade (oracle/sonic_oracle.c:263-294): n frames,
//           out[t] = (w[a + t] * (n - t) + w[a + period + t] * t) / n
//   kind 2: copy-through of n frames from window offset a
struct Pending {
  int kind, a, period, n, opos;
};

struct Chain16 {
  int* w32;
  int* ds32;
  float* rcp16;
  unsigned* sums;
  const unsigned* magic_tab;
  short* out;
  int cap;
  int lane, cGi, cSub, cG, fg, fGi0;
  int prevPeriod, prevMinDiff;
  bool timing;
  __device__ __forceinline__ int* win() const { return w32; }
  __device__ __forceinline__ int* ds() const { return ds32; }
  __device__ __forceinline__ float* rcp() const { return rcp16; }
  __device__ __forceinline__ unsigned* part() const { return sums; }
};

// ceil(2^(32+shift) / n) for 2^shift < n <= 2^(shift+1): floor from the correctly rounded
// double quotient (the true one is at least 1/n away from the integers it does not hit)
__device__ __forceinline__ unsigned division_magic(int n, int shift) {
  const double qd = __ddiv_rn((double)(1ULL << (32 + shift)), (double)n);
  return (unsigned)(unsigned long long)qd + ((n & (n - 1)) ? 1u : 0u);
}

// The first 128 elements of up to two pending operations, split into the shared-memory loads
// and the rest, so that the pitch search can put its own loads in between (amdf16.cuh: Hook).
struct EmitCtx {
  const int* w32;
  const unsigned* magic_tab;
  short* out;
  int cap, lane;
};
struct Emit {
  EmitCtx k;
  Pending p1, p2;  // p1: a cross-fade (or nothing), p2: a copy-through (or nothing)
  int x[4], y[4], z[4];
  __device__ __forceinline__ Emit(const Chain16& kk) {
    k.w32 = kk.w32;
    k.magic_tab = kk.magic_tab;
    k.out = kk.out;
    k.cap = kk.cap;
    k.lane = kk.lane;
    p1.kind = p1.a = p1.period = p1.n = p1.opos = 0;
    p2 = p1;
  }
  // (no branches up to the rare tails: the loads and stores are predicated on t < n, with
  // n = 0 for "nothing pending", so that they can be scheduled among the caller's)
  __device__ __forceinline__ void load() {
    const int* d = k.w32 + p1.a + k.lane;
    const int* u = d + p1.period;
    const int* c = k.w32 + p2.a + k.lane;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const bool in = k.lane + 32 * q < p1.n;
      x[q] = in ? d[32 * q] : 0;
      y[q] = in ? u[32 * q] : 0;
    }
#pragma unroll
    for (int q = 0; q < 4; q++) z[q] = k.lane + 32 * q < p2.n ? c[32 * q] : 0;
  }
  // overlapAdd (oracle/sonic_oracle.c:263-276), C integer arithmetic: trunc(|num| / n) ==
  // umulhi(|num|, magic) >> shift for |num| < 2^26, 1 < n < 2^11, 2^shift < n <= 2^(shift+1)
  __device__ __forceinline__ void finish() {
    const int n = p1.n;  // <= kMaxP
    const int shift = n > 1 ? 31 - __clz(n - 1) : 0;
    const unsigned magic = k.magic_tab[n];
    short* o = k.out + p1.opos;
    const int room = k.cap - p1.opos;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int t = k.lane + 32 * q;
      const int num = x[q] * (n - t) + y[q] * t;
      const int v = n == 1 ? abs(num) : (int)(__umulhi((unsigned)abs(num), magic) >> shift);
      if (t < n && t < room) o[t] = (short)(num < 0 ? -v : v);
    }
    const int m = p2.n;
    short* oc = k.out + p2.opos;
    const int room2 = k.cap - p2.opos;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int t = k.lane + 32 * q;
      if (t < m && t < room2) oc[t] = (short)z[q];
    }
    if (n > 128) {  // (lags above 128 at a speed of two or less)
      const int* d = k.w32 + p1.a;
      const int* u = d + p1.period;
      for (int t = 128 + k.lane; t < n; t += 32) {
        const int num = d[t] * (n - t) + u[t] * t;
        const int v = (int)(__umulhi((unsigned)abs(num), magic) >> shift);
        if (t < room) o[t] = (short)(num < 0 ? -v : v);
      }
    }
    if (m > 128) {
      const int* c = k.w32 + p2.a;
      for (int t = 128 + k.lane; t < m; t += 32) {
        if (t < room2) oc[t] = (short)c[t];
      }
    }
    p1.n = 0;
    p2.n = 0;
  }
  __device__ __forceinline__ void flush() {
    if (p1.n | p2.n) {
      load();
      finish();
    }
  }
};

// insertPitchPeriod (oracle/sonic_oracle.c:297-315), slow-down: the period itself, then the
// cross-fade back into it.  Not deferred (no benchmark runs below a speed of one).
__device__ __noinline__ void emit_insert(const int* w32, const unsigned* magic_tab, short* out, int cap, int lane, int a,
                                         int period, int n, int opos) {
  short* o = out + opos;
  const int room = cap - opos;
  const int* r0 = w32 + a;
  for (int t = lane; t < period; t += 32) {
    if (t < room) o[t] = (short)r0[t];
  }
  if (n <= 0) return;
  unsigned magic = 0u;
  int shift = 0;
  if (n > 1) {
    shift = 31 - __clz(n - 1);
    magic = n <= kMaxP ? magic_tab[n] : division_magic(n, shift);
  }
  const int* d = r0 + period;
  for (int t = lane; t < n; t += 32) {
    const int num = d[t] * (n - t) + r0[t] * t;
    const int v = n == 1 ? abs(num) : (int)(__umulhi((unsigned)abs(num), magic) >> shift);
    if (period + t < room) o[period + t] = (short)(num < 0 ? -v : v);
  }
}

}  // namespace

template <int N, int MINB>
__global__ void __launch_bounds__(32, MINB) k4_chain16(K4Params p) {
  using L = Lay<N>;
  constexpr int G = N - kOverlap;
  const int s = blockIdx.x;
  if (s >= p.n_streams) return;
  const int lane = threadIdx.x;

  uint64_t* bar = reinterpret_cast<uint64_t*>(chain_smem + L::bar);
  short* raw = reinterpret_cast<short*>(chain_smem + L::raw);
  unsigned* magic_tab = reinterpret_cast<unsigned*>(chain_smem + L::magic);
  Chain16 k;
  k.lane = lane;
  k.w32 = reinterpret_cast<int*>(chain_smem + L::win);
  k.ds32 = reinterpret_cast<int*>(chain_smem + L::ds);
  k.sums = reinterpret_cast<unsigned*>(chain_smem + L::part);
  k.rcp16 = reinterpret_cast<float*>(chain_smem + L::rcp);
  k.magic_tab = magic_tab;
  k.timing = false;
  k.cap = (int)p.out_capacity;
  k.out = p.out + (size_t)s * p.out_capacity;
  if (lane == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  // tables: reciprocals of the lags (the float keys of the arg-min), the overlap-add's
  // division constants, zeroed partial-sum rows and over-read pads
  for (int n = lane; n < 256; n += 32) k.rcp16[n] = n ? __frcp_rn((float)n) : 0.f;
  for (int n = lane; n <= kMaxP; n += 32) magic_tab[n] = n >= 2 ? division_magic(n, 31 - __clz(n - 1)) : 0u;
  for (int i = lane; i < amdf16::kPartWords; i += 32) k.sums[i] = 0u;
  k.w32[N + lane] = 0;
  k.ds32[128 + lane] = 0;
  {
    const unsigned m = p.lane_map[lane];
    k.cGi = (int)(m & 0xffu) - 1;
    k.cSub = (int)((m >> 8) & 0xffu);
    k.cG = (int)((m >> 16) & 0xffu);
    const int slot = lane / 3;  // fine pass: three lanes per lag group (amdf16.cuh)
    k.fg = lane - slot * 3;
    k.fGi0 = slot < 10 ? slot : (1 << 30);
  }
  __syncwarp();

  const Plan pl = make_plan(p, s);
  const long long lim_abs = pl.base + pl.data_end;
  const bool src_vec = pl.src.in != nullptr && (reinterpret_cast<size_t>(pl.src.in) & 15) == 0;

  // ---- windows ---------------------------------------------------------------
  int wk = 0, wbase = 0;       // current window and its first relative frame
  bool pre = false;            // a bulk copy of window wk + 1 is in flight (or has landed) in `raw`
  unsigned pre_parity = 0;
  // can window kk come straight from the caller's buffer as one aligned bulk copy?
  auto bulk_ok = [&](int kk) -> bool {
    const long long a0 = pl.base + (long long)kk * G;
    return src_vec && a0 >= pl.src.t_old && a0 + N <= pl.src.t_new && a0 + N <= lim_abs;
  };
  auto prefetch = [&](int kk) {
    // (every lane has finished reading `raw`: the callers synchronise the warp first)
    if ((long long)kk * G < (long long)pl.last_fed + 16 && bulk_ok(kk)) {
      if (lane == 0) {
        fence_proxy_async();
        mbar_expect_tx(bar, (unsigned)(N * sizeof(short)));
        bulk_load(raw, pl.src.in + (pl.base + (long long)kk * G - pl.src.t_old), (unsigned)(N * sizeof(short)), bar);
      }
      pre = true;
    }
  };
  auto load_window = [&](int kk) {
    __syncwarp();  // every lane is done with the old window
    if (pre) {
      mbar_wait(bar, pre_parity);
      pre_parity ^= 1u;
      pre = false;
      // widen: 8 frames per lane and round, one 16-byte load, two 16-byte stores
#pragma unroll 4
      for (int v = lane; v < N / 8; v += 32) {
        const int4 q = *reinterpret_cast<const int4*>(raw + 8 * v);
        int4* d4 = reinterpret_cast<int4*>(k.w32 + 8 * v);
        d4[0] = make_int4((short)(q.x & 0xffff), q.x >> 16, (short)(q.y & 0xffff), q.y >> 16);
        d4[1] = make_int4((short)(q.z & 0xffff), q.z >> 16, (short)(q.w & 0xffff), q.w >> 16);
      }
    } else {
      stage_mono<32, int>(pl.src, pl.base + (long long)kk * G, N, lim_abs, k.w32, nullptr, lane);
    }
    __syncwarp();
    wk = kk;
    wbase = kk * G;
    prefetch(kk + 1);
  };

  k.prevPeriod = p.st.prev_period[s];
  k.prevMinDiff = p.st.prev_min_diff[s];
  int remCopy = p.st.remaining_copy[s];
  long long outTotal = p.st.out_total[s];
  int outCount = pl.out_count0;
  const int cap = k.cap;
  int status = 0;
  float speed = p.st.sonic_speed[s];

  int pos = pl.pos0;  // the cursor, relative
  int fed_cur = pl.fed0;
  int evi = 0;
  bool doneB = false;
  // 32 speeds at a time, one per lane, the next batch already in flight
  float sp_cur = 0.f, sp_next = 0.f;
  if (pl.per_frame) {
    sp_cur = lane < pl.nA ? pl.spA[lane] : 0.f;  // rows past this launch are not ready
    sp_next = 32 + lane < pl.nA ? pl.spA[32 + lane] : 0.f;
  }
  const float one_hi = __uint_as_float(0x3F800054u);  // smallest float above 1.00001 (as a double)
  const float one_lo = __uint_as_float(0x3F7FFF58u);  // largest float below 0.99999

  Emit em(k);
  if (pl.nA > 0 || pl.hasB) load_window(0);  // (pos0 < 8)

  auto advance_out = [&](int n) {
    outTotal += n;
    if (outCount + n > cap) {
      status |= 1;  // SPEEDY_STATUS_OUTPUT_OVERFLOW
      outCount = cap;
    } else {
      outCount += n;
    }
  };
  // the span [pos, pos + 508) lies in the current window
  auto ensure = [&]() {
    if (pos - wbase >= G) {
      em.flush();  // the pending operations read the old window
      load_window(wk + 1);
    }
  };

#pragma unroll 1
  for (;;) {
    if (pos + kMaxReq > fed_cur) {
      // changeSpeed has not enough buffered for another period: processStreamInput returns,
      // the next call (the next 10 ms buffer, or the write's one call) feeds more
      if (evi < pl.nA) {
        fed_cur = pl.fedA0 + evi * kStep;
        if (pl.per_frame) {
          speed = __shfl_sync(kFull, sp_cur, evi & 31);
          if ((evi & 31) == 31) {
            sp_cur = sp_next;
            const int idx = evi + 33 + lane;
            sp_next = idx < pl.nA ? pl.spA[idx] : 0.f;
          }
        }
        evi++;
      } else if (pl.hasB && !doneB) {
        doneB = true;
        fed_cur = pl.fedB;
      } else {
        break;
      }
      if (!(speed >= one_hi || speed <= one_lo)) {
        // speed == 1: the whole FIFO goes through unmodified (oracle/sonic_oracle.c:373-376)
        while (pos < fed_cur) {
          const int n = fed_cur - pos < kMaxReq ? fed_cur - pos : kMaxReq;
          ensure();
          if (em.p2.n) em.flush();
          em.p2.a = pos - wbase;
          em.p2.n = n;
          em.p2.opos = outCount;
          advance_out(n);
          pos += n;
        }
      }
      continue;
    }
    ensure();
    if (remCopy > 0) {
      // copyThrough (oracle/sonic_oracle.c:319-327)
      const int n = remCopy < kMaxReq ? remCopy : kMaxReq;
      if (em.p2.n) em.flush();
      em.p2.a = pos - wbase;
      em.p2.n = n;
      em.p2.opos = outCount;
      advance_out(n);
      remCopy -= n;
      pos += n;
      continue;
    }
    const int off = pos - wbase;
    const int period = amdf16::find_pitch_period(k, off, em);
    int newSamples, adv;
    if (speed > 1.0f) {
      // skipPitchPeriod (oracle/sonic_oracle.c:279-294)
      if (speed >= 2.0f) {
        newSamples = (int)(long long)__fdiv_rn((float)period, __fsub_rn(speed, 1.0f));
      } else {
        newSamples = period;
        remCopy = (int)__fdiv_rn(__fmul_rn((float)period, __fsub_rn(2.0f, speed)), __fsub_rn(speed, 1.0f));
      }
      em.p1.a = off;
      em.p1.period = period;
      em.p1.n = newSamples;
      em.p1.opos = outCount;
      advance_out(newSamples);
      adv = period + newSamples;
    } else {
      // insertPitchPeriod (oracle/sonic_oracle.c:297-315)
      if (speed < 0.5f) {
        newSamples = (int)(long long)__fdiv_rn(__fmul_rn((float)period, speed), __fsub_rn(1.0f, speed));
      } else {
        newSamples = period;
        remCopy = (int)__fdiv_rn(__fmul_rn((float)period, __fsub_rn(__fmul_rn(2.0f, speed), 1.0f)),
                                 __fsub_rn(1.0f, speed));
      }
      emit_insert(k.w32, k.magic_tab, k.out, k.cap, lane, off, period, newSamples, outCount);
      advance_out(period + newSamples);
      adv = newSamples;
    }
    if (newSamples == 0) {
      // Upstream gives up on the write here (oracle/sonic_oracle.c:351) and leaves the input
      // unconsumed; a speed that leaves no room for even one sample is outside what the
      // speed law produces.  Flag it and carry on (as k4_splice.cu does).
      status |= 16;  // SPEEDY_STATUS_SPLICE_STALLED
    }
    pos += adv;
  }
  em.flush();
  // no bulk copy may still be in flight when the CTA retires
  if (pre) mbar_wait(bar, pre_parity);

  if (lane == 0) {
    p.st.sonic_head[s] = pl.base + pos;
    p.st.sonic_fed[s] = pl.base + (fed_cur > pos ? fed_cur : pos);
    p.st.out_total[s] = outTotal;
    p.st.out_count[s] = outCount;
    p.st.prev_period[s] = k.prevPeriod;
    p.st.prev_min_diff[s] = k.prevMinDiff;
    p.st.remaining_copy[s] = remCopy;
    p.st.sonic_speed[s] = speed;
    if (status) atomicOr(&p.st.status[s], status);
  }
}

// 16 kHz mono, not a flush: the chain shape.
bool k4_chain16_supported(const K4Params& p) {
  const Geometry& g = p.g;
  return !p.flush && g.channels == 1 && g.rate == 16000 && g.step == kStep && g.min_period == 40 &&
         g.max_period == kMaxP && g.max_required == kMaxReq && g.skip == 4 && p.out_capacity < (1LL << 30);
}

template <int N, int MINB>
static cudaError_t launch_chain16(const K4Params& p, cudaStream_t stream) {
  static SmemOptIn opt;  // (one per instantiation)
  if (cudaError_t e = opt.ensure(k4_chain16<N, MINB>, Lay<N>::total)) return e;
  k4_chain16<N, MINB><<<p.n_streams, 32, Lay<N>::total, stream>>>(p);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_k4_chain16(const K4Params& p0, cudaStream_t stream) {
  K4Params p = p0;
  k4_lane_map(p, 1);
  // one wave of streams stays resident: 26 KB per stream (4096-frame windows) or 14 KB (2048)
  int n = p.n_streams <= 148 * 8 ? 4096 : 2048;
  if (const char* e = getenv("SPEEDY_K4_CHAIN_N")) n = atoi(e) == 4096 ? 4096 : 2048;
  if (p.n_streams <= 148 * 8) return n == 4096 ? launch_chain16<4096, 1>(p, stream) : launch_chain16<2048, 1>(p, stream);
  return launch_chain16<2048, 16>(p, stream);
}

}  // namespace speedy
