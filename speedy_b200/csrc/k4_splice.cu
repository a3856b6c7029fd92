// K4, pipelined shape — Sonic time-scale modification with one CTA of three warps per
// stream, each warp with one role:
//
//   filler  (warp 1)  streams the stream's input into shared memory ahead of the cursor:
//                     one lane issues 1-D TMA bulk copies (cp.async.bulk, completion on an
//                     mbarrier) of 256-frame chunks into a ring of raw int16 frames, the warp
//                     widens each landed chunk into the 32-bit mono window the AMDF reads;
//   chain   (warp 0)  owns the splice cursor: decimation, coarse and fine AMDF search,
//                     previous-period rule, the float expressions that size a splice.  It
//                     never touches global memory for samples and never stores output: every
//                     splice becomes a 16-byte record in a shared-memory queue;
//   output  (warp 2)  consumes the records: overlap-add / copy-through from the raw ring into
//                     an output ring, drained to the stream's output buffer 16 bytes at a time.
//
// The cursor is strictly sequential (the next position depends on the period just found),
// so the time of a launch is the length of the chain warp's loop; the other two roles only
// have to keep up.  Hand-offs are mbarriers (full / empty per ring slot and per record),
// never a block barrier.
//
// Replaces what the reference does through upstream Sonic (soniclib.c:354, 369-370, 398,
// 547, 551 -> sonicIntSetSpeed, sonicIntWriteShortToStream; algorithm restated in
// oracle/sonic_oracle.c:169-355 and SURVEY.md Appendix A).  Results are bit-identical to
// k4_sonic.cu (the one-warp shape, still used for flush, short launches and multi-channel
// streams): integer sums are associative, and the float expressions are the same _rn
// intrinsics (the file is built with --fmad=false).
//
// Events.  The reference calls processStreamInput once per 10 ms buffer with the speed of
// that buffer (soniclib.c:354, 369-371), or once per write in the linear mode (:397-399).
// A pitch iteration at cursor `pos` runs inside the first such call whose fed total reaches
// pos + maxRequired (oracle/sonic_oracle.c:335, 352), at that call's speed; the loop below
// walks the calls ("events") in order and asks only that question, so the per-call set-up
// of the one-warp kernel is gone.
#include <stdlib.h>

#include "amdf16.cuh"
#include "k4_plan.cuh"
#include "kernels.cuh"

namespace speedy {

void k4_lane_map(K4Params& p, int NW);  // k4_sonic.cu

#ifdef K4_TIMING
__device__ unsigned long long g_k4s_cycles[16];
#define TS_BEGIN() const long long _t0 = clock64()
#define TS_END(slot) do { if (k.lane == 0 && k.timing) atomicAdd(&g_k4s_cycles[slot], (unsigned long long)(clock64() - _t0)); } while (0)
#else
#define TS_BEGIN() do {} while (0)
#define TS_END(slot) do {} while (0)
#endif

namespace {

using amdf16::kFull;
constexpr int kCF = 256;     // frames per ring chunk
constexpr int kCFShift = 8;
constexpr int kQ = 16;       // splice records in flight
constexpr int kPadW = 16;    // frames the AMDF may read past a search span
constexpr int kThreads = 96;

enum { REC_SKIP = 0, REC_INSERT = 1, REC_COPY = 2, REC_EXIT = 3 };

// Shared-memory carve-up (byte offsets, multiples of 16): computed by the launcher, and at
// compile time for the 16 kHz instantiation (every address in its splice loop is then the
// shared window's base plus a constant).
struct SpliceLayout {
  int bars, ctrl, recs, rcp, magic, ds, part, win, raw, oring, total;
  int nsr;       // raw ring slots (chunks)
  int nsw;       // window ring slots
  int rw;        // window ring frames = nsw * kCF
  int mir;       // frames at the start of the window ring mirrored behind its end
  int rr;        // raw ring frames = nsr * kCF
  int or_elems;  // output ring, int16 elements (power of two)
  int depth;     // chunks of bulk copies in flight ahead of the chunk being widened
  int out_vec;   // 1: the stream's output rows take 16-byte stores
  int poll_ns;   // back-off between the filler's / output role's polls of a barrier
};

constexpr int kDepthDefault = 6, kWindowAhead = 3;

__host__ __device__ constexpr int lay_take(int& off, int bytes) {
  const int at = off;
  off += (bytes + 15) & ~15;
  return at;
}

__host__ __device__ constexpr SpliceLayout make_layout(int max_period, int max_required, int skip, int channels,
                                                       int depth, int nsw_override) {
  SpliceLayout L{};
  const int span = max_required + kPadW;  // what one search needs in the window
  const int span_chunks = (span + kCF - 1) / kCF;
  L.depth = depth;
  L.nsr = span_chunks + 1 + depth + 1 + 6;  // (+ what the records still queued for the output role hold)
  L.nsw = nsw_override > 0 ? nsw_override : span_chunks + 1 + kWindowAhead;
  L.rw = L.nsw * kCF;
  L.rr = L.nsr * kCF;
  L.mir = (span + 15) & ~15;
  int or_elems = 1024;
  while (or_elems < max_required * channels + 64) or_elems <<= 1;
  L.or_elems = or_elems;
  L.out_vec = 0;
  L.poll_ns = 500;
  int off = 0;
  L.bars = lay_take(off, (2 * L.nsr + 2 * L.nsw + kQ) * 8);
  L.ctrl = lay_take(off, 16);
  L.recs = lay_take(off, kQ * 16);
  L.rcp = lay_take(off, ((max_period + 8) & ~3) * 4);
  L.magic = lay_take(off, ((max_period + 4) & ~3) * 4);
  L.ds = lay_take(off, (max_required / skip + 32 + 16) * 4);
  L.part = lay_take(off, amdf16::kPartWords * 4);  // per-lag partial sums: 4 rows (coarse) + 3 rows (fine), amdf16.cuh
  L.win = lay_take(off, (L.rw + L.mir) * 4);
  L.raw = lay_take(off, L.rr * channels * 2);
  L.oring = lay_take(off, L.or_elems * 2);
  L.total = off;
  return L;
}

constexpr SpliceLayout kLay16 = make_layout(246, 492, 4, 1, kDepthDefault, 0);  // 16 kHz mono

extern __shared__ __align__(128) unsigned char splice_smem[];

// ---- mbarrier / bulk-copy primitives (PTX) ---------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
// the same with a suspend-time hint (nanoseconds): a role with nothing to do sleeps in the
// barrier unit instead of spinning through the issue slots the chain warps need
__device__ __forceinline__ bool mbar_try_wait_sleep(uint64_t* bar, unsigned parity, unsigned ns) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
      : "memory");
  return ok != 0;
}
// non-blocking probe
__device__ __forceinline__ bool mbar_test(uint64_t* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
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
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, unsigned parity) {
#ifdef K4_NOSLEEP
  while (!mbar_try_wait(bar, parity)) {
  }
#else
  while (!mbar_try_wait_sleep(bar, parity, 2000u)) __nanosleep(500);
#endif
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

__device__ __forceinline__ int ld_volatile_s32(const int* p) {
  int v;
  asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_s32(int* p, int v) {
  asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}

// (what a launch has to do: Plan / make_plan, k4_plan.cuh)
__device__ __forceinline__ Plan make_splice_plan(const K4Params& p, int s) {
  Plan pl = make_plan(p, s);
  const long long span = (long long)pl.last_fed + kPadW;
  pl.nchunks = (pl.nA > 0 || pl.hasB) && span > 0 ? (int)((span + kCF - 1) >> kCFShift) : 0;
  return pl;
}

// ---------------------------------------------------------------------------
// chain role
// ---------------------------------------------------------------------------
template <bool K16>
struct Chain {
  // shared memory: byte offsets of the generic instantiation (constants for 16 kHz)
  int o_win, o_ds, o_rcp, o_recs, o_ctrl, o_wempty, o_qfull;
  int rw_, nsw_;
  __device__ __forceinline__ int* win() const { return reinterpret_cast<int*>(splice_smem + (K16 ? kLay16.win : o_win)); }
  __device__ __forceinline__ int* ds() const { return reinterpret_cast<int*>(splice_smem + (K16 ? kLay16.ds : o_ds)); }
  __device__ __forceinline__ float* rcp() const { return reinterpret_cast<float*>(splice_smem + (K16 ? kLay16.rcp : o_rcp)); }
  __device__ __forceinline__ int4* recs() const { return reinterpret_cast<int4*>(splice_smem + (K16 ? kLay16.recs : o_recs)); }
  // [0] records the output role has finished, [1] chain done, [2] window chunks widened
  __device__ __forceinline__ int* ctrl() const { return reinterpret_cast<int*>(splice_smem + (K16 ? kLay16.ctrl : o_ctrl)); }
  __device__ __forceinline__ uint64_t* wempty() const {
    return reinterpret_cast<uint64_t*>(splice_smem + (K16 ? kLay16.bars + (2 * kLay16.nsr + kLay16.nsw) * 8 : o_wempty));
  }
  __device__ __forceinline__ uint64_t* qfull() const {
    return reinterpret_cast<uint64_t*>(splice_smem + (K16 ? kLay16.bars + (2 * kLay16.nsr + 2 * kLay16.nsw) * 8 : o_qfull));
  }
  __device__ __forceinline__ unsigned* part() const { return reinterpret_cast<unsigned*>(splice_smem + kLay16.part); }
  __device__ __forceinline__ int rw() const { return K16 ? kLay16.rw : rw_; }
  __device__ __forceinline__ int nsw() const { return K16 ? kLay16.nsw : nsw_; }
  // geometry
  int S, minP, maxP, maxReq, skip;
  int c_lo, c_hi, ds_count;  // coarse lag range, decimated values per search
  unsigned dec_magic;
  // lane maps of the two searches (kernels.cuh: K4Params::lane_map)
  int lane, cGi, cSub, cG, cMaxG, fG, fg, fGi0;
  // window ring progress
  int cready, rslot, rpar;  // chunks known to be filled; slot / parity of the next one
  int released, eslot;      // chunks handed back to the filler
  // record queue
  int nposted, qslot, known_done;
  // Sonic state
  int prevPeriod, prevMinDiff, remCopy;
  bool timing;
  // geometry accessors: compile-time constants for 16 kHz (the benchmark configurations), so
  // that the splice loop carries no divisions, parameter reloads or generic paths
  __device__ __forceinline__ int gS() const { return K16 ? 160 : S; }
  __device__ __forceinline__ int gMinP() const { return K16 ? 40 : minP; }
  __device__ __forceinline__ int gMaxP() const { return K16 ? 246 : maxP; }
  __device__ __forceinline__ int gMaxReq() const { return K16 ? 492 : maxReq; }
  __device__ __forceinline__ int gSkip() const { return K16 ? 4 : skip; }
  __device__ __forceinline__ int gCLo() const { return K16 ? 10 : c_lo; }
  __device__ __forceinline__ int gCHi() const { return K16 ? 61 : c_hi; }
};

using amdf16::sad16;
using amdf16::blocks_run;
using amdf16::block_head;
using amdf16::block_tail;

// Exact resolution of a short list of candidates, in the order of the C scan
// (oracle/sonic_oracle.c:189-205): lanes ascend with the lag, so does l.  Rare (several
// lags within float rounding of the extremum, or silence), kept out of line.
__device__ __noinline__ void resolve_exact(unsigned d0, unsigned d1, unsigned d2, unsigned d3, int pg, unsigned cand,
                                           int want_min, unsigned* rd, int* rp) {
  unsigned bal = __ballot_sync(kFull, cand != 0u);
  unsigned bd = 0u;
  int bp = want_min ? 0 : 255;
  while (bal) {
    const int src = __ffs(bal) - 1;
    bal &= bal - 1;
    const unsigned c = __shfl_sync(kFull, cand, src);
    const int q = __shfl_sync(kFull, pg, src);
    const unsigned x[4] = {__shfl_sync(kFull, d0, src), __shfl_sync(kFull, d1, src), __shfl_sync(kFull, d2, src),
                           __shfl_sync(kFull, d3, src)};
#pragma unroll
    for (int l = 0; l < 4; l++) {
      if (!((c >> l) & 1u)) continue;
      const unsigned cd = x[l];
      const int cp = q + l;
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

using amdf16::udiv_small;

// AMDF over lags lo..hi on a[i] = arr[off + i] (oracle/sonic_oracle.c:185-209): lane
// `sub` of the `G` adjacent lanes of a lag group (lags pg .. pg+3, pg a multiple of four)
// takes every G-th fully valid block; lane 0 of the group also the ragged start, lane 1 (or
// 0 when alone) the ragged end.  Returns the best lag; with WANT_DIFFS also the per-sample
// difference at the best lag and what the previous-period rule needs of the worst one (the
// coarse pass needs neither: the refinement's replace them).
//
// Arg-min / arg-max of diff / lag.  The C scan compares by cross-multiplication with strict
// inequalities, so ties go to the smaller lag.  Every group leader forms float keys
// diff * (1 / lag) for its four lags (relative error < 2e-7); the lags within 2e-6 of the
// warp-wide extremum are a superset of the true extremum; nearly always that is one lag,
// otherwise the short list is resolved exactly.
//
// MAXG > 0: the largest group size is a compile-time constant (and, with UNIFORM_G, every
// group has exactly that many lanes).
template <bool K16, int MAXG, bool UNIFORM_G, bool WANT_DIFFS>
__device__ __forceinline__ int search(const Chain<K16>& k, const int* arr, int off, int lo, int hi, int pg, bool live,
                                      int sub, int G_rt, int maxG_rt, int* minDiff, int* maxDiff) {
  const int G = UNIFORM_G ? MAXG : G_rt;
  const int maxG = MAXG > 0 ? MAXG : maxG_rt;
  unsigned d[4] = {0u, 0u, 0u, 0u};
  if (live) {
    const int B0 = off & ~3, hd = off & 3;
    const int* base = arr + B0;
    const int jf1 = (hd + pg) >> 2;  // blocks jf0 .. jf1-1 are fully valid
    blocks_run(base, pg, (hd ? 1 : 0) + sub, G, jf1, d);
    if (sub == 0 && hd) block_head(base, pg, hd, d);
    if (sub == (G > 1 ? 1 : 0)) block_tail(base + 4 * jf1, pg, hd, d);
  }
  // the group's first lane collects its neighbours' partial sums (independent shuffles)
  {
    const unsigned o0 = d[0], o1 = d[1], o2 = d[2], o3 = d[3];
#pragma unroll
    for (int n = 1; n < (MAXG > 0 ? MAXG : 1); n++) {
      const unsigned t0 = __shfl_down_sync(kFull, o0, n), t1 = __shfl_down_sync(kFull, o1, n);
      const unsigned t2 = __shfl_down_sync(kFull, o2, n), t3 = __shfl_down_sync(kFull, o3, n);
      if (UNIFORM_G || n < G) {
        d[0] += t0; d[1] += t1; d[2] += t2; d[3] += t3;
      }
    }
    if (MAXG == 0) {
#pragma unroll 1
      for (int n = 1; n < maxG; n++) {
        const unsigned t0 = __shfl_down_sync(kFull, o0, n), t1 = __shfl_down_sync(kFull, o1, n);
        const unsigned t2 = __shfl_down_sync(kFull, o2, n), t3 = __shfl_down_sync(kFull, o3, n);
        if (n < G) {
          d[0] += t0; d[1] += t1; d[2] += t2; d[3] += t3;
        }
      }
    }
  }
  const bool leader = live && sub == 0;
  const float4 r4 = leader ? *reinterpret_cast<const float4*>(k.rcp() + pg) : make_float4(0.f, 0.f, 0.f, 0.f);
  // float keys of this lane's lags inside [lo, hi]
  const bool v0 = leader && pg >= lo && pg <= hi;
  const bool v1 = leader && pg + 1 >= lo && pg + 1 <= hi;
  const bool v2 = leader && pg + 2 >= lo && pg + 2 <= hi;
  const bool v3 = leader && pg + 3 >= lo && pg + 3 <= hi;
  const float k0 = __uint2float_rn(d[0]) * r4.x, k1 = __uint2float_rn(d[1]) * r4.y;
  const float k2 = __uint2float_rn(d[2]) * r4.z, k3 = __uint2float_rn(d[3]) * r4.w;
  const float big = 3.0e38f;
  const float kmin = fminf(fminf(v0 ? k0 : big, v1 ? k1 : big), fminf(v2 ? k2 : big, v3 ? k3 : big));
  const float emin = __uint_as_float(__reduce_min_sync(kFull, __float_as_uint(kmin)));
  const float tmin = emin * 1.000002f;
  const unsigned cb = (v0 && k0 <= tmin ? 1u : 0u) | (v1 && k1 <= tmin ? 2u : 0u) | (v2 && k2 <= tmin ? 4u : 0u) |
                      (v3 && k3 <= tmin ? 8u : 0u);
  const unsigned bal_b = __ballot_sync(kFull, cb != 0u);
  unsigned cw = 0u, bal_w = 0u;
  float emax = 1.0f;
  if (WANT_DIFFS) {
    const float kmax = fmaxf(fmaxf(v0 ? k0 : 0.f, v1 ? k1 : 0.f), fmaxf(v2 ? k2 : 0.f, v3 ? k3 : 0.f));
    emax = __uint_as_float(__reduce_max_sync(kFull, __float_as_uint(kmax)));
    const float tmax = emax * 0.999998f;
    cw = (v0 && k0 >= tmax ? 1u : 0u) | (v1 && k1 >= tmax ? 2u : 0u) | (v2 && k2 >= tmax ? 4u : 0u) |
         (v3 && k3 >= tmax ? 8u : 0u);
    bal_w = __ballot_sync(kFull, cw != 0u);
  }
  const unsigned multi = __ballot_sync(kFull, (cb & (cb - 1u)) != 0u || (cw & (cw - 1u)) != 0u);
  unsigned best_diff, worst_diff = 0u;
  int best, worst = 255;
  if (multi == 0u && (bal_b & (bal_b - 1u)) == 0u && (bal_w & (bal_w - 1u)) == 0u && emax > 0.f) {
    const int lb = __ffs(cb) - 1;  // this lane's candidate (if any)
    const unsigned sd_b = lb == 0 ? d[0] : lb == 1 ? d[1] : lb == 2 ? d[2] : d[3];
    const int sb = __ffs(bal_b) - 1;
    best_diff = __shfl_sync(kFull, sd_b, sb);
    best = __shfl_sync(kFull, pg + lb, sb);
    if (WANT_DIFFS) {
      const int lw = __ffs(cw) - 1;
      const unsigned sd_w = lw == 0 ? d[0] : lw == 1 ? d[1] : lw == 2 ? d[2] : d[3];
      const int sw = __ffs(bal_w) - 1;
      worst_diff = __shfl_sync(kFull, sd_w, sw);
      worst = __shfl_sync(kFull, pg + lw, sw);
    }
  } else {
    resolve_exact(d[0], d[1], d[2], d[3], pg, cb, 1, &best_diff, &best);
    if (WANT_DIFFS) resolve_exact(d[0], d[1], d[2], d[3], pg, cw, 0, &worst_diff, &worst);
  }
  if (WANT_DIFFS) {
    const int md = udiv_small(best_diff, best, k.rcp()[best]);
    *minDiff = md;
    // maxDiff = floor(worst_diff / worst) only ever meets "maxDiff > 3 * minDiff"
    // (oracle/sonic_oracle.c:217), i.e. worst_diff >= (3 * minDiff + 1) * worst: no division.
    // (The C scan starts from maxDiff = 0, worstPeriod = 255: all-zero differences give 0.)
    *maxDiff = worst_diff >= (unsigned)(3 * md + 1) * (unsigned)worst ? 3 * md + 1 : 0;
  }
  return best;
}

// Upstream downSampleInput for a mono stream (oracle/sonic_oracle.c:169-178): sum `skip`
// frames, C integer division (truncating).  16 kHz: four frames per value, every lane makes
// four consecutive values from five aligned 16-byte loads, (v + (v < 0 ? 3 : 0)) >> 2.
// Otherwise |sum| < 2^21 and the divisor is small, so the quotient is exact as
// (|sum| * ceil(2^32 / divisor)) >> 32.
template <bool K16>
__device__ __forceinline__ void decimate(const Chain<K16>& k, int off) {
  TS_BEGIN();
  __syncwarp();  // every lane is done reading the previous decimated copy
  if ((k.skip & 3) == 0) {
    // aligned 16-byte walk: the first and last vector of a value are partial
    const int count = k.ds_count;
    const int r = off & 3;
    const int nmid = (k.skip >> 2) - 1;
    const int* base = k.win() + (off & ~3);
#pragma unroll 1
    for (int i = k.lane; i < count; i += 32) {
      const int4* p = reinterpret_cast<const int4*>(base + i * k.skip);
      const int4 x = p[0];
      const int4 z = p[nmid + 1];
      int v = x.w + (r == 0 ? x.x : z.x) + (r <= 1 ? x.y : z.y) + (r <= 2 ? x.z : z.z);
#pragma unroll 1
      for (int m = 1; m <= nmid; m++) {
        const int4 t = p[m];
        v += (t.x + t.y) + (t.z + t.w);
      }
      const int qa = (int)__umulhi((unsigned)abs(v), k.dec_magic);
      k.ds()[i] = v < 0 ? -qa : qa;
    }
  } else {
    const int count = k.ds_count;
#pragma unroll 1
    for (int i = k.lane; i < count; i += 32) {
      int v = 0;
      const int* q = k.win() + off + i * k.skip;
#pragma unroll 1
      for (int j = 0; j < k.skip; j++) v += q[j];
      const int qa = (int)__umulhi((unsigned)abs(v), k.dec_magic);
      k.ds()[i] = v < 0 ? -qa : qa;
    }
  }
  __syncwarp();
  TS_END(1);
}

// findPitchPeriod (oracle/sonic_oracle.c:225-257) at window offset `off`: the coarse pass
// on the decimated copy (a static lane assignment with more lanes for the longer lags: group q
// has q + 1 blocks), then the refinement at the full rate.
template <bool K16>
__device__ __forceinline__ int find_pitch_period(Chain<K16>& k, int off) {
  if (K16) return amdf16::find_pitch_period(k, off);
  int minDiff = 0, maxDiff = 0, period = 0;
  int lo = k.gMinP(), hi = k.gMaxP();
  if (k.gSkip() != 1) {
    decimate<K16>(k, off);
    TS_BEGIN();
    const int cpg = 4 * ((k.gCLo() >> 2) + k.cGi);
    period = search<K16, 0, false, false>(k, k.ds(), 0, k.c_lo, k.c_hi, cpg, k.cGi >= 0, k.cSub, k.cG, k.cMaxG, nullptr, nullptr);
#ifdef K4_TIMING
    if (k.lane == 0 && k.timing) atomicAdd(&g_k4s_cycles[2], (unsigned long long)(clock64() - _t0 + (period & 0)));
#endif
    // refine around the coarse estimate at the full rate
    period *= k.gSkip();
    lo = period - (k.gSkip() << 2);
    hi = period + (k.gSkip() << 2);
    if (lo < k.gMinP()) lo = k.gMinP();
    if (hi > k.gMaxP()) hi = k.gMaxP();
  }
  {
    TS_BEGIN();
    const int g0 = lo >> 2;
    const int pg = 4 * (g0 + k.fGi0);
    const bool live = k.fGi0 < (hi >> 2) - g0 + 1;
    period = search<K16, 0, false, true>(k, k.win(), off, lo, hi, pg, live, k.fg, k.fG, k.fG, &minDiff, &maxDiff);
#ifdef K4_TIMING
    if (k.lane == 0 && k.timing) atomicAdd(&g_k4s_cycles[4], (unsigned long long)(clock64() - _t0 + (period & 0)));
#endif
  }
  // previousPeriodBetter(preferNew = 1), oracle/sonic_oracle.c:213-223
  const bool keep_prev =
      minDiff != 0 && k.prevPeriod != 0 && !(maxDiff > minDiff * 3) && !(minDiff * 2 <= k.prevMinDiff * 3);
  const int result = keep_prev ? k.prevPeriod : period;
  k.prevMinDiff = minDiff;
  k.prevPeriod = period;
  return result;
}

// Window chunks up to relative frame `need_end` are filled.  The filler publishes the number
// of chunks it has widened in a shared-memory word (its stores fenced before it): a plain
// load on the chain's critical path, not a barrier-unit round trip.
template <bool K16>
__device__ __forceinline__ void chain_ensure(Chain<K16>& k, int need_end) {
  const int cneed = (need_end - 1) >> kCFShift;
  if (k.cready > cneed) return;
  TS_BEGIN();
#ifdef K4_TIMING
  if (k.lane == 0 && k.timing) {
    atomicAdd(&g_k4s_cycles[11], 1ULL);
    const int have = ld_volatile_s32(k.ctrl() + 2);
    if (have <= cneed) {
      atomicAdd(&g_k4s_cycles[12], 1ULL);
      atomicAdd(&g_k4s_cycles[3], (unsigned long long)(cneed + 1 - have));           // chunks short
      atomicAdd(&g_k4s_cycles[7], (unsigned long long)(cneed - k.released));         // need - released
    }
    // how far behind is the output role
    atomicAdd(&g_k4s_cycles[14], (unsigned long long)(k.nposted - ld_volatile_s32(k.ctrl())));
  }
#endif
  do {
    k.cready = ld_volatile_s32(k.ctrl() + 2);
  } while (k.cready <= cneed);
  TS_END(0);
}

// The cursor has passed these chunks: the filler may reuse their window slots.
template <bool K16>
__device__ __forceinline__ void chain_release(Chain<K16>& k, int pos) {
  while (((k.released + 1) << kCFShift) <= pos) {
    __syncwarp();  // every lane is done reading the chunk
    if (k.lane == 0) mbar_arrive(k.wempty() + k.eslot);
    k.released++;
    if (++k.eslot == k.nsw()) k.eslot = 0;
  }
}

template <bool K16>
__device__ __forceinline__ void chain_post(Chain<K16>& k, int kind, int pos, int period, int n, int opos) {
  TS_BEGIN();
  if (k.nposted - k.known_done >= kQ) {
#ifdef K4_TIMING
    if (k.lane == 0 && k.timing) atomicAdd(&g_k4s_cycles[13], 1ULL);
#endif
    do {
      k.known_done = ld_volatile_s32(k.ctrl());
    } while (k.nposted - k.known_done >= kQ);
  }
  if (k.lane == 0) {
    k.recs()[k.qslot] = make_int4(pos, period | (kind << 16), n, opos);
    mbar_arrive(k.qfull() + k.qslot);  // release: the record is visible to whoever sees the arrival
  }
  k.nposted++;
  if (++k.qslot == kQ) k.qslot = 0;
  TS_END(6);
}

template <bool K16>
__device__ void chain_role(const K4Params& p, const SpliceLayout& L, int s, const Plan& pl) {
  const Geometry& g = p.g;
  Chain<K16> k;
  k.lane = threadIdx.x & 31;
  k.o_win = L.win;
  k.o_ds = L.ds;
  k.o_rcp = L.rcp;
  k.o_recs = L.recs;
  k.o_ctrl = L.ctrl;
  k.o_wempty = L.bars + (2 * L.nsr + L.nsw) * 8;
  k.o_qfull = L.bars + (2 * L.nsr + 2 * L.nsw) * 8;
  k.rw_ = L.rw;
  k.nsw_ = L.nsw;
  float* rcp = k.rcp();
  k.S = g.step;
  k.minP = g.min_period;
  k.maxP = g.max_period;
  k.maxReq = g.max_required;
  k.skip = g.skip;
  k.dec_magic = (unsigned)((0x100000000ULL + (unsigned)k.skip - 1) / (unsigned)k.skip);
  k.c_lo = k.minP / k.skip;
  k.c_hi = k.maxP / k.skip;
  k.ds_count = k.maxReq / k.skip;
  // reciprocals of the lags (the float keys of the arg-min), the decimated copy's pad
  for (int n = k.lane; n < ((k.maxP + 8) & ~3); n += 32) rcp[n] = n ? __frcp_rn((float)n) : 0.f;
  for (int i = k.lane; i < 32; i += 32) k.ds()[k.maxReq / k.skip + i] = 0;
  if (K16) {
    for (int i = k.lane; i < amdf16::kPartWords; i += 32) k.part()[i] = 0u;
  }
  {
    unsigned m = p.lane_map[k.lane];
    // (opaque: kept in a register, not re-read from the parameter bank with a per-lane index
    // on every pitch iteration)
    m = __shfl_sync(kFull, m, k.lane);
    k.cGi = (int)(m & 0xffu) - 1;
    k.cSub = (int)((m >> 8) & 0xffu);
    k.cG = (int)((m >> 16) & 0xffu);
    k.cMaxG = p.c_max_g;
    k.fG = p.f_g;
    __builtin_assume(k.cG >= 1 && k.cG <= 32 && k.cSub >= 0 && k.cSub < 32 && k.cMaxG >= 1 && k.cMaxG <= 32);
    __builtin_assume(k.fG >= 1 && k.fG <= 32);
    const int gpw = 32 / k.fG;  // groups per warp
    const int slot = k.lane / k.fG;
    k.fg = k.lane - slot * k.fG;
    k.fGi0 = slot < gpw ? slot : (1 << 30);  // idle lanes never match
    k.fg = __shfl_sync(kFull, k.fg, k.lane);
    k.fGi0 = __shfl_sync(kFull, k.fGi0, k.lane);
  }
  __syncwarp();
  k.cready = 0;
  k.rslot = 0;
  k.rpar = 0;
  k.released = 0;
  k.eslot = 0;
  k.nposted = 0;
  k.qslot = 0;
  k.known_done = 0;

#ifdef K4_TIMING
  k.timing = s == 0 && !p.flush;
  const long long t_kernel = clock64();
#else
  k.timing = false;
#endif
  k.prevPeriod = p.st.prev_period[s];
  k.prevMinDiff = p.st.prev_min_diff[s];
  k.remCopy = p.st.remaining_copy[s];
  long long outTotal = p.st.out_total[s];
  int outCount = pl.out_count0;
  const int cap = (int)p.out_capacity;
  int status = 0;
  float speed = p.st.sonic_speed[s];

  int pos = pl.pos0;   // the cursor, relative
  int woff = pos;      // its offset in the window ring (pos0 < 8 <= rw)
  int fed_cur = pl.fed0;
  int evi = 0;
  bool doneB = false, in_final = false;
  long long expected = 0;
  // 32 speeds at a time, one per lane, the next batch already in flight
  float sp_cur = 0.f, sp_next = 0.f;
  if (pl.per_frame) {
    sp_cur = k.lane < pl.nA ? pl.spA[k.lane] : 0.f;  // rows past this launch are not ready
    sp_next = 32 + k.lane < pl.nA ? pl.spA[32 + k.lane] : 0.f;
  }
  const float one_hi = __uint_as_float(0x3F800054u);  // smallest float above 1.00001 (as a double)
  const float one_lo = __uint_as_float(0x3F7FFF58u);  // largest float below 0.99999

#pragma unroll 1
  for (;;) {
    if (pos + k.gMaxReq() > fed_cur) {
      // changeSpeed has not enough buffered for another period: processStreamInput returns,
      // the next call (the next 10 ms buffer, or the write's / flush's one call) feeds more
      if (evi < pl.nA) {
        fed_cur = pl.fedA0 + evi * k.gS();
        if (pl.per_frame) {
          speed = __shfl_sync(kFull, sp_cur, evi & 31);
          if ((evi & 31) == 31) {
            sp_cur = sp_next;
            const int idx = evi + 33 + k.lane;
            sp_next = idx < pl.nA ? pl.spA[idx] : 0.f;
          }
        }
        evi++;
      } else if (pl.hasB && !doneB) {
        doneB = true;
        if (pl.finalB) {
          // upstream sonicFlushStream (oracle/sonic_oracle.c:443-460): the length the real
          // samples should still produce, then 2 * maxRequired frames of silence
          const int remaining = pl.fed_real - pos;
          expected = outTotal + (int)__fadd_rn(__fdiv_rn(__fdiv_rn((float)remaining, speed), 1.0f), 0.5f);
          in_final = true;
        }
        fed_cur = pl.fedB;
      } else {
        break;
      }
      if (!(speed >= one_hi || speed <= one_lo)) {
        // speed == 1: the whole FIFO goes through unmodified (oracle/sonic_oracle.c:373-376)
        while (pos < fed_cur) {
          const int n = fed_cur - pos < k.gMaxReq() ? fed_cur - pos : k.gMaxReq();
          chain_ensure(k, pos + n);
          chain_post(k, REC_COPY, pos, 0, n, outCount);
          outTotal += n;
          if (outCount + n > cap) { status |= 1; outCount = cap; } else { outCount += n; }
          pos += n;
          woff += n;
          if (woff >= k.rw()) woff -= k.rw();
          chain_release(k, pos);
        }
      }
      continue;
    }
    chain_ensure(k, pos + k.gMaxReq() + kPadW);
    int adv;
    if (k.remCopy > 0) {
      // copyThrough (oracle/sonic_oracle.c:319-327)
      const int n = k.remCopy < k.gMaxReq() ? k.remCopy : k.gMaxReq();
      chain_post(k, REC_COPY, pos, 0, n, outCount);
      outTotal += n;
      if (outCount + n > cap) { status |= 1; outCount = cap; } else { outCount += n; }
      k.remCopy -= n;
      adv = n;
    } else {
#ifdef K4_TIMING
      if (k.lane == 0 && k.timing) atomicAdd(&g_k4s_cycles[9], 1ULL);
#endif
      const int period = find_pitch_period<K16>(k, woff);
      int newSamples, produced, kind;
      if (speed > 1.0f) {
        // skipPitchPeriod (oracle/sonic_oracle.c:279-294)
        if (speed >= 2.0f) {
          newSamples = (int)(long long)__fdiv_rn((float)period, __fsub_rn(speed, 1.0f));
        } else {
          newSamples = period;
          k.remCopy = (int)__fdiv_rn(__fmul_rn((float)period, __fsub_rn(2.0f, speed)), __fsub_rn(speed, 1.0f));
        }
        kind = REC_SKIP;
        produced = newSamples;
        adv = period + newSamples;
      } else {
        // insertPitchPeriod (oracle/sonic_oracle.c:297-315)
        if (speed < 0.5f) {
          newSamples = (int)(long long)__fdiv_rn(__fmul_rn((float)period, speed), __fsub_rn(1.0f, speed));
        } else {
          newSamples = period;
          k.remCopy = (int)__fdiv_rn(__fmul_rn((float)period, __fsub_rn(__fmul_rn(2.0f, speed), 1.0f)),
                                     __fsub_rn(1.0f, speed));
        }
        kind = REC_INSERT;
        produced = period + newSamples;
        adv = newSamples;
      }
      chain_post(k, kind, pos, period, newSamples, outCount);
      outTotal += produced;
      if (outCount + produced > cap) { status |= 1; outCount = cap; } else { outCount += produced; }
      if (newSamples == 0) {
        // Upstream gives up on the write here (oracle/sonic_oracle.c:351) and leaves the input
        // unconsumed; a speed that leaves no room for even one sample is outside what the
        // speed law produces.  This shape cannot rewind its rings: flag it and carry on.
        status |= 16;  // SPEEDY_STATUS_SPLICE_STALLED
      }
    }
    pos += adv;
    woff += adv;
    if (woff >= k.rw()) woff -= k.rw();
    chain_release(k, pos);
  }

  long long head = pl.base + pos;
  long long fed = pl.base + (fed_cur > pos ? fed_cur : pos);
  int remCopy = k.remCopy;
  if (in_final) {
    if (outTotal > expected) {
      const long long excess = outTotal - expected;
      outTotal = expected;
      outCount = outCount > excess ? (int)(outCount - excess) : 0;
    }
    // the padding is not input: the stream carries on from the real end of the data
    fed = pl.base + pl.fed_real;
    head = fed;
    remCopy = 0;
    status |= 2;  // SPEEDY_STATUS_FLUSHED
  }
#ifdef K4_TIMING
  if (k.lane == 0 && k.timing) {
    atomicAdd(&g_k4s_cycles[8], (unsigned long long)(clock64() - t_kernel));
    atomicAdd(&g_k4s_cycles[10], (unsigned long long)evi);
  }
#endif
  chain_post(k, REC_EXIT, pos, 0, 0, outCount);
  if (k.lane == 0) {
    st_volatile_s32(k.ctrl() + 1, 1);  // the filler stops waiting for window slots
    p.st.sonic_head[s] = head;
    p.st.sonic_fed[s] = fed;
    p.st.out_total[s] = outTotal;
    p.st.out_count[s] = outCount;
    p.st.prev_period[s] = k.prevPeriod;
    p.st.prev_min_diff[s] = k.prevMinDiff;
    p.st.remaining_copy[s] = remCopy;
    p.st.sonic_speed[s] = speed;
    if (status) atomicOr(&p.st.status[s], status);
  }
}

// ---------------------------------------------------------------------------
// filler role
// ---------------------------------------------------------------------------
// wait on `bar` unless the chain has finished (it then never frees another slot)
__device__ __forceinline__ bool wait_or_done(uint64_t* bar, unsigned parity, const int* done_flag, unsigned poll_ns) {
#ifdef K4_NOSLEEP
  while (!mbar_try_wait(bar, parity)) {
    if (ld_volatile_s32(done_flag)) return false;
  }
#else
  while (!mbar_try_wait_sleep(bar, parity, poll_ns)) {
    if (ld_volatile_s32(done_flag)) return false;
    __nanosleep(poll_ns);
  }
#endif
  return true;
}


__device__ void filler_role(const K4Params& p, const SpliceLayout& L, const Plan& pl) {
  const int lane = threadIdx.x & 31;
  const int C = p.g.channels;
  uint64_t* bars = reinterpret_cast<uint64_t*>(splice_smem + L.bars);
  uint64_t* rfull = bars;
  uint64_t* rempty = bars + L.nsr;
  uint64_t* wfull = bars + 2 * L.nsr;
  uint64_t* wempty = wfull + L.nsw;
  int* ctrl = reinterpret_cast<int*>(splice_smem + L.ctrl);
  const int* done_flag = ctrl + 1;
  int* win = reinterpret_cast<int*>(splice_smem + L.win);
  short* raw = reinterpret_cast<short*>(splice_smem + L.raw);
  const unsigned chunk_bytes = (unsigned)(kCF * C * sizeof(short));
  const long long lim = pl.base + pl.data_end;  // absolute frames >= lim read as silence

  int issued = 0, widened = 0;
  bool live = true;
#pragma unroll 1
  for (int c = 0; c < pl.nchunks && live; c++) {
    // keep `depth` chunks of copies in flight ahead of this one
    while (issued < pl.nchunks && issued <= c + L.depth) {
      const int islot = issued % L.nsr, iuse = issued / L.nsr;
      if (iuse >= 1) {
        // the slot still holds a chunk the output role may need.  The chunk this trip widens
        // must be fetched; anything further ahead is fetched only if its slot is free already
        // (never sit on a landed chunk waiting for room to prefetch)
        if (issued > c) {
          if (!mbar_test(rempty + islot, (iuse - 1) & 1)) break;
        } else if (
#ifdef K4_TIMING
            (blockIdx.x == 0 && lane == 0 && !mbar_test(rempty + islot, (iuse - 1) & 1) ? (void)atomicAdd(&g_k4s_cycles[15], 1ULL) : (void)0),
#endif
            !wait_or_done(rempty + islot, (iuse - 1) & 1, done_flag, L.poll_ns)) {
          live = false;
          break;
        }
      }
      const long long a0 = pl.base + ((long long)issued << kCFShift);  // absolute first frame
      const long long a1 = a0 + kCF;
      short* dst = raw + (size_t)islot * kCF * C;
      const int16_t* srcp = nullptr;
      if (pl.src.in && a0 >= pl.src.t_old && a1 <= pl.src.t_new && a1 <= lim) {
        srcp = pl.src.in + (a0 - pl.src.t_old) * C;
      } else if (a0 >= pl.src.hist_base && a1 <= pl.src.t_old && a1 <= lim) {
        srcp = pl.src.hist + (a0 - pl.src.hist_base) * C;
      }
      if (srcp && (reinterpret_cast<size_t>(srcp) & 15) == 0) {
        if (lane == 0) {
          fence_proxy_async();  // earlier generic-proxy accesses to this slot are ordered before the copy
          mbar_expect_tx(rfull + islot, chunk_bytes);
          bulk_load(dst, srcp, chunk_bytes, rfull + islot);
        }
      } else {
        // ragged chunk (straddles the carried history and the caller's buffer, runs into the
        // end of the data, or the source is not 16-byte aligned): plain loads
        for (int i = lane; i < kCF * C; i += 32) {
          const int fr = C == 1 ? i : i / C;
          const int ch = C == 1 ? 0 : i - fr * C;
          const long long f = a0 + fr;
          int v = 0;
          if (f < lim && f >= pl.src.hist_base && (f < pl.src.t_old || (pl.src.in && f < pl.src.t_new))) {
            v = pl.src.raw(f, ch);
          }
          dst[i] = (short)v;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(rfull + islot);
      }
      issued++;
    }
    if (!live) break;
    const int wslot = c % L.nsw, wuse = c / L.nsw;
#ifdef K4_TIMING
    if (blockIdx.x == 0 && lane == 0 && wuse >= 1) {
      if (!mbar_test(wempty + wslot, (wuse - 1) & 1)) atomicAdd(&g_k4s_cycles[5], 1ULL);
    }
#endif
    if (wuse >= 1 && !wait_or_done(wempty + wslot, (wuse - 1) & 1, done_flag, L.poll_ns)) break;
    const int rslot = c % L.nsr;
    mbar_wait_sleep(rfull + rslot, (c / L.nsr) & 1);
    widened = c + 1;
    // widen: 8 frames per lane, one 16-byte load, two 16-byte stores (and the mirror)
    const short* rs = raw + (size_t)rslot * kCF * C;
    int* wd = win + wslot * kCF;
    if (C == 1) {
      const int4 q = *reinterpret_cast<const int4*>(rs + lane * 8);
      const int4 lo4 = make_int4((short)(q.x & 0xffff), q.x >> 16, (short)(q.y & 0xffff), q.y >> 16);
      const int4 hi4 = make_int4((short)(q.z & 0xffff), q.z >> 16, (short)(q.w & 0xffff), q.w >> 16);
      int4* d4 = reinterpret_cast<int4*>(wd + lane * 8);
      d4[0] = lo4;
      d4[1] = hi4;
      if (wslot * kCF + lane * 8 < L.mir) {
        int4* m4 = reinterpret_cast<int4*>(wd + L.rw + lane * 8);
        m4[0] = lo4;
        m4[1] = hi4;
      }
    } else {
      for (int f = lane; f < kCF; f += 32) {
        int sum = 0;
        for (int ch = 0; ch < C; ch++) sum += rs[f * C + ch];
        const int v = sum / C;
        wd[f] = v;
        if (wslot * kCF + f < L.mir) wd[L.rw + f] = v;
      }
    }
    __syncwarp();
    if (lane == 0) {
      __threadfence_block();  // the chunk is in the window before the count says so
      st_volatile_s32(ctrl + 2, c + 1);
    }
  }
  // no bulk copy may still be in flight when the CTA retires
  for (int c = widened; c < issued; c++) mbar_wait(rfull + c % L.nsr, (c / L.nsr) & 1);
}

// ---------------------------------------------------------------------------
// output role
// ---------------------------------------------------------------------------
// ceil(2^(32+shift) / n) for 2^shift < n <= 2^(shift+1): floor from the correctly rounded
// double quotient (the true one is at least 1/n away from the integers it does not hit)
__device__ __forceinline__ unsigned division_magic(int n, int shift) {
  const double qd = __ddiv_rn((double)(1ULL << (32 + shift)), (double)n);
  return (unsigned)(unsigned long long)qd + ((n & (n - 1)) ? 1u : 0u);
}

struct Output {
  const short* raw;
  short* oring;
  const unsigned* magic_tab;
  int16_t* out;  // the stream's output row
  int C, rrC, omask, maxP, lane;
  long long cap_e;  // capacity in elements
};

// ring index of raw element e0 + i (e0 < rrC, i < rrC)
__device__ __forceinline__ int raw_wrap(int idx, int rrC) { return idx >= rrC ? idx - rrC : idx; }

// `n` frames starting at raw-ring element index r0 go to output elements [oe, oe + n*C)
__device__ __forceinline__ void out_copy(const Output& o, int r0, int n, long long oe) {
  const int total = n * o.C;
  for (int i = o.lane; i < total; i += 32) {
    o.oring[(int)((oe + i) & o.omask)] = o.raw[raw_wrap(r0 + i, o.rrC)];
  }
}

// overlapAdd (oracle/sonic_oracle.c:263-276): out[t] = (down[t]*(n-t) + up[t]*t) / n per
// channel, C integer arithmetic.  trunc(|num| / n) == umulhi(|num|, magic) >> shift for
// |num| < 2^26, 1 < n < 2^11, 2^shift < n <= 2^(shift+1), magic = ceil(2^(32+shift) / n).
__device__ __forceinline__ void out_overlap_add(const Output& o, int rd, int ru, int n, long long oe) {
  if (n <= 0) return;
  unsigned magic = 0u;
  int shift = 0;
  if (n > 1) {
    shift = 31 - __clz(n - 1);
    magic = n <= o.maxP ? o.magic_tab[n] : division_magic(n, shift);
  }
  const int total = n * o.C;
  for (int i = o.lane; i < total; i += 32) {
    const int t = o.C == 1 ? i : (o.C == 2 ? i >> 1 : i / o.C);
    const int a = o.raw[raw_wrap(rd + i, o.rrC)], b = o.raw[raw_wrap(ru + i, o.rrC)];
    const int num = a * (n - t) + b * t;
    const int q = n == 1 ? abs(num) : (int)(__umulhi((unsigned)abs(num), magic) >> shift);
    o.oring[(int)((oe + i) & o.omask)] = (short)(num < 0 ? -q : q);
  }
}

__device__ void output_role(const K4Params& p, const SpliceLayout& L, int s, const Plan& pl) {
  const Geometry& g = p.g;
  Output o;
  o.lane = threadIdx.x & 31;
  o.C = g.channels;
  o.rrC = L.rr * g.channels;
  o.omask = L.or_elems - 1;
  o.maxP = g.max_period;
  o.raw = reinterpret_cast<const short*>(splice_smem + L.raw);
  o.oring = reinterpret_cast<short*>(splice_smem + L.oring);
  unsigned* magic_tab = reinterpret_cast<unsigned*>(splice_smem + L.magic);
  o.magic_tab = magic_tab;
  o.out = p.out + (size_t)s * p.out_capacity * g.channels;
  o.cap_e = p.out_capacity * g.channels;
  uint64_t* bars = reinterpret_cast<uint64_t*>(splice_smem + L.bars);
  uint64_t* rempty = bars + L.nsr;
  uint64_t* qfull = bars + 2 * L.nsr + 2 * L.nsw;
  const int4* recs = reinterpret_cast<const int4*>(splice_smem + L.recs);
  int* ctrl = reinterpret_cast<int*>(splice_smem + L.ctrl);
  // the overlap-add's division constants, once per launch
  for (int n = 2 + o.lane; n <= o.maxP; n += 32) magic_tab[n] = division_magic(n, 31 - __clz(n - 1));
  // output elements below `fe` are in the stream's row; [fe, ee) sit in the ring
  long long ee = (long long)pl.out_count0 * o.C;
  long long fe = ee;
  const bool vec = L.out_vec && ((reinterpret_cast<size_t>(o.out) & 15) == 0);
  if (vec) {
    // the partly filled 16-byte block the launch starts in comes back into the ring
    fe = ee & ~7LL;
    for (long long e = fe + o.lane; e < ee; e += 32) o.oring[(int)(e & o.omask)] = o.out[e];
  }
  __syncwarp();

  int k = 0;            // records consumed
  int rreleased = 0;    // raw chunks handed back
  int last_pos = 0, last_idx = 0;  // raw-ring frame index of relative frame last_pos
#pragma unroll 1
  for (;;) {
    const int qslot = k & (kQ - 1);
    mbar_wait_sleep(qfull + qslot, (k / kQ) & 1);
    const int4 r = recs[qslot];
    const int kind = r.y >> 16, period = r.y & 0xffff, n = r.z, pos = r.x;
    if (kind == REC_EXIT) break;
    // raw-ring index of the record's first frame (records ascend, by less than the ring)
    int idx = last_idx + (pos - last_pos);
    while (idx >= L.rr) idx -= L.rr;
    last_pos = pos;
    last_idx = idx;
    const int r0 = idx * o.C;
    const int rp = raw_wrap(r0 + period * o.C, o.rrC);
    long long oe = (long long)r.w * o.C;  // first output element of the record
    int produced, npos;
    if (kind == REC_COPY) {
      produced = n;
      npos = pos + n;
    } else if (kind == REC_SKIP) {
      produced = n;
      npos = pos + period + n;
    } else {
      produced = period + n;
      npos = pos + n;
    }
    // clamp to the capacity of the row (the chain has flagged the overflow)
    long long room = o.cap_e - oe;
    if (room < 0) room = 0;
    const int room_f = (int)(room / o.C < produced ? room / o.C : produced);
    if (kind == REC_COPY) {
      out_copy(o, r0, room_f, oe);
    } else if (kind == REC_SKIP) {
      if (room_f == n) out_overlap_add(o, r0, rp, n, oe);
      else if (room_f > 0) {
        // a clipped cross-fade still divides by the full length: element by element
        for (int i = o.lane; i < room_f * o.C; i += 32) {
          const int t = i / o.C;
          const int a = o.raw[raw_wrap(r0 + i, o.rrC)], b = o.raw[raw_wrap(rp + i, o.rrC)];
          o.oring[(int)((oe + i) & o.omask)] = (short)((a * (n - t) + b * t) / n);
        }
      }
    } else {
      // insertPitchPeriod: the period itself, then the cross-fade back into it
      const int nc = room_f < period ? room_f : period;
      out_copy(o, r0, nc, oe);
      const int left = room_f - nc;
      if (left == n) out_overlap_add(o, rp, r0, n, oe + (long long)period * o.C);
      else if (left > 0) {
        for (int i = o.lane; i < left * o.C; i += 32) {
          const int t = i / o.C;
          const int a = o.raw[raw_wrap(rp + i, o.rrC)], b = o.raw[raw_wrap(r0 + i, o.rrC)];
          o.oring[(int)((oe + (long long)period * o.C + i) & o.omask)] = (short)((a * (n - t) + b * t) / n);
        }
      }
    }
    if (room_f > 0) ee = oe + (long long)room_f * o.C;
    __syncwarp();
    // drain what is complete
    if (vec) {
      const long long hi = ee & ~7LL;
      for (long long b = fe + o.lane * 8; b < hi; b += 256) {
        *reinterpret_cast<int4*>(o.out + b) = *reinterpret_cast<const int4*>(o.oring + (int)(b & o.omask));
      }
      if (hi > fe) fe = hi;
    } else {
      for (long long e = fe + o.lane; e < ee; e += 32) o.out[e] = o.oring[(int)(e & o.omask)];
      fe = ee;
    }
    __syncwarp();
    // raw chunks entirely below the next record's first frame go back to the filler
    while (((rreleased + 1) << kCFShift) <= npos) {
      if (o.lane == 0) mbar_arrive(rempty + rreleased % L.nsr);
      rreleased++;
    }
    k++;
    if (o.lane == 0) st_volatile_s32(ctrl, k);
  }
  // the tail of the last, partly filled 16-byte block
  for (long long e = fe + o.lane; e < ee; e += 32) o.out[e] = o.oring[(int)(e & o.omask)];
}

}  // namespace

template <bool K16>
__global__ void __launch_bounds__(kThreads, 7) k4_splice(K4Params p, SpliceLayout L) {
  const int s = blockIdx.x;
  if (s >= p.n_streams) return;
  const int warp = threadIdx.x >> 5;
  uint64_t* bars = reinterpret_cast<uint64_t*>(splice_smem + L.bars);
  if (threadIdx.x == 0) {
    for (int i = 0; i < L.nsr; i++) {
      mbar_init(bars + i, 1);          // rfull: the bulk copy's bytes (or the filler's plain loads)
      mbar_init(bars + L.nsr + i, 1);  // rempty: the output role
    }
    for (int i = 0; i < L.nsw; i++) {
      mbar_init(bars + 2 * L.nsr + i, 32);         // wfull: every filler lane
      mbar_init(bars + 2 * L.nsr + L.nsw + i, 1);  // wempty: the chain
    }
    for (int i = 0; i < kQ; i++) mbar_init(bars + 2 * L.nsr + 2 * L.nsw + i, 1);  // qfull: the chain
    int* ctrl = reinterpret_cast<int*>(splice_smem + L.ctrl);
    ctrl[0] = 0;  // records the output role has finished
    ctrl[1] = 0;  // the chain is done
    ctrl[2] = 0;  // window chunks the filler has widened
    fence_barrier_init();
  }
  __syncthreads();
  const Plan pl = make_splice_plan(p, s);
  __syncthreads();  // every role has read the stream's state before the chain rewrites it
  if (warp == 0) chain_role<K16>(p, L, s, pl);
  else if (warp == 1) filler_role(p, L, pl);
  else output_role(p, L, s, pl);
}

// Can the pipelined shape take this launch?  (Mono streams, a fine search that fits one round
// of lag groups; flushes and short launches stay with the one-warp kernel.)
bool k4_splice_supported(const K4Params& p) {
  const Geometry& g = p.g;
  if (g.channels != 1) return false;
  if (g.skip != 1 && 2 * g.skip + 2 > 32) return false;
  if (g.max_period >= 2048) return false;  // record layout, division constants
  return true;
}

cudaError_t launch_k4_splice(const K4Params& p0, cudaStream_t stream) {
  K4Params p = p0;
  const Geometry& g = p.g;
  k4_lane_map(p, 1);
  int depth = kDepthDefault, nsw_override = 0;
  bool tuned = false;
  if (const char* e = getenv("SPEEDY_K4_NSW")) {
    const int v = atoi(e);
    if (v >= (g.max_required + kPadW + kCF - 1) / kCF + 2 && v <= 32) nsw_override = v, tuned = true;
  }
  if (const char* e = getenv("SPEEDY_K4_DEPTH")) {
    const int d = atoi(e);
    if (d >= 1 && d <= 16) depth = d, tuned = true;
  }
  SpliceLayout L = make_layout(g.max_period, g.max_required, g.skip, g.channels, depth, nsw_override);
  if (const char* e = getenv("SPEEDY_K4_POLL_NS")) L.poll_ns = atoi(e) > 0 ? atoi(e) : 500;
  L.out_vec = ((p.out_capacity * g.channels) % 8 == 0 && (reinterpret_cast<size_t>(p.out) & 15) == 0) ? 1 : 0;
  // 16 kHz streams take the instantiation with compile-time geometry
  const bool skip4 = !tuned && g.channels == 1 && g.rate == 16000 && g.step == 160 && g.min_period == 40 &&
                     g.max_period == 246 && g.skip == 4;  // kLay16 is this layout
  static SmemOptIn opt16, opt;
  if (cudaError_t e = skip4 ? opt16.ensure(k4_splice<true>, L.total) : opt.ensure(k4_splice<false>, L.total)) return e;
  if (skip4) k4_splice<true><<<p.n_streams, kThreads, L.total, stream>>>(p, L);
  else k4_splice<false><<<p.n_streams, kThreads, L.total, stream>>>(p, L);
  count_launch();
  return cudaGetLastError();
}

}  // namespace speedy

#ifdef K4_TIMING
// developer build only (SPEEDY_K4_TIMING=1): read / reset the chain warp's per-phase cycle counters
extern "C" void speedyDebugK4SpliceCycles(unsigned long long* out, int reset) {
  if (out) {
    cudaMemcpyFromSymbol(out, speedy::g_k4s_cycles, sizeof(unsigned long long) * 16);
    unsigned long long a[8];  // the 16 kHz search's own phases (amdf16.cuh)
    cudaMemcpyFromSymbol(a, speedy::amdf16::g_amdf_cycles, sizeof(a));
    out[1] += a[0];
    out[2] += a[1] + a[2];
    out[4] += a[3] + a[4];
  }
  if (reset) {
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(speedy::g_k4s_cycles, z, sizeof(z));
    cudaMemcpyToSymbol(speedy::amdf16::g_amdf_cycles, z, sizeof(unsigned long long) * 8);
  }
}
#endif
