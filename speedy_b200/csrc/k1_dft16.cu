// K1, 16 kHz tensor-core shape — the 480-point spectrogram of speedyAddDataShort
// (speedy.c:553-565: int16 -> float, pre-emphasis :416-425, Hamming + zero-padded FFT +
// magnitude :438-474), frame energy (:510-523) and the raw local spectral difference
// (:705-719) with the transform done as a GEMM on the 5th-generation tensor cores
// (tcgen05.mma, operands in shared memory, accumulator in tensor memory).
//
// The DFT as a GEMM that fits one SM.  Only |X[k]| is needed, so the window can be re-centred:
// with u = n - 119.5 (n = 0 .. 239 the pre-emphasised, Hamming-weighted samples v[n]),
//   |X[k]| = | sum_u v(u) e^{-2 pi i k u / 480} |
//          = | sum_{t<120} s[t] cos(th_k (t + 1/2))  -  i sum_{t<120} d[t] sin(th_k (t + 1/2)) |
//   s[t] = v[120 + t] + v[119 - t],   d[t] = v[120 + t] - v[119 - t],   th_k = 2 pi k / 480,
// which halves the inner dimension (K = 120 instead of 240), and because
//   sin(th_k (t + 1/2)) = (-1)^t cos(th_{240-k} (t + 1/2))
// the sine matrix is the cosine matrix read backwards: with d'[t] = (-1)^t d[t],
//   Re[k] = (s . C)[k],  Im[k] = -(d' . C)[240 - k],  C[t][j] = cos(2 pi j (t + 1/2) / 480).
// One matrix C (120 x 241, padded to 128 x 256) serves both parts, stays resident in shared memory
// for the whole launch, and every window contributes two rows (s, d') to the M dimension.
//
// Precision.  fp16 operands, fp32 accumulation, each operand split into two fp16 terms: the rows at
// the int16 scale / 4 (|s| < 2^15) as hi (top 11 bits, by truncation) + lo (the exact remainder
// rounded to fp16), C as hi + lo: three products a_hi c_hi + a_lo c_hi + a_hi c_lo carry ~2^-21 of
// every term, the level of an fp32 FFT of this size (tests/test_gpu_parity.py holds both to the
// same 1e-4 bars).  The accumulator holds 2^13 X; everything downstream works on log2 |X|^2
// differences and on the energy, rescaled exactly (powers of two).
//
// Shape.  One persistent CTA per SM, 20 warps.  A tile is 64 window slots = 128 rows (2w: s,
// 2w + 1: d'), four groups of 16; a group holds runs of consecutive windows of a stream, each a halo
// window (the previous window, recomputed: its spectrum is the other operand of the spectral
// difference) + up to 15 new ones: one long run, or several short ones of different streams (make_tiling).
//   warps 8-18 prepare the rows: the tile's samples arrive in shared memory as one TMA bulk copy
//              (cp.async.bulk + mbarrier; tiles that touch the carried history or pack several streams
//              are staged by the warps themselves), then pre-emphasis, window, fold and split into
//              the A operand (no-swizzle canonical layout: core matrices of 8 rows x 16 bytes), one
//              16-byte K chunk of a row pair per unit of work, each announced on the mbarrier of its
//              MMA step;
//   warp 19    one thread issues the 24 tcgen05.mma of a tile, step by step as the chunks land, and
//              commits every step to a barrier that frees its chunks for the next tile's rows, so
//              the rows of tile i + 1 are written under the MMAs of tile i;
//   warps 0-7  read the accumulator (tcgen05.ld, thread = row, two warps per 32 rows with half of
//              the columns each): lane pairs (s-row, d'-row) trade the mirrored columns by shuffle,
//              each lane takes the bins of one side: power, log2, energy, 40 dB gate, spectral
//              difference against the window two lanes down.
// The accumulator is double-buffered in tensor memory (2 x 256 columns), so the epilogue of tile i
// runs under the preparation and the MMAs of tile i + 1.  Every wait is bounded (a wrong barrier
// must not hang the device): on a time-out the kernel sets an error word and drains.
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>

#include <mutex>
#include <vector>

#include "kernels.cuh"

namespace speedy {

__device__ int g_k1_dft16_error;
#ifdef K1_TIMING
// developer build: cycles per phase of CTA 0 (rows: 1 barrier, 2 samples, 3 units; issuer: 5 wait accumulator,
// 6 wait chunks + MMA issue; epilogue: 8 wait accumulator, 9 pass 1, 10 exchange, 11 pass 2 + out; 15 tiles)
__device__ unsigned long long g_k1_cycles[16];
#define KT_DECL long long _kt = clock64()
#define KT_MARK(slot, who) do { const long long _n = clock64(); if (blockIdx.x == 0 && (who)) atomicAdd(&g_k1_cycles[slot], (unsigned long long)(_n - _kt)); _kt = _n; } while (0)
#else
#define KT_DECL do {} while (0)
#define KT_MARK(slot, who) do {} while (0)
#endif

namespace {

constexpr int kS = 160, kW = 240, kP = 80;
constexpr int kGroupNew = 15, kGroups = 4, kTileNew = kGroupNew * kGroups;  // 60 new windows per tile
constexpr int kRows = 128, kK = 128, kN = 256;
constexpr int kEpiWarps = 8, kRowWarps = 11;  // + one warp for the MMA issuer
constexpr int kThreads = (kEpiWarps + kRowWarps + 1) * 32, kRowThreads = kRowWarps * 32;
constexpr float kPreHi = 0.97f;                          // speedy.c:422

// shared memory (bytes)
constexpr int kOffBhi = 0, kOffBlo = 65536, kOffAhi = 131072, kOffAlo = 163840;
constexpr int kOffWin = 196608;            // float[240]: Hamming / 4
constexpr int kOffBars = kOffWin + 1024;   // 22 mbarriers
constexpr int kOffSlot = kOffBars + 256;   // tensor-memory base address
constexpr int kOffSamp = kOffSlot + 64;    // short[kSampCap]: the samples of the tile's runs
constexpr int kSampCap = 12800;            // 32 runs of 400 (one new window + halo each) is the largest tile
constexpr int kOffXch = kOffSamp + kSampCap * 2;       // float2[2][4][16]: the two column halves of a window meet here
constexpr int kOffSlots = kOffXch + 2 * 4 * 16 * 8;    // Slot[4][64]: what each row pair of a tile is
constexpr int kOffBulk = kOffSlots + 4 * 64 * 16;      // int[4]: the tile's samples arrive as one bulk copy
constexpr int kSmemBytes = kOffBulk + 16;
constexpr int kSpanSamples = kTileNew * kS + kW;  // one stream's 61 consecutive windows: 9840 samples

// element (r, k) of an [R x 128] K-major fp16 operand in the no-swizzle canonical layout: core
// matrices of 8 rows x 16 bytes contiguous along the rows (stride byte offset 128), the 8-element
// K chunks R * 16 bytes apart (leading byte offset)
__host__ __device__ inline int op_off(int r, int k, int R) { return (k >> 3) * (R * 16) + (r >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2; }

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(unsigned addr, unsigned lbo_bytes, unsigned sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version 1 (sm_100); layout type 0 = no swizzle
  return d;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait (a wrong barrier must not hang the device): the try sleeps in the barrier unit
// for up to the hinted time; false (and the error word set) after a few thousand tries
__device__ __forceinline__ bool mbar_try_hint(uint64_t* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(200000u)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, unsigned parity) {
  if (mbar_try(bar, parity)) return true;
#pragma unroll 1
  for (int tries = 0; tries < 20000; tries++) {
    if (mbar_try_hint(bar, parity)) return true;
  }
  atomicExch(&g_k1_dft16_error, 1);
  return false;
}
// global -> shared bulk copy (TMA, 1-D), completion counted in bytes on `bar`
// pull a span into L2 ahead of its bulk copy
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// all earlier tcgen05.mma of this thread done -> one arrival on `bar`
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_f16(unsigned tmem_d, uint64_t da, uint64_t db, unsigned idesc, unsigned accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 16 consecutive columns of this thread's tensor-memory lane
__device__ __forceinline__ void tmem_ld16(unsigned taddr, float (&v)[16]) {
  unsigned r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}

// log2 on the special-function unit alone (the powers here are never subnormal unless they are zero)
__device__ __forceinline__ float fast_log2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// x = hi + lo with hi on an fp16 grid (top 11 bits, by truncation: exact) and lo the exact
// remainder rounded to fp16; two values per call, packed as half2 bit patterns
__device__ __forceinline__ void split2(float a, float b, unsigned& hi, unsigned& lo) {
  const float ah = __uint_as_float(__float_as_uint(a) & 0xffffe000u), bh = __uint_as_float(__float_as_uint(b) & 0xffffe000u);
  const __half2 h = __floats2half2_rn(ah, bh);
  const __half2 l = __floats2half2_rn(a - ah, b - bh);
  hi = *reinterpret_cast<const unsigned*>(&h);
  lo = *reinterpret_cast<const unsigned*>(&l);
}

// How a launch is cut into tiles.  Every stream has at most n = max_new_frames new windows.  A run is
// a halo window (the predecessor of its first new window, recomputed) followed by up to 15 new ones,
// consecutive windows of one stream; the 16 window slots of a warp-sized group of rows hold one long
// run (n > 15: the stream is cut into runs of 15) or as many short ones as fit (n <= 15: one run per
// stream, 16 / (n + 1) streams per group: a 10 ms streaming write of thousands of sessions packs 32
// streams into a tile).  A tile is four groups.
struct Tiling {
  int new_per_run, slots_per_run, runs_per_group, runs_per_stream, run_samples, total_runs, n_tiles;
};
__host__ __device__ inline Tiling make_tiling(int n, int n_streams) {
  Tiling t;
  t.new_per_run = n < kGroupNew ? n : kGroupNew;
  t.slots_per_run = t.new_per_run + 1;
  t.runs_per_group = 16 / t.slots_per_run;
  t.runs_per_stream = (n + kGroupNew - 1) / kGroupNew;
  t.run_samples = t.new_per_run * kS + kW;
  t.total_runs = t.runs_per_stream * n_streams;
  const int groups = (t.total_runs + t.runs_per_group - 1) / t.runs_per_group;
  t.n_tiles = (groups + kGroups - 1) / kGroups;
  return t;
}

// One window slot (a row pair) of a tile.
struct __align__(16) Slot {
  int s;      // stream
  int k;      // window (k = -1: the halo before a stream's first window: zeros)
  int kA;     // the stream's scratch rows count from this window
  int flags;  // 1: a window of this launch (results go out)  2: rows carry data (0 <= k < windows analysed)
              // bits 8..15: position in its run, bits 16..: the run's index in the tile
};

// speedy's frames_analyzed for W = 240, S = 160 (soniclib.c:440-444), in 32 bits when the total allows
__device__ __forceinline__ int windows_after(long long total) {
  if (total < kW + 1) return 0;
  if (total < (1LL << 31)) return ((int)total - kW - 1) / kS + 1;
  return (int)((total - kW - 1) / kS) + 1;
}

__device__ __forceinline__ Slot make_slot(const K1Params& p, const Tiling& tl, int tile, int w) {
  Slot sl;
  sl.s = 0; sl.k = 0; sl.kA = 0; sl.flags = 0;
  const int g = w >> 4, wl = w & 15;
  const int r_local = wl / tl.slots_per_run, pos = wl - r_local * tl.slots_per_run;
  const int run = (tile * kGroups + g) * tl.runs_per_group + r_local;
  sl.flags = (pos << 8) | ((g * tl.runs_per_group + r_local) << 16);
  if (r_local >= tl.runs_per_group || run >= tl.total_runs) return sl;
  sl.s = run / tl.runs_per_stream;
  const int j = run - sl.s * tl.runs_per_stream;
  if (p.st.nonlinear[sl.s] == 0.0f) return sl;  // soniclib.c:397-399: Speedy is bypassed in the linear mode
  const Range rg = write_range(p.st.total, p.counts, p.frames, p.done, sl.s);
  sl.kA = windows_after(rg.t_old);
  const int kD = windows_after(rg.t_done), kB = windows_after(rg.t_new);
  sl.k = kD + kGroupNew * j + pos - 1;
  if (kD + kGroupNew * j >= kB) return sl;  // the run has no new window
  if (sl.k >= 0 && sl.k < kB) sl.flags |= 2;
  if (pos >= 1 && sl.k < kB) sl.flags |= 1;
  return sl;
}

// Long launches: can the tile's samples -- one stream, four consecutive runs, frames
// [(k0 - 1) S, (k0 + 59) S + W) -- come as one aligned bulk copy straight from the caller's buffer?
// (Not the first tile of a write: its halo window reaches into the carried history.)
__device__ __forceinline__ const int16_t* tile_bulk_src(const K1Params& p, const Tiling& tl, int tile) {
  if (tile >= tl.n_tiles || tl.runs_per_group != 1 || p.g.channels != 1 || p.in == nullptr) return nullptr;
  const int run0 = tile * kGroups;
  const int s = run0 / tl.runs_per_stream, j = run0 - s * tl.runs_per_stream;
  if (j + kGroups > tl.runs_per_stream || p.st.nonlinear[s] == 0.0f) return nullptr;
  const Range rg = write_range(p.st.total, p.counts, p.frames, p.done, s);
  const int kD = windows_after(rg.t_done);
  const long long f0 = (long long)(kD + kGroupNew * j - 1) * kS;
  if (f0 < rg.t_old || f0 + kSpanSamples > rg.t_new) return nullptr;
  const int16_t* src = p.in + (size_t)s * p.in_stride_frames + (f0 - rg.t_old);
  return (reinterpret_cast<size_t>(src) & 15) == 0 ? src : nullptr;
}

}  // namespace

template <bool TAP>
__global__ void __launch_bounds__(kThreads, 1) k1_dft16(K1Params p, const uint4* __restrict__ dft_hi_lo, Tiling tl) {
  extern __shared__ __align__(1024) unsigned char smem[];
  float* s_win = reinterpret_cast<float*>(smem + kOffWin);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBars);
  uint64_t* bar_k_ready = bars + 0;    // [8] the two 8-element K chunks of MMA step ks are in shared memory (128 units; 64 for the last)
  uint64_t* bar_k_free = bars + 8;     // [8] commit: the MMAs of step ks have read them
  uint64_t* bar_acc_full = bars + 16;  // [2] commit: accumulator b holds a tile
  uint64_t* bar_acc_free = bars + 18;  // [2] 256 arrivals: the epilogue has read accumulator b (and copied its slot)
  uint64_t* bar_smp = bars + 20;       // bulk copy of a tile's samples has landed
  uint64_t* bar_smp_free = bars + 21;  // every row thread holds its share of the tile's samples in registers
  unsigned* tmem_slot = reinterpret_cast<unsigned*>(smem + kOffSlot);
  short* smp = reinterpret_cast<short*>(smem + kOffSamp);
  // [4][64]: tile it uses table it % 4.  Four are enough without a barrier of their own: the table of
  // tile it + 1 is written while the rows of tile it are, i.e. after the MMAs of tile it - 1 have run, and
  // those were issued only after the epilogue of tile it - 3 had handed its accumulator back (it copies
  // its slot first)
  Slot* slots = reinterpret_cast<Slot*>(smem + kOffSlots);
  int* bulk_flag = reinterpret_cast<int*>(smem + kOffBulk);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_tiles = tl.n_tiles;

  // the DFT matrix (hi, then lo), already in operand layout: 128 KB, once per CTA
  {
    uint4* dst = reinterpret_cast<uint4*>(smem + kOffBhi);
    for (int i = tid; i < 131072 / 16; i += kThreads) dst[i] = dft_hi_lo[i];
    uint4* a = reinterpret_cast<uint4*>(smem + kOffAhi);  // rows start as zeros (the K padding stays zero)
    for (int i = tid; i < 65536 / 16; i += kThreads) a[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int i = tid; i < kW; i += kThreads) s_win[i] = p.window[i] * 0.25f;  // Hamming / 4: rows at the int16 scale / 4
  }
  if (tid == 0) {
    for (int ks = 0; ks < 8; ks++) {
      mbar_init(bar_k_ready + ks, ks < 7 ? 128 : 64);  // (chunk 15 is the K padding: nobody writes it)
      mbar_init(bar_k_free + ks, 1);
    }
    mbar_init(bar_acc_full + 0, 1);
    mbar_init(bar_acc_full + 1, 1);
    mbar_init(bar_acc_free + 0, 256);
    mbar_init(bar_acc_free + 1, 256);
    mbar_init(bar_smp, 1);
    mbar_init(bar_smp_free, kRowThreads);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem = *tmem_slot;

  if (warp == kEpiWarps + kRowWarps) {
    // ================================ MMA issue =====================================
    // One thread: as soon as the two K chunks of a step are in shared memory it issues the step's three
    // products and commits them to the step's "free" barrier, so the rows of the next tile are written
    // under the MMAs of this one; after the last step it starts the next tile's bulk copy.
    if (lane == 0) {
      const unsigned idesc = (1u << 4) | ((unsigned)(kN >> 3) << 17) | ((unsigned)(kRows >> 4) << 24);  // f16 x f16 -> f32, K-major both
      const unsigned ah = smem_u32(smem + kOffAhi), al = smem_u32(smem + kOffAlo);
      const unsigned bh = smem_u32(smem + kOffBhi), bl = smem_u32(smem + kOffBlo);
      const int16_t* first = tile_bulk_src(p, tl, blockIdx.x);
      if (first) bulk_load(smp, first, kSpanSamples * 2, bar_smp);
      if (const int16_t* second = tile_bulk_src(p, tl, blockIdx.x + gridDim.x)) bulk_prefetch_l2(second, kSpanSamples * 2);
      int it = 0;
      bool ok = true;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
        const int b = it & 1;
        KT_DECL;
        // the next tile's samples start towards L2 now; its bulk copy follows when this tile's rows are done
        const int16_t* nxt = tile_bulk_src(p, tl, tile + gridDim.x);
        const int16_t* nxt2 = tile_bulk_src(p, tl, tile + 2 * gridDim.x);
        if (nxt2) bulk_prefetch_l2(nxt2, kSpanSamples * 2);
        // the row threads take their samples into registers first thing: the next tile's bulk copy can
        // start as soon as they have, and lands under this tile's arithmetic
        if (ok) ok = mbar_wait(bar_smp_free, (unsigned)it & 1u);
        if (nxt && ok) bulk_load(smp, nxt, kSpanSamples * 2, bar_smp);
        if (ok && it >= 2) ok = mbar_wait(bar_acc_free + b, (unsigned)((it >> 1) - 1) & 1u);
        KT_MARK(5, true);
        const unsigned acc = tmem + (unsigned)(b * kN);
#pragma unroll 1
        for (int ks = 0; ks < kK / 16; ks++) {  // two 8-element chunks per MMA
          if (ok) ok = mbar_wait(bar_k_ready + ks, (unsigned)it & 1u);
          if (ok) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const unsigned ao = (unsigned)ks * 2u * (kRows * 16), bo = (unsigned)ks * 2u * (kN * 16);
            const uint64_t d_ah = make_desc(ah + ao, kRows * 16, 128), d_al = make_desc(al + ao, kRows * 16, 128);
            const uint64_t d_bh = make_desc(bh + bo, kN * 16, 128), d_bl = make_desc(bl + bo, kN * 16, 128);
            umma_f16(acc, d_ah, d_bh, idesc, ks > 0 ? 1u : 0u);
            umma_f16(acc, d_al, d_bh, idesc, 1u);
            umma_f16(acc, d_ah, d_bl, idesc, 1u);
          }
          umma_commit(bar_k_free + ks);  // (committed even after a time-out, so that the other roles drain too)
        }
        umma_commit(bar_acc_full + b);
        KT_MARK(6, true);
#ifdef K1_TIMING
        if (blockIdx.x == 0) atomicAdd(&g_k1_cycles[15], 1ULL);
#endif
      }
    }
  } else if (warp >= kEpiWarps) {
    // ================================== rows =========================================
    const int ptid = tid - 256, pwarp = warp - 8;
    unsigned char* a_hi = smem + kOffAhi;
    unsigned char* a_lo = smem + kOffAlo;
    if (ptid < 64) slots[ptid] = make_slot(p, tl, blockIdx.x, ptid);
    if (ptid == 64) bulk_flag[0] = tile_bulk_src(p, tl, blockIdx.x) != nullptr;
    int it = 0, n_bulk = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
      KT_DECL;
      Slot* tile_slots = slots + (it & 3) * 64;
      // (every row thread is done with the previous tile's samples; this tile's slot table is visible)
      asm volatile("bar.sync 1, %0;" ::"n"(kRowThreads) : "memory");
      KT_MARK(1, ptid == 32);
      const bool bulk = bulk_flag[it & 3] != 0;
      // ---- the samples of the tile's runs ------------------------------------------
      // bulk: smp[i] = frame (k0 - 1) S + i of the one stream; else run r of the tile at smp + r * run_samples
      const int run_stride = bulk ? kGroupNew * kS : tl.run_samples;
      if (bulk) {
        if (!mbar_wait(bar_smp, (unsigned)n_bulk & 1u)) break;
        n_bulk++;
      } else {
        for (int r = pwarp; r < kGroups * tl.runs_per_group; r += kRowWarps) {
          const int g = r / tl.runs_per_group, r_local = r - g * tl.runs_per_group;
          const Slot first = tile_slots[g * 16 + r_local * tl.slots_per_run];
          bool any = false;
          for (int q = 0; q < tl.slots_per_run; q++) any = any || (tile_slots[g * 16 + r_local * tl.slots_per_run + q].flags & 2);
          if (!any) continue;
          const Range rg = write_range(p.st.total, p.counts, p.frames, p.done, first.s);
          Source src;
          src.channels = p.g.channels;
          src.hist = p.hist + (size_t)first.s * p.hist_stride;
          src.in = p.in ? p.in + (size_t)first.s * p.in_stride_frames * p.g.channels : nullptr;
          src.hist_base = p.st.hist_base[first.s];
          src.t_old = rg.t_old;
          src.t_new = rg.t_new;
          stage_mono<32, short>(src, (long long)first.k * kS, tl.run_samples, rg.t_new, smp + r * tl.run_samples, nullptr, lane);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kRowThreads) : "memory");
      }
      KT_MARK(2, ptid == 32);
      // ---- rows: unit = (chunk of eight t, window slot); consecutive threads take consecutive slots.
      // First the raw samples of all of a thread's units (at most kMaxUnits) into registers, so that the
      // sample buffer is free for the next tile's bulk copy while the arithmetic runs.
      constexpr int kMaxUnits = (15 * 64 + kRowThreads - 1) / kRowThreads;
      uint4 raw_u[kMaxUnits], raw_l[kMaxUnits];
      float raw_pu[kMaxUnits], raw_pl[kMaxUnits];
#pragma unroll
      for (int i = 0; i < kMaxUnits; i++) {
        const int u = ptid + i * kRowThreads;
        raw_u[i] = make_uint4(0u, 0u, 0u, 0u);
        raw_l[i] = raw_u[i];
        raw_pu[i] = 0.0f;
        raw_pl[i] = 0.0f;
        if (u < 15 * 64) {
          const int c = u >> 6, w = u & 63;
          const Slot sl = tile_slots[w];
          if (sl.flags & 2) {
            const short* x = smp + (sl.flags >> 16) * run_stride + ((sl.flags >> 8) & 0xff) * kS;
            const int nu = 120 + 8 * c, nl = 112 - 8 * c;  // first sample of the upper / lower eight
            raw_u[i] = *reinterpret_cast<const uint4*>(x + nu);
            raw_l[i] = *reinterpret_cast<const uint4*>(x + nl);
            raw_pu[i] = (float)x[nu - 1];
            // the state entering sample 0 is the last sample of the previous window, i.e. sample
            // P - 1 of this one (speedy.c:416-425); 0 before the first window
            raw_pl[i] = nl > 0 ? (float)x[nl - 1] : (sl.k >= 1 ? (float)x[kP - 1] : 0.0f);
          }
        }
      }
      mbar_arrive(bar_smp_free);
      bool ok = true;
#pragma unroll
      for (int i = 0; i < kMaxUnits; i++) {
        const int u = ptid + i * kRowThreads;
        if (u >= 15 * 64) break;
        const int c = u >> 6, w = u & 63;
        const Slot sl = tile_slots[w];
        uint4 s_hi = make_uint4(0u, 0u, 0u, 0u), s_lo = s_hi, d_hi = s_hi, d_lo = s_hi;
        if (sl.flags & 2) {
          const int nu = 120 + 8 * c, nl = 112 - 8 * c;
          const uint4 qu = raw_u[i];
          const uint4 ql = raw_l[i];
          const float4 hu0 = *reinterpret_cast<const float4*>(s_win + nu), hu1 = *reinterpret_cast<const float4*>(s_win + nu + 4);
          const float4 hl0 = *reinterpret_cast<const float4*>(s_win + nl), hl1 = *reinterpret_cast<const float4*>(s_win + nl + 4);
          float xu[9], xl[9];  // [0] = the sample before
          xu[0] = raw_pu[i];
          xl[0] = raw_pl[i];
          const unsigned wu[4] = {qu.x, qu.y, qu.z, qu.w}, wlw[4] = {ql.x, ql.y, ql.z, ql.w};
#pragma unroll
          for (int i = 0; i < 4; i++) {
            xu[1 + 2 * i] = (float)(short)(wu[i] & 0xffffu);
            xu[2 + 2 * i] = (float)(short)(wu[i] >> 16);
            xl[1 + 2 * i] = (float)(short)(wlw[i] & 0xffffu);
            xl[2 + 2 * i] = (float)(short)(wlw[i] >> 16);
          }
          const float hu[8] = {hu0.x, hu0.y, hu0.z, hu0.w, hu1.x, hu1.y, hu1.z, hu1.w};
          const float hl[8] = {hl0.x, hl0.y, hl0.z, hl0.w, hl1.x, hl1.y, hl1.z, hl1.w};
          float vu[8], vl[8];
#pragma unroll
          for (int i = 0; i < 8; i++) {
            // y = x - 0.97 * previous (speedy.c:422, evaluated there in double).  One fused multiply-add
            // with the float constant: its 2e-8 relative error is far below the 2^-21 of the fp16 split
            // that follows (the FFT kernels, which keep fp32 throughout, carry the constant's remainder too)
            vu[i] = __fmul_rn(__fmaf_rn(-kPreHi, xu[i], xu[i + 1]), hu[i]);
            vl[i] = __fmul_rn(__fmaf_rn(-kPreHi, xl[i], xl[i + 1]), hl[i]);
          }
          // t = 8c + i pairs sample 120 + t (vu[i]) with sample 119 - t (vl[7 - i])
          float sv[8], dv[8];
#pragma unroll
          for (int i = 0; i < 8; i++) {
            sv[i] = vu[i] + vl[7 - i];
            dv[i] = (i & 1) ? (vl[7 - i] - vu[i]) : (vu[i] - vl[7 - i]);  // (-1)^t d[t]
          }
          split2(sv[0], sv[1], s_hi.x, s_lo.x);
          split2(sv[2], sv[3], s_hi.y, s_lo.y);
          split2(sv[4], sv[5], s_hi.z, s_lo.z);
          split2(sv[6], sv[7], s_hi.w, s_lo.w);
          split2(dv[0], dv[1], d_hi.x, d_lo.x);
          split2(dv[2], dv[3], d_hi.y, d_lo.y);
          split2(dv[4], dv[5], d_hi.z, d_lo.z);
          split2(dv[6], dv[7], d_hi.w, d_lo.w);
        }
        // the previous tile's MMAs of this K step have read the chunk
        if (it > 0 && ok) ok = mbar_wait(bar_k_free + (c >> 1), (unsigned)(it - 1) & 1u);
        const int o = op_off(2 * w, 8 * c, kRows);  // (row 2w + 1 is the next 16 bytes)
        *reinterpret_cast<uint4*>(a_hi + o) = s_hi;
        *reinterpret_cast<uint4*>(a_hi + o + 16) = d_hi;
        *reinterpret_cast<uint4*>(a_lo + o) = s_lo;
        *reinterpret_cast<uint4*>(a_lo + o + 16) = d_lo;
        // the tensor core reads shared memory through the async proxy
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(bar_k_ready + (c >> 1));
      }
      KT_MARK(3, ptid == 32);
      // the next tile's slot table, off the critical path
      if (tile + gridDim.x < n_tiles) {
        if (ptid < 64) slots[((it + 1) & 3) * 64 + ptid] = make_slot(p, tl, tile + gridDim.x, ptid);
        if (ptid == 64) bulk_flag[(it + 1) & 3] = tile_bulk_src(p, tl, tile + gridDim.x) != nullptr;
      }
      if (!ok) break;
    }
  } else {
    // =============================== epilogue =====================================
    // thread = accumulator row: lane pair (2 wl, 2 wl + 1) = (s-row, d'-row) of window wl of group `grp`;
    // a lane's bins: role 0: k = i, role 1: k = 240 - i, and the i range 0 .. 120 is cut in two halves
    // (0 .. 60, 61 .. 120) taken by warps grp and grp + 4, which read the same tensor-memory lanes
    const int grp = warp & 3, half = warp >> 2;
    const int role = lane & 1, wl = lane >> 1;
    const int i0 = 61 * half;
    float2* xch = reinterpret_cast<float2*>(smem + kOffXch);  // [2 halves][4 groups][16 windows]: (energy, peak), then lsd
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
      const int b = it & 1;
      KT_DECL;
      if (!mbar_wait(bar_acc_full + b, (unsigned)(it >> 1) & 1u)) break;
      KT_MARK(8, tid == 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const unsigned row_base = tmem + ((unsigned)(grp * 32) << 16) + (unsigned)(b * kN);
      const Slot sl = slots[(it & 3) * 64 + grp * 16 + wl];
      const bool out = (sl.flags & 1) != 0;
      float* tap = nullptr;
      if (TAP && out) tap = p.tap_spec + ((size_t)sl.s * p.tap_stride + (sl.k - sl.kA)) * 480;
      // pass 1: power and log2 power of this lane's bins; energy and peak (bins 0 and N/2 stay out of both)
      float l[64];
      float e = 0.0f, mx = 0.0f;
#pragma unroll
      for (int c = 0; c < 4; c++) {
        float lo[16], hi[16];
        tmem_ld16(row_base + (unsigned)(i0 + 16 * c), lo);         // own columns i = i0 + 16c + j
        tmem_ld16(row_base + (unsigned)(225 - i0 - 16 * c), hi);   // own columns 240 - i for those i, backwards
#pragma unroll
        for (int j = 0; j < 16; j++) {
          const int ii = 16 * c + j;  // i - i0
          if (ii > 60) continue;
          const int i = i0 + ii;
          // the partner's column 240 - i: Re and Im of the same bin meet here
          const float other = __shfl_xor_sync(0xffffffffu, hi[15 - j], 1);
          const float q = __fmaf_rn(lo[j], lo[j], __fmul_rn(other, other));  // 2^26 |X|^2
          l[ii] = fast_log2(q);
          // bins 0 (ii = 0 of the lower half) and N/2 stay out; bin 120 counts once (role 0); ii = 60 of the
          // upper half (i = 121) is padding: only these three positions need a test
          bool counted = true;
          if (ii == 0) counted = half != 0;
          if (ii == 59) counted = half == 0 || role == 0;
          if (ii == 60) counted = half == 0;
          const float qc = counted ? q : 0.0f;
          e += qc;
          mx = fmaxf(mx, qc);
          if (TAP) {
            if (tap && i <= 120 && (i < 120 || role == 0)) {
              const float m = __fsqrt_rn(q) * 1.220703125e-4f;  // 2^-13
              const int kbin = role ? 240 - i : i;
              tap[kbin] = m;
              if (kbin != 0 && kbin != 240) tap[480 - kbin] = m;
            }
          }
        }
      }
      // the accumulator is in registers now: hand it back to the MMA issuer
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(bar_acc_free + b);
      KT_MARK(9, tid == 0);
      // energy and peak of the window: the two roles by shuffle, the two halves through shared memory
      e += __shfl_xor_sync(0xffffffffu, e, 1);
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      if (role == 0) xch[(half * 4 + grp) * 16 + wl] = make_float2(e, mx);
      asm volatile("bar.sync %0, 64;" ::"r"(2 + grp) : "memory");
      {
        const float2 o = xch[((half ^ 1) * 4 + grp) * 16 + wl];
        e = (e + o.x) * 1.4901161193847656e-08f;  // 2^-26: the /32768 scale of speedy.c:558
        mx = fmaxf(mx, o.y);
      }
      KT_MARK(10, tid == 0);
      // pass 2, speedy.c:705-719 in the log2 domain: with n_i = |X_i| / (sqrt(E) + eps),
      //   log(n_c / n_l) = ln2 * (0.5 (lp_c - lp_l) + (linv_c - linv_l)),  lp = log2 |X|^2, linv = -log2(sqrt(E) + eps);
      //   |X_i| > max|X| / 100  <=>  lp_i > log2(max p) - log2(1e4)
      const float linv = -fast_log2(__fsqrt_rn(e) + 2.2204e-16f);
      const float linv_last = __shfl_up_sync(0xffffffffu, linv, 2);
      const float thr = fast_log2(mx) - 13.287712379549449f;
      const float d2 = 2.0f * (linv - linv_last);
      float acc = 0.0f;
#pragma unroll
      for (int ii = 0; ii <= 60; ii++) {
        const float ll = __shfl_up_sync(0xffffffffu, l[ii], 2);  // the same bin of the previous window
        const float term = fabsf((l[ii] - ll) + d2);
        bool counted = true;  // (as in pass 1)
        if (ii == 0) counted = half != 0;
        if (ii == 59) counted = half == 0 || role == 0;
        if (ii == 60) counted = half == 0;
        if (counted && l[ii] > thr && ll > thr) acc += term;
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      asm volatile("bar.sync %0, 64;" ::"r"(2 + grp) : "memory");  // (both halves have read the energies)
      if (half == 1 && role == 0) xch[(4 + grp) * 16 + wl].x = acc;
      asm volatile("bar.sync %0, 64;" ::"r"(2 + grp) : "memory");
      if (half == 0 && out && role == 0) {
        acc += xch[(4 + grp) * 16 + wl].x;
        p.feat[(size_t)sl.s * p.feat_stride + (sl.k - sl.kA)] = make_float2(e, acc * 0.34657359027997264f);  // ln2 / 2
      }
      asm volatile("bar.sync %0, 64;" ::"r"(2 + grp) : "memory");  // (the exchange slots are free for the next tile)
      KT_MARK(11, tid == 0);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

// ---- host side -------------------------------------------------------------------

bool k1_dft16_supported(const K1Params& p) {
  const Geometry& g = p.g;
  // (any channel count: the mono down-mix of soniclib.c:271-274 happens while the samples are staged)
  return g.fft == 480 && g.window == kW && g.step == kS && g.partial == kP && g.channels >= 1;
}

// C[t][j] = cos(2 pi j (t + 1/2) / 480) as fp16 hi + lo in operand layout (row = j, K = t), per device
static const uint4* dft_matrix(cudaError_t* err) {
  static std::mutex mu;
  static const uint4* table[kMaxDevices];
  int dev = 0;
  if ((*err = cudaGetDevice(&dev)) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  if (dev < 0 || dev >= kMaxDevices) {
    *err = cudaErrorInvalidDevice;
    return nullptr;
  }
  if (table[dev]) return table[dev];
  std::vector<unsigned char> host(131072, 0);
  for (int j = 0; j <= 240; j++) {
    for (int t = 0; t < 120; t++) {
      const double c = cos(2.0 * M_PI * (double)j * ((double)t + 0.5) / 480.0);
      const __half hi = __float2half_rn((float)c);
      const __half lo = __float2half_rn((float)(c - (double)__half2float(hi)));
      *reinterpret_cast<__half*>(host.data() + op_off(j, t, kN)) = hi;
      *reinterpret_cast<__half*>(host.data() + 65536 + op_off(j, t, kN)) = lo;
    }
  }
  void* d = nullptr;
  if ((*err = cudaMalloc(&d, host.size())) != cudaSuccess) return nullptr;
  if ((*err = cudaMemcpy(d, host.data(), host.size(), cudaMemcpyHostToDevice)) != cudaSuccess) return nullptr;
  table[dev] = reinterpret_cast<const uint4*>(d);
  return table[dev];
}

// The matrix is built and uploaded when a batch is created (a synchronous allocation and copy have no
// place on a launch path that may be under stream capture).
cudaError_t k1_dft16_prepare() {
  cudaError_t e = cudaSuccess;
  return dft_matrix(&e) ? cudaSuccess : e;
}

cudaError_t launch_k1_dft16(const K1Params& p, cudaStream_t stream) {
  cudaError_t e = cudaSuccess;
  const uint4* dft = dft_matrix(&e);
  if (!dft) return e;
  static SmemOptIn opt, opt_tap;
  if ((e = (p.tap_spec ? opt_tap.ensure(k1_dft16<true>, kSmemBytes) : opt.ensure(k1_dft16<false>, kSmemBytes))) != cudaSuccess) return e;
  static int sms[kMaxDevices];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < kMaxDevices && sms[dev] == 0) cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
  const int n_sm = (dev >= 0 && dev < kMaxDevices && sms[dev] > 0) ? sms[dev] : 148;
  // new windows per stream in this launch: at most one per step of new frames
  long long n = (p.frames - p.done + kS - 1) / kS;
  if (n > p.max_new_frames) n = p.max_new_frames;
  if (n < 1) n = 1;
  const Tiling tl = make_tiling((int)n, p.n_streams);
  const int grid = tl.n_tiles < n_sm ? tl.n_tiles : n_sm;
  if (p.tap_spec) k1_dft16<true><<<grid, kThreads, kSmemBytes, stream>>>(p, dft, tl);
  else k1_dft16<false><<<grid, kThreads, kSmemBytes, stream>>>(p, dft, tl);
  count_launch();
  return cudaGetLastError();
}

}  // namespace speedy

#ifdef K1_TIMING
extern "C" void speedyDebugK1Cycles(unsigned long long* out, int reset) {
  if (out) cudaMemcpyFromSymbol(out, speedy::g_k1_cycles, sizeof(unsigned long long) * 16);
  if (reset) {
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(speedy::g_k1_cycles, z, sizeof(z));
  }
}
#endif

extern "C" int speedyDebugK1Dft16Error(void) {
  int v = 0;
  cudaMemcpyFromSymbol(&v, speedy::g_k1_dft16_error, sizeof(v));
  return v;
}
