// Kernel parameter blocks and host-side launchers (one .cu per kernel family).
//
// A write of L frames per stream may be processed as several launches over growing
// prefixes (done, frames] of the same write: the analysis kernels of a later prefix
// then overlap the resynthesis kernel of an earlier one (they only meet through the
// speeds array), and in speedyBatchProcess a prefix can be processed while the rest
// of the write is still crossing PCIe.  `done` = 0, `frames` = L is the one-launch case.
// Scratch rows stay indexed from the first new frame of the whole write.
#pragma once

#include <atomic>

#include "common.cuh"

#ifndef K1_RUN
#define K1_RUN 15  // new analysis windows per warp/CTA (plus one halo window)
#endif

namespace speedy {

constexpr int kMaxFactors = 16;

void count_launch();

// Opt a kernel in to `bytes` of dynamic shared memory on the CURRENT device.  The attribute
// is per device and launches may come from several host threads, so the largest size
// already granted is cached per device in atomics (one cache per call site).
constexpr int kMaxDevices = 64;
struct SmemOptIn {
  std::atomic<int> granted[kMaxDevices];
  template <class Kernel>
  cudaError_t ensure(Kernel kernel, size_t bytes) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const bool cached = dev >= 0 && dev < kMaxDevices;
    if (cached && granted[dev].load(std::memory_order_acquire) >= (int)bytes) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
    if (cached) {
      int seen = granted[dev].load(std::memory_order_relaxed);
      while (seen < (int)bytes && !granted[dev].compare_exchange_weak(seen, (int)bytes, std::memory_order_release)) {
      }
    }
    return cudaSuccess;
  }
};

// ---- K1: spectrogram, frame energy, raw spectral difference ---------------
struct K1Params {
  Geometry g;
  StreamState st;
  int n_streams;
  // input of this write
  const int16_t* hist;           // [n][hist_stride] carried tail (current buffer)
  long long hist_stride;         // int16 elements per stream
  const int16_t* in;             // caller's device buffer (may be null if frames == 0)
  long long in_stride_frames;
  const int32_t* counts;         // optional per-stream frame counts
  long long frames;              // frames of this write visible so far (a prefix)
  long long done;                // prefix already handled by earlier launches of this write
  // per-call scratch out: [n][feat_stride] (energy, raw spectral difference)
  float2* feat;
  int feat_stride;
  int max_new_frames;
  int runs_per_stream;           // filled by the launcher
  // tables
  const float* window;           // [W] Hamming (speedy.c:256-258)
  const float2* tw_n;            // [N] W_N^k
  const float2* tw_half;         // [N/2] W_{N/2}^k
  int n_factors;
  int factors[kMaxFactors];      // generic path: radices of N; mixed path: radices of M = W/2
  int plan_m[kMaxFactors];       // mixed path: sub-transform length after each stage
  int plan_per[kMaxFactors];     // mixed path: butterflies per transform in each stage
  // chirp-z (Bluestein) path for windows the mixed-radix kernel cannot factor (44.1 kHz: W = 661):
  int bl_L;                      // convolution length (0: not used); the plan above is then for L
  const float2* bl_tw;           // [L] W_L^k
  const float2* bl_B;            // [L] FFT_L of the chirp filter e^{+i pi m^2 / N}, divided by L
  const float2* bl_chirp;        // [W] e^{-i pi n^2 / N}
  // optional tap: [n][tap_stride][N]
  float* tap_spec;
  int tap_stride;
};
cudaError_t launch_k1(const K1Params& p, cudaStream_t stream);
// the 16 kHz tensor-core shape (k1_dft16.cu): the transform as a tcgen05 GEMM, one persistent CTA per SM
bool k1_dft16_supported(const K1Params& p);
bool k1_uses_dft16(const Geometry& g);
cudaError_t k1_dft16_prepare();  // per device, once: the DFT matrix
cudaError_t launch_k1_dft16(const K1Params& p, cudaStream_t stream);

// ---- K2/K3: recurrences, hysteresis, tension, speed ------------------------
struct K2Params {
  Geometry g;
  StreamState st;
  int n_streams;
  const int32_t* counts;
  long long frames;              // visible prefix of this write
  long long done;                // prefix already handled
  const float2* feat;            // from K1, [n][feat_stride]
  int feat_stride;
  int max_new_frames;
  float* speeds;                 // out: [n][speeds_stride]
  int speeds_stride;
  const float* override_speeds;  // optional [n][override_stride], indexed by tension frame
  long long override_stride;
  // optional taps, rows indexed by the j-th new tension of this write
  float* tap_features;           // [n][max_new_frames][15]
  float* tap_tension;            // [n][max_new_frames]
  float* tap_speed;              // [n][max_new_frames]
  float* tap_energy;             // [n][max_new_frames] by new analysis frame
};
cudaError_t launch_k2(const K2Params& p, cudaStream_t stream);

// ---- K4: Sonic AMDF pitch search + overlap-add ----------------------------
struct K4Params {
  Geometry g;
  StreamState st;
  int n_streams;
  const int16_t* hist;
  long long hist_stride;
  const int16_t* in;
  long long in_stride_frames;
  const int32_t* counts;
  long long frames;              // visible prefix of this write
  long long done;                // prefix already handled
  const float* speeds;           // from K2
  int speeds_stride;
  int flush;                     // 1: this launch is sonicFlushStream
  const int32_t* flush_mask;     // flush only: streams with a zero entry are left alone (null: all)
  int16_t* out;                  // [n][out_capacity][C]
  long long out_capacity;        // sample frames per stream
  int threads_per_stream;
  int buf_frames;                // shared-memory window, filled by the launcher
  // lane assignment of the AMDF searches, filled by the launcher (k4_sonic.cu):
  // per lane cGi + 1 (0 = idle) | cSub << 8 | cG << 16, then the per-kernel scalars
  unsigned lane_map[128];
  int c_max_g, f_g, f_per_round;
};
cudaError_t launch_k4(const K4Params& p, cudaStream_t stream);
// the pipelined shape (k4_splice.cu): chain / filler / output warps per stream
bool k4_splice_supported(const K4Params& p);
cudaError_t launch_k4_splice(const K4Params& p, cudaStream_t stream);
// the 16 kHz mono chain shape (k4_chain16.cu): one warp per stream, position-independent
// windows prefetched by TMA bulk copies, output fused into the next search
bool k4_chain16_supported(const K4Params& p);
cudaError_t launch_k4_chain16(const K4Params& p, cudaStream_t stream);

// ---- bookkeeping after a write: carry the input tail, advance totals ------
struct TailParams {
  Geometry g;
  StreamState st;
  int n_streams;
  const int16_t* hist_src;       // current history buffer
  int16_t* hist_dst;             // the other one
  long long hist_stride;
  const int16_t* in;
  long long in_stride_frames;
  const int32_t* counts;
  long long frames;
};
cudaError_t launch_tail(const TailParams& p, cudaStream_t stream);

}  // namespace speedy
