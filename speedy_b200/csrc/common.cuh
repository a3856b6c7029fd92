// Shared definitions for the speedy_b200 CUDA kernels (sm_100a).
//
// Everything on the device is addressed in ABSOLUTE sample-frame indices since
// stream creation, so the closed-form schedule of the reference's shim
// (soniclib.c:427-450, 246-373) can be evaluated independently per frame:
//   window k          = samples [k*S, k*S+W)           at_time = k+1
//   analysed after T  = (T - W - 1)/S + 1 windows      (soniclib.c:440-444)
//   tension r ready   when r + Future <= at_time       (speedy.c:755)
//   Sonic is fed      buffer r = samples [r*S, (r+1)*S) at speed[r]
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace speedy {

constexpr int kRing = 32;          // per-stream feature ring (>= Future+Past+1)
constexpr int kFeatureCount = 15;  // speedy.c:124

// Frame geometry of one sample rate (speedy.c:213-214, 335-338; speedy.h:136-146;
// upstream Sonic's SONIC_MIN/MAX_PITCH and SONIC_AMDF_FREQ).
struct Geometry {
  int rate;
  int channels;
  int window;        // W
  int fft;           // N = 2W
  int step;          // S
  int partial;       // P = W - S
  int future;        // F
  int past;          // B
  int min_period;
  int max_period;
  int max_required;  // 2 * max_period
  int skip;          // AMDF decimation
  int hist_frames;   // capacity of the carried-over input tail, sample frames
};

// Where a kernel reads input samples: the tail carried over from earlier writes
// ("hist", absolute frames [hist_base, t_old)) followed by the caller's buffer
// of this write (absolute frames [t_old, t_new)).
struct Source {
  const int16_t* hist;  // this stream's history buffer
  const int16_t* in;    // this stream's slice of the caller's buffer (or null)
  long long hist_base;
  long long t_old;
  long long t_new;
  int channels;

  __device__ __forceinline__ int raw(long long frame, int c) const {
    if (frame >= t_old) return in[(frame - t_old) * channels + c];
    return hist[(frame - hist_base) * channels + c];
  }
  // Mono value Speedy analyses: (sum over channels) / channels with C integer
  // division, soniclib.c:271-274.  Frames outside [hist_base, t_new) read 0.
  __device__ __forceinline__ int mono(long long frame) const {
    if (frame < hist_base || frame >= t_new) return 0;
    if (channels == 1) return raw(frame, 0);
    int sum = 0;
    for (int c = 0; c < channels; c++) sum += raw(frame, c);
    return sum / channels;
  }
};

// Per-stream device state, structure of arrays (index = stream).
struct StreamState {
  // totals
  long long* total;       // T: sample frames written
  int* status;            // SPEEDY_STATUS_* bits
  // parameters
  float* speed;           // R_g
  float* nonlinear;
  float* feedback;
  // analysis recurrences (speedy.c:166-171)
  float* lp_energy;
  float* lp_diff;
  float* cur_dur;
  float* des_dur;
  float* ring_comp;       // [kRing][n] compressed energy by at_time & 31
  float* ring_energy;     // [kRing][n] frame energy
  float* ring_lsd;        // [kRing][n] raw local spectral difference
  float* ring_lp;         // [kRing][n] energy low-pass (features tap)
  // Sonic state
  long long* sonic_head;  // absolute frame of the FIFO head
  long long* sonic_fed;   // absolute frame one past the FIFO tail
  int* prev_period;
  int* prev_min_diff;
  int* remaining_copy;
  float* sonic_speed;     // last speed handed to Sonic
  long long* out_total;   // output frames produced since creation
  int* out_count;         // output frames pending in the out buffer
  // carried input tail
  long long* hist_base;
};

__host__ __device__ inline int frames_analyzed(const Geometry& g, long long total) {
  if (total < g.window + 1) return 0;
  return (int)((total - g.window - 1) / g.step) + 1;
}

__host__ __device__ inline int tensions_ready(const Geometry& g, int analyzed) {
  int n = analyzed - g.future + 1;
  return n > 0 ? n : 0;
}

}  // namespace speedy
