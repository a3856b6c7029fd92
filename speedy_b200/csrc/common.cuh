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
  int time_base;     // at_time of window 0: 1 through the shim (soniclib.c:296), 0 when frames are
                     // handed to speedyAddData directly (the white-box hook, speedy_test.cc:912)
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

// One mono value (channel average with C integer division, soniclib.c:271-274) of an
// interleaved frame; raw16 (optional) receives the frame's samples too.
template <typename T>
__device__ __forceinline__ void stage_frame(const int16_t* p, int C, T* dst, short* raw16) {
  if (C == 1) {
    *dst = (T)p[0];
  } else {
    int sum = 0;
    for (int c = 0; c < C; c++) {
      const int x = p[c];
      sum += x;
      if (raw16) raw16[c] = (short)x;
    }
    *dst = (T)(sum / C);
  }
}

// Frames [a, b) of one contiguous source (ptr[0] is the first sample of frame f0) into
// dst[f - start]: 16-byte loads (8 int16, four in flight per thread) over the part
// that is 16-byte aligned in the source, one frame at a time at its ragged ends.
// INFLIGHT: 16-byte loads a thread issues before it consumes the first (a lone warp on a latency-bound
// chain asks for 16: a whole 4096-frame window in one round trip to HBM instead of four).
template <int THREADS, typename T, int INFLIGHT = 4>
__device__ __forceinline__ void stage_span(const int16_t* ptr, long long f0, int C, long long a, long long b,
                                           long long start, T* dst, short* raw16, int tid) {
  if (b <= a) return;
  long long v0 = a, v1 = a;  // frames served by the vector path: [v0, v1)
  if ((C == 1 || C == 2) && b - a >= 64 && ((reinterpret_cast<size_t>(ptr) & (C == 2 ? 3 : 1)) == 0)) {
    const int fpv = 8 / C;  // frames per 16-byte vector
    const long long a0 = (long long)((reinterpret_cast<size_t>(ptr) >> 1) & 7) / C;  // frame offset of ptr in its vector
    const long long g0 = a - f0 + a0, g1 = b - f0 + a0;
    const long long c0 = (g0 + fpv - 1) / fpv, c1 = g1 / fpv;
    if (c1 > c0) {
      const long long fv0 = f0 + c0 * fpv - a0;  // first frame of vector c0
      const int4* vp = reinterpret_cast<const int4*>(ptr + (fv0 - f0) * C);
      const int nvec = (int)(c1 - c0);
      T* d = dst + (fv0 - start);
      const bool d_aligned = (reinterpret_cast<size_t>(d) & 15) == 0;
      short* r = raw16 ? raw16 + (fv0 - start) * C : nullptr;
      for (int vb = tid; vb < nvec; vb += INFLIGHT * THREADS) {
        int4 q4[INFLIGHT];
#pragma unroll
        for (int u = 0; u < INFLIGHT; u++) {  // INFLIGHT 16-byte loads in flight per thread
          const int v = vb + u * THREADS;
          q4[u] = v < nvec ? vp[v] : make_int4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < INFLIGHT; u++) {
          const int v = vb + u * THREADS;
          if (v >= nvec) continue;
          const int w[4] = {q4[u].x, q4[u].y, q4[u].z, q4[u].w};
          if (C == 1) {
            // 16-byte stores when the destination of a vector is 16-byte aligned (scalar stores
            // of a widened vector are 8-way bank conflicts: lanes are 32 bytes apart)
            if (sizeof(T) == 2 && d_aligned) {
              *reinterpret_cast<int4*>(d + v * 8) = q4[u];
            } else if (sizeof(T) == 4 && d_aligned) {
              int4* d4 = reinterpret_cast<int4*>(d + v * 8);
              d4[0] = make_int4((int)(T)(short)(w[0] & 0xffff), (int)(T)(short)(w[0] >> 16), (int)(T)(short)(w[1] & 0xffff),
                                (int)(T)(short)(w[1] >> 16));
              d4[1] = make_int4((int)(T)(short)(w[2] & 0xffff), (int)(T)(short)(w[2] >> 16), (int)(T)(short)(w[3] & 0xffff),
                                (int)(T)(short)(w[3] >> 16));
            } else {
#pragma unroll
              for (int j = 0; j < 4; j++) {
                d[v * 8 + 2 * j] = (T)(short)(w[j] & 0xffff);
                d[v * 8 + 2 * j + 1] = (T)(short)(w[j] >> 16);
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 4; j++) {
              const int l = (short)(w[j] & 0xffff), h = (short)(w[j] >> 16);
              d[v * 4 + j] = (T)((l + h) / 2);
              if (r) {
                r[(v * 4 + j) * 2] = (short)l;
                r[(v * 4 + j) * 2 + 1] = (short)h;
              }
            }
          }
        }
      }
      v0 = fv0;
      v1 = fv0 + (long long)nvec * fpv;
    }
  }
  // the ragged ends (or everything): [a, v0) and [v1, b)
  const int n_head = (int)(v0 - a), n_tail = (int)(b - v1);
  for (int i = tid; i < n_head + n_tail; i += THREADS) {
    const long long f = i < n_head ? a + i : v1 + (i - n_head);
    stage_frame<T>(ptr + (f - f0) * C, C, dst + (f - start), raw16 ? raw16 + (f - start) * C : nullptr);
  }
}

// Stage `count` mono samples (channel average, soniclib.c:271-274) of frames
// [start, start + count) into shared memory, dst[i] = frame start + i.  Frames
// outside [src.hist_base, lim) read as 0.  The carried history and the caller's
// buffer are each one contiguous span.  raw16 (optional, channels > 1) receives the
// interleaved samples too.
template <int THREADS, typename T, int INFLIGHT = 4>
__device__ __forceinline__ void stage_mono(const Source& src, long long start, int count, long long lim,
                                           T* dst, short* raw16, int tid) {
  const int C = src.channels;
  if (lim > src.t_new) lim = src.t_new;
  const long long end = start + count;
  // [h0, h1) from the history, [i0, i1) from the caller's buffer, zeros elsewhere
  long long h0 = start > src.hist_base ? start : src.hist_base;
  long long h1 = end < src.t_old ? end : src.t_old;
  if (h1 > lim) h1 = lim;
  if (h1 < h0) h1 = h0;
  long long i0 = start > src.t_old ? start : src.t_old;
  long long i1 = end < lim ? end : lim;
  if (i1 < i0) i1 = i0;
  stage_span<THREADS, T>(src.hist, src.hist_base, C, h0, h1, start, dst, raw16, tid);
  if (src.in) stage_span<THREADS, T, INFLIGHT>(src.in, src.t_old, C, i0, i1, start, dst, raw16, tid);
  // zeros: before the history begins, and from the end of the data on
  const long long z0 = h0 < end ? h0 : end;                        // [start, z0)
  const long long z1 = (src.in && i1 > i0) ? i1 : (h1 > start ? h1 : start);  // [z1, end)
  const int nz0 = (int)(z0 - start), nz1 = (int)(end - (z1 < end ? z1 : end));
  for (int i = tid; i < nz0 + nz1; i += THREADS) {
    const long long f = i < nz0 ? start + i : z1 + (i - nz0);
    dst[f - start] = (T)0;
    if (raw16 && C > 1) {
      for (int c = 0; c < C; c++) raw16[(f - start) * C + c] = 0;
    }
  }
}

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
  float* ring_comp;       // [n][kRing] compressed energy by at_time & 31
  float* ring_energy;     // [n][kRing] frame energy
  float* ring_lsd;        // [n][kRing] raw local spectral difference
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

// Frames of a stream visible in a launch over the prefix (done, frames] of a write.
struct Range {
  long long t_old, t_done, t_new;
};
__device__ __forceinline__ Range write_range(const long long* total, const int32_t* counts, long long frames,
                                             long long done, int s) {
  Range r;
  r.t_old = total[s];
  const long long c = counts ? counts[s] : frames;
  r.t_done = r.t_old + (c < done ? c : done);
  r.t_new = r.t_old + (c < frames ? c : frames);
  return r;
}

__host__ __device__ inline int frames_analyzed(const Geometry& g, long long total) {
  if (total < g.window + 1) return 0;
  return (int)((total - g.window - 1) / g.step) + 1;
}

__host__ __device__ inline int tensions_ready(const Geometry& g, int analyzed) {
  int n = analyzed - g.future + g.time_base;
  return n > 0 ? n : 0;
}

}  // namespace speedy
