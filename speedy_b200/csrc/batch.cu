// Batch context: device-resident state of N independent streams and the C-ABI
// entry points of include/speedy_b200.h (section 2).  Host logic only, plus the
// three bookkeeping kernels (input-tail carry, output read, synthetic input).
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <stdlib.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "../../include/speedy_b200.h"
#include "kernels.cuh"
#include "synth.h"

namespace speedy {

constexpr int kPipeEvents = 16;

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static thread_local std::string g_error;
static void set_error(const std::string& e) { g_error = e; }

#define CU_TRY(expr)                                                              \
  do {                                                                            \
    cudaError_t _e = (expr);                                                      \
    if (_e != cudaSuccess) {                                                      \
      set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));              \
      return 0;                                                                   \
    }                                                                             \
  } while (0)

// ---------------------------------------------------------------------------
// bookkeeping kernels
// ---------------------------------------------------------------------------

// After a write: keep the last hist_frames sample frames of every stream (the
// delayed audio Sonic has not consumed yet plus the analysis overlap), advance
// the stream total.  soniclib.c keeps the same data in its ring of 10 ms buffers
// (soniclib.c:186-233) and upstream Sonic in its input FIFO.
// count int16 elements, 16 bytes at a time where source and destination are equally aligned
__device__ __forceinline__ void copy_shorts(int16_t* d, const int16_t* src, long long total, int tid, int nthreads) {
  const size_t ms = reinterpret_cast<size_t>(src) & 15, md = reinterpret_cast<size_t>(d) & 15;
  if (ms == md) {
    long long head = ms ? (long long)((16 - ms) >> 1) : 0;
    if (head > total) head = total;
    const long long nv = (total - head) / 8;
    for (long long i = tid; i < head; i += nthreads) d[i] = src[i];
    const int4* sv = reinterpret_cast<const int4*>(src + head);
    int4* dv = reinterpret_cast<int4*>(d + head);
    for (long long i = tid; i < nv; i += nthreads) dv[i] = sv[i];
    for (long long i = head + nv * 8 + tid; i < total; i += nthreads) d[i] = src[i];
  } else {
    for (long long i = tid; i < total; i += nthreads) d[i] = src[i];
  }
}

__global__ void __launch_bounds__(128) tail_kernel(TailParams p) {
  const int s = blockIdx.x;
  const Geometry& g = p.g;
  const long long t_old = p.st.total[s];
  const long long t_new = t_old + (p.counts ? p.counts[s] : p.frames);
  const long long old_base = p.st.hist_base[s];
  long long nb = t_new - g.hist_frames;
  if (nb < old_base) nb = old_base;
  Source src;
  src.channels = g.channels;
  src.hist = p.hist_src + (size_t)s * p.hist_stride;
  src.in = p.in ? p.in + (size_t)s * p.in_stride_frames * g.channels : nullptr;
  src.hist_base = old_base;
  src.t_old = t_old;
  src.t_new = t_new;
  int16_t* dst = p.hist_dst + (size_t)s * p.hist_stride;
  // two contiguous pieces: what stays of the old tail, then the end of this write
  const int C = g.channels;
  const long long f0 = nb > t_old ? nb : t_old;  // first frame taken from the caller's buffer
  if (nb < t_old) copy_shorts(dst, src.hist + (nb - old_base) * C, (t_old - nb) * C, threadIdx.x, blockDim.x);
  if (t_new > f0) {
    copy_shorts(dst + (f0 - nb) * C, src.in + (f0 - t_old) * C, (t_new - f0) * C, threadIdx.x, blockDim.x);
  }
  __syncthreads();  // all reads of total/hist_base above are done
  if (threadIdx.x == 0) {
    if (p.st.sonic_head[s] < nb) atomicOr(&p.st.status[s], SPEEDY_STATUS_INPUT_OVERFLOW);
    p.st.total[s] = t_new;
    p.st.hist_base[s] = nb;
  }
}

cudaError_t launch_tail(const TailParams& p, cudaStream_t stream) {
  tail_kernel<<<p.n_streams, 128, 0, stream>>>(p);
  count_launch();
  return cudaGetLastError();
}

// sonicReadShortFromStream for every stream: copy the pending output.
__global__ void __launch_bounds__(256) read_copy_kernel(const int16_t* out, long long cap, int channels,
                                                        const int* pending, int16_t* dst,
                                                        long long dst_stride) {
  const int s = blockIdx.y;
  long long n = pending[s];
  if (n > dst_stride) n = dst_stride;
  const long long total = n * channels;
  const int16_t* src = out + (size_t)s * cap * channels;
  int16_t* d = dst + (size_t)s * dst_stride * channels;
  // 16-byte vectors when both rows are aligned
  const bool aligned = ((((size_t)src) | ((size_t)d)) & 15) == 0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (aligned) {
    const long long nv = total / 8;
    const int4* sv = reinterpret_cast<const int4*>(src);
    int4* dv = reinterpret_cast<int4*>(d);
    for (long long i = tid; i < nv; i += stride) dv[i] = sv[i];
    for (long long i = nv * 8 + tid; i < total; i += stride) d[i] = src[i];
  } else {
    for (long long i = tid; i < total; i += stride) d[i] = src[i];
  }
}

// The same with one warp per stream: many streams with little pending output each (the 10 ms
// streaming step: a grid of 32 x n blocks would be half a million nearly empty CTAs).
__global__ void __launch_bounds__(256) read_copy_warp_kernel(const int16_t* out, long long cap, int channels,
                                                             const int* pending, int16_t* dst, long long dst_stride,
                                                             int n) {
  const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (s >= n) return;
  const int lane = threadIdx.x & 31;
  long long c = pending[s];
  if (c > dst_stride) c = dst_stride;
  const long long total = c * channels;
  const int16_t* src = out + (size_t)s * cap * channels;
  int16_t* d = dst + (size_t)s * dst_stride * channels;
  if (((((size_t)src) | ((size_t)d)) & 15) == 0) {
    const long long nv = total / 8;
    const int4* sv = reinterpret_cast<const int4*>(src);
    int4* dv = reinterpret_cast<int4*>(d);
    for (long long i = lane; i < nv; i += 32) dv[i] = sv[i];
    for (long long i = nv * 8 + lane; i < total; i += 32) d[i] = src[i];
  } else {
    for (long long i = lane; i < total; i += 32) d[i] = src[i];
  }
}

// speedyBatchProcess: move what every stream produced since the last call,
// out[s][done[s] .. upto[s]), straight into the caller's pinned host buffer (a
// device-accessible pointer under unified addressing) with 16-byte stores where the
// two sides are equally aligned; done[s] advances to upto[s].
__global__ void __launch_bounds__(256) scatter_out_kernel(const int16_t* out, long long cap, int channels,
                                                          const int* upto, const int* done, int16_t* dst,
                                                          long long dst_stride, int n) {
  // A small grid (the copy is PCIe-bound; it must not crowd the compute kernels out of the SMs,
  // nor the host-to-device copies out of the link: see the launch): block b walks streams b,
  // b + gridDim.x, ...
  for (int s = blockIdx.x; s < n; s += gridDim.x) {
    long long a = done[s], b = upto[s];
    if (b > dst_stride) b = dst_stride;
    const int16_t* src = out + ((size_t)s * cap + a) * channels;
    int16_t* d = dst + ((size_t)s * dst_stride + a) * channels;
    const long long total = (b - a) * channels;
    if (total <= 0) continue;
    const size_t ms = reinterpret_cast<size_t>(src) & 15, md = reinterpret_cast<size_t>(d) & 15;
    if (ms == md) {
      long long head = ms ? (long long)((16 - ms) >> 1) : 0;
      if (head > total) head = total;
      const long long nv = (total - head) / 8;
      for (long long i = threadIdx.x; i < head; i += blockDim.x) d[i] = src[i];
      const int4* sv = reinterpret_cast<const int4*>(src + head);
      int4* dv = reinterpret_cast<int4*>(d + head);
      for (long long i = threadIdx.x; i < nv; i += blockDim.x) dv[i] = sv[i];
      for (long long i = head + nv * 8 + threadIdx.x; i < total; i += blockDim.x) d[i] = src[i];
    } else {
      for (long long i = threadIdx.x; i < total; i += blockDim.x) d[i] = src[i];
    }
  }
}

__global__ void scatter_advance_kernel(int n, const int* upto, int* done, long long dst_stride) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  long long b = upto[s];
  if (b > dst_stride) b = dst_stride;
  done[s] = (int)b;
}

__global__ void read_finish_kernel(int n, int* pending, int* status, int* counts, long long dst_stride) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  int c = pending[s];
  if (counts) {
    if (c > dst_stride) {
      atomicOr(&status[s], SPEEDY_STATUS_READ_TRUNCATED);
      c = (int)dst_stride;
    }
    counts[s] = c;
  }
  pending[s] = 0;
}

// Back to the just-created state for the streams with a non-zero mask entry (what reset_state
// does for all of them): a slot of the session pool being handed to a new session.
__global__ void reset_streams_kernel(int n, const int32_t* mask, StreamState s) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || mask[i] == 0) return;
  s.total[i] = 0;
  s.status[i] = 0;
  s.lp_energy[i] = 2.14204f;  // speedy.c:287-292
  s.lp_diff[i] = 123.837f;
  s.cur_dur[i] = 0.0f;
  s.des_dur[i] = 0.0f;
  for (int r = 0; r < kRing; r++) {
    s.ring_comp[(size_t)i * kRing + r] = 0.0f;
    s.ring_energy[(size_t)i * kRing + r] = 0.0f;
    s.ring_lsd[(size_t)i * kRing + r] = 0.0f;
  }
  s.sonic_head[i] = 0;
  s.sonic_fed[i] = 0;
  s.prev_period[i] = 0;
  s.prev_min_diff[i] = 0;
  s.remaining_copy[i] = 0;
  s.out_total[i] = 0;
  s.out_count[i] = 0;
  s.hist_base[i] = 0;
}

__global__ void __launch_bounds__(256) synth_kernel(int16_t* out, unsigned long long first_id, int rate,
                                                    int channels, long long frames) {
  const int s = blockIdx.y;
  int16_t* row = out + (size_t)s * frames * channels;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < frames; n += stride) {
    int32_t x = synth_mono(first_id + s, rate, n);
    for (int c = 0; c < channels; c++) {
      int32_t v = x;
      if (channels > 1) v = (c & 1) ? (x * 11) / 10 : (x * 9) / 10;
      if (v > 32767) v = 32767;
      if (v < -32768) v = -32768;
      row[n * channels + c] = (int16_t)v;
    }
  }
}

__global__ void fill_float_kernel(float* p, long long n, float v) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// ---------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------
static void make_geometry(int rate, int channels, int match_matlab, Geometry* g) {
  g->rate = rate;
  g->channels = channels;
  g->window = (int)(1.5 * rate / (float)100.0);  // speedy.c:213
  g->fft = 2 * g->window;                        // speedy.c:214
  g->step = (int)(rate / 100.0);                 // speedy.c:335-338
  g->partial = g->window - (g->window / g->step) * g->step;  // soniclib.c:410-411
  g->future = match_matlab ? 8 : 12;             // speedy.h:136-146
  g->past = match_matlab ? 12 : 8;
  g->min_period = rate / 400;                    // SONIC_MAX_PITCH
  g->max_period = rate / 65;                     // SONIC_MIN_PITCH
  g->max_required = 2 * g->max_period;
  g->skip = rate > 4000 ? rate / 4000 : 1;       // SONIC_AMDF_FREQ
  g->hist_frames = g->max_required + (g->future + 4) * g->step + g->window + 64;
  g->time_base = 1;  // soniclib.c:296
}

static int factorize(int n, int* f) {
  int c = 0;
  while (n % 4 == 0) { f[c++] = 4; n /= 4; }
  while (n % 2 == 0) { f[c++] = 2; n /= 2; }
  for (int p = 3; n > 1; p += 2) {
    while (n % p == 0) {
      if (c >= kMaxFactors) return -1;
      f[c++] = p;
      n /= p;
    }
  }
  return c;
}

// Does the mixed-radix spectrogram kernel take this window (k1_spectral.cu: even, and
// W / 2 a product of 2, 3, 5, 7, 11, 13)?
static bool mixed_radix_window(int window) {
  if (window % 2 != 0 || window / 2 < 4) return false;
  int rest = window / 2;
  const int pref[] = {2, 3, 5, 7, 11, 13};
  for (int r : pref) {
    while (rest % r == 0) rest /= r;
  }
  return rest == 1;
}

// Convolution length for the chirp-z kernel: a 7-smooth L >= n with few radix stages
// when taken as 8, 4, 5, 3, 2, 7 (cost ~ L x stages).
static int chirp_length(int n) {
  int best = 0;
  long long best_cost = 0;
  for (int L = n; L <= 2 * n; L++) {
    int rest = L, stages = 0;
    const int pref[] = {8, 4, 5, 3, 2, 7};
    for (int r : pref) {
      while (rest % r == 0) {
        rest /= r;
        stages++;
      }
    }
    if (rest != 1 || stages > kMaxFactors) continue;
    const long long cost = (long long)L * stages;
    if (!best || cost < best_cost) {
      best = L;
      best_cost = cost;
    }
  }
  return best;
}

}  // namespace speedy

using namespace speedy;

struct speedyBatchStruct {
  speedyBatchConfig cfg;
  Geometry g;
  int n;
  cudaStream_t own_stream;
  StreamState st;
  std::vector<void*> allocs;
  // tables
  float* d_window;
  float2* d_tw_n;
  float2* d_tw_half;
  int n_factors;
  int factors[kMaxFactors];
  // chirp-z tables (windows the mixed-radix kernel cannot factor, e.g. 44.1 kHz)
  int bl_L;
  float2* d_bl_tw;
  float2* d_bl_B;
  float2* d_bl_chirp;
  // history (ping-pong)
  int16_t* d_hist[2];
  int hist_cur;
  long long hist_stride;
  // scratch
  int max_new_frames;
  float2* d_feat;
  float* d_speeds;
  int speeds_stride;
  // output
  int16_t* d_out;
  long long out_capacity;
  // override speeds
  float* d_override;           // current rows (null: none)
  float* d_override_buf;       // the allocation behind them, reused while it is large enough
  size_t override_capacity;    // floats
  long long override_stride;
  // taps
  float *d_tap_spec, *d_tap_energy, *d_tap_features, *d_tap_tension, *d_tap_speed;
  long long last_frames;           // the last write, for GetTaps
  const int32_t* last_d_counts;
  // host staging for the host-pointer entry points
  int16_t* d_stage;
  long long stage_frames;
  int32_t* d_counts_stage;
  int32_t* d_mask_stage;   // per-stream masks of FlushStreams / ResetStreams (allocated on first use)
  int32_t* h_pinned_counts;
  // speedyBatchProcess: double-buffered input chunks and the copy streams
  int pipe_ready;
  cudaStream_t s_h2d, s_d2h;
  int* d_snap[3];       // out_count snapshots per chunk
  int* d_done;          // frames already delivered to the host
  // per-kernel timing (speedyBatchSetProfiling): (start, stop) event pairs
  int profiling;
  std::vector<cudaEvent_t> prof_events;
  std::vector<std::pair<int, int>> prof_marks;  // (slot, is_begin), parallel to the events in use
  int prof_used;
  // internal stream the resynthesis kernel runs on when a write is split, fork/join events
  cudaStream_t s_sonic;
  cudaEvent_t ev_pipe[16];
};

namespace {

template <typename T>
bool dev_alloc(speedyBatch b, T** p, size_t count) {
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, count * sizeof(T) > 0 ? count * sizeof(T) : 16);
  if (e != cudaSuccess) {
    set_error(std::string("cudaMalloc: ") + cudaGetErrorString(e));
    return false;
  }
  b->allocs.push_back(q);
  *p = reinterpret_cast<T*>(q);
  return true;
}

cudaStream_t pick_stream(speedyBatch b, void* s) { return s ? (cudaStream_t)s : b->own_stream; }

int fill_floats(speedyBatch b, float* d, const float* values, float uniform, cudaStream_t st) {
  // parameters change between launches, never under one: kernels queued on a caller's stream finish first
  CU_TRY(cudaDeviceSynchronize());
  if (values) {
    CU_TRY(cudaMemcpyAsync(d, values, sizeof(float) * b->n, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaStreamSynchronize(st));
  } else {
    fill_float_kernel<<<(b->n + 255) / 256, 256, 0, st>>>(d, b->n, uniform);
    CU_TRY(cudaGetLastError());
  }
  return 1;
}

int reset_state(speedyBatch b, cudaStream_t st) {
  const int n = b->n;
  StreamState& s = b->st;
  CU_TRY(cudaMemsetAsync(s.total, 0, sizeof(long long) * n, st));
  CU_TRY(cudaMemsetAsync(s.status, 0, sizeof(int) * n, st));
  // speedy.c:287-292: both one-pole filters start at the long-term means
  fill_float_kernel<<<(n + 255) / 256, 256, 0, st>>>(s.lp_energy, n, 2.14204f);
  fill_float_kernel<<<(n + 255) / 256, 256, 0, st>>>(s.lp_diff, n, 123.837f);
  CU_TRY(cudaMemsetAsync(s.cur_dur, 0, sizeof(float) * n, st));
  CU_TRY(cudaMemsetAsync(s.des_dur, 0, sizeof(float) * n, st));
  CU_TRY(cudaMemsetAsync(s.ring_comp, 0, sizeof(float) * n * kRing, st));
  CU_TRY(cudaMemsetAsync(s.ring_energy, 0, sizeof(float) * n * kRing, st));
  CU_TRY(cudaMemsetAsync(s.ring_lsd, 0, sizeof(float) * n * kRing, st));
  CU_TRY(cudaMemsetAsync(s.sonic_head, 0, sizeof(long long) * n, st));
  CU_TRY(cudaMemsetAsync(s.sonic_fed, 0, sizeof(long long) * n, st));
  CU_TRY(cudaMemsetAsync(s.prev_period, 0, sizeof(int) * n, st));
  CU_TRY(cudaMemsetAsync(s.prev_min_diff, 0, sizeof(int) * n, st));
  CU_TRY(cudaMemsetAsync(s.remaining_copy, 0, sizeof(int) * n, st));
  CU_TRY(cudaMemsetAsync(s.out_total, 0, sizeof(long long) * n, st));
  CU_TRY(cudaMemsetAsync(s.out_count, 0, sizeof(int) * n, st));
  CU_TRY(cudaMemsetAsync(s.hist_base, 0, sizeof(long long) * n, st));
  CU_TRY(cudaGetLastError());
  b->hist_cur = 0;
  return 1;
}

}  // namespace

extern "C" {

const char* speedyBatchLastError(void) { return g_error.c_str(); }

int64_t speedyBatchKernelLaunches(void) { return g_launches.load(); }

const char* speedyBatchBuildInfo(void) {
  return "speedy_b200 sm_100a: k1_dft16 (16 kHz spectrogram as a tcgen05 GEMM, TMEM accumulator, TMA-fed) | "
         "k1_spectral_480<4 warps> (radix-8 x radix-15 real FFT, SPEEDY_K1_TC=0) | "
         "k1_spectral_mixed<128> (packed half-length Stockham FFT, any even window) | k1_spectral_bluestein<128> (chirp-z, "
         "prime windows) | "
         "k2_tension | k4_sonic<1|2|4 warps per stream, mono specialisation> | tail | read | synth";
}

void speedyBatchDefaultConfig(speedyBatchConfig* cfg) {
  memset(cfg, 0, sizeof(*cfg));
  cfg->sample_rate = 16000;
  cfg->num_channels = 1;
  cfg->num_streams = 1;
  cfg->match_matlab = 0;
  cfg->speed = 1.0f;              // soniclib.c:114
  cfg->nonlinear_factor = 0.0f;   // soniclib.c:117
  cfg->feedback_strength = 0.1f;  // soniclib.c:122
  cfg->device = 0;
  cfg->max_write_frames = 16000;
  cfg->out_capacity = 0;
  cfg->taps = 0;
  cfg->threads_per_stream = 0;
}

int speedyBatchFrameGeometry(int sample_rate, int* window, int* fft, int* step) {
  if (sample_rate < 800) return 0;
  Geometry g;
  make_geometry(sample_rate, 1, 0, &g);
  if (window) *window = g.window;
  if (fft) *fft = g.fft;
  if (step) *step = g.step;
  return 1;
}

int speedyBatchNumStreams(speedyBatch b) { return b ? b->n : 0; }

void speedyBatchDestroy(speedyBatch b) {
  if (!b) return;
  cudaSetDevice(b->cfg.device);
  if (b->s_h2d) cudaStreamDestroy(b->s_h2d);
  if (b->s_d2h) cudaStreamDestroy(b->s_d2h);
  for (void* p : b->allocs) cudaFree(p);
  if (b->h_pinned_counts) cudaFreeHost(b->h_pinned_counts);
  for (cudaEvent_t e : b->prof_events) cudaEventDestroy(e);
  for (int i = 0; i < kPipeEvents; i++) if (b->ev_pipe[i]) cudaEventDestroy(b->ev_pipe[i]);
  if (b->s_sonic) cudaStreamDestroy(b->s_sonic);
  if (b->own_stream) cudaStreamDestroy(b->own_stream);
  delete b;
}

speedyBatch speedyBatchCreate(const speedyBatchConfig* cfg) {
  if (!cfg || cfg->num_streams < 1 || cfg->num_channels < 1 || cfg->sample_rate < 800 ||
      cfg->max_write_frames < 1) {
    set_error("speedyBatchCreate: bad configuration");
    return nullptr;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || cfg->device >= ndev) {
    set_error("speedyBatchCreate: no CUDA device (this library has no CPU fallback)");
    return nullptr;
  }
  if (cudaSetDevice(cfg->device) != cudaSuccess) {
    set_error("speedyBatchCreate: cudaSetDevice failed");
    return nullptr;
  }
  speedyBatch b = new speedyBatchStruct();
  b->cfg = *cfg;
  b->n = cfg->num_streams;
  b->own_stream = nullptr;
  b->h_pinned_counts = nullptr;
  b->d_stage = nullptr;
  b->stage_frames = 0;
  b->d_counts_stage = nullptr;
  b->d_mask_stage = nullptr;
  b->d_override = nullptr;
  b->d_override_buf = nullptr;
  b->override_capacity = 0;
  b->override_stride = 0;
  b->pipe_ready = 0;
  b->s_h2d = b->s_d2h = nullptr;
  b->d_snap[0] = b->d_snap[1] = b->d_snap[2] = nullptr;
  b->d_done = nullptr;
  b->profiling = 0;
  b->prof_used = 0;
  b->s_sonic = nullptr;
  for (int i = 0; i < kPipeEvents; i++) b->ev_pipe[i] = nullptr;
  b->d_tap_spec = b->d_tap_energy = b->d_tap_features = b->d_tap_tension = b->d_tap_speed = nullptr;
  make_geometry(cfg->sample_rate, cfg->num_channels, cfg->match_matlab, &b->g);
  if (k1_uses_dft16(b->g) && cfg->analysis_frame_step <= 0 && k1_dft16_prepare() != cudaSuccess) {
    set_error("speedyBatchCreate: the DFT matrix of the tensor-core spectrogram kernel could not be uploaded");
    delete b;
    return nullptr;
  }
  if (cfg->analysis_frame_step > 0) {  // white-box hook: explicit analysis frames (speedy_b200.h)
    b->g.step = cfg->analysis_frame_step;
    b->g.partial = b->g.window - (b->g.window / b->g.step) * b->g.step;
    b->g.hist_frames = b->g.max_required + (b->g.future + 4) * b->g.step + b->g.window + 64;
    b->g.time_base = 0;  // frames numbered as the test numbers its speedyAddData calls (speedy_test.cc:912)
  }
  const Geometry& g = b->g;
  b->n_factors = factorize(g.fft, b->factors);
  bool ok = b->n_factors > 0;
  if (!ok) set_error("speedyBatchCreate: unsupported FFT size");
  for (int i = 0; ok && i < b->n_factors; i++) {
    if (b->factors[i] > 2047) {  // the generic kernel loops over the radix: any prime works, slowly
      ok = false;
      set_error("speedyBatchCreate: FFT size has a large prime factor");
    }
  }
  if (ok && cudaStreamCreateWithFlags(&b->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
    ok = false;
    set_error("speedyBatchCreate: cudaStreamCreate failed");
  }
  const int n = b->n;
  StreamState& s = b->st;
  ok = ok && dev_alloc(b, &s.total, n) && dev_alloc(b, &s.status, n) && dev_alloc(b, &s.speed, n) &&
       dev_alloc(b, &s.nonlinear, n) && dev_alloc(b, &s.feedback, n) && dev_alloc(b, &s.lp_energy, n) &&
       dev_alloc(b, &s.lp_diff, n) && dev_alloc(b, &s.cur_dur, n) && dev_alloc(b, &s.des_dur, n) &&
       dev_alloc(b, &s.ring_comp, (size_t)n * kRing) && dev_alloc(b, &s.ring_energy, (size_t)n * kRing) &&
       dev_alloc(b, &s.ring_lsd, (size_t)n * kRing) &&
       dev_alloc(b, &s.sonic_head, n) && dev_alloc(b, &s.sonic_fed, n) && dev_alloc(b, &s.prev_period, n) &&
       dev_alloc(b, &s.prev_min_diff, n) && dev_alloc(b, &s.remaining_copy, n) &&
       dev_alloc(b, &s.sonic_speed, n) && dev_alloc(b, &s.out_total, n) && dev_alloc(b, &s.out_count, n) &&
       dev_alloc(b, &s.hist_base, n);
  // tables (speedy.c:256-258 Hamming in double, stored float; FFT roots likewise)
  if (ok) {
    std::vector<float> w(g.window);
    for (int i = 0; i < g.window; i++) w[i] = (float)(0.54 - 0.46 * cos(2 * M_PI * i / (g.window - 1.0)));
    std::vector<float2> tn(g.fft), th(g.fft / 2);
    for (int k = 0; k < g.fft; k++) {
      double ph = -2.0 * M_PI * k / g.fft;
      tn[k] = make_float2((float)cos(ph), (float)sin(ph));
    }
    for (int k = 0; k < g.fft / 2; k++) {
      double ph = -2.0 * M_PI * k / (g.fft / 2);
      th[k] = make_float2((float)cos(ph), (float)sin(ph));
    }
    ok = dev_alloc(b, &b->d_window, g.window) && dev_alloc(b, &b->d_tw_n, g.fft) &&
         dev_alloc(b, &b->d_tw_half, g.fft / 2);
    if (ok) {
      cudaMemcpy(b->d_window, w.data(), sizeof(float) * g.window, cudaMemcpyHostToDevice);
      cudaMemcpy(b->d_tw_n, tn.data(), sizeof(float2) * g.fft, cudaMemcpyHostToDevice);
      cudaMemcpy(b->d_tw_half, th.data(), sizeof(float2) * (g.fft / 2), cudaMemcpyHostToDevice);
    }
  }
  b->bl_L = 0;
  b->d_bl_tw = b->d_bl_B = b->d_bl_chirp = nullptr;
  if (ok && g.fft != 480 && !mixed_radix_window(g.window)) {
    // chirp-z tables, in double: c[n] = e^{-i pi n^2 / N}, the filter e^{+i pi m^2 / N} for
    // m = -(W-1) .. W laid out modulo L, its L-point transform divided by L
    const int N = g.fft, W = g.window, L = chirp_length(N);
    if (L > 0) {
      std::vector<float2> tw(L), B(L), ch(W);
      std::vector<double> br(L, 0.0), bi(L, 0.0);
      for (int k = 0; k < L; k++) {
        const double ph = -2.0 * M_PI * k / L;
        tw[k] = make_float2((float)cos(ph), (float)sin(ph));
      }
      for (int n = 0; n < W; n++) {
        const double ph = -M_PI * (double)(((long long)n * n) % (2LL * N)) / N;
        ch[n] = make_float2((float)cos(ph), (float)sin(ph));
      }
      for (int m = -(W - 1); m <= W; m++) {
        const double ph = M_PI * (double)(((long long)m * m) % (2LL * N)) / N;
        const int at = ((m % L) + L) % L;
        br[at] = cos(ph);
        bi[at] = sin(ph);
      }
      std::vector<double> cr(L), ci(L);
      for (int k = 0; k < L; k++) {
        const double ph = -2.0 * M_PI * k / L;
        cr[k] = cos(ph);
        ci[k] = sin(ph);
      }
      for (int k = 0; k < L; k++) {
        double sr = 0.0, si = 0.0;
        for (int j = 0; j < L; j++) {
          if (br[j] == 0.0 && bi[j] == 0.0) continue;
          const int e = (int)(((long long)j * k) % L);
          sr += br[j] * cr[e] - bi[j] * ci[e];
          si += br[j] * ci[e] + bi[j] * cr[e];
        }
        B[k] = make_float2((float)(sr / L), (float)(si / L));
      }
      ok = dev_alloc(b, &b->d_bl_tw, L) && dev_alloc(b, &b->d_bl_B, L) && dev_alloc(b, &b->d_bl_chirp, W);
      if (ok) {
        cudaMemcpy(b->d_bl_tw, tw.data(), sizeof(float2) * L, cudaMemcpyHostToDevice);
        cudaMemcpy(b->d_bl_B, B.data(), sizeof(float2) * L, cudaMemcpyHostToDevice);
        cudaMemcpy(b->d_bl_chirp, ch.data(), sizeof(float2) * W, cudaMemcpyHostToDevice);
        b->bl_L = L;
      }
    }
  }
  // history, scratch, output
  b->hist_stride = (long long)g.hist_frames * g.channels;
  b->max_new_frames = (int)(cfg->max_write_frames / g.step) + 2;
  b->speeds_stride = b->max_new_frames;
  b->out_capacity = cfg->out_capacity > 0 ? cfg->out_capacity
                                          : cfg->max_write_frames + 4 * (long long)g.max_required;
  ok = ok && dev_alloc(b, &b->d_hist[0], (size_t)n * b->hist_stride) &&
       dev_alloc(b, &b->d_hist[1], (size_t)n * b->hist_stride) &&
       dev_alloc(b, &b->d_feat, (size_t)n * b->max_new_frames) &&
       dev_alloc(b, &b->d_speeds, (size_t)n * b->speeds_stride) &&
       dev_alloc(b, &b->d_out, (size_t)n * b->out_capacity * g.channels) &&
       dev_alloc(b, &b->d_counts_stage, n);
  const size_t tap_rows = (size_t)n * b->max_new_frames;
  if (ok && (cfg->taps & SPEEDY_TAP_SPECTROGRAM)) ok = dev_alloc(b, &b->d_tap_spec, tap_rows * g.fft);
  if (ok && (cfg->taps & SPEEDY_TAP_ENERGY)) ok = dev_alloc(b, &b->d_tap_energy, tap_rows);
  if (ok && (cfg->taps & SPEEDY_TAP_FEATURES)) ok = dev_alloc(b, &b->d_tap_features, tap_rows * kFeatureCount);
  if (ok && (cfg->taps & SPEEDY_TAP_TENSION)) ok = dev_alloc(b, &b->d_tap_tension, tap_rows);
  if (ok && (cfg->taps & SPEEDY_TAP_SPEED)) ok = dev_alloc(b, &b->d_tap_speed, tap_rows);
  if (ok && cudaMallocHost((void**)&b->h_pinned_counts, sizeof(int32_t) * 2 * n) != cudaSuccess) {
    ok = false;
    set_error("speedyBatchCreate: cudaMallocHost failed");
  }
  if (ok) {
    ok = reset_state(b, b->own_stream) && fill_floats(b, s.speed, nullptr, cfg->speed, b->own_stream) &&
         fill_floats(b, s.sonic_speed, nullptr, cfg->speed, b->own_stream) &&
         fill_floats(b, s.nonlinear, nullptr, cfg->nonlinear_factor, b->own_stream) &&
         fill_floats(b, s.feedback, nullptr, cfg->feedback_strength, b->own_stream);
    if (ok && cudaStreamSynchronize(b->own_stream) != cudaSuccess) {
      ok = false;
      set_error("speedyBatchCreate: initialisation kernels failed");
    }
  }
  b->last_frames = 0;
  b->last_d_counts = nullptr;
  if (!ok) {
    speedyBatchDestroy(b);
    return nullptr;
  }
  return b;
}

int speedyBatchReset(speedyBatch b, void* cuda_stream) {
  if (!b) return 0;
  cudaStream_t st = pick_stream(b, cuda_stream);
  CU_TRY(cudaSetDevice(b->cfg.device));
  if (!reset_state(b, st)) return 0;
  // sonicSetSpeed also hands the speed to Sonic (soniclib.c:181-182)
  CU_TRY(cudaMemcpyAsync(b->st.sonic_speed, b->st.speed, sizeof(float) * b->n, cudaMemcpyDeviceToDevice, st));
  b->last_frames = 0;
  b->last_d_counts = nullptr;
  return 1;
}

int speedyBatchSetSpeed(speedyBatch b, const float* values, float uniform) {
  if (!b) return 0;
  CU_TRY(cudaSetDevice(b->cfg.device));
  // soniclib.c:177-183: remembered as R_g and handed to Sonic at once
  return fill_floats(b, b->st.speed, values, uniform, b->own_stream) &&
         fill_floats(b, b->st.sonic_speed, values, uniform, b->own_stream) &&
         cudaStreamSynchronize(b->own_stream) == cudaSuccess;
}

int speedyBatchSetNonlinear(speedyBatch b, const float* values, float uniform) {
  if (!b) return 0;
  CU_TRY(cudaSetDevice(b->cfg.device));
  return fill_floats(b, b->st.nonlinear, values, uniform, b->own_stream) &&
         cudaStreamSynchronize(b->own_stream) == cudaSuccess;
}

int speedyBatchSetFeedback(speedyBatch b, const float* values, float uniform) {
  if (!b) return 0;
  CU_TRY(cudaSetDevice(b->cfg.device));
  return fill_floats(b, b->st.feedback, values, uniform, b->own_stream) &&
         cudaStreamSynchronize(b->own_stream) == cudaSuccess;
}

int speedyBatchOverrideSpeeds(speedyBatch b, const float* speeds, int64_t frames_per_stream) {
  if (!b) return 0;
  CU_TRY(cudaSetDevice(b->cfg.device));
  if (!speeds) {
    b->d_override = nullptr;
    b->override_stride = 0;
    return 1;
  }
  if (frames_per_stream < 1) {
    set_error("speedyBatchOverrideSpeeds: frames_per_stream must be positive");
    return 0;
  }
  const size_t need = (size_t)b->n * frames_per_stream;
  // kernels of earlier calls may still be reading the previous rows
  CU_TRY(cudaDeviceSynchronize());
  if (need > b->override_capacity) {
    if (b->d_override_buf) {
      cudaFree(b->d_override_buf);
      b->allocs.erase(std::remove(b->allocs.begin(), b->allocs.end(), (void*)b->d_override_buf), b->allocs.end());
      b->d_override_buf = nullptr;
      b->override_capacity = 0;
    }
    float* d = nullptr;
    if (!dev_alloc(b, &d, need)) return 0;
    b->d_override_buf = d;
    b->override_capacity = need;
  }
  CU_TRY(cudaMemcpy(b->d_override_buf, speeds, sizeof(float) * need, cudaMemcpyHostToDevice));
  b->d_override = b->d_override_buf;
  b->override_stride = frames_per_stream;
  return 1;
}

// ---- one write, possibly as several launches over growing prefixes ----------
namespace {

struct WriteCall {
  speedyBatch b;
  const int16_t* d_in;
  long long stride_frames, frames;
  const int32_t* d_counts;
};

// per-kernel timing: accumulate event pairs (start, stop, slot) while profiling
void prof_mark(speedyBatch b, cudaStream_t st, int slot, bool begin) {
  if (!b->profiling) return;
  cudaEvent_t e = nullptr;
  if (b->prof_used < (int)b->prof_events.size()) {
    e = b->prof_events[b->prof_used];
  } else {
    if (cudaEventCreate(&e) != cudaSuccess) return;
    b->prof_events.push_back(e);
  }
  b->prof_used++;
  cudaEventRecord(e, st);
  b->prof_marks.push_back({slot, begin ? 1 : 0});
}

// analysis (K1 + K2) of the prefix (done, prefix] on stream sa
int launch_analysis(const WriteCall& w, long long done, long long prefix, cudaStream_t sa) {
  speedyBatch b = w.b;
  const Geometry& g = b->g;
  K1Params k1;
  memset(&k1, 0, sizeof(k1));
  k1.g = g;
  k1.st = b->st;
  k1.n_streams = b->n;
  k1.hist = b->d_hist[b->hist_cur];
  k1.hist_stride = b->hist_stride;
  k1.in = w.d_in;
  k1.in_stride_frames = w.stride_frames;
  k1.counts = w.d_counts;
  k1.frames = prefix;
  k1.done = done;
  k1.feat = b->d_feat;
  k1.feat_stride = b->max_new_frames;
  // at most this many analysis windows can complete in this launch
  k1.max_new_frames = (int)((prefix - done) / g.step) + 2;
  if (k1.max_new_frames > b->max_new_frames) k1.max_new_frames = b->max_new_frames;
  k1.window = b->d_window;
  k1.tw_n = b->d_tw_n;
  k1.tw_half = b->d_tw_half;
  k1.n_factors = b->n_factors;
  memcpy(k1.factors, b->factors, sizeof(k1.factors));
  k1.bl_L = b->bl_L;
  k1.bl_tw = b->d_bl_tw;
  k1.bl_B = b->d_bl_B;
  k1.bl_chirp = b->d_bl_chirp;
  k1.tap_spec = b->d_tap_spec;
  k1.tap_stride = b->max_new_frames;
  prof_mark(b, sa, 0, true);
  CU_TRY(launch_k1(k1, sa));  // streams with nonlinear factor 0 skip themselves
  prof_mark(b, sa, 0, false);

  K2Params k2;
  memset(&k2, 0, sizeof(k2));
  k2.g = g;
  k2.st = b->st;
  k2.n_streams = b->n;
  k2.counts = w.d_counts;
  k2.frames = prefix;
  k2.done = done;
  k2.feat = b->d_feat;
  k2.feat_stride = b->max_new_frames;
  k2.max_new_frames = b->max_new_frames;
  k2.speeds = b->d_speeds;
  k2.speeds_stride = b->speeds_stride;
  k2.override_speeds = b->d_override;
  k2.override_stride = b->override_stride;
  k2.tap_features = b->d_tap_features;
  k2.tap_tension = b->d_tap_tension;
  k2.tap_speed = b->d_tap_speed;
  k2.tap_energy = b->d_tap_energy;
  prof_mark(b, sa, 1, true);
  CU_TRY(launch_k2(k2, sa));
  prof_mark(b, sa, 1, false);
  return 1;
}

// resynthesis (K4) of the prefix (done, prefix] on stream ss
int launch_resynthesis(const WriteCall& w, long long done, long long prefix, cudaStream_t ss) {
  speedyBatch b = w.b;
  K4Params k4;
  memset(&k4, 0, sizeof(k4));
  k4.g = b->g;
  k4.st = b->st;
  k4.n_streams = b->n;
  k4.hist = b->d_hist[b->hist_cur];
  k4.hist_stride = b->hist_stride;
  k4.in = w.d_in;
  k4.in_stride_frames = w.stride_frames;
  k4.counts = w.d_counts;
  k4.frames = prefix;
  k4.done = done;
  k4.speeds = b->d_speeds;
  k4.speeds_stride = b->speeds_stride;
  k4.flush = 0;
  k4.out = b->d_out;
  k4.out_capacity = b->out_capacity;
  k4.threads_per_stream = b->cfg.threads_per_stream;
  prof_mark(b, ss, 2, true);
  CU_TRY(launch_k4(k4, ss));
  prof_mark(b, ss, 2, false);
  return 1;
}

// after the last prefix: carry the input tail, advance the totals
int launch_write_tail(const WriteCall& w, cudaStream_t st) {
  speedyBatch b = w.b;
  TailParams tp;
  memset(&tp, 0, sizeof(tp));
  tp.g = b->g;
  tp.st = b->st;
  tp.n_streams = b->n;
  tp.hist_src = b->d_hist[b->hist_cur];
  tp.hist_dst = b->d_hist[b->hist_cur ^ 1];
  tp.hist_stride = b->hist_stride;
  tp.in = w.d_in;
  tp.in_stride_frames = w.stride_frames;
  tp.counts = w.d_counts;
  tp.frames = w.frames;
  prof_mark(b, st, 3, true);
  CU_TRY(launch_tail(tp, st));
  prof_mark(b, st, 3, false);
  b->hist_cur ^= 1;
  b->last_frames = w.frames;
  b->last_d_counts = w.d_counts;
  return 1;
}

int ensure_aux_streams(speedyBatch b) {
  if (b->s_sonic) return 1;
  // the resynthesis chain is latency bound: its CTAs go first whenever an SM has room
  int prio_lo = 0, prio_hi = 0;
  CU_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  if (getenv("SPEEDY_B200_NO_PRIO")) prio_hi = prio_lo;
  CU_TRY(cudaStreamCreateWithPriority(&b->s_sonic, cudaStreamNonBlocking, prio_hi));
  for (int i = 0; i < kPipeEvents; i++) CU_TRY(cudaEventCreateWithFlags(&b->ev_pipe[i], cudaEventDisableTiming));
  return 1;
}

}  // namespace

// A long write is cut into a few prefixes: the analysis kernels of prefix c+1 run on
// the caller's stream while the resynthesis kernel of prefix c runs on an internal
// one (they meet only through the speeds array), which hides the analysis behind the
// latency-bound splice chain.  Everything is ordered after earlier work on the
// caller's stream and the caller's stream waits for the internal one before the call
// returns control to it, so the call still behaves like work on one stream.
int speedyBatchWriteDevice(speedyBatch b, const int16_t* d_in, int64_t stride_frames, int64_t frames,
                           const int32_t* d_counts, void* cuda_stream) {
  if (!b) return 0;
  if (frames < 0 || frames > b->cfg.max_write_frames) {
    set_error("speedyBatchWriteDevice: frames exceeds max_write_frames");
    return 0;
  }
  if (frames == 0) return 1;
  if (!d_in || stride_frames < frames) {
    set_error("speedyBatchWriteDevice: null input or stride_frames < frames");
    return 0;
  }
  CU_TRY(cudaSetDevice(b->cfg.device));
  cudaStream_t st = pick_stream(b, cuda_stream);
  WriteCall w = {b, d_in, stride_frames, frames, d_counts};
  int parts = 1;
  if (frames >= 8LL * b->g.rate) parts = 6;       // >= 8 s of audio per stream
  else if (frames >= 2LL * b->g.rate) parts = 3;
  if (k1_uses_dft16(b->g)) {
    // the tensor-core analysis kernel does not share SMs with the resynthesis (it fills an SM's shared
    // memory), so prefixes only pay while the resynthesis leaves SMs idle: four at the latency-bound
    // stream counts (measured 15.6 ms against 15.7 with three or six and 15.9 with one at 1024 x 60 s;
    // giving the analysis of a later prefix a few SMs of its own beside the resynthesis: 16.8 - 19.9 ms),
    // one when the streams fill the machine (54.9 ms against 57.0 with six at 8192 x 30 s)
    if (parts > 4) parts = 4;
    if (b->n >= 148 * 24) parts = 1;
  }
  if (const char* e = getenv("SPEEDY_B200_WRITE_PARTS")) parts = atoi(e) > 0 ? atoi(e) : parts;
  if (parts > kPipeEvents - 2) parts = kPipeEvents - 2;
  if (parts == 1) {
    // (Replaying these four launches as a CUDA graph was measured on the 10 ms streaming step,
    // 16384 sessions: 0.264 ms per step with the graph, 0.265 without.  Queued asynchronously the
    // step is bound by the four kernels' own device time, not by launch overhead; not kept.)
    if (!launch_analysis(w, 0, frames, st) || !launch_resynthesis(w, 0, frames, st)) return 0;
    return launch_write_tail(w, st);
  }
  if (!ensure_aux_streams(b)) return 0;
  cudaStream_t ss = b->s_sonic;
  // fork: the internal stream starts after everything already queued on the caller's
  CU_TRY(cudaEventRecord(b->ev_pipe[kPipeEvents - 2], st));
  CU_TRY(cudaStreamWaitEvent(ss, b->ev_pipe[kPipeEvents - 2], 0));
  const long long step = b->g.step;
  long long done = 0;
  for (int c = 0; c < parts; c++) {
    long long prefix = c == parts - 1 ? (long long)frames : ((long long)frames * (c + 1) / parts) / step * step;
    if (prefix <= done) continue;
    if (!launch_analysis(w, done, prefix, st)) return 0;
    CU_TRY(cudaEventRecord(b->ev_pipe[c], st));
    CU_TRY(cudaStreamWaitEvent(ss, b->ev_pipe[c], 0));
    if (!launch_resynthesis(w, done, prefix, ss)) return 0;
    done = prefix;
  }
  // join
  CU_TRY(cudaEventRecord(b->ev_pipe[kPipeEvents - 1], ss));
  CU_TRY(cudaStreamWaitEvent(st, b->ev_pipe[kPipeEvents - 1], 0));
  return launch_write_tail(w, st);
}

int speedyBatchFlushDevice(speedyBatch b, void* cuda_stream) {
  if (!b) return 0;
  CU_TRY(cudaSetDevice(b->cfg.device));
  cudaStream_t st = pick_stream(b, cuda_stream);
  K4Params k4;
  memset(&k4, 0, sizeof(k4));
  k4.g = b->g;
  k4.st = b->st;
  k4.n_streams = b->n;
  k4.hist = b->d_hist[b->hist_cur];
  k4.hist_stride = b->hist_stride;
  k4.in = nullptr;
  k4.frames = 0;
  k4.speeds = nullptr;
  k4.flush = 1;
  k4.out = b->d_out;
  k4.out_capacity = b->out_capacity;
  k4.threads_per_stream = b->cfg.threads_per_stream;
  prof_mark(b, st, 4, true);
  CU_TRY(launch_k4(k4, st));
  prof_mark(b, st, 4, false);
  return 1;
}

// sonicFlushStream for the streams with a non-zero mask entry only (the others keep their
// pending input): what a session pool needs when one of its handles is flushed.
static int upload_mask(speedyBatch b, const int32_t* h_mask, cudaStream_t st) {
  if (!b->d_mask_stage && !dev_alloc(b, &b->d_mask_stage, (size_t)b->n)) return 0;
  CU_TRY(cudaMemcpyAsync(b->d_mask_stage, h_mask, sizeof(int32_t) * b->n, cudaMemcpyHostToDevice, st));
  return 1;
}

int speedyBatchFlushStreams(speedyBatch b, const int32_t* h_mask) {
  if (!b || !h_mask) return 0;
  CU_TRY(cudaSetDevice(b->cfg.device));
  cudaStream_t st = b->own_stream;
  if (!upload_mask(b, h_mask, st)) return 0;
  K4Params k4;
  memset(&k4, 0, sizeof(k4));
  k4.g = b->g;
  k4.st = b->st;
  k4.n_streams = b->n;
  k4.hist = b->d_hist[b->hist_cur];
  k4.hist_stride = b->hist_stride;
  k4.flush = 1;
  k4.flush_mask = b->d_mask_stage;
  k4.out = b->d_out;
  k4.out_capacity = b->out_capacity;
  k4.threads_per_stream = b->cfg.threads_per_stream;
  CU_TRY(launch_k4(k4, st));
  CU_TRY(cudaStreamSynchronize(st));
  return 1;
}

int speedyBatchResetStreams(speedyBatch b, const int32_t* h_mask) {
  if (!b || !h_mask) return 0;
  CU_TRY(cudaSetDevice(b->cfg.device));
  cudaStream_t st = b->own_stream;
  if (!upload_mask(b, h_mask, st)) return 0;
  reset_streams_kernel<<<(b->n + 127) / 128, 128, 0, st>>>(b->n, b->d_mask_stage, b->st);
  count_launch();
  CU_TRY(cudaGetLastError());
  CU_TRY(cudaStreamSynchronize(st));
  return 1;
}

int speedyBatchReadDevice(speedyBatch b, int16_t* d_out, int64_t stride_frames, int32_t* d_counts,
                          void* cuda_stream) {
  if (!b) return 0;
  CU_TRY(cudaSetDevice(b->cfg.device));
  cudaStream_t st = pick_stream(b, cuda_stream);
  if (d_out) {
    if (b->n >= 2048 && b->out_capacity * b->g.channels <= 65536) {
      // many streams, little room per stream: one warp each
      read_copy_warp_kernel<<<(b->n + 7) / 8, 256, 0, st>>>(b->d_out, b->out_capacity, b->g.channels, b->st.out_count,
                                                           d_out, stride_frames, b->n);
    } else {
      dim3 grid(32, b->n);
      read_copy_kernel<<<grid, 256, 0, st>>>(b->d_out, b->out_capacity, b->g.channels, b->st.out_count, d_out,
                                             stride_frames);
    }
    count_launch();
  }
  read_finish_kernel<<<(b->n + 127) / 128, 128, 0, st>>>(b->n, b->st.out_count, b->st.status, d_counts,
                                                         d_out ? stride_frames : (1LL << 40));
  count_launch();
  CU_TRY(cudaGetLastError());
  return 1;
}

int speedyBatchPeekOutputDevice(speedyBatch b, const int16_t** d_out, const int32_t** d_counts,
                                int64_t* capacity_frames) {
  if (!b) return 0;
  if (d_out) *d_out = b->d_out;
  if (d_counts) *d_counts = b->st.out_count;
  if (capacity_frames) *capacity_frames = b->out_capacity;
  return 1;
}

int speedyBatchDiscardOutput(speedyBatch b, void* cuda_stream) {
  return speedyBatchReadDevice(b, nullptr, 0, nullptr, cuda_stream);
}

// ---- host-pointer variants ------------------------------------------------

static int ensure_stage(speedyBatch b, int64_t frames) {
  if (b->d_stage && b->stage_frames >= frames) return 1;
  int16_t* d = nullptr;
  const long long want = b->cfg.max_write_frames;
  if (!dev_alloc(b, &d, (size_t)b->n * want * b->g.channels)) return 0;
  b->d_stage = d;
  b->stage_frames = want;
  return 1;
}

int speedyBatchWrite(speedyBatch b, const int16_t* h_in, int64_t stride_frames, int64_t frames,
                     const int32_t* h_counts) {
  if (!b) return 0;
  if (frames <= 0) return frames == 0;
  if (frames > b->cfg.max_write_frames) {
    set_error("speedyBatchWrite: frames exceeds max_write_frames");
    return 0;
  }
  CU_TRY(cudaSetDevice(b->cfg.device));
  if (!ensure_stage(b, frames)) return 0;
  cudaStream_t st = b->own_stream;
  const size_t row = (size_t)frames * b->g.channels * sizeof(int16_t);
  CU_TRY(cudaMemcpy2DAsync(b->d_stage, row, h_in, (size_t)stride_frames * b->g.channels * sizeof(int16_t), row,
                           b->n, cudaMemcpyHostToDevice, st));
  const int32_t* d_counts = nullptr;
  if (h_counts) {
    CU_TRY(cudaMemcpyAsync(b->d_counts_stage, h_counts, sizeof(int32_t) * b->n, cudaMemcpyHostToDevice, st));
    d_counts = b->d_counts_stage;
  }
  if (!speedyBatchWriteDevice(b, b->d_stage, frames, frames, d_counts, st)) return 0;
  CU_TRY(cudaStreamSynchronize(st));
  return 1;
}

int speedyBatchFlush(speedyBatch b) {
  if (!b) return 0;
  if (!speedyBatchFlushDevice(b, b->own_stream)) return 0;
  CU_TRY(cudaStreamSynchronize(b->own_stream));
  return 1;
}

int speedyBatchRead(speedyBatch b, int16_t* h_out, int64_t stride_frames, int32_t* h_counts) {
  if (!b) return 0;
  CU_TRY(cudaSetDevice(b->cfg.device));
  cudaStream_t st = b->own_stream;
  CU_TRY(cudaMemcpyAsync(b->h_pinned_counts, b->st.out_count, sizeof(int32_t) * b->n, cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaStreamSynchronize(st));
  long long mx = 0;
  for (int s = 0; s < b->n; s++) {
    long long c = b->h_pinned_counts[s];
    if (c > stride_frames) c = stride_frames;
    if (h_counts) h_counts[s] = (int32_t)c;
    if (c > mx) mx = c;
  }
  if (mx > 0 && h_out) {
    const size_t row = (size_t)mx * b->g.channels * sizeof(int16_t);
    CU_TRY(cudaMemcpy2DAsync(h_out, (size_t)stride_frames * b->g.channels * sizeof(int16_t), b->d_out,
                             (size_t)b->out_capacity * b->g.channels * sizeof(int16_t), row, b->n,
                             cudaMemcpyDeviceToHost, st));
  }
  read_finish_kernel<<<(b->n + 127) / 128, 128, 0, st>>>(b->n, b->st.out_count, b->st.status, b->d_counts_stage,
                                                         stride_frames);
  count_launch();
  CU_TRY(cudaGetLastError());
  CU_TRY(cudaStreamSynchronize(st));
  return 1;
}

int speedyBatchSetProfiling(speedyBatch b, int on) {
  if (!b) return 0;
  CU_TRY(cudaSetDevice(b->cfg.device));
  b->profiling = on ? 1 : 0;
  b->prof_used = 0;
  b->prof_marks.clear();
  return 1;
}

// Device time of the kernels launched since profiling was switched on or this
// function was last called: {spectral, tension, sonic, tail, flush-sonic}, each the
// sum of its launches' durations in milliseconds (CUDA events on the launching
// streams; launches on different streams may overlap in wall time).
int speedyBatchGetKernelTimes(speedyBatch b, float* ms5) {
  if (!b || !ms5 || !b->profiling) return 0;
  CU_TRY(cudaSetDevice(b->cfg.device));
  for (int i = 0; i < 5; i++) ms5[i] = 0.0f;
  CU_TRY(cudaDeviceSynchronize());
  for (size_t i = 0; i + 1 < b->prof_marks.size(); i += 2) {
    float ms = 0.0f;
    if (cudaEventElapsedTime(&ms, b->prof_events[i], b->prof_events[i + 1]) == cudaSuccess) {
      ms5[b->prof_marks[i].first] += ms;
    }
  }
  b->prof_used = 0;
  b->prof_marks.clear();
  return 1;
}

// One-shot over host buffers, pipelined in TIME.  The whole call is ONE logical
// write whose data arrives progressively: the input crosses PCIe in chunks of a few
// seconds into a full-size device staging buffer, and as soon as a chunk has landed
// the kernels are launched over the prefix it completes (see kernels.cuh) - analysis
// on one stream, resynthesis on another, the output each resynthesis launch added
// going back to the host on a third while later chunks are still arriving.
int speedyBatchProcess(speedyBatch b, const int16_t* h_in, int64_t frames, int16_t* h_out,
                       int64_t out_stride_frames, int32_t* h_out_counts) {
  if (!b || !h_in || !h_out || frames < 1) return 0;
  if (frames > b->cfg.max_write_frames) {
    set_error("speedyBatchProcess: frames exceeds max_write_frames");
    return 0;
  }
  CU_TRY(cudaSetDevice(b->cfg.device));
  const int n = b->n, C = b->g.channels;
  // Chunk boundaries.  The kernels outrun PCIe, so what the call adds to the transfer
  // time is the work left when the last byte lands: the last chunks taper off
  // (..., 1, 1, 0.8, 0.6, 0.4, 0.2 of the regular size) to keep that tail short.
  long long n_regular = 10;
  if (const char* e = getenv("SPEEDY_B200_CHUNKS")) {
    if (atoi(e) > 0) n_regular = atoi(e);
  }
  long long chunk = (frames + n_regular - 1) / n_regular;
  bool taper = true;
  if (const char* e = getenv("SPEEDY_B200_CHUNK_FRAMES")) {
    if (atoll(e) > 0) {
      chunk = atoll(e);
      taper = false;
    }
  }
  const long long min_chunk = 4 * b->g.rate / 10;  // at least 0.4 s of audio
  if (chunk < min_chunk) chunk = min_chunk;
  chunk = (chunk + 7) & ~7LL;
  if (chunk > frames) chunk = frames;
  std::vector<long long> bound(1, 0);
  {
    long long at = 0;
    const double tail[4] = {0.8, 0.6, 0.4, 0.2};
    const long long tapered = taper ? 2 * chunk : 0;  // the four tapered chunks cover two regular ones
    while (at < frames) {
      long long next = at + chunk;
      if (taper && frames - at <= tapered + 8 && frames - at > min_chunk * 4) {
        const long long left = frames - at;
        for (int i = 0; i < 4 && at < frames; i++) {
          long long sz = ((long long)(left * tail[i] / 2.0) + 7) & ~7LL;
          if (sz < min_chunk) sz = min_chunk;
          next = i == 3 || at + sz > frames ? (long long)frames : at + sz;
          bound.push_back(next);
          at = next;
        }
        break;
      }
      if (next > frames) next = frames;
      bound.push_back(next);
      at = next;
    }
  }
  const int nchunks = (int)bound.size() - 1;
  if (!ensure_stage(b, frames) || !ensure_aux_streams(b)) return 0;
  if (!b->pipe_ready) {
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);  // the copy-out kernel yields to compute
    CU_TRY(cudaStreamCreateWithFlags(&b->s_h2d, cudaStreamNonBlocking));
    CU_TRY(cudaStreamCreateWithPriority(&b->s_d2h, cudaStreamNonBlocking, prio_lo));
    b->pipe_ready = 1;
  }
  cudaStream_t sa = b->own_stream;  // analysis
  cudaStream_t ss = b->s_sonic;     // resynthesis
  if (!speedyBatchReset(b, sa)) return 0;
  std::vector<cudaEvent_t> ev_in(nchunks), ev_k2(nchunks), ev_done(nchunks), ev_out(nchunks);
  const bool trace = getenv("SPEEDY_B200_TRACE") != nullptr;  // developer aid: per-chunk timeline
  const unsigned ev_flags = trace ? cudaEventDefault : cudaEventDisableTiming;
  cudaEvent_t ev_t0 = nullptr;
  for (int c = 0; c < nchunks; c++) {
    CU_TRY(cudaEventCreateWithFlags(&ev_in[c], ev_flags));
    CU_TRY(cudaEventCreateWithFlags(&ev_k2[c], ev_flags));
    CU_TRY(cudaEventCreateWithFlags(&ev_done[c], ev_flags));
    CU_TRY(cudaEventCreateWithFlags(&ev_out[c], ev_flags));
  }
  CU_TRY(cudaEventCreate(&ev_t0));
  CU_TRY(cudaEventRecord(ev_t0, sa));  // after the reset
  CU_TRY(cudaStreamWaitEvent(b->s_h2d, ev_t0, 0));
  CU_TRY(cudaStreamWaitEvent(ss, ev_t0, 0));
  std::vector<long long> prev(n, 0);
  int ok = 1;
  const size_t in_pitch = (size_t)frames * C * sizeof(int16_t);
  const size_t stage_pitch = (size_t)b->stage_frames * C * sizeof(int16_t);
  const size_t out_dpitch = (size_t)out_stride_frames * C * sizeof(int16_t);
  const size_t out_spitch = (size_t)b->out_capacity * C * sizeof(int16_t);
  // Is the caller's output buffer pinned, i.e. addressable from the device?  Then
  // each chunk's new output is stored straight into it by a kernel, ragged starts
  // and all, with no host involvement; otherwise fall back to one rectangular
  // device->host copy per chunk (which re-copies the spread between streams).
  int16_t* h_out_dev = nullptr;
  {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, h_out) == cudaSuccess && attr.type == cudaMemoryTypeHost &&
        attr.devicePointer != nullptr && !getenv("SPEEDY_B200_NO_ZEROCOPY")) {
      h_out_dev = static_cast<int16_t*>(attr.devicePointer);
    }
    cudaGetLastError();
  }
  if (h_out_dev && !b->d_done) {
    for (int i = 0; i < 3; i++) {
      if (!dev_alloc(b, &b->d_snap[i], n)) return 0;
    }
    if (!dev_alloc(b, &b->d_done, n)) return 0;
  }
  if (h_out_dev) CU_TRY(cudaMemsetAsync(b->d_done, 0, sizeof(int) * n, ss));
  auto drain = [&](int c) -> int {
    if (h_out_dev) {
      // device-side scatter on the copy-out stream, ordered after chunk c's kernels
      if (cudaStreamWaitEvent(b->s_d2h, ev_done[c], 0) != cudaSuccess) return 0;
      // Eight blocks: enough to fill the link's device-to-host direction (a block sustains about
      // 6.6 GB/s), few enough that the stores do not arrive in bursts that starve the read requests
      // of the host-to-device copies running beside them (1024 x 60 s end to end: 2 / 4 / 8 / 16 /
      // 48 / 148 blocks = 73.2 / 42.3 / 40.0 / 40.8 / 43.6 / 44.3 ms, profiles/README.md)
      static const int scatter_blocks = [] {
        const char* e = getenv("SPEEDY_B200_SCATTER_BLOCKS");
        return e && atoi(e) > 0 ? atoi(e) : 8;
      }();
      const int blocks = n < scatter_blocks ? n : scatter_blocks;
      scatter_out_kernel<<<blocks, 256, 0, b->s_d2h>>>(b->d_out, b->out_capacity, C, b->d_snap[c % 3], b->d_done,
                                                       h_out_dev, out_stride_frames, n);
      scatter_advance_kernel<<<(n + 127) / 128, 128, 0, b->s_d2h>>>(n, b->d_snap[c % 3], b->d_done,
                                                                    out_stride_frames);
      count_launch();
      count_launch();
      return cudaGetLastError() == cudaSuccess;
    }
    if (cudaEventSynchronize(ev_done[c]) != cudaSuccess) return 0;
    const int32_t* cnt = b->h_pinned_counts + (size_t)(c & 1) * n;
    long long lo = 1LL << 60, hi = 0;
    for (int s = 0; s < n; s++) {
      long long now = cnt[s];
      if (now > out_stride_frames) now = out_stride_frames;
      if (prev[s] < lo) lo = prev[s];
      if (now > hi) hi = now;
      prev[s] = now;
    }
    if (hi > lo) {
      if (cudaMemcpy2DAsync(h_out + lo * C, out_dpitch, b->d_out + lo * C, out_spitch,
                            (size_t)(hi - lo) * C * sizeof(int16_t), n, cudaMemcpyDeviceToHost,
                            b->s_d2h) != cudaSuccess)
        return 0;
    }
    return 1;
  };
  // all host->device copies are queued up front: the copy stream never runs dry
  for (int c = 0; c < nchunks; c++) {
    const long long f0 = bound[c];
    const long long fc = bound[c + 1] - f0;
    CU_TRY(cudaMemcpy2DAsync(b->d_stage + f0 * C, stage_pitch, h_in + f0 * C, in_pitch,
                             (size_t)fc * C * sizeof(int16_t), n, cudaMemcpyHostToDevice, b->s_h2d));
    CU_TRY(cudaEventRecord(ev_in[c], b->s_h2d));
  }
  WriteCall w = {b, b->d_stage, b->stage_frames, frames, nullptr};
  for (int c = 0; c < nchunks && ok; c++) {
    const long long done = bound[c];
    const long long prefix = bound[c + 1];
    CU_TRY(cudaStreamWaitEvent(sa, ev_in[c], 0));
    if (!launch_analysis(w, done, prefix, sa)) ok = 0;
    CU_TRY(cudaEventRecord(ev_k2[c], sa));
    CU_TRY(cudaStreamWaitEvent(ss, ev_k2[c], 0));
    if (ok && !launch_resynthesis(w, done, prefix, ss)) ok = 0;
    if (ok && c == nchunks - 1) {
      // the flush needs the totals the tail advances: tail on sa after this K4, flush after it
      CU_TRY(cudaEventRecord(ev_t0, ss));
      CU_TRY(cudaStreamWaitEvent(sa, ev_t0, 0));
      if (!launch_write_tail(w, sa)) ok = 0;
      CU_TRY(cudaEventRecord(ev_t0, sa));
      CU_TRY(cudaStreamWaitEvent(ss, ev_t0, 0));
      if (ok && !speedyBatchFlushDevice(b, ss)) ok = 0;
    }
    if (h_out_dev) {
      // the scatter of chunk c reads this snapshot while later chunks advance out_count;
      // snapshot c % 3 is free again once the scatter of chunk c - 3 has run
      if (c >= 3) CU_TRY(cudaStreamWaitEvent(ss, ev_out[c - 3], 0));
      CU_TRY(cudaMemcpyAsync(b->d_snap[c % 3], b->st.out_count, sizeof(int32_t) * n, cudaMemcpyDeviceToDevice, ss));
    }
    if (!h_out_dev || c == nchunks - 1) {
      CU_TRY(cudaMemcpyAsync(b->h_pinned_counts + (size_t)(c & 1) * n, b->st.out_count, sizeof(int32_t) * n,
                             cudaMemcpyDeviceToHost, ss));
    }
    CU_TRY(cudaEventRecord(ev_done[c], ss));
    if (h_out_dev) {
      if (ok) ok = drain(c);
      CU_TRY(cudaEventRecord(ev_out[c], b->s_d2h));
    } else if (ok && c >= 1) {
      ok = drain(c - 1);
    }
  }
  if (ok && !h_out_dev) ok = drain(nchunks - 1);
  if (ok && h_out_dev) {
    if (cudaEventSynchronize(ev_done[nchunks - 1]) != cudaSuccess) ok = 0;
    const int32_t* cnt = b->h_pinned_counts + (size_t)((nchunks - 1) & 1) * n;
    for (int s = 0; s < n; s++) prev[s] = cnt[s] < out_stride_frames ? cnt[s] : out_stride_frames;
  }
  if (h_out_counts) {
    for (int s = 0; s < n; s++) h_out_counts[s] = (int32_t)prev[s];
  }
  if (cudaStreamSynchronize(b->s_d2h) != cudaSuccess) ok = 0;
  read_finish_kernel<<<(n + 127) / 128, 128, 0, ss>>>(n, b->st.out_count, b->st.status, b->d_counts_stage,
                                                      out_stride_frames);
  count_launch();
  if (cudaStreamSynchronize(b->s_h2d) != cudaSuccess) ok = 0;
  if (cudaStreamSynchronize(ss) != cudaSuccess) ok = 0;
  if (cudaStreamSynchronize(sa) != cudaSuccess) ok = 0;
  if (trace) {
    for (int c = 0; c < nchunks; c++) {
      float a = 0, k = 0, d = 0, o = 0;
      cudaEventElapsedTime(&a, ev_in[0], ev_in[c]);
      cudaEventElapsedTime(&k, ev_in[0], ev_k2[c]);
      cudaEventElapsedTime(&d, ev_in[0], ev_done[c]);
      if (h_out_dev) cudaEventElapsedTime(&o, ev_in[0], ev_out[c]);
      fprintf(stderr, "[speedyBatchProcess] chunk %2d (ms after the first h2d): h2d %7.2f analysis %7.2f "
              "resynthesis %7.2f out %7.2f\n", c, a, k, d, o);
    }
  }
  cudaEventDestroy(ev_t0);
  for (int c = 0; c < nchunks; c++) {
    cudaEventDestroy(ev_in[c]);
    cudaEventDestroy(ev_k2[c]);
    cudaEventDestroy(ev_done[c]);
    cudaEventDestroy(ev_out[c]);
  }
  if (!ok) set_error(std::string("speedyBatchProcess: ") + cudaGetErrorString(cudaGetLastError()));
  return ok;
}

int speedyBatchGetStatus(speedyBatch b, int32_t* status) {
  if (!b || !status) return 0;
  CU_TRY(cudaSetDevice(b->cfg.device));
  CU_TRY(cudaDeviceSynchronize());  // (work may be queued on a caller's stream as well as on ours)
  CU_TRY(cudaMemcpy(status, b->st.status, sizeof(int32_t) * b->n, cudaMemcpyDeviceToHost));
  return 1;
}

int speedyBatchGetTaps(speedyBatch b, int64_t max_frames, int32_t* n_analysis, int32_t* n_tension,
                       float* spectrogram, float* energy, float* features, float* tension, float* speed) {
  if (!b) return 0;
  CU_TRY(cudaSetDevice(b->cfg.device));
  CU_TRY(cudaDeviceSynchronize());
  const int n = b->n;
  const long long rows = b->max_new_frames;
  const long long take = max_frames < rows ? max_frames : rows;
  auto pull = [&](float* h, const float* d, int width) -> int {
    if (!h) return 1;
    if (!d) {
      set_error("speedyBatchGetTaps: tap was not enabled at creation");
      return 0;
    }
    CU_TRY(cudaMemcpy2D(h, (size_t)max_frames * width * sizeof(float), d, (size_t)rows * width * sizeof(float),
                        (size_t)take * width * sizeof(float), n, cudaMemcpyDeviceToHost));
    return 1;
  };
  if (!pull(spectrogram, b->d_tap_spec, b->g.fft) || !pull(energy, b->d_tap_energy, 1) ||
      !pull(features, b->d_tap_features, kFeatureCount) || !pull(tension, b->d_tap_tension, 1) ||
      !pull(speed, b->d_tap_speed, 1))
    return 0;
  if (n_analysis || n_tension) {
    // frames produced by the last write = f(total) - f(total - frames written)
    std::vector<long long> tot(n);
    std::vector<int32_t> cnt(n);
    CU_TRY(cudaMemcpy(tot.data(), b->st.total, sizeof(long long) * n, cudaMemcpyDeviceToHost));
    if (b->last_d_counts) {
      CU_TRY(cudaMemcpy(cnt.data(), b->last_d_counts, sizeof(int32_t) * n, cudaMemcpyDeviceToHost));
    }
    for (int s = 0; s < n; s++) {
      const long long before = tot[s] - (b->last_d_counts ? cnt[s] : b->last_frames);
      const int a1 = frames_analyzed(b->g, tot[s]);
      const int a0 = frames_analyzed(b->g, before);
      if (n_analysis) n_analysis[s] = a1 - a0;
      if (n_tension) n_tension[s] = tensions_ready(b->g, a1) - tensions_ready(b->g, a0);
    }
  }
  return 1;
}

void* speedyBatchHostAlloc(size_t bytes, int write_combined) {
  void* p = nullptr;
  const unsigned flags = cudaHostAllocPortable | (write_combined ? cudaHostAllocWriteCombined : 0u);
  if (cudaHostAlloc(&p, bytes, flags) != cudaSuccess) {
    cudaGetLastError();
    set_error("speedyBatchHostAlloc: cudaHostAlloc failed");
    return nullptr;
  }
  return p;
}

void speedyBatchHostFree(void* p) {
  if (p) cudaFreeHost(p);
}

int speedyBatchSynthDevice(int16_t* d_out, uint64_t first_id, int32_t num_streams, int32_t sample_rate,
                           int32_t channels, int64_t frames, void* cuda_stream) {
  if (!d_out || num_streams < 1 || frames < 1) return 0;
  long long blocks = (frames + 255) / 256;
  if (blocks > 1024) blocks = 1024;
  dim3 grid((unsigned)blocks, num_streams);
  synth_kernel<<<grid, 256, 0, (cudaStream_t)cuda_stream>>>(d_out, first_id, sample_rate, channels, frames);
  count_launch();
  CU_TRY(cudaGetLastError());
  return 1;
}

}  // extern "C"
