/* speedy_b200 — C ABI of the B200-native nonlinear speech speed-up path.
 *
 * Two surfaces, both plain C (pointers and sizes only, no C++/torch types):
 *
 *  1. The Sonic/Speedy drop-in: the exact names and signatures a client of the
 *     reference's libspeedy.so links against (/root/reference/sonic2.h:54-125,
 *     implemented there by soniclib.c).  One handle = one stream = a batch of one
 *     on the GPU; correct, not fast.
 *
 *  2. The batched multi-stream entry points (speedyBatch*, new; SURVEY.md §8b
 *     "New"): N independent streams with device-resident state, processed by
 *     the four sm_100a kernels (spectrogram FFT, features/tension/speed
 *     recurrences, Sonic AMDF + overlap-add).  This is the hot path.
 *
 * There is no CPU fallback: every entry point that computes needs a CUDA device
 * and fails (NULL / 0 / negative) without one.
 */
#ifndef SPEEDY_B200_H_
#define SPEEDY_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ======================================================================== *
 * 1. Sonic / Speedy drop-in (replaces soniclib.c + speedy.c + upstream Sonic)
 * ======================================================================== */

struct sonicStreamStruct;
typedef struct sonicStreamStruct* sonicStream;

/* sonic2.h:54, soniclib.c:93-134.  NULL on failure (no device, out of memory). */
sonicStream sonicCreateStream(int sampleRate, int numChannels);
/* sonic2.h:55, soniclib.c:141-167 */
void sonicDestroyStream(sonicStream stream);
/* sonic2.h:60-61, soniclib.c:391-452.  sampleCount is in multi-channel sample
 * frames, data interleaved.  Returns 1 on success, 0 on failure. */
int sonicWriteShortToStream(sonicStream stream, const short* inBuffer,
                            int sampleCount);
/* sonic2.h:62-63, soniclib.c:519-522.  Returns the number of sample frames
 * copied to outBuffer (0 = nothing ready). */
int sonicReadShortFromStream(sonicStream stream, short* outBuffer,
                             int bufferSize);
/* sonic2.h:65-68, soniclib.c:457-517, 524-527.  Floats in (-1, 1). */
int sonicWriteFloatToStream(sonicStream stream, const float* inBuffer,
                            int sampleCount);
int sonicReadFloatFromStream(sonicStream stream, float* outBuffer,
                             int bufferSize);
/* sonic2.h:70, soniclib.c:169-175 (forwards to upstream Sonic's sonicSetRate).  The
 * frames the speed change produced are resampled by upstream Sonic's classic
 * linear-interpolation rate change (output frames = input / (speed * rate), pitch
 * scaled by rate), on the host between the device output and the read FIFO; rate 1
 * costs nothing.  Parity unpinned: upstream Sonic is not part of the reference tree
 * and no reference test sets a rate. */
void sonicSetRate(sonicStream stream, float rate);
/* sonic2.h:71, soniclib.c:177-183.  Global speed R_g. */
void sonicSetSpeed(sonicStream stream, float speed);
/* sonic2.h:72, soniclib.c:529-552. */
int sonicFlushStream(sonicStream stream);
/* sonic2.h:83-84, soniclib.c:555-562.  0 = linear Sonic, 1 = Speedy. */
void sonicEnableNonlinearSpeedup(sonicStream stream, float nonlinearFactor);
/* sonic2.h:92, soniclib.c:565-571.  Library default 0.1 (soniclib.c:122). */
void sonicSetDurationFeedbackStrength(sonicStream stream, float factor);
/* sonic2.h:97, soniclib.c:672-680.  0 before the first nonlinear write. */
int getSonicBufferSize(sonicStream stream);
/* sonic2.h:125, soniclib.c:661-669. */
int sonicSpectrogramSize(sonicStream stream);

/* Debug/parity taps, fired synchronously inside sonicWrite*, once per 10 ms
 * frame, in the reference's order (soniclib.c:297-353): spectrogram,
 * normalized spectrogram, tension, features, speed.  sonic2.h:100-124. */
typedef void (*tensionFunction)(sonicStream stream, int time, float tension);
typedef void (*speedFunction)(sonicStream stream, int time, float speed);
typedef void (*featuresFunction)(sonicStream stream, int time, float* features);
typedef void (*spectrogramFunction)(sonicStream stream, int time,
                                    float* spectrogram);
void sonicTensionCallback(sonicStream stream, tensionFunction fn);
tensionFunction getSonicTensionCallback(sonicStream stream);
void sonicSpeedCallback(sonicStream stream, speedFunction fn);
speedFunction getSonicSpeedCallback(sonicStream stream);
void sonicFeaturesCallback(sonicStream stream, featuresFunction fn);
featuresFunction getSonicFeaturesCallback(sonicStream stream);
void sonicSpectrogramCallback(sonicStream stream, spectrogramFunction fn);
spectrogramFunction getSonicSpectrogramCallback(sonicStream stream);
void sonicNormalizedSpectrogramCallback(sonicStream stream,
                                        spectrogramFunction fn);
spectrogramFunction getSonicNormalizedSpectrogramCallback(sonicStream stream);

/* Upstream-Sonic accessors the reference's clients use under their renamed
 * names (sonic2.h:22-35; sonic_test.cc:370). */
int sonicIntGetNumChannels(sonicStream stream);
int sonicIntGetSampleRate(sonicStream stream);
float sonicIntGetSpeed(sonicStream stream);
int sonicIntSamplesAvailable(sonicStream stream);
/* Plain Sonic on the same handle (sonic_test.cc:729-752).  Valid while the nonlinear
 * factor is 0; otherwise the write and flush return 0. */
void sonicIntSetSpeed(sonicStream stream, float speed);
int sonicIntWriteShortToStream(sonicStream stream, const short* inBuffer, int sampleCount);
int sonicIntReadShortFromStream(sonicStream stream, short* outBuffer, int bufferSize);
int sonicIntFlushStream(sonicStream stream);

/* ------------------------------------------------------------------------ *
 * 1b. Session pool: many sonicStream handles multiplexed onto ONE device batch.
 *
 * A libsonic client with thousands of live handles (BASELINE.json config 5: 16384 sessions
 * fed 10 ms chunks through sonicWriteShortToStream, soniclib.c:391-452) would otherwise be
 * thousands of batches of one.  Handles opened from a pool share one speedyBatch: a write
 * only queues the samples in the pool's page-locked staging row of that handle; one
 * coalesced step (one host->device copy, the four kernels over every session with pending
 * input, one device->host read) runs when
 *   - a handle with queued input is read, flushed or asked how much it has available,
 *   - a handle's staging row is full,
 *   - `auto_step_sessions` handles have queued input (0: never by count), or
 *   - speedySessionPoolStep is called.
 * Each session's result is bit-identical to feeding the same samples to its own stream
 * (chunking never changes the output, tests/test_gpu_parity.py); only WHEN output becomes
 * readable differs from the reference.  All drop-in calls of section 1 work on pooled
 * handles except the five debug callbacks (a pooled handle ignores them: taps of thousands
 * of sessions are what the batched API's taps are for).
 * sonicDestroyStream closes the session and frees its slot.  Calls are serialised by a
 * mutex per pool.  Setting the environment variable SPEEDY_B200_POOL_SESSIONS=<n> makes
 * plain sonicCreateStream open its handles from an implicit pool of n sessions per
 * (sample rate, channels), so an unmodified client is multiplexed too.
 * ------------------------------------------------------------------------ */
struct speedySessionPoolStruct;
typedef struct speedySessionPoolStruct* speedySessionPool;

typedef struct {
  int32_t sample_rate;
  int32_t num_channels;
  int32_t max_sessions;        /* slots of the shared batch */
  int32_t device;
  int32_t max_pending_frames;  /* staging row per session; default 100 ms of audio */
  float min_speed;             /* sizes the per-step output room (frames / min_speed); default 0.25 */
  int32_t auto_step_sessions;  /* step as soon as this many sessions have queued input; 0 = off */
} speedySessionPoolConfig;

typedef struct {
  int64_t steps;            /* coalesced steps run so far */
  int64_t session_writes;   /* sonicWrite* calls absorbed */
  int64_t sessions_served;  /* sum over steps of sessions that had input */
  int32_t open_sessions;
  int32_t pending_sessions; /* sessions with queued input right now */
  double last_step_ms;      /* host wall time of the last step */
} speedySessionPoolStats;

void speedySessionPoolDefaultConfig(speedySessionPoolConfig* cfg);
/* NULL on failure (speedyBatchLastError()). */
speedySessionPool speedySessionPoolCreate(const speedySessionPoolConfig* cfg);
/* Closes every handle still open, then frees the batch. */
void speedySessionPoolDestroy(speedySessionPool pool);
/* A new session (library defaults: speed 1, linear, feedback 0.1); NULL when the pool is full. */
sonicStream speedySessionPoolOpen(speedySessionPool pool);
/* Process everything queued now.  Returns the number of sessions served, -1 on failure. */
int speedySessionPoolStep(speedySessionPool pool);
int speedySessionPoolGetStats(speedySessionPool pool, speedySessionPoolStats* stats);

/* ======================================================================== *
 * 2. Batched multi-stream API (new)
 * ======================================================================== */

struct speedyBatchStruct;
typedef struct speedyBatchStruct* speedyBatch;

#define SPEEDY_FEATURE_COUNT 15 /* speedy.h:129, speedy.c:106-124 */

/* Which per-frame values to keep for speedyBatchGetTaps (parity/debug). */
#define SPEEDY_TAP_TENSION 1
#define SPEEDY_TAP_SPEED 2
#define SPEEDY_TAP_FEATURES 4
#define SPEEDY_TAP_SPECTROGRAM 8
#define SPEEDY_TAP_ENERGY 16

typedef struct {
  int32_t sample_rate;     /* Hz; frame geometry follows speedy.c:213-214 */
  int32_t num_channels;    /* interleaved channels per sample frame */
  int32_t num_streams;     /* independent streams in this batch */
  int32_t match_matlab;    /* 1: hysteresis Future=8/Past=12 (-DMATCH_MATLAB,
                              speedy.h:136-146); 0: 12/8 (shipped library) */
  float speed;             /* R_g for every stream (sonicSetSpeed) */
  float nonlinear_factor;  /* sonicEnableNonlinearSpeedup; 0 = linear Sonic */
  float feedback_strength; /* sonicSetDurationFeedbackStrength */
  int32_t device;          /* CUDA device ordinal */
  int64_t max_write_frames;/* largest per-stream write, in sample frames */
  int64_t out_capacity;    /* per-stream output buffer, in sample frames */
  int32_t taps;            /* SPEEDY_TAP_* mask */
  int32_t threads_per_stream; /* resynthesis kernel: 0 = choose, or 32/64/128 */
  /* White-box analysis hook (0 = off), for driving the analysis the way the reference's own
   * tests drive speedyAddData (speedy_test.cc:859-941: explicit 330-sample frames starting at
   * round(t * 220.5), numbered from 0): analysis frames advance by this many samples instead
   * of rate / 100 (speedy.c:335-338) and are numbered from 0 instead of the shim's 1
   * (soniclib.c:296).  With the window length itself, consecutive windows are disjoint and a
   * stream that is the concatenation of explicit frames reproduces speedyAddDataShort frame by
   * frame, pre-emphasis state included (speedy.c:416-425, 553-565). */
  int32_t analysis_frame_step;
} speedyBatchConfig;

/* Fills cfg with the library defaults (soniclib.c:114-122: speed 1, nonlinear
 * off, feedback 0.1). */
void speedyBatchDefaultConfig(speedyBatchConfig* cfg);

/* NULL on failure; speedyBatchLastError() says why. */
speedyBatch speedyBatchCreate(const speedyBatchConfig* cfg);
void speedyBatchDestroy(speedyBatch batch);
const char* speedyBatchLastError(void);

/* Back to the just-created state (all streams empty; speed, nonlinear factor and
 * feedback strength are kept).  Return 1/0. */
int speedyBatchReset(speedyBatch batch, void* cuda_stream);

/* Per-stream parameters; `values` is a host array of num_streams floats or
 * NULL to set every stream to `uniform`.  Same meaning as the per-handle
 * setters above.  Return 1/0. */
int speedyBatchSetSpeed(speedyBatch batch, const float* values, float uniform);
int speedyBatchSetNonlinear(speedyBatch batch, const float* values,
                            float uniform);
int speedyBatchSetFeedback(speedyBatch batch, const float* values,
                           float uniform);

/* Replace the computed per-frame speeds with caller-supplied ones for every
 * following write (test hook: "integer stages bit-exact given identical
 * per-frame speeds").  speeds: host array [num_streams][frames_per_stream],
 * row s holds the speeds of stream s for tension frames 0,1,...  NULL turns the
 * override off. */
int speedyBatchOverrideSpeeds(speedyBatch batch, const float* speeds,
                              int64_t frames_per_stream);

/* Work is enqueued on `cuda_stream` (a cudaStream_t passed as void*; NULL = the
 * batch's own stream).  The *Device variants take device pointers and do not
 * synchronise; the host variants copy through pinned staging and return when
 * the data is safe to reuse.
 *
 * Write: `frames` sample frames for every stream (or counts[s] <= frames when
 * `counts` is non-NULL), stream s starting at in + s * stride_frames *
 * num_channels.  Equivalent to calling sonicWriteShortToStream once per stream
 * (soniclib.c:391-452).  Return 1/0. */
int speedyBatchWriteDevice(speedyBatch batch, const int16_t* d_in,
                           int64_t stride_frames, int64_t frames,
                           const int32_t* d_counts, void* cuda_stream);
int speedyBatchWrite(speedyBatch batch, const int16_t* h_in,
                     int64_t stride_frames, int64_t frames,
                     const int32_t* h_counts);

/* sonicFlushStream for every stream (soniclib.c:529-552). */
int speedyBatchFlushDevice(speedyBatch batch, void* cuda_stream);
int speedyBatchFlush(speedyBatch batch);
/* The same for the streams with a non-zero entry in the host array mask[num_streams]
 * only; the others keep their pending input (the session pool's per-handle flush). */
int speedyBatchFlushStreams(speedyBatch batch, const int32_t* h_mask);
/* Back to the just-created state for the masked streams only (parameters are kept). */
int speedyBatchResetStreams(speedyBatch batch, const int32_t* h_mask);

/* Read: moves every stream's pending output (sonicReadShortFromStream drained
 * to empty) to out + s * stride_frames * num_channels and stores the number of
 * sample frames in counts[s].  A stream whose pending output exceeds
 * stride_frames is truncated and flagged (speedyBatchGetStatus).  Return 1/0. */
int speedyBatchReadDevice(speedyBatch batch, int16_t* d_out,
                          int64_t stride_frames, int32_t* d_counts,
                          void* cuda_stream);
int speedyBatchRead(speedyBatch batch, int16_t* h_out, int64_t stride_frames,
                    int32_t* h_counts);

/* Zero-copy view of the pending output: device pointer to
 * [num_streams][out_capacity][num_channels] int16 and to the per-stream pending
 * counts.  Valid until the next write/read on the batch. */
int speedyBatchPeekOutputDevice(speedyBatch batch, const int16_t** d_out,
                                const int32_t** d_counts,
                                int64_t* capacity_frames);
/* Drop the pending output of every stream without copying it. */
int speedyBatchDiscardOutput(speedyBatch batch, void* cuda_stream);

/* One-shot convenience over host buffers: RESET (every stream starts empty: state and
 * pending output of earlier calls are discarded; parameters are kept), write `frames`
 * per stream, flush, read; host<->device copies are pipelined against the kernels in
 * chunks of time.  out_counts[s] receives the frames produced.  Return 1/0. */
int speedyBatchProcess(speedyBatch batch, const int16_t* h_in, int64_t frames,
                       int16_t* h_out, int64_t out_stride_frames,
                       int32_t* h_out_counts);

/* Per-kernel device timing for benchmarks.  When on, every write/flush brackets
 * its kernels with CUDA events on the launching stream; GetKernelTimes returns
 * the milliseconds of {spectral, tension, sonic, tail, flush-sonic} of the last
 * write and flush (synchronises on those events).  Return 1/0. */
int speedyBatchSetProfiling(speedyBatch batch, int on);
int speedyBatchGetKernelTimes(speedyBatch batch, float* ms5);

/* Taps of the LAST write call, copied to host arrays (any may be NULL):
 *   n_analysis[s], n_tension[s]  frames produced by that write
 *   spectrogram [s][max_frames][fft]   row j = j-th new analysis frame
 *   energy      [s][max_frames]
 *   features    [s][max_frames][15], tension/speed [s][max_frames]
 * Only taps enabled at creation are available.  Return 1/0. */
int speedyBatchGetTaps(speedyBatch batch, int64_t max_frames,
                       int32_t* n_analysis, int32_t* n_tension,
                       float* spectrogram, float* energy, float* features,
                       float* tension, float* speed);

/* Per-stream status bits (host array of num_streams).  Ordering contract of the host-side
 * helpers (GetStatus, GetTaps, SetSpeed / SetNonlinear / SetFeedback with a host array): they
 * wait for ALL work queued on the device, including *Device calls on a caller's stream. */
#define SPEEDY_STATUS_OUTPUT_OVERFLOW 1 /* output did not fit out_capacity */
#define SPEEDY_STATUS_FLUSHED 2
#define SPEEDY_STATUS_INPUT_OVERFLOW 4  /* Sonic FIFO outgrew the history */
#define SPEEDY_STATUS_READ_TRUNCATED 8
#define SPEEDY_STATUS_SPLICE_STALLED 16 /* a speed left no room for one output sample per pitch period */
int speedyBatchGetStatus(speedyBatch batch, int32_t* status);

/* Frame geometry for a sample rate (speedy.c:213-214, 335-338). */
int speedyBatchFrameGeometry(int sample_rate, int* window, int* fft, int* step);
int speedyBatchNumStreams(speedyBatch batch);

/* Number of kernels this library has launched since the process started. */
int64_t speedyBatchKernelLaunches(void);
/* Name of the kernel variants compiled in (for logs). */
const char* speedyBatchBuildInfo(void);

/* Fill a device buffer [num_streams][frames][channels] with the deterministic
 * integer "speech-shaped" test signal; stream s gets id
 * first_id + s.  Benchmark/test helper, not part of the reference API. */
int speedyBatchSynthDevice(int16_t* d_out, uint64_t first_id, int32_t num_streams,
                           int32_t sample_rate, int32_t channels, int64_t frames,
                           void* cuda_stream);

/* Page-locked host buffers for speedyBatchProcess / Write / Read (what cudaHostAlloc
 * gives, without making the client link the CUDA runtime).  write_combined = 1 suits an
 * INPUT buffer the host only writes (the device reads it without cache snoops); never
 * use it for a buffer the host reads.  Returns NULL on failure. */
void* speedyBatchHostAlloc(size_t bytes, int write_combined);
void speedyBatchHostFree(void* p);

#ifdef __cplusplus
}
#endif
#endif /* SPEEDY_B200_H_ */
