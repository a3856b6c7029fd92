/* TEST INFRASTRUCTURE (oracle) — not part of the product path.
 *
 * Driver around the reference's OWN public API.  It is linked, together with
 * the unmodified /root/reference/speedy.c and soniclib.c (compiled from where
 * they lie, see the Makefile), into oracle/_ref/libspeedy_ref_*.so.  It only
 * does what a client of the library does (speedy_wave.cc:154-242,
 * sonic_test.cc:364-403): create a stream, set speed / nonlinear factor /
 * feedback, register the debug callbacks (sonic2.h:100-125), write the samples
 * in chunks, read, flush, drain.
 */
#include <fcntl.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "sonic2.h" /* /root/reference/sonic2.h, via -I */
#include "speedy.h" /* /root/reference/speedy.h */

typedef struct {
  int max_frames;
  int fft;
  int n_tension, n_speed, n_features, n_spec, n_norm;
  float* tension;     /* [max_frames] */
  float* speed;       /* [max_frames] */
  float* features;    /* [max_frames][15] */
  float* spectrogram; /* [max_frames][fft] */
  float* normalized;  /* [max_frames][fft/2], as handed to the callback */
  int* tension_time;  /* [max_frames] frame index passed to the callback */
  int* spec_time;     /* [max_frames] */
} ref_taps;

static __thread ref_taps* tls_taps;

static void on_tension(sonicStream s, int time, float tension) {
  (void)s;
  ref_taps* t = tls_taps;
  if (t && t->n_tension < t->max_frames) {
    if (t->tension) t->tension[t->n_tension] = tension;
    if (t->tension_time) t->tension_time[t->n_tension] = time;
  }
  if (t) t->n_tension++;
}
static void on_speed(sonicStream s, int time, float speed) {
  (void)s; (void)time;
  ref_taps* t = tls_taps;
  if (t && t->speed && t->n_speed < t->max_frames) t->speed[t->n_speed] = speed;
  if (t) t->n_speed++;
}
static void on_features(sonicStream s, int time, float* f) {
  (void)s; (void)time;
  ref_taps* t = tls_taps;
  if (t && t->features && t->n_features < t->max_frames) {
    memcpy(t->features + (size_t)t->n_features * kFeatureValueCount, f,
           sizeof(float) * kFeatureValueCount);
  }
  if (t) t->n_features++;
}
static void on_spectrogram(sonicStream s, int time, float* spec) {
  (void)s;
  ref_taps* t = tls_taps;
  if (t && t->n_spec < t->max_frames) {
    if (t->spectrogram) {
      memcpy(t->spectrogram + (size_t)t->n_spec * t->fft, spec,
             sizeof(float) * t->fft);
    }
    if (t->spec_time) t->spec_time[t->n_spec] = time;
  }
  if (t) t->n_spec++;
}
static void on_normalized(sonicStream s, int time, float* spec) {
  (void)s; (void)time;
  ref_taps* t = tls_taps;
  if (t && t->normalized && t->n_norm < t->max_frames) {
    memcpy(t->normalized + (size_t)t->n_norm * (t->fft / 2), spec,
           sizeof(float) * (t->fft / 2));
  }
  if (t) t->n_norm++;
}

/* soniclib.c:201,219 print to stdout on the first write of every stream. */
static int saved_stdout = -1;
void ref_quiet(int on) {
  fflush(stdout);
  if (on && saved_stdout < 0) {
    saved_stdout = dup(1);
    int devnull = open("/dev/null", O_WRONLY);
    dup2(devnull, 1);
    close(devnull);
  } else if (!on && saved_stdout >= 0) {
    dup2(saved_stdout, 1);
    close(saved_stdout);
    saved_stdout = -1;
  }
}

int ref_future_frames(void) { return kTemporalHysteresisFuture; }
int ref_past_frames(void) { return kTemporalHysteresisPast; }

/* One stream through sonicCreateStream / Write / Read / Flush, `chunk` sample
 * frames per write (chunk <= 0: one write).  Returns the number of output
 * sample frames produced (only out_cap are stored). */
long ref_run_stream(const short* in, long n_frames, int rate, int channels,
                    float speed, float nonlinear, float feedback, int chunk,
                    short* out, long out_cap, ref_taps* taps) {
  sonicStream s = sonicCreateStream(rate, channels);
  if (!s) return -1;
  sonicSetSpeed(s, speed);
  sonicEnableNonlinearSpeedup(s, nonlinear);
  sonicSetDurationFeedbackStrength(s, feedback);
  tls_taps = taps;
  if (taps) {
    taps->fft = sonicSpectrogramSize(s);
    taps->n_tension = taps->n_speed = taps->n_features = 0;
    taps->n_spec = taps->n_norm = 0;
    sonicTensionCallback(s, on_tension);
    sonicSpeedCallback(s, on_speed);
    sonicFeaturesCallback(s, on_features);
    sonicSpectrogramCallback(s, on_spectrogram);
    sonicNormalizedSpectrogramCallback(s, on_normalized);
  }
  if (chunk <= 0) chunk = n_frames > 0 ? (int)n_frames : 1;
  long produced = 0;
  short* tmp = (short*)malloc(sizeof(short) * (size_t)channels * 4096);
  for (long t = 0; t < n_frames; t += chunk) {
    int count = (int)(n_frames - t < chunk ? n_frames - t : chunk);
    sonicWriteShortToStream(s, in + (size_t)t * channels, count);
    for (;;) {
      int got = sonicReadShortFromStream(s, tmp, 4096);
      if (got <= 0) break;
      long room = out_cap - produced;
      long take = got < room ? got : (room > 0 ? room : 0);
      if (out) memcpy(out + produced * channels, tmp, sizeof(short) * (size_t)take * channels);
      produced += got;
    }
  }
  sonicFlushStream(s);
  for (;;) {
    int got = sonicReadShortFromStream(s, tmp, 4096);
    if (got <= 0) break;
    long room = out_cap - produced;
    long take = got < room ? got : (room > 0 ? room : 0);
    if (out) memcpy(out + produced * channels, tmp, sizeof(short) * (size_t)take * channels);
    produced += got;
  }
  free(tmp);
  tls_taps = NULL;
  sonicDestroyStream(s);
  return produced;
}

typedef struct {
  const short* in;
  long n_frames;
  int n_streams, rate, channels;
  float speed, nonlinear, feedback;
  int chunk;
  short* out;
  long out_cap;
  long* out_counts;
  int tid, n_threads;
} ref_job;

static void* ref_worker(void* arg) {
  ref_job* j = (ref_job*)arg;
  for (int i = j->tid; i < j->n_streams; i += j->n_threads) {
    j->out_counts[i] = ref_run_stream(
        j->in + (size_t)i * j->n_frames * j->channels, j->n_frames, j->rate,
        j->channels, j->speed, j->nonlinear, j->feedback, j->chunk,
        j->out ? j->out + (size_t)i * j->out_cap * j->channels : NULL,
        j->out ? j->out_cap : 0, NULL);
  }
  return NULL;
}

/* Independent streams, one OS thread per shard (the CPU baseline, kind
 * "reference").  in: [n_streams][n_frames][channels]. */
int ref_run_batch(const short* in, long n_frames, int n_streams, int rate,
                  int channels, float speed, float nonlinear, float feedback,
                  int chunk, short* out, long out_cap, long* out_counts,
                  int n_threads) {
  if (n_threads < 1) n_threads = 1;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * n_threads);
  ref_job* jobs = (ref_job*)malloc(sizeof(ref_job) * n_threads);
  ref_quiet(1);
  for (int t = 0; t < n_threads; t++) {
    ref_job j = {in, n_frames, n_streams, rate, channels, speed, nonlinear,
                 feedback, chunk, out, out_cap, out_counts, t, n_threads};
    jobs[t] = j;
    pthread_create(&th[t], NULL, ref_worker, &jobs[t]);
  }
  for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
  ref_quiet(0);
  free(th);
  free(jobs);
  return 0;
}
