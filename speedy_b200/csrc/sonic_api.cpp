// The Sonic/Speedy drop-in surface (include/speedy_b200.h section 1): the names
// and signatures of /root/reference/sonic2.h:54-125, implemented on top of the
// batched CUDA path with a batch of one stream.  Host logic only: parameter
// bookkeeping, the host-side output FIFO sonicRead* pops from, and the five
// debug callbacks replayed in the reference's order (soniclib.c:297-353).
//
// Correct, not fast: every write is a host->device copy, four kernel launches
// and a device->host read.  Throughput comes from speedyBatch*.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/speedy_b200.h"

namespace {
constexpr int kChunkFrames = 8192;  // largest single batch write
}

struct sonicStreamStruct {
  int sample_rate;
  int channels;
  float speed;
  float rate;
  float nonlinear;
  float feedback;
  speedyBatch batch;
  int window, fft, step;
  int buffer_size;  // 0 until the first nonlinear write (soniclib.c:195, 672-680)
  bool started;
  int last_mode;     // -1 nothing written yet, 0 last write was linear, 1 nonlinear
  bool flushed;      // nothing pending since the last sonicFlushStream
  long long out_capacity;
  std::vector<short> fifo;     // produced, not yet read (interleaved)
  std::vector<short> scratch;  // one batch read
  // callbacks (sonic2.h:100-124)
  tensionFunction on_tension;
  speedFunction on_speed;
  featuresFunction on_features;
  spectrogramFunction on_spectrogram;
  spectrogramFunction on_normalized;
  // callback replay state
  int spec_time;      // at_time of the last spectrogram reported
  int tension_time;   // next tension frame index
  int future;
  std::vector<std::vector<float>> spec_ring;  // last 32 spectrogram frames by at_time & 31
  std::vector<float> normalized;              // what speedyGetNormalizedSpectrogram would hold
  std::vector<float> tap_spec, tap_feat, tap_tension, tap_speed;
};

static bool ensure_batch(sonicStream s) {
  if (s->batch) return true;
  speedyBatchConfig cfg;
  speedyBatchDefaultConfig(&cfg);
  cfg.sample_rate = s->sample_rate;
  cfg.num_channels = s->channels;
  cfg.num_streams = 1;
  cfg.match_matlab = 0;  // the shipped library: Future = 12, Past = 8 (speedy.h:142-146)
  cfg.speed = s->speed;
  cfg.nonlinear_factor = s->nonlinear;
  cfg.feedback_strength = s->feedback;
  cfg.max_write_frames = kChunkFrames;
  // slow-down can expand up to 1/kMinimumSpeed = 100x (speedy.c:92)
  s->out_capacity = (long long)kChunkFrames * 102 + 8 * (s->sample_rate / 65);
  cfg.out_capacity = s->out_capacity;
  cfg.taps = SPEEDY_TAP_TENSION | SPEEDY_TAP_SPEED | SPEEDY_TAP_FEATURES | SPEEDY_TAP_SPECTROGRAM;
  cfg.threads_per_stream = 128;
  s->batch = speedyBatchCreate(&cfg);
  return s->batch != nullptr;
}

static bool drain_device_output(sonicStream s) {
  int32_t count = 0;
  s->scratch.resize((size_t)s->out_capacity * s->channels);
  if (!speedyBatchRead(s->batch, s->scratch.data(), s->out_capacity, &count)) return false;
  s->fifo.insert(s->fifo.end(), s->scratch.begin(), s->scratch.begin() + (size_t)count * s->channels);
  return true;
}

// Replay the debug callbacks for the frames the last batch write produced, in
// the order soniclib.c:297-353 fires them.
static bool replay_callbacks(sonicStream s) {
  const int rows = kChunkFrames / s->step + 2;
  int32_t n_analysis = 0, n_tension = 0;
  const bool any = s->on_tension || s->on_speed || s->on_features || s->on_spectrogram || s->on_normalized;
  s->tap_spec.resize((size_t)rows * s->fft);
  s->tap_feat.resize((size_t)rows * SPEEDY_FEATURE_COUNT);
  s->tap_tension.resize(rows);
  s->tap_speed.resize(rows);
  if (!speedyBatchGetTaps(s->batch, rows, &n_analysis, &n_tension, any ? s->tap_spec.data() : nullptr, nullptr,
                          any ? s->tap_feat.data() : nullptr, any ? s->tap_tension.data() : nullptr,
                          any ? s->tap_speed.data() : nullptr))
    return false;
  int jt = 0;
  for (int j = 0; j < n_analysis; j++) {
    const int at = ++s->spec_time;
    float* spec = s->tap_spec.data() + (size_t)j * s->fft;
    if (any) s->spec_ring[at & 31].assign(spec, spec + s->fft);
    if (s->on_spectrogram) s->on_spectrogram(s, at, spec);
    if (s->on_normalized) s->on_normalized(s, at, s->normalized.data());
    if (at - s->future >= s->tension_time && jt < n_tension) {
      const int r = s->tension_time++;
      if (any) {
        // speedy.c:673-675: the spectrum of at_time r scaled by 1/(sqrt(E)+eps)
        const float energy = s->tap_feat[(size_t)jt * SPEEDY_FEATURE_COUNT + 0];
        const float eps = 2.2204e-16f;
        const float inv = (float)(1.0 / (sqrt((double)energy) + (double)eps));
        const std::vector<float>& src = s->spec_ring[r & 31];
        for (int i = 0; i < s->fft / 2; i++) s->normalized[i] = (r >= 1 && !src.empty()) ? src[i] * inv : 0.0f;
      }
      if (s->on_tension) s->on_tension(s, r, s->tap_tension[jt]);
      if (s->on_features) s->on_features(s, r, s->tap_feat.data() + (size_t)jt * SPEEDY_FEATURE_COUNT);
      if (s->on_speed) s->on_speed(s, r, s->tap_speed[jt]);
      jt++;
    }
  }
  return true;
}

static int write_frames(sonicStream s, const short* in, int count) {
  if (!s) return 0;
  if (s->rate != 1.0f) return 0;  // playback-rate conversion is not on this path
  if (!ensure_batch(s)) return 0;
  s->started = true;
  // A flushed handle that switches between the linear short circuit and Speedy starts
  // Speedy's clock at zero, as the reference does (its Speedy stream only ever sees the
  // nonlinear writes, soniclib.c:397-399; the flushed inner Sonic FIFO is empty).  Sonic's
  // previous-period memory is the one thing not carried across.
  const int mode = s->nonlinear != 0.0f ? 1 : 0;
  if (s->last_mode >= 0 && s->last_mode != mode && s->flushed) {
    if (!speedyBatchReset(s->batch, nullptr)) return 0;
    s->spec_time = 0;
    s->tension_time = 0;
    for (auto& row : s->spec_ring) row.clear();
    std::fill(s->normalized.begin(), s->normalized.end(), 0.0f);
  }
  s->last_mode = mode;
  s->flushed = false;
  if (s->nonlinear != 0.0f && s->buffer_size == 0) s->buffer_size = s->step;
  for (int done = 0; done < count;) {
    const int n = std::min(count - done, kChunkFrames);
    if (!speedyBatchWrite(s->batch, in + (size_t)done * s->channels, n, n, nullptr)) return 0;
    if (s->nonlinear != 0.0f && !replay_callbacks(s)) return 0;
    if (!drain_device_output(s)) return 0;
    done += n;
  }
  return 1;
}

extern "C" {

sonicStream sonicCreateStream(int sampleRate, int numChannels) {
  if (sampleRate < 800 || numChannels < 1) return nullptr;
  sonicStream s = new sonicStreamStruct();
  s->sample_rate = sampleRate;
  s->channels = numChannels;
  s->speed = 1.0f;      // soniclib.c:114
  s->rate = 1.0f;
  s->nonlinear = 0.0f;  // soniclib.c:117
  s->feedback = 0.1f;   // soniclib.c:122
  s->batch = nullptr;
  s->buffer_size = 0;
  s->started = false;
  s->last_mode = -1;
  s->flushed = false;
  s->on_tension = nullptr;
  s->on_speed = nullptr;
  s->on_features = nullptr;
  s->on_spectrogram = nullptr;
  s->on_normalized = nullptr;
  s->spec_time = 0;
  s->tension_time = 0;
  s->future = 12;
  speedyBatchFrameGeometry(sampleRate, &s->window, &s->fft, &s->step);
  s->spec_ring.resize(32);
  s->normalized.assign(s->fft, 0.0f);
  if (!ensure_batch(s)) {  // fails without a CUDA device: there is no CPU fallback
    delete s;
    return nullptr;
  }
  return s;
}

void sonicDestroyStream(sonicStream s) {
  if (!s) return;
  speedyBatchDestroy(s->batch);
  delete s;
}

int sonicWriteShortToStream(sonicStream s, const short* inBuffer, int sampleCount) {
  if (!s) return 0;
  if (!inBuffer || sampleCount <= 0) return 1;
  return write_frames(s, inBuffer, sampleCount);
}

int sonicWriteFloatToStream(sonicStream s, const float* inBuffer, int sampleCount) {
  if (!s) return 0;
  if (!inBuffer || sampleCount <= 0) return 1;
  std::vector<short> tmp((size_t)sampleCount * s->channels);
  if (s->nonlinear != 0.0f) {
    // soniclib.c:496
    for (size_t i = 0; i < tmp.size(); i++) tmp[i] = (short)(inBuffer[i] * 32768.0);
  } else {
    // upstream sonicWriteFloatToStream (soniclib.c:464)
    for (size_t i = 0; i < tmp.size(); i++) tmp[i] = (short)(inBuffer[i] * 32767.0f);
  }
  return write_frames(s, tmp.data(), sampleCount);
}

int sonicReadShortFromStream(sonicStream s, short* outBuffer, int bufferSize) {
  if (!s || !outBuffer || bufferSize <= 0) return 0;
  const size_t have = s->fifo.size() / s->channels;
  const size_t n = std::min(have, (size_t)bufferSize);
  if (n == 0) return 0;
  memcpy(outBuffer, s->fifo.data(), n * s->channels * sizeof(short));
  s->fifo.erase(s->fifo.begin(), s->fifo.begin() + n * s->channels);
  return (int)n;
}

int sonicReadFloatFromStream(sonicStream s, float* outBuffer, int bufferSize) {
  if (!s || !outBuffer || bufferSize <= 0) return 0;
  const size_t have = s->fifo.size() / s->channels;
  const size_t n = std::min(have, (size_t)bufferSize);
  if (n == 0) return 0;
  for (size_t i = 0; i < n * s->channels; i++) outBuffer[i] = s->fifo[i] / 32767.0f;
  s->fifo.erase(s->fifo.begin(), s->fifo.begin() + n * s->channels);
  return (int)n;
}

void sonicSetRate(sonicStream s, float rate) {
  if (s) s->rate = rate;
}

void sonicSetSpeed(sonicStream s, float speed) {
  if (!s) return;
  s->speed = speed;
  if (s->batch) speedyBatchSetSpeed(s->batch, nullptr, speed);
}

int sonicFlushStream(sonicStream s) {
  if (!s || !ensure_batch(s)) return 0;
  if (!speedyBatchFlush(s->batch)) return 0;
  s->flushed = true;
  return drain_device_output(s) ? 1 : 0;
}

void sonicEnableNonlinearSpeedup(sonicStream s, float nonlinearFactor) {
  if (!s) return;
  s->nonlinear = nonlinearFactor;
  if (s->batch) speedyBatchSetNonlinear(s->batch, nullptr, nonlinearFactor);
}

void sonicSetDurationFeedbackStrength(sonicStream s, float factor) {
  if (!s) return;
  s->feedback = factor;
  if (s->batch) speedyBatchSetFeedback(s->batch, nullptr, factor);
}

int getSonicBufferSize(sonicStream s) { return s ? s->buffer_size : 0; }
int sonicSpectrogramSize(sonicStream s) { return s ? s->fft : 0; }

void sonicTensionCallback(sonicStream s, tensionFunction fn) { if (s) s->on_tension = fn; }
tensionFunction getSonicTensionCallback(sonicStream s) { return s ? s->on_tension : nullptr; }
void sonicSpeedCallback(sonicStream s, speedFunction fn) { if (s) s->on_speed = fn; }
speedFunction getSonicSpeedCallback(sonicStream s) { return s ? s->on_speed : nullptr; }
void sonicFeaturesCallback(sonicStream s, featuresFunction fn) { if (s) s->on_features = fn; }
featuresFunction getSonicFeaturesCallback(sonicStream s) { return s ? s->on_features : nullptr; }
void sonicSpectrogramCallback(sonicStream s, spectrogramFunction fn) { if (s) s->on_spectrogram = fn; }
spectrogramFunction getSonicSpectrogramCallback(sonicStream s) { return s ? s->on_spectrogram : nullptr; }
void sonicNormalizedSpectrogramCallback(sonicStream s, spectrogramFunction fn) { if (s) s->on_normalized = fn; }
spectrogramFunction getSonicNormalizedSpectrogramCallback(sonicStream s) {
  return s ? s->on_normalized : nullptr;
}

int sonicIntGetNumChannels(sonicStream s) { return s ? s->channels : 0; }
int sonicIntGetSampleRate(sonicStream s) { return s ? s->sample_rate : 0; }
float sonicIntGetSpeed(sonicStream s) { return s ? s->speed : 0.0f; }
int sonicIntSamplesAvailable(sonicStream s) { return s ? (int)(s->fifo.size() / s->channels) : 0; }

// The inner Sonic stream under its SONIC_INTERNAL names (sonic2.h:22-35), as
// sonic_test.cc:729-752 drives it: plain Sonic at the handle's speed.  Only valid while
// Speedy is switched off; with a nonlinear factor set these refuse (return 0) rather
// than bypass the analysis.
void sonicIntSetSpeed(sonicStream s, float speed) { sonicSetSpeed(s, speed); }
int sonicIntWriteShortToStream(sonicStream s, const short* inBuffer, int sampleCount) {
  if (!s || s->nonlinear != 0.0f) return 0;
  return sonicWriteShortToStream(s, inBuffer, sampleCount);
}
int sonicIntReadShortFromStream(sonicStream s, short* outBuffer, int bufferSize) {
  return sonicReadShortFromStream(s, outBuffer, bufferSize);
}
int sonicIntFlushStream(sonicStream s) {
  if (!s || s->nonlinear != 0.0f) return 0;
  return sonicFlushStream(s);
}

}  // extern "C"
