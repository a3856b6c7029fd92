// The Sonic/Speedy drop-in surface (include/speedy_b200.h section 1): the names
// and signatures of /root/reference/sonic2.h:54-125, implemented on top of the
// batched CUDA path with a batch of one stream.  Host logic only: parameter
// bookkeeping, the host-side output FIFO sonicRead* pops from, and the five
// debug callbacks replayed in the reference's order (soniclib.c:297-353).
//
// A handle made by sonicCreateStream owns a batch of one: correct, not fast (every write is a
// host->device copy, four kernel launches and a device->host read).  Handles opened from a
// session pool (speedy_b200.h section 1b) share one batch: writes queue in page-locked
// staging rows and one coalesced step serves every session with pending input.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <map>
#include <mutex>
#include <utility>
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
  speedySessionPool pool;  // non-null: a pooled session, `slot` of the pool's batch
  int slot;
  int window, fft, step;
  int buffer_size;  // 0 until the first nonlinear write (soniclib.c:195, 672-680)
  bool started;
  int last_mode;     // -1 nothing written yet, 0 last write was linear, 1 nonlinear
  bool flushed;      // nothing pending since the last sonicFlushStream
  long long out_capacity;
  std::vector<short> fifo;     // produced, not yet read (interleaved)
  // sonicSetRate: upstream Sonic resamples what the speed change produced (adjustRate)
  std::vector<short> rate_pitch;  // speed-changed frames not yet resampled (upstream's pitch buffer)
  int old_rate_pos, new_rate_pos;
  size_t flush_mark;              // FIFO length when the current flush began (rate != 1 only)
  size_t last_drain_frames;       // frames the last drain took from the device
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

// ---- playback rate (sonicSetRate, sonic2.h:70, soniclib.c:169-175 -> upstream sonicSetRate) ----
// Upstream Sonic changes the playback rate after the speed change, by resampling its output: the
// classic implementation's adjustRate / interpolate (linear interpolation between neighbouring
// frames, integer positions old * newRate against new * oldRate with both rates halved until they
// fit 14 bits), restated here on the host because the drop-in API hands frames out through a host
// FIFO anyway.  It is outside the hot path (no BASELINE configuration sets a rate) and, like the
// rest of upstream Sonic, unpinned: the reference clones upstream at HEAD, whose later revisions
// resample with a windowed sinc instead.
static void rate_resample(sonicStream s) {
  const int C = s->channels;
  int new_rate = (int)((float)s->sample_rate / s->rate), old_rate = s->sample_rate;
  while (new_rate > (1 << 14) || old_rate > (1 << 14)) {
    new_rate >>= 1;
    old_rate >>= 1;
  }
  if (new_rate < 1) new_rate = 1;
  if (old_rate < 1) old_rate = 1;
  const size_t n = s->rate_pitch.size() / C;
  size_t position = 0;
  for (; position + 1 < n; position++) {  // (one frame stays behind: the right neighbour of the next output)
    while ((long long)(s->old_rate_pos + 1) * new_rate > (long long)s->new_rate_pos * old_rate) {
      const short* in = s->rate_pitch.data() + position * C;
      for (int c = 0; c < C; c++) {
        const int left = in[c], right = in[c + C];
        const int pos = s->new_rate_pos * old_rate;
        const int left_pos = s->old_rate_pos * new_rate, right_pos = (s->old_rate_pos + 1) * new_rate;
        const int ratio = right_pos - pos, width = right_pos - left_pos;
        s->fifo.push_back((short)((ratio * left + (width - ratio) * right) / width));
      }
      s->new_rate_pos++;
    }
    s->old_rate_pos++;
    if (s->old_rate_pos == old_rate) {
      s->old_rate_pos = 0;
      s->new_rate_pos = 0;
    }
  }
  s->rate_pitch.erase(s->rate_pitch.begin(), s->rate_pitch.begin() + position * C);
}

// frames the device produced for this handle -> the FIFO sonicRead* pops from
static void append_output(sonicStream s, const short* src, size_t frames) {
  if (s->rate == 1.0f && s->rate_pitch.empty()) {
    s->fifo.insert(s->fifo.end(), src, src + frames * s->channels);
    return;
  }
  s->rate_pitch.insert(s->rate_pitch.end(), src, src + frames * s->channels);
  if (s->rate == 1.0f) {  // the rate went back to 1: what was waiting for a right neighbour goes out as it is
    s->fifo.insert(s->fifo.end(), s->rate_pitch.begin(), s->rate_pitch.end());
    s->rate_pitch.clear();
    return;
  }
  rate_resample(s);
}

// sonicFlushStream with a rate: upstream expects (remaining / speed + pitch frames) / rate + 0.5 more
// frames, resamples the silence it pads with as well, trims to that count and empties the pitch buffer
static void begin_rate_flush(sonicStream s) { s->flush_mark = s->fifo.size(); }
static void finish_rate_flush(sonicStream s, size_t pitch_before, size_t flushed_frames) {
  if (s->rate == 1.0f) return;
  const int C = s->channels;
  const size_t expected = (size_t)((float)(flushed_frames + pitch_before) / s->rate + 0.5f);
  const short zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int guard = 0; (s->fifo.size() - s->flush_mark) / C < expected && guard < (1 << 20); guard++) {
    for (int c = 0; c < C; c += 8) s->rate_pitch.insert(s->rate_pitch.end(), zero, zero + (C - c < 8 ? C - c : 8));
    rate_resample(s);
  }
  if ((s->fifo.size() - s->flush_mark) / C > expected) s->fifo.resize(s->flush_mark + expected * C);
  s->rate_pitch.clear();
}

static bool drain_device_output(sonicStream s) {
  int32_t count = 0;
  s->scratch.resize((size_t)s->out_capacity * s->channels);
  if (!speedyBatchRead(s->batch, s->scratch.data(), s->out_capacity, &count)) return false;
  s->last_drain_frames = (size_t)count;
  append_output(s, s->scratch.data(), (size_t)count);
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


// ---- session pool ------------------------------------------------------------

struct speedySessionPoolStruct {
  speedySessionPoolConfig cfg;
  speedyBatch batch = nullptr;
  int n = 0, channels = 1;
  long long row_frames = 0;  // staging row per session
  long long out_cap = 0;     // output room per session and step
  std::mutex mu;
  int16_t* h_in = nullptr;   // page-locked [n][row_frames][C]
  int16_t* h_out = nullptr;  // page-locked [n][out_cap][C]
  std::vector<int32_t> counts, out_counts, mask, status;
  std::vector<float> speed, nonlinear, feedback;
  bool params_dirty = false;
  std::vector<sonicStream> slots;  // null: free
  std::vector<int> free_slots;
  std::vector<char> needs_reset;   // the slot's device state is a closed session's
  int pending = 0;                 // sessions with queued input
  speedySessionPoolStats stats = {};
};

namespace {

bool pool_push_params(speedySessionPool p) {
  if (!p->params_dirty) return true;
  p->params_dirty = false;
  return speedyBatchSetSpeed(p->batch, p->speed.data(), 0.0f) &&
         speedyBatchSetNonlinear(p->batch, p->nonlinear.data(), 0.0f) &&
         speedyBatchSetFeedback(p->batch, p->feedback.data(), 0.0f);
}

// move what the batch produced into the sessions' FIFOs
bool pool_collect(speedySessionPool p) {
  if (!speedyBatchRead(p->batch, p->h_out, p->out_cap, p->out_counts.data())) return false;
  bool full = false;
  for (int i = 0; i < p->n; i++) {
    const int32_t c = p->out_counts[i];
    if (c <= 0) continue;
    sonicStream h = p->slots[i];
    if (h) {
      const short* src = p->h_out + (size_t)i * p->out_cap * p->channels;
      h->last_drain_frames = (size_t)c;
      append_output(h, src, (size_t)c);
    }
    if (c >= p->out_cap) full = true;
  }
  if (full) {
    // a row that came back full may have lost output (min_speed too optimistic)
    if (!speedyBatchGetStatus(p->batch, p->status.data())) return false;
    for (int i = 0; i < p->n; i++) {
      if (p->status[i] & (SPEEDY_STATUS_OUTPUT_OVERFLOW | SPEEDY_STATUS_READ_TRUNCATED)) return false;
    }
  }
  return true;
}

// one coalesced step over every session with queued input (mutex held); sessions served or -1
int pool_step(speedySessionPool p) {
  if (p->pending == 0) return 0;
  const auto t0 = std::chrono::steady_clock::now();
  if (!pool_push_params(p)) return -1;
  int32_t most = 0;
  for (int i = 0; i < p->n; i++) most = std::max(most, p->counts[i]);
  if (!speedyBatchWrite(p->batch, p->h_in, p->row_frames, most, p->counts.data())) return -1;
  if (!pool_collect(p)) return -1;
  const int served = p->pending;
  std::fill(p->counts.begin(), p->counts.end(), 0);
  p->pending = 0;
  p->stats.steps++;
  p->stats.sessions_served += served;
  p->stats.last_step_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  return served;
}

bool pool_step_if_pending(sonicStream s) {
  speedySessionPool p = s->pool;
  return p->counts[s->slot] == 0 || pool_step(p) >= 0;
}

bool pool_reset_slot(speedySessionPool p, int slot) {
  std::fill(p->mask.begin(), p->mask.end(), 0);
  p->mask[slot] = 1;
  return speedyBatchResetStreams(p->batch, p->mask.data()) != 0;
}

int pool_write(sonicStream s, const short* in, int count) {
  speedySessionPool p = s->pool;
  std::lock_guard<std::mutex> lock(p->mu);
  s->started = true;
  // (the same rule as a private handle: a flushed session that switches between the linear
  // short circuit and Speedy starts Speedy's clock at zero, soniclib.c:397-399)
  const int mode = s->nonlinear != 0.0f ? 1 : 0;
  if (s->last_mode >= 0 && s->last_mode != mode && s->flushed) {
    if (!pool_step_if_pending(s) || !pool_reset_slot(p, s->slot)) return 0;
  }
  s->last_mode = mode;
  s->flushed = false;
  if (mode && s->buffer_size == 0) s->buffer_size = s->step;
  p->stats.session_writes++;
  for (int done = 0; done < count;) {
    if (p->counts[s->slot] == p->row_frames && pool_step(p) < 0) return 0;
    int32_t& have = p->counts[s->slot];
    const int n = (int)std::min<long long>(count - done, p->row_frames - have);
    memcpy(p->h_in + ((size_t)s->slot * p->row_frames + have) * p->channels, in + (size_t)done * p->channels,
           (size_t)n * p->channels * sizeof(short));
    if (have == 0) p->pending++;
    have += n;
    done += n;
  }
  if (p->cfg.auto_step_sessions > 0 && p->pending >= p->cfg.auto_step_sessions && pool_step(p) < 0) return 0;
  return 1;
}

int pool_flush(sonicStream s) {
  speedySessionPool p = s->pool;
  std::lock_guard<std::mutex> lock(p->mu);
  if (pool_step(p) < 0) return 0;  // everything queued goes first, so that the read below is this flush's only
  if (!pool_push_params(p)) return 0;
  std::fill(p->mask.begin(), p->mask.end(), 0);
  p->mask[s->slot] = 1;
  if (!speedyBatchFlushStreams(p->batch, p->mask.data())) return 0;
  s->flushed = true;
  const size_t pitch_before = s->rate_pitch.size() / s->channels;
  begin_rate_flush(s);
  s->last_drain_frames = 0;
  if (!pool_collect(p)) return 0;
  finish_rate_flush(s, pitch_before, s->last_drain_frames);
  return 1;
}

void pool_set_param(sonicStream s, std::vector<float> speedySessionPoolStruct::*field, float value) {
  speedySessionPool p = s->pool;
  std::lock_guard<std::mutex> lock(p->mu);
  if ((p->*field)[s->slot] == value) return;
  pool_step_if_pending(s);  // queued samples are processed under the parameters they were written with
  (p->*field)[s->slot] = value;
  p->params_dirty = true;
}

void pool_close(sonicStream s) {
  speedySessionPool p = s->pool;
  std::lock_guard<std::mutex> lock(p->mu);
  if (p->counts[s->slot] > 0) {  // queued input of a closed session is dropped, as the reference's destroy does
    p->counts[s->slot] = 0;
    p->pending--;
  }
  p->slots[s->slot] = nullptr;
  p->needs_reset[s->slot] = 1;
  p->free_slots.push_back(s->slot);
  p->stats.open_sessions--;
}

// SPEEDY_B200_POOL_SESSIONS=<n>: sonicCreateStream opens from an implicit pool per (rate, channels)
std::mutex g_implicit_mu;
std::map<std::pair<int, int>, speedySessionPool> g_implicit_pools;

sonicStream implicit_pool_open(int rate, int channels) {
  const char* e = getenv("SPEEDY_B200_POOL_SESSIONS");
  const int n = e ? atoi(e) : 0;
  if (n <= 0) return nullptr;
  std::lock_guard<std::mutex> lock(g_implicit_mu);
  speedySessionPool& p = g_implicit_pools[{rate, channels}];
  if (!p) {
    speedySessionPoolConfig cfg;
    speedySessionPoolDefaultConfig(&cfg);
    cfg.sample_rate = rate;
    cfg.num_channels = channels;
    cfg.max_sessions = n;
    if (const char* d = getenv("SPEEDY_B200_POOL_DEVICE")) cfg.device = atoi(d);
    p = speedySessionPoolCreate(&cfg);
    if (!p) return nullptr;
  }
  return speedySessionPoolOpen(p);  // null when full: the caller falls back to a private batch
}

}  // namespace

static int write_frames(sonicStream s, const short* in, int count) {
  if (!s) return 0;
  if (s->pool) return pool_write(s, in, count);
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

static sonicStream new_handle(int sampleRate, int numChannels);

sonicStream sonicCreateStream(int sampleRate, int numChannels) {
  if (sampleRate < 800 || numChannels < 1) return nullptr;
  if (sonicStream pooled = implicit_pool_open(sampleRate, numChannels)) return pooled;
  sonicStream s = new_handle(sampleRate, numChannels);
  if (!ensure_batch(s)) {  // fails without a CUDA device: there is no CPU fallback
    delete s;
    return nullptr;
  }
  return s;
}

static sonicStream new_handle(int sampleRate, int numChannels) {
  sonicStream s = new sonicStreamStruct();
  s->pool = nullptr;
  s->slot = -1;
  s->sample_rate = sampleRate;
  s->channels = numChannels;
  s->speed = 1.0f;      // soniclib.c:114
  s->rate = 1.0f;
  s->old_rate_pos = 0;
  s->new_rate_pos = 0;
  s->flush_mark = 0;
  s->last_drain_frames = 0;
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
  return s;
}

void sonicDestroyStream(sonicStream s) {
  if (!s) return;
  if (s->pool) pool_close(s);
  else speedyBatchDestroy(s->batch);
  delete s;
}

void speedySessionPoolDefaultConfig(speedySessionPoolConfig* cfg) {
  if (!cfg) return;
  memset(cfg, 0, sizeof(*cfg));
  cfg->sample_rate = 16000;
  cfg->num_channels = 1;
  cfg->max_sessions = 1024;
  cfg->device = 0;
  cfg->max_pending_frames = 0;  // 100 ms
  cfg->min_speed = 0.25f;
  cfg->auto_step_sessions = 0;
}

speedySessionPool speedySessionPoolCreate(const speedySessionPoolConfig* c) {
  if (!c || c->sample_rate < 800 || c->num_channels < 1 || c->max_sessions < 1) return nullptr;
  speedySessionPool p = new speedySessionPoolStruct();
  p->cfg = *c;
  p->n = c->max_sessions;
  p->channels = c->num_channels;
  p->row_frames = c->max_pending_frames > 0 ? c->max_pending_frames : c->sample_rate / 10;
  const float min_speed = c->min_speed > 0.0f ? std::max(c->min_speed, 0.01f) : 0.25f;  // speedy.c:92: speeds stop at 0.01
  // one step can emit what it was fed plus what Sonic's FIFO still held (it never keeps more
  // than it needs for one search and one 10 ms buffer), stretched by 1 / speed, plus a flush's padding
  const long long max_required = 2LL * (c->sample_rate / 65);
  p->out_cap = (long long)((double)(p->row_frames + 2 * max_required + c->sample_rate / 100) / min_speed) + 2 * max_required;
  speedyBatchConfig cfg;
  speedyBatchDefaultConfig(&cfg);
  cfg.sample_rate = c->sample_rate;
  cfg.num_channels = c->num_channels;
  cfg.num_streams = c->max_sessions;
  cfg.device = c->device;
  cfg.match_matlab = 0;
  cfg.max_write_frames = p->row_frames;
  cfg.out_capacity = p->out_cap;
  cfg.taps = 0;
  p->batch = speedyBatchCreate(&cfg);
  const size_t in_bytes = (size_t)p->n * p->row_frames * p->channels * sizeof(int16_t);
  const size_t out_bytes = (size_t)p->n * p->out_cap * p->channels * sizeof(int16_t);
  if (p->batch) {
    p->h_in = (int16_t*)speedyBatchHostAlloc(in_bytes, 0);
    p->h_out = (int16_t*)speedyBatchHostAlloc(out_bytes, 0);
  }
  if (!p->batch || !p->h_in || !p->h_out) {
    speedySessionPoolDestroy(p);
    return nullptr;
  }
  p->counts.assign(p->n, 0);
  p->out_counts.assign(p->n, 0);
  p->mask.assign(p->n, 0);
  p->status.assign(p->n, 0);
  p->speed.assign(p->n, 1.0f);      // soniclib.c:114
  p->nonlinear.assign(p->n, 0.0f);  // soniclib.c:117
  p->feedback.assign(p->n, 0.1f);   // soniclib.c:122
  p->params_dirty = true;
  p->slots.assign(p->n, nullptr);
  p->needs_reset.assign(p->n, 0);
  for (int i = p->n - 1; i >= 0; i--) p->free_slots.push_back(i);
  return p;
}

void speedySessionPoolDestroy(speedySessionPool p) {
  if (!p) return;
  for (sonicStream h : p->slots) {
    if (h) delete h;
  }
  if (p->h_in) speedyBatchHostFree(p->h_in);
  if (p->h_out) speedyBatchHostFree(p->h_out);
  speedyBatchDestroy(p->batch);
  delete p;
}

sonicStream speedySessionPoolOpen(speedySessionPool p) {
  if (!p) return nullptr;
  std::lock_guard<std::mutex> lock(p->mu);
  if (p->free_slots.empty()) return nullptr;
  const int slot = p->free_slots.back();
  if (p->needs_reset[slot]) {
    if (p->counts[slot] != 0 || !pool_reset_slot(p, slot)) return nullptr;
    p->needs_reset[slot] = 0;
  }
  p->free_slots.pop_back();
  sonicStream s = new_handle(p->cfg.sample_rate, p->cfg.num_channels);
  s->pool = p;
  s->slot = slot;
  p->slots[slot] = s;
  if (p->speed[slot] != 1.0f || p->nonlinear[slot] != 0.0f || p->feedback[slot] != 0.1f) p->params_dirty = true;
  p->speed[slot] = 1.0f;
  p->nonlinear[slot] = 0.0f;
  p->feedback[slot] = 0.1f;
  p->stats.open_sessions++;
  return s;
}

int speedySessionPoolStep(speedySessionPool p) {
  if (!p) return -1;
  std::lock_guard<std::mutex> lock(p->mu);
  return pool_step(p);
}

int speedySessionPoolGetStats(speedySessionPool p, speedySessionPoolStats* stats) {
  if (!p || !stats) return 0;
  std::lock_guard<std::mutex> lock(p->mu);
  *stats = p->stats;
  stats->pending_sessions = p->pending;
  return 1;
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

// a pooled session with queued input is served before its FIFO is looked at
static bool settle(sonicStream s) {
  if (!s->pool) return true;
  std::lock_guard<std::mutex> lock(s->pool->mu);
  return pool_step_if_pending(s);
}

int sonicReadShortFromStream(sonicStream s, short* outBuffer, int bufferSize) {
  if (!s || !outBuffer || bufferSize <= 0 || !settle(s)) return 0;
  const size_t have = s->fifo.size() / s->channels;
  const size_t n = std::min(have, (size_t)bufferSize);
  if (n == 0) return 0;
  memcpy(outBuffer, s->fifo.data(), n * s->channels * sizeof(short));
  s->fifo.erase(s->fifo.begin(), s->fifo.begin() + n * s->channels);
  return (int)n;
}

int sonicReadFloatFromStream(sonicStream s, float* outBuffer, int bufferSize) {
  if (!s || !outBuffer || bufferSize <= 0 || !settle(s)) return 0;
  const size_t have = s->fifo.size() / s->channels;
  const size_t n = std::min(have, (size_t)bufferSize);
  if (n == 0) return 0;
  for (size_t i = 0; i < n * s->channels; i++) outBuffer[i] = s->fifo[i] / 32767.0f;
  s->fifo.erase(s->fifo.begin(), s->fifo.begin() + n * s->channels);
  return (int)n;
}

void sonicSetRate(sonicStream s, float rate) {
  if (!s || !(rate > 0.0f)) return;
  std::unique_lock<std::mutex> lock;
  if (s->pool) {  // queued samples are processed (and resampled) at the rate they were written with
    lock = std::unique_lock<std::mutex>(s->pool->mu);
    pool_step_if_pending(s);
  }
  s->rate = rate;
  s->old_rate_pos = 0;  // upstream sonicSetRate
  s->new_rate_pos = 0;
}

void sonicSetSpeed(sonicStream s, float speed) {
  if (!s) return;
  s->speed = speed;
  if (s->pool) pool_set_param(s, &speedySessionPoolStruct::speed, speed);
  else if (s->batch) speedyBatchSetSpeed(s->batch, nullptr, speed);
}

int sonicFlushStream(sonicStream s) {
  if (s && s->pool) return pool_flush(s);
  if (!s || !ensure_batch(s)) return 0;
  if (!speedyBatchFlush(s->batch)) return 0;
  s->flushed = true;
  const size_t pitch_before = s->rate_pitch.size() / s->channels;
  begin_rate_flush(s);
  if (!drain_device_output(s)) return 0;
  finish_rate_flush(s, pitch_before, s->last_drain_frames);
  return 1;
}

void sonicEnableNonlinearSpeedup(sonicStream s, float nonlinearFactor) {
  if (!s) return;
  s->nonlinear = nonlinearFactor;
  if (s->pool) pool_set_param(s, &speedySessionPoolStruct::nonlinear, nonlinearFactor);
  else if (s->batch) speedyBatchSetNonlinear(s->batch, nullptr, nonlinearFactor);
}

void sonicSetDurationFeedbackStrength(sonicStream s, float factor) {
  if (!s) return;
  s->feedback = factor;
  if (s->pool) pool_set_param(s, &speedySessionPoolStruct::feedback, factor);
  else if (s->batch) speedyBatchSetFeedback(s->batch, nullptr, factor);
}

int getSonicBufferSize(sonicStream s) { return s ? s->buffer_size : 0; }
int sonicSpectrogramSize(sonicStream s) { return s ? s->fft : 0; }

void sonicTensionCallback(sonicStream s, tensionFunction fn) { if (s && !s->pool) s->on_tension = fn; }
tensionFunction getSonicTensionCallback(sonicStream s) { return s ? s->on_tension : nullptr; }
void sonicSpeedCallback(sonicStream s, speedFunction fn) { if (s && !s->pool) s->on_speed = fn; }
speedFunction getSonicSpeedCallback(sonicStream s) { return s ? s->on_speed : nullptr; }
void sonicFeaturesCallback(sonicStream s, featuresFunction fn) { if (s && !s->pool) s->on_features = fn; }
featuresFunction getSonicFeaturesCallback(sonicStream s) { return s ? s->on_features : nullptr; }
void sonicSpectrogramCallback(sonicStream s, spectrogramFunction fn) { if (s && !s->pool) s->on_spectrogram = fn; }
spectrogramFunction getSonicSpectrogramCallback(sonicStream s) { return s ? s->on_spectrogram : nullptr; }
void sonicNormalizedSpectrogramCallback(sonicStream s, spectrogramFunction fn) { if (s && !s->pool) s->on_normalized = fn; }
spectrogramFunction getSonicNormalizedSpectrogramCallback(sonicStream s) {
  return s ? s->on_normalized : nullptr;
}

int sonicIntGetNumChannels(sonicStream s) { return s ? s->channels : 0; }
int sonicIntGetSampleRate(sonicStream s) { return s ? s->sample_rate : 0; }
float sonicIntGetSpeed(sonicStream s) { return s ? s->speed : 0.0f; }
int sonicIntSamplesAvailable(sonicStream s) { return s && settle(s) ? (int)(s->fifo.size() / s->channels) : 0; }

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
