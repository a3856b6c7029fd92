// stream_bench -- BASELINE.json config 5 through the API it names: N live sessions, each a
// sonicStream handle, every session fed one 10 ms chunk per tick with
// sonicWriteShortToStream and drained with sonicReadShortFromStream (soniclib.c:391-452,
// 519-522), exactly as a libsonic client would.  The handles come from a session pool
// (speedy_b200.h section 1b), so a tick is one coalesced device step.  Reports the latency
// of a tick (first write of the tick -> every session's output read back: what the slowest
// session of the tick waits), p50 / p90 / p99 / max, and the real-time factor.
//
//   stream_bench [--sessions 16384] [--rate 16000] [--chunk 160] [--ticks 400] [--warmup 100]
//                [--speed 2.0] [--nonlinear 1.0] [--feedback 0.1] [--device 0] [--implicit]
//
// --implicit: open the handles with plain sonicCreateStream under
// SPEEDY_B200_POOL_SESSIONS (an unmodified client) instead of speedySessionPoolOpen.
// Host C++ only; links libspeedy_b200.so; prints one JSON object.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include "../include/speedy_b200.h"
#include "../speedy_b200/csrc/synth.h"

static double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char** argv) {
  int sessions = 16384, rate = 16000, chunk = 160, ticks = 400, warmup = 100, device = 0;
  float speed = 2.0f, nonlinear = 1.0f, feedback = 0.1f;
  bool implicit = false;
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    auto next = [&]() { return i + 1 < argc ? argv[++i] : "0"; };
    if (a == "--sessions") sessions = atoi(next());
    else if (a == "--rate") rate = atoi(next());
    else if (a == "--chunk") chunk = atoi(next());
    else if (a == "--ticks") ticks = atoi(next());
    else if (a == "--warmup") warmup = atoi(next());
    else if (a == "--device") device = atoi(next());
    else if (a == "--speed") speed = (float)atof(next());
    else if (a == "--nonlinear") nonlinear = (float)atof(next());
    else if (a == "--feedback") feedback = (float)atof(next());
    else if (a == "--implicit") implicit = true;
    else { fprintf(stderr, "unknown flag %s\n", a.c_str()); return 2; }
  }
  // half a second of synthetic speech per session, cycled (the generator is the bench's own, synth.h)
  const int loop_ticks = std::max(1, rate / 2 / chunk);
  const size_t per_session = (size_t)loop_ticks * chunk;
  std::vector<short> pcm((size_t)sessions * per_session);
  {
    const unsigned nt = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; t++) {
      th.emplace_back([&, t]() {
        for (int s = (int)t; s < sessions; s += (int)nt) {
          short* row = pcm.data() + (size_t)s * per_session;
          for (size_t n = 0; n < per_session; n++) row[n] = synth_sample((uint64_t)s, rate, 1, 0, (int64_t)n);
        }
      });
    }
    for (auto& x : th) x.join();
  }

  speedySessionPool pool = nullptr;
  if (implicit) {
    setenv("SPEEDY_B200_POOL_SESSIONS", std::to_string(sessions).c_str(), 1);
    setenv("SPEEDY_B200_POOL_DEVICE", std::to_string(device).c_str(), 1);
  } else {
    speedySessionPoolConfig cfg;
    speedySessionPoolDefaultConfig(&cfg);
    cfg.sample_rate = rate;
    cfg.num_channels = 1;
    cfg.max_sessions = sessions;
    cfg.device = device;
    cfg.max_pending_frames = 2 * chunk;
    cfg.min_speed = 1.0f;
    pool = speedySessionPoolCreate(&cfg);
    if (!pool) { fprintf(stderr, "pool: %s\n", speedyBatchLastError()); return 1; }
  }
  std::vector<sonicStream> hs(sessions);
  for (int s = 0; s < sessions; s++) {
    hs[s] = implicit ? sonicCreateStream(rate, 1) : speedySessionPoolOpen(pool);
    if (!hs[s]) { fprintf(stderr, "open failed at session %d: %s\n", s, speedyBatchLastError()); return 1; }
    sonicSetSpeed(hs[s], speed);
    sonicEnableNonlinearSpeedup(hs[s], nonlinear);
    sonicSetDurationFeedbackStrength(hs[s], feedback);
  }
  std::vector<short> out(8192);
  std::vector<double> lat;
  lat.reserve(ticks);
  long long out_frames = 0, checksum = 0;
  const int64_t launches0 = speedyBatchKernelLaunches();
  double t_begin = 0.0;
  for (int k = 0; k < warmup + ticks; k++) {
    if (k == warmup) t_begin = now_ms();
    const size_t off = (size_t)(k % loop_ticks) * chunk;
    const double t0 = now_ms();
    for (int s = 0; s < sessions; s++) {
      if (!sonicWriteShortToStream(hs[s], pcm.data() + (size_t)s * per_session + off, chunk)) {
        fprintf(stderr, "write failed: %s\n", speedyBatchLastError());
        return 1;
      }
    }
    for (int s = 0; s < sessions; s++) {
      int n;
      while ((n = sonicReadShortFromStream(hs[s], out.data(), (int)out.size())) > 0) {
        if (k >= warmup) {
          out_frames += n;
          checksum += out[0] + out[n - 1];
        }
      }
    }
    if (k >= warmup) lat.push_back(now_ms() - t0);
  }
  const double total_ms = now_ms() - t_begin;
  const int64_t launches = speedyBatchKernelLaunches() - launches0;
  speedySessionPoolStats st;
  memset(&st, 0, sizeof(st));
  if (pool) speedySessionPoolGetStats(pool, &st);
  std::sort(lat.begin(), lat.end());
  auto pick = [&](double q) { return lat[std::min(lat.size() - 1, (size_t)(q * lat.size()))]; };
  const double mean = total_ms / ticks;
  const double tick_audio_ms = 1e3 * chunk / rate;
  printf("{\"api\": \"sonicWriteShortToStream + sonicReadShortFromStream on %s handles\", \"sessions\": %d, "
         "\"chunk_frames\": %d, \"ticks\": %d, \"warmup_ticks\": %d, \"tick_ms\": {\"p50\": %.4f, \"p90\": %.4f, "
         "\"p99\": %.4f, \"max\": %.4f, \"mean\": %.4f}, \"real_time_budget_ms\": %.3f, \"real_time_factor\": %.3f, "
         "\"audio_s_per_s\": %.1f, \"out_frames_per_tick\": %.1f, \"kernel_launches_per_tick\": %.2f, "
         "\"pool_steps\": %lld, \"pool_last_step_ms\": %.4f, \"checksum\": %lld}\n",
         implicit ? "sonicCreateStream (SPEEDY_B200_POOL_SESSIONS)" : "speedySessionPoolOpen", sessions, chunk, ticks,
         warmup, pick(0.50), pick(0.90), pick(0.99), lat.back(), mean, tick_audio_ms, tick_audio_ms / mean,
         sessions * (chunk / (double)rate) / (mean / 1e3), out_frames / (double)ticks, launches / (double)(warmup + ticks),
         (long long)st.steps, st.last_step_ms, checksum);
  for (sonicStream h : hs) sonicDestroyStream(h);
  if (pool) speedySessionPoolDestroy(pool);
  return 0;
}
