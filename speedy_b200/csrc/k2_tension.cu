// K2/K3 — per-stream recurrences across frames: energy low-pass and compression,
// temporal hysteresis, low-energy gate, difference low-pass, tension, speed with
// duration feedback, nonlinear blend.
//
// Replaces speedyComputeLocalEnergy (speedy.c:517-522), speedyEvaluateHysteresis
// (:590-610), the scalar part of speedyComputeSpectralDifference (:672-703,
// :720-728), speedyComputeTension (:752-766), speedyComputeSpeedFromTension
// (:768-788) and the blend in sonicSendDataToSpeedy (soniclib.c:339-345).
//
// Every expression keeps the reference's C evaluation types (float vs double
// promotion) and rounding (no FMA contraction: this file is compiled with
// --fmad=false and uses explicit _rn intrinsics), so that given the same frame
// energies and raw spectral differences the speeds are bit-identical.
//
// One warp owns one stream and walks its new frames in tiles of 32 (lane = frame).
// Three things are strictly sequential in the frame index — the two one-pole
// filters (speedy.c:73-76) and the duration feedback (speedy.c:778-785); they run
// as 32-step warp-uniform chains fed by shuffles, in the reference's order, so no
// rounding changes.  Everything else (compression, the 21-tap hysteresis maximum,
// gates, tension, the speed law) is evaluated by all 32 lanes at once from a
// 64-entry shared-memory ring per stream.  "AddData" of at_time a and
// "ComputeTension" of r = a - Future stay interleaved exactly as
// soniclib.c:295-371 does: within a tile, lane i handles both for a = a0 + i.
#include "kernels.cuh"

namespace speedy {

namespace {
constexpr int kWarps = 4;       // streams per CTA
constexpr int kSmemRing = 64;   // >= 32 (tile) + Future + Past + 1
}  // namespace

__global__ void __launch_bounds__(kWarps * 32) k2_tension(K2Params p, float alpha) {
  __shared__ float s_comp[kWarps][kSmemRing];
  __shared__ float s_energy[kWarps][kSmemRing];
  __shared__ float s_lsd[kWarps][kSmemRing];
  // the hysteresis triangles (speedy.c:590-610): (F - i) / F and (B - i) / B, the same correctly rounded
  // quotients the per-tap divisions gave, computed once instead of 34 times per frame
  __shared__ float s_tri_f[32], s_tri_b[32];
  if (threadIdx.x < 32) {
    const int i = threadIdx.x;
    s_tri_f[i] = i <= p.g.future ? __fdiv_rn((float)(p.g.future - i), (float)p.g.future) : 0.0f;
    s_tri_b[i] = i <= p.g.past ? __fdiv_rn((float)(p.g.past - i), (float)p.g.past) : 0.0f;
  }
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int s = blockIdx.x * kWarps + warp;
  if (s >= p.n_streams) return;
  const Geometry& g = p.g;
  const float nonlinear = p.st.nonlinear[s];
  if (nonlinear == 0.0f) return;  // soniclib.c:397-399: Speedy is bypassed

  const Range rg = write_range(p.st.total, p.counts, p.frames, p.done, s);
  const int kA = frames_analyzed(g, rg.t_old);   // scratch rows count from here
  const int kD = frames_analyzed(g, rg.t_done);  // first window of this launch
  const int kB = frames_analyzed(g, rg.t_new);
  const int rA = tensions_ready(g, kA);
  const int F = g.future, B = g.past;
  if (kB == kD) return;

  float lp_e = p.st.lp_energy[s];
  float lp_d = p.st.lp_diff[s];
  float cur_dur = p.st.cur_dur[s];
  float des_dur = p.st.des_dur[s];
  const float Rg = p.st.speed[s];
  const float fb = p.st.feedback[s];
  const float one_minus_alpha = __fsub_rn(1.0f, alpha);

  // speedy.c:263-267
  const float mean_lpf = 123.979f;
  const float mean_rsd = 0.971975f;
  const float max_hyst = 1.41421f;
  const float low_thr = (float)(0.04 * (double)max_hyst);  // :682
  const float change_cap = __fmul_rn(4.0f, mean_rsd);       // :728
  const float frame_duration = (float)(1.0 / 100.0);        // :783
  const float des_step = __fdiv_rn(frame_duration, Rg);

  float* rc = s_comp[warp];
  float* re = s_energy[warp];
  float* rl = s_lsd[warp];
  // bring in the carried ring: the last kRing at_times (slot = at_time & 31)
  {
    const size_t base = (size_t)s * kRing;
    const int a_last = kD - 1 + g.time_base;  // newest at_time already stored
    // entry for at_time t lives at ring[t & 63]; only t in (a_last-32, a_last] exist
    const int t = a_last - lane;
    const bool ok = t >= g.time_base;
    const int gslot = t & (kRing - 1);
    const float c = ok ? p.st.ring_comp[base + gslot] : 0.0f;
    const float e = ok ? p.st.ring_energy[base + gslot] : 0.0f;
    const float l = ok ? p.st.ring_lsd[base + gslot] : 0.0f;
    // at_times <= 0 read as zero (the reference's rings start zeroed)
    rc[t & (kSmemRing - 1)] = c;
    re[t & (kSmemRing - 1)] = e;
    rl[t & (kSmemRing - 1)] = l;
    rc[(t - 32) & (kSmemRing - 1)] = 0.0f;
    re[(t - 32) & (kSmemRing - 1)] = 0.0f;
    rl[(t - 32) & (kSmemRing - 1)] = 0.0f;
  }
  __syncwarp();

  const float2* feat = p.feat + (size_t)s * p.feat_stride;
  float* speeds = p.speeds + (size_t)s * p.speeds_stride;

  for (int j0 = kD - kA; j0 < kB - kA; j0 += 32) {
    const int j = j0 + lane;
    const bool have = j < kB - kA;
    const int a = kA + j + g.time_base;  // at_time of window kA + j (soniclib.c:296: k + 1)
    const float2 ef = have ? feat[j] : make_float2(0.0f, 0.0f);

    // ---- energy low-pass, sequential (speedy.c:517-518) ----------------------
    float my_lp = 0.0f;
    const int n_have = min(32, kB - kA - j0);
#pragma unroll 8
    for (int i = 0; i < n_have; i++) {
      const float e_i = __shfl_sync(0xffffffffu, ef.x, i);
      lp_e = __fadd_rn(__fmul_rn(one_minus_alpha, e_i), __fmul_rn(alpha, lp_e));
      if (i == lane) my_lp = lp_e;
    }
    // ---- compression (speedy.c:519-520), ring update ---------------------------
    float local = 0.0f, comp = 0.0f;
    if (have) {
      local = __fdiv_rn(ef.x, my_lp);
      comp = (float)sqrt(local > 2.0f ? 2.0 : (double)local);
      rc[a & (kSmemRing - 1)] = comp;
      re[a & (kSmemRing - 1)] = ef.x;
      rl[a & (kSmemRing - 1)] = ef.y;
      if (p.tap_energy) p.tap_energy[(size_t)s * p.max_new_frames + j] = ef.x;
    }
    __syncwarp();

    // ---- tension frame r = a - Future (speedy.c:755) ---------------------------
    const int r = a - F;
    const bool live = have && r >= 0;
    float hyst = 0.0f, e_r = 0.0f, lsd_raw = 0.0f;
    bool low = true;
    if (live) {
      // speedy.c:590-610 (at_times <= 0 hold the initial zeros)
      float future_max = 0.0f, past_max = 0.0f;
      for (int i = 0; i <= F; i++) {
        const int t = r + i;
        float v = t >= g.time_base ? rc[t & (kSmemRing - 1)] : 0.0f;
        v = __fmul_rn(v, s_tri_f[i]);
        if (v > future_max) future_max = v;
      }
      for (int i = 0; i <= B; i++) {
        const int t = r - i;
        float v = t >= g.time_base ? rc[t & (kSmemRing - 1)] : 0.0f;
        v = __fmul_rn(v, s_tri_b[i]);
        if (v > past_max) past_max = v;
      }
      hyst = __fmul_rn(__fadd_rn(past_max, future_max), 0.5f);
      // speedy.c:673-703: spectrum of at_time r; at_time 0 is the all-zero row
      e_r = r >= g.time_base ? re[r & (kSmemRing - 1)] : 0.0f;
      lsd_raw = r >= g.time_base ? rl[r & (kSmemRing - 1)] : 0.0f;
      low = e_r <= low_thr;
    }
    const float lsd = low ? 0.0f : lsd_raw;
    const float ewld = low ? 0.0f : __fmul_rn(lsd, hyst);  // :720

    // ---- difference low-pass, sequential (speedy.c:698-699, 722-724) -----------
    float my_lpd = 0.0f;
#pragma unroll 8
    for (int i = 0; i < n_have; i++) {
      const float x_i = __shfl_sync(0xffffffffu, ewld, i);
      const int live_i = __shfl_sync(0xffffffffu, (int)live, i);
      if (live_i) lp_d = __fadd_rn(__fmul_rn(one_minus_alpha, x_i), __fmul_rn(alpha, lp_d));
      if (i == lane) my_lpd = lp_d;
    }
    float rel = 0.0f, changes = 0.0f;
    if (live && !low) {
      rel = (float)((double)ewld / ((double)my_lpd + 0.01 * (double)mean_lpf));  // :725
      changes = (float)fmin((double)rel, (double)change_cap);                   // :727
    }
    // speedy.c:754-762
    const float tension = __fadd_rn(__fmul_rn(0.5f, __fsub_rn(hyst, 0.7f)),
                                    __fmul_rn(0.25f, __fsub_rn(changes, 1.0f)));
    // speedy.c:773-777
    float v0;
    const float slope = __fmul_rn(__fsub_rn(1.0f, Rg), tension);
    if ((double)Rg > 1.0) {
      v0 = (float)fmax(1.0, (double)__fadd_rn(Rg, slope));
    } else {
      v0 = (float)fmax(0.01, fmin(1.0, (double)__fsub_rn(Rg, slope)));
    }
    // ---- duration feedback, sequential (speedy.c:778-785) ----------------------
    //   requested += fmax(0.01, strength * excess)   (double)      excess = current - desired duration
    //   current += frame / requested;  desired += frame / R_g
    // Only `excess` carries from frame to frame.  Everything that does not depend on it is worked out by
    // the 32 lanes beforehand, off the chain: fmax picks the constant 0.01 unless strength * excess exceeds
    // it (a float exceeds the double 0.01 iff it exceeds the largest float below it, 0x3C23D70A), and then
    // the frame's speed w = (float)((double)v0 + 0.01) and its duration frame / w are known in advance;
    // otherwise the double sum of the two floats is exact, so its rounding is the float sum's.  The chain
    // keeps a subtract, a multiply, a compare and an add per frame (and a division in the rare branch).
    const float fb_floor = __uint_as_float(0x3C23D70Au);
    const float w0 = fb > 0.0f ? (float)((double)v0 + 0.01) : v0;
    const float q0 = __fdiv_rn(frame_duration, w0);
    float v = v0;
#pragma unroll 8
    for (int i = 0; i < n_have; i++) {
      const int live_i = __shfl_sync(0xffffffffu, (int)live, i);
      float v_i = __shfl_sync(0xffffffffu, w0, i);
      float q_i = __shfl_sync(0xffffffffu, q0, i);
      const float v0_i = __shfl_sync(0xffffffffu, v0, i);
      if (live_i) {
        if (fb > 0.0f) {
          const float fbx = __fmul_rn(fb, __fsub_rn(cur_dur, des_dur));
          if (fbx > fb_floor) {
            v_i = __fadd_rn(v0_i, fbx);
            q_i = __fdiv_rn(frame_duration, v_i);
          }
        }
        cur_dur = __fadd_rn(cur_dur, q_i);
        des_dur = __fadd_rn(des_dur, des_step);
        if (i == lane) v = v_i;
      }
    }
    if (live) {
      // soniclib.c:343-345
      float rate = __fadd_rn(__fmul_rn(v, nonlinear), __fmul_rn(Rg, __fsub_rn(1.0f, nonlinear)));
      const int jr = r - rA;  // index among this write's new tensions
      // (a stream that runs past the rows it was given falls back to its own speeds)
      if (p.override_speeds && r < p.override_stride) rate = p.override_speeds[(size_t)s * p.override_stride + r];
      speeds[jr] = rate;
      if (p.tap_tension) p.tap_tension[(size_t)s * p.max_new_frames + jr] = tension;
      if (p.tap_speed) p.tap_speed[(size_t)s * p.max_new_frames + jr] = rate;
      if (p.tap_features) {
        float* f = p.tap_features + ((size_t)s * p.max_new_frames + jr) * kFeatureCount;
        f[0] = e_r;  f[1] = my_lp;  f[2] = local;  f[3] = comp;  f[4] = hyst;
        f[5] = low ? 1.0f : 0.0f;  f[6] = lsd;  f[7] = ewld;  f[8] = my_lpd;  f[9] = rel;
        f[10] = changes;  f[11] = tension;  f[12] = (float)a;  f[13] = (float)r;  f[14] = low_thr;
      }
    }
    __syncwarp();
  }

  // carry the newest kRing at_times and the recurrences to the next write
  {
    const size_t base = (size_t)s * kRing;
    const int t = kB - 1 + g.time_base - lane;  // the newest 32 at_times
    if (t >= g.time_base) {
      p.st.ring_comp[base + (t & (kRing - 1))] = rc[t & (kSmemRing - 1)];
      p.st.ring_energy[base + (t & (kRing - 1))] = re[t & (kSmemRing - 1)];
      p.st.ring_lsd[base + (t & (kRing - 1))] = rl[t & (kSmemRing - 1)];
    }
  }
  if (lane == 0) {
    p.st.lp_energy[s] = lp_e;
    p.st.lp_diff[s] = lp_d;
    p.st.cur_dur[s] = cur_dur;
    p.st.des_dur[s] = des_dur;
  }
}

cudaError_t launch_k2(const K2Params& p, cudaStream_t stream) {
  if (p.max_new_frames <= 0) return cudaSuccess;
  // speedy.c:67: alpha = exp(-1.0 / time_constant), time constant 100 frames
  const float alpha = (float)exp(-1.0 / (double)100.0f);
  const int blocks = (p.n_streams + kWarps - 1) / kWarps;
  k2_tension<<<blocks, kWarps * 32, 0, stream>>>(p, alpha);
  count_launch();
  return cudaGetLastError();
}

}  // namespace speedy
