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
// The three recurrences (two one-pole filters, the duration feedback) are
// strictly sequential in the frame index; one thread owns one stream and walks
// its new frames in order, interleaving "AddData" for at_time a with
// "ComputeTension" for r = a - Future exactly as soniclib.c:295-371 does.
// Per-stream rings (32 entries by at_time) live in global memory, laid out
// [slot][stream] so that a warp's accesses coalesce.
#include "kernels.cuh"

namespace speedy {

__global__ void __launch_bounds__(64) k2_tension(K2Params p, float alpha) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= p.n_streams) return;
  const Geometry& g = p.g;
  const int n = p.n_streams;
  const float nonlinear = p.st.nonlinear[s];
  if (nonlinear == 0.0f) return;  // soniclib.c:397-399: Speedy is bypassed

  const long long t_old = p.st.total[s];
  const long long t_new = t_old + (p.counts ? p.counts[s] : p.frames);
  const int kA = frames_analyzed(g, t_old);
  const int kB = frames_analyzed(g, t_new);
  const int rA = tensions_ready(g, kA);
  const int F = g.future, B = g.past;

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
  const float low_thr = (float)(0.04 * (double)max_hyst);       // :682
  const float change_cap = __fmul_rn(4.0f, mean_rsd);            // :728
  const float frame_duration = (float)(1.0 / 100.0);             // :783

  float* ring_c = p.st.ring_comp;
  float* ring_e = p.st.ring_energy;
  float* ring_l = p.st.ring_lsd;

  for (int j = 0; j < kB - kA; j++) {
    const int a = kA + j + 1;  // at_time of window kA + j (soniclib.c:296)
    const float2 ef = p.feat[(size_t)j * n + s];
    const float e = ef.x;
    // speedy.c:517-520
    lp_e = __fadd_rn(__fmul_rn(one_minus_alpha, e), __fmul_rn(alpha, lp_e));
    const float local = __fdiv_rn(e, lp_e);
    const float comp = (float)sqrt(local > 2.0f ? 2.0 : (double)local);
    const int slot = a & (kRing - 1);
    ring_c[(size_t)slot * n + s] = comp;
    ring_e[(size_t)slot * n + s] = e;
    ring_l[(size_t)slot * n + s] = ef.y;
    if (p.tap_energy) p.tap_energy[(size_t)s * p.max_new_frames + j] = e;

    const int r = a - F;  // speedy.c:755: ready once r + Future <= current_time
    if (r < 0) continue;

    // speedy.c:590-610 (ring reads of at_time <= 0 return the initial zeros)
    float future_max = 0.0f, past_max = 0.0f;
    for (int i = 0; i <= F; i++) {
      const int t = r + i;
      float v = t >= 1 ? ring_c[(size_t)(t & (kRing - 1)) * n + s] : 0.0f;
      v = __fmul_rn(v, __fdiv_rn((float)(F - i), (float)F));
      if (v > future_max) future_max = v;
    }
    for (int i = 0; i <= B; i++) {
      const int t = r - i;
      float v = t >= 1 ? ring_c[(size_t)(t & (kRing - 1)) * n + s] : 0.0f;
      v = __fmul_rn(v, __fdiv_rn((float)(B - i), (float)B));
      if (v > past_max) past_max = v;
    }
    const float hyst = __fmul_rn(__fadd_rn(past_max, future_max), 0.5f);

    // speedy.c:673-703: spectrum of at_time r; at_time 0 is the all-zero row
    const float e_r = r >= 1 ? ring_e[(size_t)(r & (kRing - 1)) * n + s] : 0.0f;
    const float lsd_raw = r >= 1 ? ring_l[(size_t)(r & (kRing - 1)) * n + s] : 0.0f;
    const bool low = e_r <= low_thr;
    float lsd = 0.0f, ewld = 0.0f, rel = 0.0f, changes = 0.0f;
    if (low) {
      lp_d = __fadd_rn(__fmul_rn(one_minus_alpha, 0.0f), __fmul_rn(alpha, lp_d));
    } else {
      lsd = lsd_raw;
      ewld = __fmul_rn(lsd, hyst);                                         // :720
      lp_d = __fadd_rn(__fmul_rn(one_minus_alpha, ewld), __fmul_rn(alpha, lp_d));
      rel = (float)((double)ewld / ((double)lp_d + 0.01 * (double)mean_lpf));  // :725
      changes = (float)fmin((double)rel, (double)change_cap);             // :727
    }
    // speedy.c:754-762
    const float tension = __fadd_rn(__fmul_rn(0.5f, __fsub_rn(hyst, 0.7f)),
                                    __fmul_rn(0.25f, __fsub_rn(changes, 1.0f)));
    // speedy.c:773-785
    float v;
    const float slope = __fmul_rn(__fsub_rn(1.0f, Rg), tension);
    if ((double)Rg > 1.0) {
      v = (float)fmax(1.0, (double)__fadd_rn(Rg, slope));
    } else {
      v = (float)fmax(0.01, fmin(1.0, (double)__fsub_rn(Rg, slope)));
    }
    if (fb > 0.0f) {
      const float excess = __fsub_rn(cur_dur, des_dur);
      v = (float)((double)v + fmax(0.01, (double)__fmul_rn(fb, excess)));
    }
    cur_dur = __fadd_rn(cur_dur, __fdiv_rn(frame_duration, v));
    des_dur = __fadd_rn(des_dur, __fdiv_rn(frame_duration, Rg));
    // soniclib.c:343-345
    float rate = __fadd_rn(__fmul_rn(v, nonlinear), __fmul_rn(Rg, __fsub_rn(1.0f, nonlinear)));

    const int jr = r - rA;  // index among this write's new tensions
    if (p.override_speeds) rate = p.override_speeds[(size_t)s * p.override_stride + r];
    p.speeds[(size_t)s * p.speeds_stride + jr] = rate;
    if (p.tap_tension) p.tap_tension[(size_t)s * p.max_new_frames + jr] = tension;
    if (p.tap_speed) p.tap_speed[(size_t)s * p.max_new_frames + jr] = rate;
    if (p.tap_features) {
      float* f = p.tap_features + ((size_t)s * p.max_new_frames + jr) * kFeatureCount;
      f[0] = e_r;  f[1] = lp_e;  f[2] = local;  f[3] = comp;  f[4] = hyst;
      f[5] = low ? 1.0f : 0.0f;  f[6] = lsd;  f[7] = ewld;  f[8] = lp_d;  f[9] = rel;
      f[10] = changes;  f[11] = tension;  f[12] = (float)a;  f[13] = (float)r;  f[14] = low_thr;
    }
  }
  p.st.lp_energy[s] = lp_e;
  p.st.lp_diff[s] = lp_d;
  p.st.cur_dur[s] = cur_dur;
  p.st.des_dur[s] = des_dur;
}

cudaError_t launch_k2(const K2Params& p, cudaStream_t stream) {
  if (p.max_new_frames <= 0) return cudaSuccess;
  // speedy.c:67: alpha = exp(-1.0 / time_constant), time constant 100 frames
  const float alpha = (float)exp(-1.0 / (double)100.0f);
  const int threads = 64;
  const int blocks = (p.n_streams + threads - 1) / threads;
  k2_tension<<<blocks, threads, 0, stream>>>(p, alpha);
  count_launch();
  return cudaGetLastError();
}

}  // namespace speedy
