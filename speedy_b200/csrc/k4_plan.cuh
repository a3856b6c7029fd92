// What one K4 launch has to do for one stream, in coordinates relative to a base frame: the
// feed events of the reference's shim (soniclib.c:354, 369-371, 397-399, 538-551) as two
// arithmetic runs, and where the samples come from.  Shared by the pipelined kernel
// (k4_splice.cu) and the one-warp 16 kHz chain kernel (k4_chain16.cu).
//
// Events.  The reference calls processStreamInput once per 10 ms buffer with the speed of
// that buffer (soniclib.c:354, 369-371), or once per write in the linear mode (:397-399).
// A pitch iteration at cursor `pos` runs inside the first such call whose fed total reaches
// pos + maxRequired (oracle/sonic_oracle.c:335, 352), at that call's speed.
#pragma once

#include "kernels.cuh"

namespace speedy {

struct Plan {
  long long base;      // absolute frame of relative frame 0 (== the source's start modulo 8: 16-byte alignment)
  int pos0;            // Sonic's FIFO head, relative
  int fed0;            // frames handed to Sonic so far, relative
  int nA;              // events with fed = fedA0 + i * S: one per 10 ms buffer
  int fedA0;
  bool per_frame;      // their speeds come from the speeds rows (else: the carried speed)
  const float* spA;    // row of event 0
  bool hasB;           // then one event with fed = fedB: a linear write, or the final flush
  bool finalB;
  int fedB;
  int fed_real;        // flush: fed before the padding (relative)
  int data_end;        // relative frames >= this read as silence
  int last_fed;        // the largest fed of the launch (relative)
  int nchunks;         // k4_splice: chunks its filler produces
  Source src;
  long long t_old;
  int out_count0;      // output frames pending in the stream's row when the launch starts
};

__device__ __forceinline__ Plan make_plan(const K4Params& p, int s) {
  const Geometry& g = p.g;
  Plan pl;
  const Range rg = write_range(p.st.total, p.counts, p.frames, p.done, s);
  const long long t_old = rg.t_old;
  const long long t_new = p.flush ? t_old : rg.t_new;
  const long long t_done = p.flush ? t_old : rg.t_done;
  pl.t_old = t_old;
  pl.out_count0 = p.st.out_count[s];
  pl.src.channels = g.channels;
  pl.src.hist = p.hist + (size_t)s * p.hist_stride;
  pl.src.in = p.in ? p.in + (size_t)s * p.in_stride_frames * g.channels : nullptr;
  pl.src.hist_base = p.st.hist_base[s];
  pl.src.t_old = t_old;
  pl.src.t_new = t_new;
  const long long head = p.st.sonic_head[s];
  const long long fed = p.st.sonic_fed[s];
  const bool nonlinear = p.st.nonlinear[s] != 0.0f;
  // base <= head, congruent to the start of the contiguous source the bulk copies read
  const long long anchor = p.flush ? pl.src.hist_base : t_old;
  long long m = (head - anchor) % 8;
  if (m < 0) m += 8;
  pl.base = head - m;
  pl.pos0 = (int)(head - pl.base);
  pl.fed0 = (int)(fed - pl.base);
  pl.nA = 0;
  pl.fedA0 = 0;
  pl.per_frame = false;
  pl.spA = nullptr;
  pl.hasB = false;
  pl.finalB = false;
  pl.fedB = 0;
  pl.fed_real = 0;
  long long data_end = t_new;
  long long last_fed = fed;
  if (!p.flush) {
    if (nonlinear) {
      const int rA = tensions_ready(g, frames_analyzed(g, t_old));  // the speeds rows count from here
      const int evA = tensions_ready(g, frames_analyzed(g, t_done));
      const int evE = tensions_ready(g, frames_analyzed(g, t_new));
      pl.nA = evE - evA;
      pl.fedA0 = (int)((long long)(evA + 1) * g.step - pl.base);
      pl.per_frame = true;
      pl.spA = p.speeds + (size_t)s * p.speeds_stride + (evA - rA);
      if (pl.nA > 0) last_fed = (long long)evE * g.step;
    } else if (t_new > t_done) {
      pl.hasB = true;
      pl.fedB = (int)(t_new - pl.base);
      last_fed = t_new;
    }
  } else {
    long long fed_real = fed;
    if (nonlinear) {
      const long long ev0 = fed / g.step;
      long long evE = t_old / g.step;
      if (evE < ev0) evE = ev0;
      pl.nA = (int)(evE - ev0);
      pl.fedA0 = (int)((ev0 + 1) * g.step - pl.base);
      if (pl.nA > 0) fed_real = evE * g.step;
    }
    pl.hasB = true;
    pl.finalB = true;
    pl.fedB = (int)(fed_real + 2 * g.max_required - pl.base);
    data_end = fed_real;
    last_fed = fed_real + 2 * g.max_required;
    pl.fed_real = (int)(fed_real - pl.base);
  }
  pl.data_end = (int)(data_end - pl.base);
  pl.last_fed = (int)(last_fed - pl.base);
  pl.nchunks = 0;
  return pl;
}

}  // namespace speedy

