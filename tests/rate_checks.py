"""Checker for sonicSetRate: upstream Sonic's classic playback-rate change (adjustRate /
interpolate: linear interpolation between neighbouring frames on integer positions), restated
in numpy.  Test infrastructure only.  PARITY UNPINNED: upstream Sonic is not vendored in the
reference (its build clones it), no reference test sets a rate, and upstream's later revisions
resample with a windowed sinc; what the tests pin is that the product applies exactly this
resampler to exactly the frames the speed change produced.
"""
import numpy as np


def resample(frames, sample_rate, rate):
    """frames: (n, channels) int16 as produced at rate 1.  Returns the frames a stream with
    sonicSetRate(rate) hands out before its flush (the last input frame stays behind)."""
    new_rate, old_rate = int(np.float32(sample_rate) / np.float32(rate)), sample_rate
    while new_rate > (1 << 14) or old_rate > (1 << 14):
        new_rate >>= 1
        old_rate >>= 1
    new_rate = max(new_rate, 1)
    x = frames.astype(np.int64)
    out, old_pos, new_pos = [], 0, 0
    for position in range(len(x) - 1):
        while (old_pos + 1) * new_rate > new_pos * old_rate:
            pos = new_pos * old_rate
            left_pos, right_pos = old_pos * new_rate, (old_pos + 1) * new_rate
            ratio, width = right_pos - pos, right_pos - left_pos
            v = ratio * x[position] + (width - ratio) * x[position + 1]
            out.append(np.trunc(v / width).astype(np.int16))  # C division truncates toward zero
            new_pos += 1
        old_pos += 1
        if old_pos == old_rate:
            old_pos = new_pos = 0
    return np.array(out, np.int16).reshape(-1, frames.shape[1])
