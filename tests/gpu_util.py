"""Helpers for the -m gpu parity tests: drive the C ABI through speedy_b200."""
import numpy as np

import speedy_b200 as sb

ALL_TAPS = sb.TAP_TENSION | sb.TAP_SPEED | sb.TAP_FEATURES | sb.TAP_SPECTROGRAM | sb.TAP_ENERGY


def out_capacity(frames, speed, nonlinear, max_required):
    if speed > 1.0:
        worst = 1.0
    elif nonlinear != 0:
        worst = 0.01
    else:
        worst = speed
    return int(frames / worst * 1.02) + 4 * max_required


def gpu_process(pcm, rate, speed, nonlinear=1.0, feedback=0.1, match_matlab=False, chunk=None,
                taps=ALL_TAPS, override=None, tps=0, flush=True, counts=None):
    """pcm: int16 [n, frames, channels].  Writes in `chunk`-frame pieces (None = one
    write), flushes, reads.  Returns (list of per-stream int16 outputs, taps dict with
    per-stream arrays concatenated over the writes, status)."""
    pcm = np.ascontiguousarray(pcm, np.int16)
    n, frames, channels = pcm.shape
    chunk = frames if chunk is None else chunk
    cap = out_capacity(frames, speed, nonlinear, 2 * (rate // 65))
    b = sb.Batch(n, rate, channels, speed=speed, nonlinear=nonlinear, feedback=feedback,
                 match_matlab=match_matlab, max_write_frames=max(chunk, 1), out_capacity=cap,
                 taps=taps, threads_per_stream=tps)
    if override is not None:
        b.override_speeds(override)
    acc = {k: [[] for _ in range(n)] for k in ("spectrogram", "energy", "features", "tension", "speed")}
    outs = [[] for _ in range(n)]
    for t in range(0, frames, chunk):
        piece = pcm[:, t:t + chunk]
        if counts is not None:
            c = np.clip(np.asarray(counts) - t, 0, piece.shape[1]).astype(np.int32)
            b.write(piece, c)
        else:
            b.write(piece)
        if taps and nonlinear != 0:
            tp = b.taps()
            for k in acc:
                if k in tp:
                    for s in range(n):
                        acc[k][s].append(tp[k][s].copy())
        if chunk != frames:
            o, c = b.read(cap)
            for s in range(n):
                outs[s].append(o[s, :c[s]].copy())
    if flush:
        b.flush()
    o, c = b.read(cap)
    for s in range(n):
        outs[s].append(o[s, :c[s]].copy())
    status = b.status()
    b.close()
    outs = [np.concatenate(x, axis=0) for x in outs]
    res = {}
    for k, v in acc.items():
        if any(len(x) for x in v):
            res[k] = [np.concatenate(x, axis=0) for x in v]
    return outs, res, status
