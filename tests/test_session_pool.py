"""GPU: the session pool (speedy_b200.h section 1b) -- many drop-in sonicStream handles
multiplexed onto one device batch.  Every session's output must be bit-identical to feeding
the same samples to a stream of its own (soniclib.c:391-452 call semantics: write, read,
flush per handle), whatever the other sessions are doing."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol
import speedy_b200 as sb
from gpu_util import gpu_process

pytestmark = pytest.mark.gpu


def drain(L, h, buf):
    got = []
    while True:
        n = L.sonicReadShortFromStream(h, buf.ctypes.data, len(buf))
        if n == 0:
            return got
        got.append(buf[:n].copy())


@pytest.mark.parametrize("nonlinear", [1.0, 0.0])
def test_pooled_sessions_equal_one_shot(nonlinear):
    rate, n, frames, chunk = 16000, 12, 16000 * 3, 160
    pcm = ol.synth(900, n, rate, 1, frames)
    speeds = [2.0, 3.5, 1.3, 0.7, 2.0, 1.0, 2.5, 1.7, 3.0, 0.6, 2.0, 4.0]
    L = sb.lib()
    pool = sb.SessionPool(rate, 1, max_sessions=16, max_pending_frames=400, min_speed=0.25)
    hs = [pool.open() for _ in range(n)]
    for h, sp in zip(hs, speeds):
        L.sonicSetSpeed(h, sp)
        L.sonicEnableNonlinearSpeedup(h, nonlinear)
    buf = np.zeros(4096, np.int16)
    outs = [[] for _ in range(n)]
    # session s stops writing after lens[s] frames; session 3 is flushed early, while the others run on
    lens = [frames - 777 * s for s in range(n)]
    flushed = [False] * n
    for t in range(0, frames, chunk):
        for s, h in enumerate(hs):
            piece = np.ascontiguousarray(pcm[s, t:min(t + chunk, lens[s]), 0])
            if len(piece):
                assert L.sonicWriteShortToStream(h, piece.ctypes.data, len(piece)) == 1
            elif not flushed[s]:
                flushed[s] = True
                assert L.sonicFlushStream(h) == 1
        if (t // chunk) % 2 == 1:  # read every other tick: two chunks queue per session in between
            for s, h in enumerate(hs):
                outs[s] += drain(L, h, buf)
    st = pool.stats()
    assert st["open_sessions"] == n and st["session_writes"] > 0
    # far fewer coalesced steps than writes: one per read tick plus the early flushes
    assert st["steps"] <= frames // chunk // 2 + 2 * n + 2, st
    for s, h in enumerate(hs):
        if not flushed[s]:
            assert L.sonicFlushStream(h) == 1
        outs[s] += drain(L, h, buf)
        L.sonicDestroyStream(h)
    for s in range(n):
        want, _, _ = gpu_process(pcm[s:s + 1, :lens[s]], rate, speeds[s], nonlinear=nonlinear, taps=0)
        got = np.concatenate(outs[s]) if outs[s] else np.zeros(0, np.int16)
        assert np.array_equal(got, want[0][:, 0]), (s, len(got), len(want[0]))
    pool.close()


def test_slot_reuse_and_parameter_changes():
    rate, frames = 16000, 16000 * 2
    pcm = ol.synth(77, 2, rate, 1, frames)
    L = sb.lib()
    pool = sb.SessionPool(rate, 1, max_sessions=2, max_pending_frames=1600)
    buf = np.zeros(8192, np.int16)

    def run(h, x, speed):
        L.sonicSetSpeed(h, speed)
        L.sonicEnableNonlinearSpeedup(h, 1.0)
        out = []
        for t in range(0, len(x), 1000):
            piece = np.ascontiguousarray(x[t:t + 1000])
            assert L.sonicWriteShortToStream(h, piece.ctypes.data, len(piece)) == 1
            assert L.sonicIntSamplesAvailable(h) >= 0
            out += drain(L, h, buf)
        assert L.sonicFlushStream(h) == 1
        out += drain(L, h, buf)
        return np.concatenate(out)

    a = pool.open()
    b = pool.open()
    with pytest.raises(RuntimeError):
        pool.open()  # full
    got_a = run(a, pcm[0, :, 0], 2.0)
    L.sonicDestroyStream(a)
    # the freed slot serves a new session from a clean state, next to a live one
    assert L.sonicWriteShortToStream(b, np.ascontiguousarray(pcm[1, :5000, 0]).ctypes.data, 5000) == 1
    c = pool.open()
    got_c = run(c, pcm[0, :, 0], 2.0)
    assert np.array_equal(got_a, got_c)
    want, _, _ = gpu_process(pcm[0:1], rate, 2.0, taps=0)
    assert np.array_equal(got_a, want[0][:, 0])
    # callbacks are not carried by pooled handles
    L.sonicTensionCallback(c, sb.tensionFunction(lambda *_: None))
    assert not L.getSonicTensionCallback(c)
    pool.close()
