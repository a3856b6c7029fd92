"""sonicSetRate (sonic2.h:70, soniclib.c:169-175): the playback-rate change upstream Sonic applies
to what its speed change produced.  The resampler is checked against tests/rate_checks.py
(parity unpinned, see there) and through properties: output length n / (speed * rate), a
sinusoid's frequency scaled by the rate, rate 1 untouched, pooled handles equal private ones.
"""
import ctypes as C

import numpy as np
import pytest

import speedy_b200 as sb
from rate_checks import resample

pytestmark = pytest.mark.gpu


def run(x, sample_rate, speed, rate, nonlinear=0.0, flush=True, chunk=1000, handle=None, rate_after=None):
    L = sb.lib()
    x = np.ascontiguousarray(x.reshape(len(x), -1))
    ch = x.shape[1]
    h = handle or L.sonicCreateStream(sample_rate, ch)
    assert h
    try:
        L.sonicSetSpeed(h, speed)
        L.sonicSetRate(h, rate)
        L.sonicEnableNonlinearSpeedup(h, nonlinear)
        out, buf = [], np.zeros((8192, ch), np.int16)

        def drain():
            while True:
                n = L.sonicReadShortFromStream(h, buf.ctypes.data, len(buf))
                if n <= 0:
                    return
                out.append(buf[:n].copy())

        for t in range(0, len(x), chunk):
            piece = np.ascontiguousarray(x[t:t + chunk])
            assert L.sonicWriteShortToStream(h, piece.ctypes.data, len(piece)) == 1
            drain()
            if rate_after is not None and t + chunk >= rate_after[0] > t:
                L.sonicSetRate(h, rate_after[1])
        if flush:
            assert L.sonicFlushStream(h) == 1
            drain()
        return np.concatenate(out) if out else np.zeros((0, ch), np.int16)
    finally:
        L.sonicDestroyStream(h)


def tone(sample_rate, hz, seconds, channels=1):
    t = np.arange(int(sample_rate * seconds))
    x = (8000 * np.sin(2 * np.pi * hz * t / sample_rate)).astype(np.int16)
    return np.repeat(x[:, None], channels, 1)


def dominant_hz(x, sample_rate):
    x = x[:, 0].astype(np.float64)
    spec = np.abs(np.fft.rfft(x * np.hanning(len(x))))
    return np.argmax(spec) * sample_rate / len(x)


@pytest.mark.parametrize("sample_rate,channels", [(16000, 1), (22050, 2), (48000, 1)])
@pytest.mark.parametrize("speed,rate", [(1.0, 1.5), (2.0, 0.8), (1.5, 2.0), (0.7, 0.5)])
def test_rate_is_the_linear_resampler_over_the_speed_changed_frames(sample_rate, channels, speed, rate):
    rng = np.random.default_rng(7)
    x = tone(sample_rate, 180.0, 1.0, channels) + rng.integers(-300, 300, (sample_rate, channels)).astype(np.int16)
    plain = run(x, sample_rate, speed, 1.0, flush=False)
    got = run(x, sample_rate, speed, rate, flush=False)
    want = resample(plain, sample_rate, rate)
    assert len(got) == len(want)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("speed,rate,nonlinear", [(1.0, 1.5, 0.0), (2.0, 0.8, 0.0), (2.0, 1.25, 1.0), (3.0, 2.0, 0.0)])
def test_flushed_length_and_pitch(speed, rate, nonlinear):
    sample_rate, hz = 16000, 200.0
    x = tone(sample_rate, hz, 2.0)
    y = run(x, sample_rate, speed, rate, nonlinear)
    plain = run(x, sample_rate, speed, 1.0, nonlinear)
    # upstream's flush hands out (what the speed change still owed) / rate + 0.5 frames, rounded per flush
    assert abs(len(y) - len(plain) / rate) <= 2 + 0.01 * len(y)
    body = y[len(y) // 8: len(y) * 7 // 8]
    assert abs(dominant_hz(body, sample_rate) - hz * rate) <= 2.0 * sample_rate / len(body)
    assert np.abs(y.astype(np.int32)).max() <= 8000 + 400


def test_rate_one_is_untouched_and_going_back_to_one_keeps_every_frame():
    sample_rate = 16000
    x = tone(sample_rate, 150.0, 1.0)
    a = run(x, sample_rate, 2.0, 1.0)
    b = run(x, sample_rate, 2.0, 1.0, rate_after=(4000, 1.0))
    assert np.array_equal(a, b)
    c = run(x, sample_rate, 1.0, 2.0, rate_after=(8000, 1.0))
    # first half an octave up in half the frames, second half as written
    assert abs(len(c) - (4000 + 8000)) <= 4
    assert np.array_equal(c[-4000:], x[-4000:])


def test_pooled_handle_equals_private_handle():
    sample_rate = 16000
    x = tone(sample_rate, 170.0, 1.0) + np.random.default_rng(3).integers(-200, 200, (sample_rate, 1)).astype(np.int16)
    private = run(x, sample_rate, 2.0, 1.5, 1.0)
    pool = sb.SessionPool(sample_rate, 1, max_sessions=4)
    try:
        pooled = run(x, sample_rate, 2.0, 1.5, 1.0, handle=pool.open())
    finally:
        pool.close()
    assert np.array_equal(private, pooled)
