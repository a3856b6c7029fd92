"""CPU: the reference's property tests for upstream Sonic, re-expressed against
our restatement (oracle/sonic_oracle.c).  No value-level golden vectors exist for
this stage anywhere in the reference ("parity unpinned"); these
are the properties the reference itself checks."""
import numpy as np
import pytest

import oracle_lib as ol


def linear(pcm, rate, speed, channels=1):
    c = ol.cfg(rate, channels, speed, 0.0, 0.1)
    return ol.port_process(c, pcm)["out"]


def teager(x):
    x = x.astype(np.float64)
    return x[1:-1] ** 2 - x[:-2] * x[2:]


@pytest.mark.parametrize("speed,pitch_hz", [(3.0, 100), (0.5, 100), (2.0, 237), (0.4, 237)])
def test_sinusoid_length_and_teager(speed, pitch_hz):
    # sonic_classic_test.cc:167-288, sonic_test.cc:479-589
    rate = 22050
    period = rate // pitch_hz
    one = (32000 * np.sin(np.arange(period) * 2 * np.pi / period)).astype(np.int16)
    pcm = np.tile(one, 100)
    out = linear(pcm, rate, speed)[:, 0]
    expected = len(pcm) / speed
    assert abs(len(out) - expected) < 0.015 * expected
    t_in, t_out = teager(one.astype(np.int16)), teager(out[:len(out) - 1000])
    assert abs(t_out.mean() - t_in.mean()) < 0.01 * t_in.mean()
    assert np.sqrt(t_out.var()) / t_out.mean() < 0.02


def test_speech_length_over_speed_range(golden_inputs):
    # sonic_classic_test.cc:519-535: within 14 ms of the expected length
    pcm, rate = golden_inputs["tapestry16k"]
    speed = 1.1
    while speed < 6.3:
        out = linear(pcm, rate, np.float32(speed))
        assert abs(len(out) - int(len(pcm) / speed)) <= 14 * rate // 1000, speed
        speed += 0.25


def test_noise_length_over_speed_range():
    # sonic_classic_test.cc:558-576
    rng = np.random.default_rng(0)
    pcm = np.clip(rng.normal(0, 1, 50000) * 8096, -32000, 32000).astype(np.int16)
    speed = 1.1
    while speed < 6.3:
        out = linear(pcm, 16000, np.float32(speed))
        assert abs(len(out) - int(len(pcm) / speed)) <= 1.5 * 16000 / 100, speed
        speed += 0.25


def test_mono_equals_each_stereo_channel_exactly(golden_inputs):
    # sonic_classic_test.cc:659-665, 703-708
    pcm, rate = golden_inputs["tapestry16k"]
    mono = linear(pcm, rate, 2.0)[:, 0]
    stereo = linear(np.repeat(pcm, 2, axis=1), rate, 2.0, channels=2)
    assert len(stereo) == len(mono)
    assert np.array_equal(stereo[:, 0], mono) and np.array_equal(stereo[:, 1], mono)
    sine = (16000 * np.sin(2 * np.pi * 440 * np.arange(16000) / np.float32(16000))).astype(np.int16)
    mono = linear(sine, 16000, 2.0)[:, 0]
    stereo = linear(np.stack([sine, sine], axis=1), 16000, 2.0, channels=2)
    assert np.array_equal(stereo[:, 0], mono) and np.array_equal(stereo[:, 1], mono)


def test_silent_channel_stays_silent():
    # sonic_test.cc:828-861: dichotic input through the nonlinear path
    rate = 22050
    x = (32000 * np.sin(np.arange(rate) * 2 * np.pi * 237 / rate)).astype(np.int16)
    pcm = np.stack([x, np.zeros_like(x)], axis=1)
    c = ol.cfg(rate, 2, 3.0, 1e-5, 0.1, match_matlab=True)
    out = ol.port_process(c, pcm)["out"]
    assert abs(len(out) - rate / 3.0) < 0.01 * rate
    assert not out[:, 1].any()
    assert out[:, 0].any()


def test_stereo_tapestry_matches_mono_within_one(golden_inputs):
    # sonic_test.cc:871-947
    pcm, rate = golden_inputs["tapestry16k"]
    c1 = ol.cfg(rate, 1, 3.0, 1.0, 0.1, match_matlab=True)
    c2 = ol.cfg(rate, 2, 3.0, 1.0, 0.1, match_matlab=True)
    mono = ol.port_process(c1, pcm)
    st = np.stack([pcm[:, 0] - 50, pcm[:, 0] + 50], axis=1).astype(np.int16)
    stereo = ol.port_process(c2, st)
    assert len(stereo["out"]) == len(mono["out"])
    assert np.allclose(stereo["tension"], mono["tension"], rtol=1e-5, atol=0)
    avg = (stereo["out"][:, 0].astype(int) + stereo["out"][:, 1]) // 2
    assert np.max(np.abs(avg - mono["out"][:, 0])) <= 1


def test_duration_feedback_reduces_excess(golden_inputs):
    # speedy_test.cc:653-711 (30 concatenations of tapestry instead of 100: the
    # ordering only emerges once the open-loop excess has had time to accumulate)
    pcm, rate = golden_inputs["tapestry16k"]
    long = np.tile(pcm[:, 0], 30)
    excess = []
    for fb in (0.0, 0.1, 0.2, 0.4):
        out = ol.port_process(ol.cfg(rate, 1, 3.0, 1.0, fb, match_matlab=True), long, taps=False)["out"]
        excess.append(abs(len(long) / 3.0 - len(out)) / rate)
    assert excess[1] < excess[0] and excess[2] < excess[1] and excess[3] < excess[2], excess


def test_chunked_writes_equal_one_write():
    """Feeding Sonic in pieces does not change its output at constant speed."""
    import ctypes as C
    lib = ol.port()
    pcm = ol.synth(9, 1, 16000, 1, 48000)[0]
    whole = linear(pcm, 16000, 1.7)
    lib.sonicIntCreateStream.restype = C.c_void_p
    lib.sonicIntCreateStream.argtypes = [C.c_int, C.c_int]
    for name, args in (("sonicIntSetSpeed", [C.c_void_p, C.c_float]),
                       ("sonicIntWriteShortToStream", [C.c_void_p, ol.c_short_p, C.c_int]),
                       ("sonicIntReadShortFromStream", [C.c_void_p, ol.c_short_p, C.c_int]),
                       ("sonicIntFlushStream", [C.c_void_p]), ("sonicIntDestroyStream", [C.c_void_p])):
        getattr(lib, name).argtypes = args
    s = lib.sonicIntCreateStream(16000, 1)
    lib.sonicIntSetSpeed(s, 1.7)
    got, buf = [], np.zeros(65536, np.int16)
    for t in range(0, len(pcm), 333):
        piece = np.ascontiguousarray(pcm[t:t + 333])
        lib.sonicIntWriteShortToStream(s, ol.sptr(piece), len(piece))
        n = lib.sonicIntReadShortFromStream(s, ol.sptr(buf), len(buf))
        got.append(buf[:n].copy())
    lib.sonicIntFlushStream(s)
    n = lib.sonicIntReadShortFromStream(s, ol.sptr(buf), len(buf))
    got.append(buf[:n].copy())
    lib.sonicIntDestroyStream(s)
    assert np.array_equal(np.concatenate(got), whole[:, 0])


def test_speech_sample_dtw_with_reference_thresholds(golden_inputs):
    """sonic_test.cc:641-724 on the CPU restatement: the DTW path between the original
    and the time-compressed spectrograms has slope 1/speed within the reference's own
    tolerances — a statistical pin of the restated Sonic (its upstream source is absent)."""
    from speedy_b200 import evaluation as ev
    from test_sonic_suite import compute_spectrogram
    pcm, rate = golden_inputs["tapestry16k"]
    speed = 3.0
    linear = ol.port_process(ol.cfg(rate, speed=speed, nonlinear=0.0), pcm, taps=False)["out"][:, 0]
    speedy = ol.port_process(ol.cfg(rate, speed=speed, nonlinear=1.0), pcm, taps=False)["out"][:, 0]
    assert abs(len(linear) - 50381 / speed) <= 140
    original_spec = compute_spectrogram(pcm[:, 0], rate)
    for result, tol, max_cost in ((linear, 0.02, 13000000), (speedy, 0.1, None)):
        cost, p1, p2 = ev.dtw(original_spec, compute_spectrogram(result, rate))
        if max_cost:
            assert cost < max_cost
        slope = ev.linear_slope(p1, p2)
        assert abs(slope - 1.0 / speed) <= tol
        slopes = ev.linear_slope_everywhere(p1, p2, 10)
        assert abs(ev.mean(slopes) - slope) <= 0.02
        assert ev.standard_deviation(slopes) < 0.2
