"""CPU: pin the oracle.

1. The known-answer tests of /root/reference/speedy_test.cc, re-expressed against
   the reference's OWN speedy.c compiled unmodified into oracle/_ref (so they also
   pin our FFT stand-in behind kiss_fft.h / fftw3.h).  Each test cites the lines it
   re-expresses.  Needs oracle/_ref (built here from /root/reference; on a box
   without the reference sources the prebuilt .so is used, else these skip).
2. Our restatement (oracle/speedy_oracle.c) against the compiled reference, bit
   for bit, and against the committed fixtures in tests/golden.
"""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol

KISS = pytest.mark.skipif(not ol.ref_available("kiss"), reason="oracle/_ref not built")
FFTW = pytest.mark.skipif(not ol.ref_available("fftw"), reason="oracle/_ref not built")
RATE = 22050  # speedy_test.cc:192


def f32(a):
    return np.ascontiguousarray(a, np.float32)


def arr(ptr, n):
    return np.ctypeslib.as_array(ptr, shape=(n,)).copy()


@pytest.fixture
def speedy():
    lib = ol.ref("kiss")
    s = lib.speedyCreateStream(RATE)
    yield lib, s
    lib.speedyDestroyStream(s)


@KISS
def test_first_order_filter():  # speedy_test.cc:135-156
    lib = ol.ref("kiss")
    fof = lib.CreateFirstOrderFilter(10.0)
    first = out = lib.IterateFirstOrderFilter(fof, 1.0)
    for _ in range(10):
        out = lib.IterateFirstOrderFilter(fof, 0.0)
    assert abs(first * np.exp(-1) - out) < 1e-7
    lib.ResetFirstOrderFilter(fof)
    assert abs(lib.IterateFirstOrderFilter(fof, 0.0)) < 1e-7
    lib.DeleteFirstOrderFilter(fof)


@KISS
def test_spectrogram_calculation(speedy):  # speedy_test.cc:197-218
    lib, s = speedy
    n = lib.speedyFFTSize(s) // 2
    x = np.zeros(2 * n, np.float32)
    x[:n] = np.sin(10 * np.arange(n) / np.float32(n) * np.pi)
    lib.speedySpectrogram(s, ol.fptr(x))
    spec = arr(lib.speedyGetSpectrogram(s), 2 * n)
    assert abs(spec[10] - 88.8677) < 1e-3
    assert np.argmax(spec[:n]) == 10
    far = np.abs(np.arange(n) - 10) > 3
    assert np.all(spec[:n][far] <= spec[1])


@KISS
def test_spectrogram_sinusoid(speedy):  # speedy_test.cc:222-254
    lib, s = speedy
    w = lib.speedyInputFrameSize(s)
    assert w == 330 and lib.speedyFFTSize(s) == 660
    x = f32(np.sin(2 * np.pi * np.arange(w) / np.float32(RATE) * 2200.0))
    spec = arr(lib.speedySpectrogram(s, ol.fptr(x)), 660)
    pos = int(np.argmax(spec[:330]))
    assert pos == lib.speedyFreqToBin(s, 2200.0) == 66
    assert abs(spec[pos] - 88.4847412109375) < 1e-3
    assert abs(spec[pos - 1] - 76.9396) < 1e-1
    assert abs(spec[pos + 1] - 68.0196) < 1e-1


@KISS
def test_preemphasis(speedy):  # speedy_test.cc:259-284
    lib, s = speedy
    x = f32([1, 0, 0, 0])
    lib.speedyPreemphasisFilter(s, ol.fptr(x), 4)
    assert np.allclose(x, [1.0, -0.97, 0, 0], atol=1e-7)
    s2 = lib.speedyCreateStream(RATE)
    got = []
    for v in (1.0, 0.0, 0.0, 0.0):
        one = f32([v])
        lib.speedyPreemphasisFilter(s2, ol.fptr(one), 1)
        got.append(one[0])
    lib.speedyDestroyStream(s2)
    assert np.allclose(got, [1.0, -0.97, 0, 0], atol=1e-7)


def _hysteresis_triangle(kind):
    lib = ol.ref(kind)
    s = lib.speedyCreateStream(RATE)
    for i in range(32):
        lib.speedyAddToHysteresisBuffer(s, float(i == 16), i)
    got = [lib.speedyEvaluateHysteresis(s, i) for i in range(32)]
    lib.speedyDestroyStream(s)
    return np.array(got)


@KISS
def test_hysteresis_match_matlab():  # speedy_test.cc:288-313, MATCH_MATLAB branch
    up = [i / 16.0 for i in range(1, 8)]
    down = [i / 24.0 for i in range(11, 0, -1)]
    correct = [0] * 9 + up + [1] + down + [0] * 4
    assert np.allclose(_hysteresis_triangle("kiss"), correct, atol=1e-8)


@FFTW
def test_hysteresis_paper_order():  # speedy_test.cc:296-301, default branch
    up = [i / 24.0 for i in range(1, 12)]
    down = [i / 16.0 for i in range(7, 0, -1)]
    correct = [0] * 5 + up + [1] + down + [0] * 8
    assert np.allclose(_hysteresis_triangle("fftw"), correct, atol=1e-8)


@KISS
def test_normalize_by_energy():  # speedy_test.cc:317-328
    lib = ol.ref("kiss")
    x, y = f32([0, 0, 1, 0, 1]), np.zeros(5, np.float32)
    e = lib.speedyNormalizeByEnergy(ol.fptr(x), ol.fptr(y), 5)
    assert abs(e - 2.0) < 1e-7
    assert np.allclose(y, [0, 0, np.sqrt(0.5), 0, np.sqrt(0.5)], atol=1e-7)


@KISS
def test_add_data_history(speedy):  # speedy_test.cc:331-373
    lib, s = speedy
    n = lib.speedyInputFrameSize(s)
    half = lib.speedyFFTSize(s) // 2
    for t, cyc in ((0, 1), (1, 2)):
        x = f32(np.sin(2 * np.pi * cyc * np.arange(n) / np.float32(n)))
        lib.speedyAddData(s, ol.fptr(x), t)
        assert lib.speedyGetCurrentTime(s) == t
    for t, peak in ((0, 2), (1, 4)):
        spec = arr(lib.speedyGetSpectrogramAtTime(s, t), half)
        assert int(np.argmax(spec)) == peak


@KISS
def test_local_energy(speedy):  # speedy_test.cc:380-412
    lib, s = speedy
    n = lib.speedyInputFrameSize(s)
    amp, at_max = 1.0, 0
    for t in range(100):
        x = f32(np.sin(2 * np.pi * np.arange(n) / np.float32(n)) * np.float32(amp))
        lib.speedyAddData(s, ol.fptr(x), t)
        lib.speedyComputeLocalEnergy(s, lib.speedyGetSpectrogramAtTime(s, t), t)
        at_max += lib.speedyGetEnergyCompressed(s) > 1.414
        amp = np.float32(amp * np.float32(0.9))
    assert at_max == 6
    assert abs(lib.speedyGetEnergyCompressed(s) - 1.7745e-04) < 1e-8


@KISS
def test_spectral_difference(speedy):  # speedy_test.cc:418-453
    lib, s = speedy
    n = lib.speedyInputFrameSize(s)
    amp, last = np.float32(1.0), None
    for t in range(100):
        freq = np.float32(t / 2.0)
        x = f32(np.sin(2 * np.pi * freq * np.arange(n) / np.float32(n)) * amp)
        lib.speedyAddData(s, ol.fptr(x), t)
        lib.speedyComputeSpectralDifference(s, lib.speedyGetSpectrogramAtTime(s, t),
                                            lib.speedyGetSpectrogramAtTime(s, t - 1), t)
        last = lib.speedyGetSpeechChanges(s)
        amp = np.float32(amp * np.float32(0.9))
    assert abs(last) < 1e-6


@KISS
def test_tension_known_answers(speedy):  # speedy_test.cc:457-530
    lib, s = speedy
    count = RATE
    i = np.arange(count)
    start = np.float32(0.15) * np.float32(RATE)  # float arithmetic, as in the C++ test
    decay = np.exp(-(i.astype(np.float32) - start) / np.float32(RATE * 0.5))  # std::exp(float)
    x = decay.astype(np.float64) * np.sin(2 * np.pi * 220.0 * i / np.float32(RATE))
    x[i < int(start)] = 0
    x = f32(x)
    window = lib.speedyInputFrameSize(s)
    step = np.float32(RATE / np.float32(100))
    frames = int((count - window) / step + 1)
    tension, out_t = [], 0
    for t in range(frames):
        begin = int(np.floor(float(t * step) + 0.5))  # std::round: halves away from zero
        seg = f32(x[begin:begin + window])
        lib.speedyAddData(s, ol.fptr(seg), t)
        v = C.c_float()
        if lib.speedyComputeTension(s, out_t, C.byref(v)):
            tension.append(v.value)
            out_t += 1
    tension = np.array(tension)
    assert abs(tension.min() - (-0.6)) < 1e-5
    assert abs(tension.max() - 0.14273257553577423) < 1e-6
    assert abs(tension[-1] - (-0.31351470947265625)) < 1e-5


@KISS
def test_feature_return(speedy):  # speedy_test.cc:714-757
    lib = ol.ref("kiss")
    s = lib.speedyCreateStream(16000)
    x = f32(np.cos(2 * np.pi * 440.0 * np.arange(8000) / np.float32(16000)))
    window = lib.speedyInputFrameSize(s)
    frames = int((8000 - window) / 160.0 + 1)
    peak = int(440.0 / (16000 // lib.speedyFFTSize(s)))
    out_t = 0
    for t in range(frames):
        seg = f32(x[t * 160:t * 160 + window])
        lib.speedyAddData(s, ol.fptr(seg), t)
        v = C.c_float()
        if lib.speedyComputeTension(s, out_t, C.byref(v)):
            out_t += 1
            assert arr(lib.speedyGetInternalState(s), 15)[11] == v.value
            spec = arr(lib.speedyGetInternalSpectrogram(s), 480)
            assert spec[peak] > spec[peak - 1] and spec[peak] > spec[peak + 1]
    assert frames == out_t + lib.ref_future_frames()
    lib.speedyDestroyStream(s)


@KISS
def test_tapestry_frame_counts(golden_inputs):  # speedy_test.cc:859-941
    pcm, rate = golden_inputs["tapestry22k"]
    assert pcm.shape[0] == 69431 and rate == 22050
    lib = ol.ref("kiss")
    s = lib.speedyCreateStream(rate)
    x = f32(pcm[:, 0] / 32768.0)
    assert abs(x.max() - 0.41369) < 1e-3
    window = lib.speedyInputFrameSize(s)
    step = np.float32(rate / np.float32(100))
    frames = int((len(x) - window) / step + 1)
    n_tension, energy = 0, []
    for t in range(frames):
        begin = int(np.floor(float(t * step) + 0.5))
        lib.speedyAddData(s, ol.fptr(f32(x[begin:begin + window])), t)
        v = C.c_float()
        if lib.speedyComputeTension(s, n_tension, C.byref(v)):
            n_tension += 1
            norm = arr(lib.speedyGetNormalizedSpectrogram(s), 330)
            energy.append(float(np.sum(norm.astype(np.float64) ** 2)))
    lib.speedyDestroyStream(s)
    assert frames == 314 and n_tension == 306
    # every normalised frame has unit energy (speedy_test.cc:975-978); frames the
    # low-energy gate skipped keep a stale buffer and are still unit-norm
    assert np.all(np.abs(np.array(energy[1:]) - 1) < 4e-3)


@KISS
def test_tapestry_feature_computations_matlab(golden_inputs):  # speedy_test.cc:859-1057
    """The reference-held Matlab dumps, with the reference's own SNR floors and best-delay
    table, on the compiled reference (its speedy.c over our FFT restatement): this is what
    pins the FFT restatement and the build of oracle/_ref to values the reference owns."""
    import matlab_checks as mc
    pcm, rate = golden_inputs["tapestry22k"]
    lib = ol.ref("kiss")
    s = lib.speedyCreateStream(rate)
    x = f32(pcm[:, 0] / 32768.0)
    window = lib.speedyInputFrameSize(s)
    assert window == 330 and lib.speedyFFTSize(s) == 660
    step = np.float32(rate / np.float32(100))
    frames = int((len(x) - window) / step + 1)
    spec, norm, feat, n_tension = [], [], [], 0
    for t in range(frames):
        begin = int(np.floor(float(t * step) + 0.5))
        lib.speedyAddData(s, ol.fptr(f32(x[begin:begin + window])), t)
        spec.append(arr(lib.speedyGetSpectrogram(s), 330))
        v = C.c_float()
        if lib.speedyComputeTension(s, n_tension, C.byref(v)):
            n_tension += 1
            norm.append(arr(lib.speedyGetNormalizedSpectrogram(s), 330))
            feat.append(arr(lib.speedyGetInternalState(s), 15))
    lib.speedyDestroyStream(s)
    got = mc.check(np.array(spec), np.array(norm), np.array(feat))
    assert got["spectrogram_snr_db"] > 27 and got["Audio Tension"][0] == 0


# ---- 2. the restatement against the compiled reference and the fixtures ------

CASES = [("kiss", True, False), ("fftw", False, True)]


@pytest.mark.parametrize("kind,match_matlab,fft_double", CASES)
@pytest.mark.parametrize("key,speed,feedback,chunk", [
    ("tapestry16k", 3.0, 0.1, 128), ("tapestry16k", 3.5, 0.0, 1000), ("tapestry22k", 2.0, 0.1, 1000),
    ("negative24k", 0.25, 0.1, 0), ("tapestry16k", 0.7, 0.2, 137)])
def test_port_equals_compiled_reference(golden_inputs, kind, match_matlab, fft_double, key, speed, feedback, chunk):
    if not ol.ref_available(kind):
        pytest.skip("oracle/_ref not built")
    pcm, rate = golden_inputs[key]
    r = ol.ref_process(kind, pcm, rate, 1, speed, 1.0, feedback, chunk=chunk)
    p = ol.port_process(ol.cfg(rate, 1, speed, 1.0, feedback, match_matlab, fft_double), pcm)
    assert np.array_equal(r["spectrogram"], p["spectrogram"])
    assert np.array_equal(r["tension"], p["tension"])
    assert np.array_equal(r["features"], p["features"])
    assert np.array_equal(r["speed"], p["speed"])
    assert np.array_equal(r["out"], p["out"])
    # callback times: spectrogram at_time starts at 1, tension at 0
    assert list(r["spec_time"][:3]) == [1, 2, 3] and list(r["tension_time"][:3]) == [0, 1, 2]
    # the normalised-spectrogram callback is stale by one tension step (soniclib.c:303-310)
    future = 8 if match_matlab else 12
    n_t = len(p["tension"])
    assert np.array_equal(r["normalized"][future:future + n_t - 1], p["normalized"][:n_t - 1])


@pytest.mark.parametrize("chunk", [137, 160, 1000, 16000])
def test_reference_schedule_is_chunk_invariant(golden_inputs, chunk):
    """SURVEY.md §3.2 [probe]: speeds and output do not depend on the write size."""
    if not ol.ref_available("fftw"):
        pytest.skip("oracle/_ref not built")
    pcm, rate = golden_inputs["tapestry16k"]
    a = ol.ref_process("fftw", pcm, rate, 1, 2.5, 1.0, 0.1, chunk=0)
    b = ol.ref_process("fftw", pcm, rate, 1, 2.5, 1.0, 0.1, chunk=chunk)
    assert np.array_equal(a["speed"], b["speed"]) and np.array_equal(a["out"], b["out"])


@pytest.mark.parametrize("rate,kind,match_matlab,fft_double", [(44100, "fftw", False, True), (44100, "kiss", True, False),
                                                               (11025, "fftw", False, True), (32000, "kiss", True, False)])
def test_port_equals_reference_at_prime_and_odd_windows(rate, kind, match_matlab, fft_double):
    """44.1 kHz has a prime 661-sample window (N = 1322 = 2 * 661), 11.025 kHz an odd one:
    the FFT restatement's generic butterfly (oracle/fft_oracle.c) carries them; the port
    still equals the reference's own code bit for bit there."""
    if not ol.ref_available(kind):
        pytest.skip("oracle/_ref not built")
    pcm = ol.synth(11, 1, rate, 1, rate)[0]
    r = ol.ref_process(kind, pcm, rate, 1, 2.0, 1.0, 0.1, chunk=1000)
    p = ol.port_process(ol.cfg(rate, 1, 2.0, 1.0, 0.1, match_matlab, fft_double), pcm)
    assert np.array_equal(r["spectrogram"], p["spectrogram"])
    assert np.array_equal(r["tension"], p["tension"])
    assert np.array_equal(r["speed"], p["speed"])
    assert np.array_equal(r["out"], p["out"])


def test_port_stereo_equals_reference():
    if not ol.ref_available("fftw"):
        pytest.skip("oracle/_ref not built")
    pcm = ol.synth(3, 1, 48000, 2, 48000)[0]
    r = ol.ref_process("fftw", pcm, 48000, 2, 1.5, 1.0, 0.1, chunk=480)
    p = ol.port_process(ol.cfg(48000, 2, 1.5, 1.0, 0.1, False, True), pcm)
    assert np.array_equal(r["speed"], p["speed"]) and np.array_equal(r["out"], p["out"])


def test_port_against_committed_fixtures(golden_inputs, golden_outputs):
    """tests/golden/reference_outputs.npz was produced by the compiled reference
    (tests/golden/make_golden.py); the restatement must reproduce it exactly."""
    for name, case in golden_outputs.items():
        rate, channels, speed, nonlinear, feedback, chunk, kiss = case["params"]
        key = "tapestry16k" if "tapestry16k" in name else ("tapestry22k" if "tapestry22k" in name else "negative24k")
        pcm, _ = golden_inputs[key]
        c = ol.cfg(int(rate), int(channels), float(speed), float(nonlinear), float(feedback),
                   match_matlab=bool(kiss), fft_double=not bool(kiss))
        p = ol.port_process(c, pcm)
        assert np.array_equal(p["out"], case["out"]), name
        if nonlinear != 0:
            assert np.array_equal(p["tension"], case["tension"]), name
            assert np.array_equal(p["speed"], case["speed"]), name
            assert np.array_equal(p["features"], case["features"]), name
            assert np.array_equal(p["spectrogram"][case["spec_rows"]], case["spec"]), name
            assert p["spectrogram"].shape[0] == int(case["n_spec"])


def test_schedule_closed_forms():
    """100 frames of input -> 99 spectrogram frames, 88 / 92 tensions; first audio
    reaches Sonic after Future*step + partial + 1 samples."""
    for mm, n_t, first in ((False, 88, 2001), (True, 92, 1361)):
        g = ol.geometry(16000, mm)
        assert (g.window, g.fft, g.step, g.partial) == (240, 480, 160, 80)
        nA = ol.port().oracle_frames_analyzed(C.byref(g), 16000)
        assert nA == 99 and ol.port().oracle_tensions_ready(C.byref(g), nA) == n_t
        for total in (first - 1, first):
            a = ol.port().oracle_frames_analyzed(C.byref(g), total)
            assert (ol.port().oracle_tensions_ready(C.byref(g), a) > 0) == (total == first)
    g = ol.geometry(22050)
    assert (g.window, g.fft, g.step, g.partial) == (330, 660, 220, 110)
    g = ol.geometry(48000)
    assert (g.window, g.fft, g.step) == (720, 1440, 480)
    assert (g.min_period, g.max_period, g.max_required, g.skip) == (120, 738, 1476, 12)
