"""Reference-owned tests VERDICT r1 listed as not yet re-expressed, with the reference's own
thresholds:
  TestRealSpeech / TestRealSpeechNormalized   /root/reference/speedy_test.cc:534-651
  TestChirpSpeedup                            /root/reference/sonic_classic_test.cc:303-395
  TestLongStereoSpeechRange                   /root/reference/sonic_classic_test.cc:539-556
CPU legs run on the compiled reference (oracle/_ref) exactly as the reference's test drives it;
GPU legs run the CUDA path through the C ABI."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol

f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
RG = 2.1  # speedy_test.cc:579


def chirp(rate=22050):
    """sonic_classic_test.cc:304-325: 137 Hz rising by 47 Hz over three seconds."""
    t = (np.arange(3 * rate) / np.float32(rate)).astype(np.float32)
    phase = np.float32(137) * t + np.float32(47) / 3 * t * t / np.float32(2.0)
    return (32000 * np.sin(2 * np.pi * phase.astype(np.float64))).astype(np.int16)


def linear_slope(y):  # sonic_classic_test.cc:137-160
    y = np.asarray(y, np.float64)
    x = np.arange(len(y), dtype=np.float64)
    n = float(len(y))
    return (n * (x * y).sum() - x.sum() * y.sum()) / (n * (x * x).sum() - x.sum() ** 2)


def check_chirp(out):
    """:357-391: sqrt of the Teager energy is proportional to frequency; the slope of the
    middle half (played at half the speed-up) is half that of the outer quarters."""
    o = out.astype(np.float32)
    teager = np.sqrt(o[1:-1] * o[1:-1] - o[:-2] * o[2:])
    n = len(teager)
    s1 = linear_slope(teager[:n // 4])
    s2 = linear_slope(teager[n // 4:n * 3 // 4])
    s3 = linear_slope(teager[n * 3 // 4:n - 1000])
    assert abs(s1 - s3) <= s1 * 0.05, (s1, s2, s3)
    assert abs(s2 - s1 / 2) <= s1 * 0.01, (s1, s2, s3)
    return s1, s2, s3


def drive_chirp(lib, cast):
    """:333-355 through a Sonic C API (the compiled reference's or the drop-in's)."""
    rate, x = 22050, chirp()
    h = lib.sonicCreateStream(rate, 1)
    assert h
    out, total = np.zeros(3 * rate, np.int16), 0
    for k, speed in enumerate((3.0, 1.5, 3.0)):
        lib.sonicSetSpeed(h, speed)
        piece = np.ascontiguousarray(x[k * rate:(k + 1) * rate])
        assert lib.sonicWriteShortToStream(h, cast(piece), rate)
    for _ in range(100):
        total += lib.sonicReadShortFromStream(h, cast(out[total:]), 3 * rate - total)
    assert lib.sonicFlushStream(h)
    while True:
        n = lib.sonicReadShortFromStream(h, cast(out[total:]), 3 * rate - total)
        total += n
        if n <= 0:
            break
    lib.sonicDestroyStream(h)
    return out[:total]


def check_real_speech(tension, speeds, strict):
    """speedy_test.cc:568-590 (strict: the extra bound of the un-normalised variant, :590)."""
    tension = np.asarray(tension, np.float32)
    assert tension.min() < -0.4 and tension.max() > 0.75
    assert abs(float(np.mean(tension.astype(np.float64)))) <= tension.max() / 6.0
    avg = float(np.mean(np.asarray(speeds, np.float64)))
    assert abs(avg - RG) <= RG / 10.0, avg
    if strict:
        assert avg <= RG - RG / 20.0, avg


# ---- CPU: the compiled reference, driven as the reference's tests drive it ----

@pytest.mark.parametrize("kind", ["kiss", "fftw"])
def test_real_speech_on_compiled_reference(golden_inputs, kind):
    if not ol.ref_available(kind):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    pcm, rate = golden_inputs["tapestry16k"]
    assert pcm.shape[0] == 50381 and pcm[0, 0] == 15  # :541-543
    lib = ol.ref(kind)
    s = lib.speedyCreateStream(rate)
    x = f32(pcm[:, 0])
    window = lib.speedyInputFrameSize(s)
    step = np.float32(rate / np.float32(100))
    frames = int((len(x) - window) / step + 1)
    tension, out_t = [], 0
    for t in range(frames):
        begin = int(np.floor(float(t * step) + 0.5))
        lib.speedyAddData(s, ol.fptr(f32(x[begin:begin + window])), t)
        v = C.c_float()
        if lib.speedyComputeTension(s, out_t, C.byref(v)):
            tension.append(v.value)
            out_t = 0  # (as the reference's test does, :563)
    speeds = [lib.speedyComputeSpeedFromTension(float(t), RG, 0.0, s) for t in tension]
    lib.speedyDestroyStream(s)
    check_real_speech(tension, speeds, strict=True)


def test_chirp_speedup_on_compiled_reference():
    if not ol.ref_available("fftw"):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    out = drive_chirp(ol.ref("fftw"), ol.sptr)
    check_chirp(out)


# ---- GPU: the same tests on the CUDA path ------------------------------------

@pytest.mark.gpu
def test_real_speech_gpu_taps(golden_inputs):
    """The tension and speed series of the shipped call sequence (sonicWriteShortToStream ->
    Speedy), speed R_g = 2.1 without feedback as the test's speedyComputeSpeedFromTension(t, Rg, 0)."""
    from gpu_util import gpu_process
    pcm, rate = golden_inputs["tapestry16k"]
    _, taps, _ = gpu_process(pcm[None], rate, RG, nonlinear=1.0, feedback=0.0, match_matlab=True)
    check_real_speech(taps["tension"][0], taps["speed"][0], strict=False)
    # and the formula itself, speedy.c:768-777, on the tapped tension
    t = taps["tension"][0].astype(np.float32)
    want = np.maximum(np.float32(1), np.float32(RG) + (np.float32(1) - np.float32(RG)) * t)
    assert np.allclose(taps["speed"][0], want, rtol=1e-6, atol=0)


@pytest.mark.gpu
def test_chirp_speedup_gpu():
    import speedy_b200 as sb
    out = drive_chirp(sb.lib(), lambda a: a.ctypes.data)
    s = check_chirp(out)
    if ol.ref_available("fftw"):  # and the CPU reference's samples, bit for bit (linear path)
        assert np.array_equal(out, drive_chirp(ol.ref("fftw"), ol.sptr)), s


@pytest.mark.gpu
def test_long_stereo_speech_range_gpu():
    """sonic_classic_test.cc:539-556 on a stand-in for capture_1_00x.wav (not in the reference's
    test_data): 20 s of the synthetic 48 kHz stereo signal, speeds 1.1 .. 6.1 in steps of 0.5 as
    eleven streams of one batch, final length within 300 ms."""
    from gpu_util import gpu_process
    import speedy_b200 as sb
    rate, frames = 48000, 48000 * 20
    speeds = [1.1 + 0.5 * i for i in range(11)]
    pcm = np.repeat(ol.synth(5, 1, rate, 2, frames), len(speeds), axis=0)
    cap = frames + 8192
    b = sb.Batch(len(speeds), rate, 2, speed=2.0, nonlinear=0.0, max_write_frames=frames, out_capacity=cap, taps=0)
    b.set_speed(np.array(speeds, np.float32))
    b.write(pcm)
    b.flush()
    out, counts = b.read(cap)
    b.close()
    for speed, n in zip(speeds, counts):
        assert abs(int(n) - int(frames / speed)) <= 300 * rate // 1000, (speed, int(n))
