"""-m gpu, round 2: robustness and parity hardening through the C ABI.

  * the pipelined resynthesis kernel (k4_splice.cu) and the one-warp kernel produce the
    same bytes;
  * write after flush (upstream sonicFlushStream leaves the stream usable);
  * two batches on two devices in one process;
  * end-to-end bit-equality RATE with the library's own speeds on >= 64 streams per
    BASELINE.json configuration (recorded under profiles/ when SPEEDY_RECORD_PARITY is set);
  * element-wise relative error (with a stated absolute floor) beside the scale-relative bar.
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

import speedy_b200 as sb  # noqa: E402
from gpu_util import gpu_process, out_capacity  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def oracle_run(pcm, rate, speed, nonlinear=1.0, feedback=0.1, override=None):
    c = ol.cfg(rate, pcm.shape[1], speed, nonlinear, feedback, False, False)
    return ol.port_process(c, pcm, speed_override=override)


class _env:
    def __init__(self, **kv):
        self.kv = kv

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kv}
        for k, v in self.kv.items():
            os.environ[k] = str(v)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("rate,speed,nonlinear", [(16000, 2.0, 1.0), (16000, 3.5, 1.0), (16000, 1.3, 1.0), (16000, 0.6, 1.0),
                                                  (16000, 2.0, 0.0), (16000, 1.0, 0.0), (8000, 2.0, 1.0),
                                                  (22050, 2.5, 1.0), (32000, 1.7, 1.0), (48000, 1.5, 1.0)])
def test_pipelined_kernel_equals_one_warp_kernel(rate, speed, nonlinear):
    """k4_splice.cu (chain / filler / output warps, TMA bulk loads, selected with
    SPEEDY_K4_PIPELINE=1) against the default one-warp kernel: same output bytes, same
    stream state afterwards (a second write continues identically)."""
    n = 64 if rate == 16000 else 8
    frames = rate * 3 + 123
    pcm = ol.synth(2024, n, rate, 1, frames)
    launches = sb.kernel_launches()
    a, _, st_a = gpu_process(pcm, rate, speed, nonlinear=nonlinear, taps=0, chunk=frames // 2 + 7)
    with _env(SPEEDY_K4_PIPELINE=1, SPEEDY_K4_SPLICE_MIN=0):
        b, _, st_b = gpu_process(pcm, rate, speed, nonlinear=nonlinear, taps=0, chunk=frames // 2 + 7)
    assert sb.kernel_launches() > launches
    assert np.array_equal(st_a, st_b)
    for s in range(n):
        assert np.array_equal(a[s], b[s]), (s, len(a[s]), len(b[s]))


@pytest.mark.parametrize("speed,nonlinear,chunk", [(2.0, 1.0, None), (3.5, 1.0, 40000), (1.3, 1.0, 33333), (0.6, 1.0, 50001),
                                                   (2.0, 0.0, 48000), (1.0, 0.0, 40008), (2.5, 0.5, 17777), (6.0, 1.0, None)])
def test_chain_kernel_equals_one_warp_kernel(speed, nonlinear, chunk):
    """k4_chain16.cu (SPEEDY_K4_CHAIN=1, 16 kHz mono writes of a second or more: fixed overlapping
    windows prefetched by TMA bulk copies, output deferred into the next search) against the
    general one-warp kernel (SPEEDY_K4_CHAIN=0): same output bytes and status for whole writes,
    for chunks that leave the stream at frame counts that are not multiples of eight (the bulk
    copies then fall back to plain loads) and for rows that are not 16-byte aligned."""
    n, rate = 48, 16000
    frames = rate * 10 + (0 if chunk is None else 123)
    pcm = ol.synth(777, n, rate, 1, frames)
    launches = sb.kernel_launches()
    with _env(SPEEDY_K4_CHAIN=1):
        a, _, st_a = gpu_process(pcm, rate, speed, nonlinear=nonlinear, taps=0, chunk=chunk)
    with _env(SPEEDY_K4_CHAIN=0):
        b, _, st_b = gpu_process(pcm, rate, speed, nonlinear=nonlinear, taps=0, chunk=chunk)
    assert sb.kernel_launches() > launches
    assert np.array_equal(st_a, st_b)
    for s in range(n):
        assert np.array_equal(a[s], b[s]), (s, len(a[s]), len(b[s]))


@pytest.mark.parametrize("speed", [2.0, 0.7])
def test_write_after_flush_linear(speed):
    """sonicFlushStream leaves the stream usable (upstream sets numInputSamples = 0): a write
    after a flush loses nothing.  Checked against the compiled reference driven the same way."""
    if not ol.ref_available("kiss"):
        pytest.skip("compiled reference not built")
    import ctypes as C
    R = ol.ref("kiss")
    rate = 16000
    pcm = ol.synth(99, 1, rate, 1, 12000)[0]
    pieces = [pcm[:5000], pcm[5000:]]
    h = R.sonicCreateStream(rate, 1)
    R.sonicSetSpeed(h, speed)
    want = []
    buf = np.zeros(40000, np.int16)
    for piece in pieces:
        x = np.ascontiguousarray(piece[:, 0])
        assert R.sonicWriteShortToStream(h, ol.sptr(x), len(x)) == 1
        assert R.sonicFlushStream(h) == 1
        got = R.sonicReadShortFromStream(h, ol.sptr(buf), len(buf))
        want.append(buf[:got].copy())
    R.sonicDestroyStream(h)
    cap = out_capacity(12000, speed, 0.0, 2 * (rate // 65))
    b = sb.Batch(1, rate, 1, speed=speed, nonlinear=0.0, max_write_frames=8000, out_capacity=cap)
    for i, piece in enumerate(pieces):
        b.write(np.ascontiguousarray(piece[None]))
        b.flush()
        o, c = b.read(cap)
        assert np.array_equal(o[0, :c[0], 0], want[i]), (i, c[0], len(want[i]))
    b.close()


def test_two_devices_in_one_process():
    """One batch per device in the same process (the per-device shared-memory opt-in):
    identical bytes from both devices."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    rate, n, frames = 16000, 16, 16000 * 3
    pcm = ol.synth(4242, n, rate, 1, frames)
    outs = []
    for dev in (0, 1):
        cap = out_capacity(frames, 2.0, 1.0, 2 * (rate // 65))
        b = sb.Batch(n, rate, 1, speed=2.0, nonlinear=1.0, feedback=0.1, max_write_frames=frames,
                     out_capacity=cap, device=dev)
        b.write(pcm)
        b.flush()
        o, c = b.read(cap)
        outs.append([o[s, :c[s]].copy() for s in range(n)])
        b.close()
    for s in range(n):
        assert np.array_equal(outs[0][s], outs[1][s]), s
    # 48 kHz stereo needs the larger opt-ins (mixed-radix spectrogram kernel, wider Sonic window):
    # first use on device 1, then on device 0
    pcm2 = ol.synth(7, 2, 48000, 2, 48000)
    cap = out_capacity(48000, 1.5, 1.0, 2 * (48000 // 65))
    outs = []
    for dev in (1, 0):
        b = sb.Batch(2, 48000, 2, speed=1.5, nonlinear=1.0, feedback=0.1, max_write_frames=48000,
                     out_capacity=cap, device=dev)
        b.write(pcm2)
        b.flush()
        o, c = b.read(cap)
        outs.append([o[s, :c[s]].copy() for s in range(2)])
        b.close()
    for s in range(2):
        assert np.array_equal(outs[0][s], outs[1][s]), s


# (name, rate, channels, speed, seconds, chunk): the shapes of BASELINE.json configs 2-5
CONFIG_SHAPES = [
    ("config2_16k_mono_2.0x", 16000, 1, 2.0, 6, None),
    ("config3_16k_mono_3.5x", 16000, 1, 3.5, 6, None),
    ("config4_48k_stereo_1.5x", 48000, 2, 1.5, 2, None),
    ("config5_16k_mono_2.5x_10ms_chunks", 16000, 1, 2.5, 3, 160),
]


@pytest.mark.parametrize("name,rate,channels,speed,secs,chunk", CONFIG_SHAPES)
def test_end_to_end_equality_rate(name, rate, channels, speed, secs, chunk):
    """With the library's OWN per-frame speeds (no override) a stream's output is bit-identical
    to the CPU reference path's unless a float-stage rounding difference moved a speed: count
    both on 64 streams per configuration, find the first diverging frame, and hold the float
    stage to the 1e-4 bar on every stream.  The integer stage is then checked bit-exactly on
    the library's speeds for every stream."""
    n = 64
    frames = rate * secs
    pcm = ol.synth(31337, n, rate, channels, frames)
    outs, taps, status = gpu_process(pcm, rate, speed, taps=sb.TAP_SPEED | sb.TAP_TENSION, chunk=chunk)
    speed_equal = out_equal = 0
    first_div = []
    worst_rel = 0.0
    n_frames = n_out = n_attr = 0
    N = int(ol.geometry(rate).fft)
    for s in range(n):
        o = oracle_run(pcm[s], rate, speed)
        sp, spo = taps["speed"][s], o["speed"]
        assert len(sp) == len(spo), s
        rel = np.abs(sp.astype(np.float64) - spo) / np.maximum(np.abs(spo), 1e-30)  # speeds are O(1): no floor
        n_frames += len(rel)
        bad = np.nonzero(rel >= 1e-4)[0]
        if len(bad):
            # The one discontinuity of the float stage that a 1e-7 difference in a magnitude can
            # trip: a bin enters the spectral difference only above max / 100 (speedy.c:705-719), so
            # a bin within rounding of that threshold can be in on one side and out on the other
            # (SURVEY.md section 7).  The first frame off by more than 1e-4 must be such a frame
            # (later ones carry it in the one-pole filter's memory); they are counted, and bounded.
            r = int(bad[0])
            spec = o["spectrogram"].astype(np.float64)
            near = 1.0
            for row in (r - 1, r - 2):  # at_time r and r - 1 are rows r - 1 and r - 2
                if 0 <= row < len(spec):
                    half = spec[row, 1:N // 2]
                    thr = half.max() / 100.0
                    if thr > 0:
                        near = min(near, float(np.min(np.abs(half / thr - 1.0))))
            n_attr += int(near < 5e-5)
            assert near < 5e-5, (s, r, float(rel[r]), near)
            n_out += len(bad)
            assert float(rel.max()) < 5e-2, (s, float(rel.max()))
        worst_rel = max(worst_rel, float(rel.max()))
        same = np.array_equal(sp, spo)
        speed_equal += int(same)
        if not same:
            first_div.append(int(np.nonzero(sp != spo)[0][0]))
        out_equal += int(outs[s].shape == o["out"].shape and np.array_equal(outs[s], o["out"]))
        o2 = oracle_run(pcm[s], rate, speed, override=sp)
        assert np.array_equal(outs[s], o2["out"]), s
    assert n_out <= max(2, n_frames // 500), (n_out, n_frames)
    rec = {"config": name, "streams": n, "seconds": secs, "frames_per_stream": int(len(taps["speed"][0])),
           "streams_with_bit_identical_speeds": speed_equal, "streams_with_bit_identical_output": out_equal,
           "first_diverging_frame_min": min(first_div) if first_div else None,
           "first_diverging_frame_median": float(np.median(first_div)) if first_div else None,
           "worst_speed_relative_error": worst_rel, "frames": n_frames, "frames_off_by_1e-4_or_more": n_out,
           "streams_whose_first_such_frame_sits_on_the_bin_gate": n_attr}
    print(json.dumps(rec))
    if os.environ.get("SPEEDY_RECORD_PARITY"):
        path = os.path.join(ROOT, "gpurun_out", "r02_parity_rates.jsonl")
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "a") as f:
            f.write(json.dumps(rec) + "\n")
    # a stream whose speeds are identical must give identical bytes
    assert out_equal >= speed_equal


def test_elementwise_relative_error(golden_inputs):
    """north_star: "within 1e-4 relative".  Element by element, with an absolute floor where an
    element can be arbitrarily small next to its neighbours:
      spectrogram  |d| <= 1e-4 |ref| + 4e-7 * (frame peak)   float32 FFT round-off is a few ulp of
                                                             the LARGEST bin, whatever the bin's size
      energy       |d| <= 1e-4 |ref|                          (sum of squares: no cancellation)
      tension      |d| <= 1e-4 |ref| + 2e-5                   tension = 0.5 (h - 0.7) + 0.25 (c - 1)
                                                             crosses zero: the floor is 1e-4 of the
                                                             terms' O(0.2) size
      speed        |d| <= 1e-4 |ref|
    and the plain element-wise figures are reported for the bins above -60 dB."""
    pcm, rate = golden_inputs["tapestry16k"]
    o = oracle_run(pcm, rate, 3.0)
    _, taps, _ = gpu_process(pcm[None], rate, 3.0)
    spec, ref = taps["spectrogram"][0].astype(np.float64), o["spectrogram"].astype(np.float64)
    peak = ref.max(axis=1, keepdims=True)
    d = np.abs(spec - ref)
    assert np.all(d <= 1e-4 * np.abs(ref) + 4e-7 * peak), float(np.max(d - 1e-4 * np.abs(ref) - 4e-7 * peak))
    loud = ref > 1e-3 * peak
    rel_loud = float(np.max(d[loud] / ref[loud]))
    assert rel_loud < 1e-4, rel_loud
    e, eo = taps["energy"][0].astype(np.float64), o["energy"].astype(np.float64)
    assert np.all(np.abs(e - eo) <= 1e-4 * np.abs(eo) + 1e-30)
    t, to = taps["tension"][0].astype(np.float64), o["tension"].astype(np.float64)
    assert np.all(np.abs(t - to) <= 1e-4 * np.abs(to) + 2e-5), float(np.max(np.abs(t - to)))
    sp, spo = taps["speed"][0].astype(np.float64), o["speed"].astype(np.float64)
    assert np.all(np.abs(sp - spo) <= 1e-4 * np.abs(spo))
    print("element-wise: spectrogram (bins above -60 dB) %.2e, tension abs %.2e, speed rel %.2e"
          % (rel_loud, float(np.max(np.abs(t - to))), float(np.max(np.abs(sp - spo) / np.abs(spo)))))


def test_tensor_core_spectrogram_equals_fft_kernel():
    """The 16 kHz spectrogram as a tcgen05 GEMM (k1_dft16.cu, the default) against the FFT kernel
    (k1_spectral_480, SPEEDY_K1_TC=0) on the same streams: every tap within 1e-5 of the other's scale
    (measured 2.5e-6: split-fp16 products carry ~2^-21 per term), in long writes (one stream per tile,
    bulk-copied samples), 10 ms chunks (32 streams per tile, staged samples), ragged stream lengths and
    stereo (down-mix while staging); a silent and a very quiet stream included."""
    n, rate, frames = 6, 16000, 16000 * 4
    pcm = ol.synth(11, n, rate, 1, frames)
    pcm[1, :5000] = 0
    pcm[2] = (pcm[2].astype(np.int32) // 64).astype(np.int16)
    stereo = ol.synth(12, 3, rate, 2, frames // 2)
    cases = [("one write", pcm, dict()), ("10 ms chunks", pcm[:, :16000], dict(chunk=160)),
             ("ragged", pcm, dict(counts=[frames, frames - 777, 300, 0, frames // 2, 1234])), ("stereo", stereo, dict(chunk=4000))]
    old = os.environ.get("SPEEDY_K1_TC")
    try:
        for name, x, kw in cases:
            res = {}
            for tc in ("0", "1"):
                os.environ["SPEEDY_K1_TC"] = tc
                res[tc] = gpu_process(x, rate, 2.0, **kw)
            a, b = res["0"][1], res["1"][1]
            for key in ("spectrogram", "energy", "features", "tension", "speed"):
                for s in range(x.shape[0]):
                    u, v = a[key][s].astype(np.float64), b[key][s].astype(np.float64)
                    assert u.shape == v.shape, (name, key, s)
                    if u.size == 0:
                        continue
                    if key == "spectrogram":
                        d = (np.abs(u - v) / np.maximum(u.max(axis=1, keepdims=True), 1e-30)).max()
                    elif key == "features":
                        d = (np.abs(u - v) / np.maximum(np.abs(u).max(axis=0, keepdims=True), 1e-30)).max()
                    else:
                        d = np.abs(u - v).max() / max(np.abs(u).max(), 1e-30)
                    assert d < 1e-5, (name, key, s, d)
    finally:
        if old is None:
            os.environ.pop("SPEEDY_K1_TC", None)
        else:
            os.environ["SPEEDY_K1_TC"] = old
    L = sb.lib()
    L.speedyDebugK1Dft16Error.restype = C.c_int
    assert L.speedyDebugK1Dft16Error() == 0  # no barrier of the kernel ever timed out
