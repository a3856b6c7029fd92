"""speedy_wave-compatible command-line tool (tools/speedy_wave.cpp) — SURVEY.md §8f-1.

The reference's tool is /root/reference/speedy_wave.cc; its flags, defaults and the
two-pass --match_nonlinear / --length calibration are restated in ours.  CPU tests
check the argument handling (no device needed before a stream is created); the GPU
tests run whole files through it and compare with the batched path and the
committed outputs of the reference's own code.
"""
import os
import re
import subprocess
import wave

import numpy as np
import pytest

import speedy_b200 as sb
from gpu_util import gpu_process

TOOL = os.path.join(os.path.dirname(sb.LIB_PATH), "speedy_wave")


def write_wav(path, pcm, rate):
    with wave.open(str(path), "wb") as w:
        w.setnchannels(pcm.shape[1]); w.setsampwidth(2); w.setframerate(rate)
        w.writeframes(np.ascontiguousarray(pcm, dtype="<i2").tobytes())


def read_wav(path):
    with wave.open(str(path), "rb") as w:
        assert w.getsampwidth() == 2
        x = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2")
        return x.reshape(-1, w.getnchannels()), w.getframerate()


def run(*args, check=True):
    p = subprocess.run([TOOL, *map(str, args)], capture_output=True, text=True, timeout=600)
    if check:
        assert p.returncode == 0, p.stdout + p.stderr
    return p


def test_cli_usage_and_argument_errors(tmp_path):
    assert os.path.exists(TOOL), "python -m speedy_b200.build builds the tool"
    p = run(check=False)
    assert p.returncode == 255 and "--input sound.wav --output fastsound.wav" in p.stderr
    assert run("--help").stdout.startswith("Usage:")
    p = run("--input", "x.wav", check=False)           # speedy_wave.cc:405-408
    assert p.returncode == 1 and "Must specify an output file name" in p.stdout
    p = run("--output", "y.wav", check=False)          # speedy_wave.cc:409-412
    assert p.returncode == 1 and "Must specify an input file name" in p.stdout
    p = run("--input", tmp_path / "missing.wav", "--output", tmp_path / "y.wav", check=False)
    assert p.returncode == 255 and "Can't open" in p.stderr
    p = run("--input", "x.wav", "--output", "y.wav", "--speed", "-1", check=False)
    assert p.returncode == 1


@pytest.mark.gpu
def test_cli_nonlinear_matches_batch_and_reference(tmp_path, golden_inputs, golden_outputs):
    """BASELINE.json configs[0]: speedy_wave --speed 3.5 on the 22 kHz test file."""
    pcm, rate = golden_inputs["tapestry22k"]
    src, dst = tmp_path / "in.wav", tmp_path / "out.wav"
    write_wav(src, pcm, rate)
    p = run("--input", src, "--output", dst, "--speed", 3.5, "--tension_file", tmp_path / "t.txt",
            "--speed_file", tmp_path / "s.txt", "--features_file", tmp_path / "f.txt",
            "--spectrogram_file", tmp_path / "g.txt", "--normalized_spectrogram_file", tmp_path / "n.txt")
    out, r = read_wav(dst)
    assert r == rate and out.shape[1] == 1
    m = re.search(r"read (\d+) frames, and output (\d+) frames with nonlinear=1", p.stdout)
    assert m and int(m.group(1)) == len(pcm) and int(m.group(2)) == len(out)
    # same library through the batched API: identical samples
    outs, taps, _ = gpu_process(pcm[None], rate, 3.5, feedback=0.0)
    assert np.array_equal(out, outs[0])
    # the reference's own run of this command (tests/golden, FFTW build)
    case = golden_outputs["cfg1_tapestry22k_fftw"]
    tension = np.loadtxt(tmp_path / "t.txt")
    speed = np.loadtxt(tmp_path / "s.txt")
    assert len(tension) == len(case["tension"]) == len(speed)
    scale = np.abs(case["tension"]).max()
    assert np.abs(tension - case["tension"]).max() / scale < 2e-4        # "%g" keeps 6 digits
    assert np.loadtxt(tmp_path / "f.txt").shape == (len(tension), 15)
    g = np.loadtxt(tmp_path / "g.txt")
    assert g.shape == (int(case["n_spec"]), case["spec"].shape[1])
    rows = case["spec_rows"]
    peak = np.maximum(case["spec"].max(axis=1, keepdims=True), 1e-12)
    assert (np.abs(g[rows] - case["spec"]) / peak).max() < 2e-4
    assert np.loadtxt(tmp_path / "n.txt").shape == g.shape
    assert abs(len(out) - len(case["out"])) <= 0.002 * len(case["out"])


@pytest.mark.gpu
def test_cli_linear_match_and_length(tmp_path, golden_inputs):
    pcm, rate = golden_inputs["tapestry16k"]
    src = tmp_path / "in.wav"
    write_wav(src, pcm, rate)
    # --linear: plain Sonic at the global speed
    run("--input", src, "--output", tmp_path / "lin.wav", "--speed", 2.0, "--linear")
    lin, _ = read_wav(tmp_path / "lin.wav")
    outs, _, _ = gpu_process(pcm[None], rate, 2.0, nonlinear=0.0, taps=0)
    assert np.array_equal(lin, outs[0])
    # --match_nonlinear: linear speed-up by what the nonlinear pass achieved (speedy_wave.cc:424-427)
    p = run("--input", src, "--output", tmp_path / "m.wav", "--speed", 3.0, "--match_nonlinear", "--nonlinear", 0)
    m, _ = read_wav(tmp_path / "m.wav")
    non, _, _ = gpu_process(pcm[None], rate, 3.0, feedback=0.0, taps=0)
    achieved = len(pcm) / len(non[0])
    got = float(re.search(r"linearly by ([0-9.eE+-]+)X", p.stdout).group(1))
    assert abs(got - achieved) < 1e-4 * achieved
    assert abs(len(m) - len(non[0])) <= 0.02 * len(non[0])
    # --length: two-pass calibration towards a duration (speedy_wave.cc:428-462)
    want = len(pcm) / rate / 2.5
    run("--input", src, "--output", tmp_path / "len.wav", "--length", want)
    ln, _ = read_wav(tmp_path / "len.wav")
    assert abs(len(ln) / rate - want) < 0.05 * want


@pytest.mark.gpu
def test_cli_stereo_round_trip(tmp_path, golden_inputs):
    pcm, rate = golden_inputs["tapestry16k"]
    st = np.stack([pcm[:, 0], (pcm[:, 0] // 2).astype(np.int16)], axis=1)
    write_wav(tmp_path / "st.wav", st, rate)
    run("--input", tmp_path / "st.wav", "--output", tmp_path / "o.wav", "--speed", 2.0)
    out, r = read_wav(tmp_path / "o.wav")
    outs, _, _ = gpu_process(st[None], rate, 2.0, feedback=0.0, taps=0)
    assert r == rate and np.array_equal(out, outs[0])
