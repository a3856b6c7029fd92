"""Generate tests/golden/*.npz (run HERE, where /root/reference exists).

Inputs are the reference's own test recordings (test_data/*.wav, read in place);
outputs come from the reference's UNMODIFIED speedy.c + soniclib.c compiled into
oracle/_ref (see oracle/Makefile) and driven through its public Sonic API with
the debug callbacks registered, i.e. what speedy_wave / the reference tests do.
The third-party halves behind it (FFT, Sonic) are our restatements, see
oracle/fft_oracle.c and oracle/sonic_oracle.c.

    python tests/golden/make_golden.py
"""
import os
import sys
import wave

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402

REF_DATA = "/root/reference/test_data"


def read_wav(name):
    w = wave.open(os.path.join(REF_DATA, name))
    assert w.getsampwidth() == 2
    pcm = np.frombuffer(w.readframes(w.getnframes()), dtype=np.int16).reshape(-1, w.getnchannels())
    return pcm.copy(), w.getframerate()


def main():
    ol.build()
    inputs = {}
    for key, name in (("tapestry16k", "tapestry.wav"), ("tapestry22k", "tapestry22050.wav"),
                      ("negative24k", "negative_speed.wav")):
        pcm, rate = read_wav(name)
        inputs[key] = (pcm, rate)
    np.savez_compressed(os.path.join(HERE, "inputs.npz"),
                        **{k: v[0] for k, v in inputs.items()},
                        **{k + "_rate": np.int32(v[1]) for k, v in inputs.items()})

    # (case name, input, kind, speed, nonlinear, feedback, chunk)
    cases = [
        # config #1 of BASELINE.json: speedy_wave --speed 3.5 --nonlinear 1 (feedback 0, 1000-frame writes)
        ("cfg1_tapestry22k_fftw", "tapestry22k", "fftw", 3.5, 1.0, 0.0, 1000),
        ("cfg1_tapestry16k_fftw", "tapestry16k", "fftw", 3.5, 1.0, 0.0, 1000),
        # how the reference's own tests are built (-DKISS_FFT -DMATCH_MATLAB), library default feedback
        ("tapestry16k_kiss_3x", "tapestry16k", "kiss", 3.0, 1.0, 0.1, 128),
        ("tapestry16k_fftw_2x", "tapestry16k", "fftw", 2.0, 1.0, 0.1, 160),
        ("tapestry22k_kiss_3x", "tapestry22k", "kiss", 3.0, 1.0, 0.1, 128),
        # linear Sonic path and slow-down
        ("tapestry16k_linear_2x", "tapestry16k", "fftw", 2.0, 0.0, 0.1, 1024),
        ("tapestry16k_linear_1p5x", "tapestry16k", "fftw", 1.5, 0.0, 0.1, 1024),
        ("tapestry16k_fftw_0p7x", "tapestry16k", "fftw", 0.7, 1.0, 0.1, 1000),
        ("negative24k_fftw_0p25x", "negative24k", "fftw", 0.25, 1.0, 0.1, 0),
    ]
    out = {}
    for name, key, kind, speed, nonlinear, feedback, chunk in cases:
        pcm, rate = inputs[key]
        r = ol.ref_process(kind, pcm, rate, pcm.shape[1], speed, nonlinear, feedback, chunk=chunk,
                           taps=nonlinear != 0)
        out[name + "/out"] = r["out"]
        out[name + "/params"] = np.array([rate, pcm.shape[1], speed, nonlinear, feedback, chunk,
                                          1.0 if kind == "kiss" else 0.0], np.float64)
        if nonlinear != 0:
            out[name + "/tension"] = r["tension"]
            out[name + "/speed"] = r["speed"]
            out[name + "/features"] = r["features"]
            # frame energy is features[0] delayed; keep the spectrogram of a few frames only
            spec = r["spectrogram"]
            rows = sorted({0, 1, min(50, spec.shape[0] - 1), min(150, spec.shape[0] - 1), spec.shape[0] - 1})
            out[name + "/spec_rows"] = np.array(rows, np.int32)
            out[name + "/spec"] = spec[rows]
            out[name + "/n_spec"] = np.int32(spec.shape[0])
        print(name, "out frames", r["out"].shape[0],
              "tensions", len(r.get("tension", [])))
    np.savez_compressed(os.path.join(HERE, "reference_outputs.npz"), **out)
    matlab_goldens()
    print("wrote", os.listdir(HERE))


def matlab_goldens():
    """The reference-held Matlab dumps of TestTapestryFeatureComputations
    (speedy_test.cc:859-1057; test_data/tapestry_{spectrogram,normalized_spectrogram,
    features}_data.txt): the test reads the twelve feature columns in full but only time
    step 150 of the two spectrograms, so that is what the fixture keeps (float32, as the
    test's ReadFloatMatrix parses them)."""
    spec = np.loadtxt(os.path.join(REF_DATA, "tapestry_spectrogram_data.txt"), dtype=np.float32)
    norm = np.loadtxt(os.path.join(REF_DATA, "tapestry_normalized_spectrogram_data.txt"), dtype=np.float32)
    feat = np.loadtxt(os.path.join(REF_DATA, "tapestry_features_data.txt"), dtype=np.float32)
    assert spec.shape == (314, 330) and norm.shape == (314, 330) and feat.shape == (314, 12)
    np.savez_compressed(os.path.join(HERE, "matlab_tapestry.npz"), spectrogram_row150=spec[150],
                        normalized_row150=norm[150], features=feat, shape=np.array(spec.shape, np.int32))


if __name__ == "__main__":
    main()
