"""The assertions of the reference's TestTapestryFeatureComputations
(/root/reference/speedy_test.cc:859-1057) against the Matlab dumps the reference holds
(tests/golden/matlab_tapestry.npz, made by tests/golden/make_golden.py), with the reference's
own helpers restated: ComputeSNR (:807-811, float accumulation), ExtractPortion (:836-843,
which drops the last element of the range), FindCrossCorrelation (:845-857), the SNR floors
(27 dB at zero delay, larger than at every other delay in -20..19) and the best-delay /
threshold table (:1007-1020)."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

FEATURES = [  # (name, best delay, SNR floor) -- speedy_test.cc:1007-1020
    ("Spectrogram energy", 0, 2e5), ("Energy Lowpass", 8, 7e5), ("Energy Local", 8, 4e4),
    ("Energy Compressed", 8, 9e5), ("Energy Hysteresis", 0, 320), ("Low Energy Frame", 0, 1e8),
    ("Local Spectral Difference", 0, 19), ("Emphasis Weighted Local Difference", 0, 29),
    ("Emphasis Weighted Lowpass Filter", -1, 2300), ("Relative Spectral Difference", 0, 28),
    ("Speech Changes", 0, 7), ("Audio Tension", 0, 8),
]


def load():
    return np.load(os.path.join(HERE, "golden", "matlab_tapestry.npz"))


def snr(signal, estimate):
    s = np.ascontiguousarray(signal, np.float32)
    e = np.ascontiguousarray(estimate, np.float32)
    err = np.sum((s - e) ** 2, dtype=np.float32)
    with np.errstate(divide="ignore"):
        return float(np.float32(np.sum(s * s, dtype=np.float32)) / err)


def portion(a, start, count):
    end = min(start + count, len(a))
    return a[start:end - 1]


def cross_correlation(a, b, num_delays):
    out = []
    for delay in range(-num_delays, num_delays + 1):
        if delay < 0:
            n = len(a) + delay
            out.append(snr(portion(a, -delay, n), portion(b, 0, n)))
        else:
            n = len(a) - delay
            out.append(snr(portion(a, 0, n), portion(b, delay, n)))
    return out


def check(spectrogram, normalized, features, gold=None, floors=None):
    """spectrogram [314][>=330], normalized [306][330], features [306][>=12] as the reference's
    test collects them.  Returns the measured figures; asserts the reference's bars
    (`floors`: optional replacement {name: floor} for callers that state their own)."""
    gold = gold or load()
    assert spectrogram.shape[0] == 314 and normalized.shape[0] == 306 and features.shape[0] == 306
    got = {}
    col, max_delay = 150, 20
    for name, want, have in (("spectrogram", gold["spectrogram_row150"], spectrogram[:, :330]),
                             ("normalized", gold["normalized_row150"], normalized[:, :330])):
        db = [10 * np.log10(snr(want, have[col + d])) for d in range(-max_delay, max_delay)]
        got[name + "_snr_db"] = db[max_delay]
        assert db[max_delay] > (floors or {}).get(name, 27), (name, db[max_delay])
        assert int(np.argmax(db)) == max_delay, (name, int(np.argmax(db)) - max_delay)
    energy = np.sum(normalized.astype(np.float32) ** 2, axis=1, dtype=np.float32)
    assert np.all(np.abs(energy - 1) < 4e-3)  # :975-978
    for i, (name, delay, floor) in enumerate(FEATURES):
        cc = cross_correlation(features[:, i], gold["features"][:, i], 10)
        best = int(np.argmax(cc))  # (first maximum, as the reference's strict > scan)
        got[name] = (best - 10, cc[best])
        assert best - 10 == delay, (name, best - 10, delay)
        assert cc[best] > (floors or {}).get(name, floor), (name, cc[best], floor)
    return got
