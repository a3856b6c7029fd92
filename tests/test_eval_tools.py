"""The evaluation tools of include/speedy_eval.h against the
reference's own known answers: dynamic_time_warping_test.cc:31-81, and the Teager /
slope helpers of sonic_test.cc:85-197 against numpy restatements."""
import numpy as np

from speedy_b200 import evaluation as ev


def seq(*v):
    return np.array(v, np.float32)[:, None]


def test_dtw_identical_sequences():
    s = seq(0, 1, 2, 3, 4)
    cost, p1, p2 = ev.dtw(s, s)
    assert cost == 0.0 and np.array_equal(p1, p2) and list(p1) == [0, 1, 2, 3, 4]


def test_dtw_shifted_sequences():
    cost, p1, p2 = ev.dtw(seq(0, 1, 2, 3, 4), seq(-2, -1, 0, 1, 2))
    assert cost == 6.0
    assert list(p1) == [0, 0, 0, 1, 2, 3, 4] and list(p2) == [0, 1, 2, 3, 4, 4, 4]


def test_dtw_downsampled_sequence():
    cost, p1, p2 = ev.dtw(seq(0, 1, 2, 3, 4), seq(0, 2, 4))
    assert cost == 2.0
    assert list(p1) == [0, 1, 2, 3, 4] and list(p2) == [0, 0, 1, 1, 2]


def test_dtw_path_constraints_random():
    rng = np.random.default_rng(3)
    a, b = rng.normal(size=(37, 6)).astype(np.float32), rng.normal(size=(23, 6)).astype(np.float32)
    cost, p1, p2 = ev.dtw(a, b)
    assert (p1[0], p2[0]) == (0, 0) and (p1[-1], p2[-1]) == (36, 22)
    steps = set(zip(np.diff(p1), np.diff(p2)))
    assert steps <= {(1, 0), (0, 1), (1, 1)}
    d = np.sqrt(((a[p1] - b[p2]) ** 2).sum(axis=1))
    assert abs(d.sum() - cost) < 1e-3 * cost
    # brute-force optimal cost
    D = np.sqrt(((a[:, None] - b[None]) ** 2).sum(-1)).astype(np.float64)
    acc = np.full(D.shape, np.inf)
    acc[0] = np.cumsum(D[0]); acc[:, 0] = np.cumsum(D[:, 0])
    for i in range(1, D.shape[0]):
        for j in range(1, D.shape[1]):
            acc[i, j] = D[i, j] + min(acc[i - 1, j], acc[i, j - 1], acc[i - 1, j - 1])
    assert abs(acc[-1, -1] - cost) < 1e-4 * cost


def sinusoid(rate=22050, seconds=1.0, pitch=237.0):
    period = np.float32(rate) / np.float32(pitch)
    i = np.arange(int(seconds * rate))
    return (32000 * np.sin(i * 2 * np.pi / period)).astype(np.int16)   # sonic_test.cc:296-316


def test_teager_of_a_sinusoid_is_constant():
    x = sinusoid()
    m, v = ev.teager_variance(x)
    t = x[1:-1].astype(np.float64) ** 2 - x[:-2].astype(np.float64) * x[2:]
    assert abs(m - t.mean()) < 1e-4 * t.mean()
    assert abs(v - t.var()) < 2e-2 * t.var()
    assert np.sqrt(v) / m < 0.01                                          # sonic_test.cc:531
    assert np.abs(ev.teager(x) - t).max() <= 128      # float products of ~1e9, as in the reference
    mf, vf = ev.teager_variance((x / 32768.0).astype(np.float32))
    assert abs(mf * 32768.0 ** 2 - m) < 1e-3 * m and np.sqrt(vf) / mf < 0.01
    # a phase jump shows up as outliers
    broken = np.concatenate([x[:5000], x[5037:]])
    assert ev.teager_outlier_count(x, 0.05) == 0
    assert ev.teager_outlier_count(broken, 0.05) >= 1


def test_linear_slope_helpers():
    x = np.arange(100)
    y = (x * 0.5).astype(np.int32)
    assert abs(ev.linear_slope(x, y) - 0.5) < 0.01
    s = ev.linear_slope_everywhere(x, y, 10)
    assert len(s) == 80 and abs(ev.mean(s) - 0.5) < 0.02
    assert abs(ev.standard_deviation(s) - float(np.std(s))) < 1e-6
    assert len(ev.linear_slope_everywhere(x[:10], y[:10], 10)) == 0
