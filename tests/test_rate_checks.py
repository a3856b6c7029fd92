"""CPU checks of the numpy restatement of upstream Sonic's classic rate change (tests/rate_checks.py),
the checker tests/test_set_rate.py holds the library's sonicSetRate against on the GPU."""
import numpy as np
import pytest

from rate_checks import resample


def tone(sample_rate, hz, n, channels=1):
    x = (8000 * np.sin(2 * np.pi * hz * np.arange(n) / sample_rate)).astype(np.int16)
    return np.repeat(x[:, None], channels, 1)


@pytest.mark.parametrize("sample_rate", [8000, 16000, 22050, 48000])
@pytest.mark.parametrize("rate", [0.5, 0.8, 1.25, 1.5, 2.0, 3.0])
def test_length_and_pitch(sample_rate, rate):
    n = sample_rate // 2
    x = tone(sample_rate, 200.0, n)
    y = resample(x, sample_rate, rate)
    assert abs(len(y) - (n - 1) / rate) <= 2 + 0.002 * n  # (rates above 2^14 are halved: int truncation)
    spec = np.abs(np.fft.rfft(y[:, 0] * np.hanning(len(y))))
    assert abs(np.argmax(spec) * sample_rate / len(y) - 200.0 * rate) <= 2.0 * sample_rate / len(y)


def test_rate_one_is_the_identity_less_the_frame_kept_back():
    x = tone(16000, 150.0, 4000, 2) + np.arange(8000, dtype=np.int16).reshape(4000, 2) % 7
    assert np.array_equal(resample(x, 16000, 1.0), x[:-1])


def test_interpolation_stays_between_neighbours_and_channels_are_independent():
    rng = np.random.default_rng(5)
    x = rng.integers(-32768, 32767, (3000, 2)).astype(np.int16)
    y = resample(x, 16000, 0.7)
    assert y.min() >= x.min() and y.max() <= x.max()
    assert np.array_equal(resample(x[:, :1], 16000, 0.7), y[:, :1])


def test_chunked_resampling_equals_one_shot():
    """The library resamples whatever each step produced and keeps one frame back: the positions
    carry across calls, so feeding the frames in pieces gives the same stream."""
    x = tone(16000, 310.0, 5000)
    whole = resample(x, 16000, 1.5)
    # restate the carry by hand: resample(prefix) is a prefix of resample(whole)
    for cut in (1, 2, 333, 2500, 4999):
        part = resample(x[:cut + 1], 16000, 1.5)
        assert np.array_equal(part, whole[:len(part)])
