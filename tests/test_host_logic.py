"""Host-side (run-once) logic of the product against the oracle: filter designers, channel calculator."""
import numpy as np
import pytest

import oracle
from sdrtrunk_b200 import native
from sdrtrunk_b200.dsp import ChannelCalculator, ComplexPolyphaseChannelizerM2, FilterFactory, TunerChannel, WindowType


@pytest.mark.parametrize("channels,length", [(96, 864), (400, 3600), (800, 7200), (114, 1026)])
def test_channelizer_prototype_matches_oracle(channels, length):
    got = FilterFactory.getSincM2Channelizer(25000.0, channels, 9)
    want = oracle.sinc_m2_channelizer(25000.0, channels, 9)
    assert got.size == length
    assert np.array_equal(got, want)


def test_synthesizer_and_half_band_match_oracle():
    assert np.array_equal(FilterFactory.getSincM2Synthesizer(50000.0, 25000.0, 2, 9),
                          oracle.sinc_m2_synthesizer(50000.0, 25000.0, 2, 9))
    for length, w, name in [(63, WindowType.HAMMING, "hamming"), (23, WindowType.BLACKMAN, "blackman"),
                            (15, WindowType.BLACKMAN, "blackman"), (11, WindowType.BLACKMAN, "blackman")]:
        assert np.array_equal(FilterFactory.getHalfBand(length, w), oracle.half_band(length, name))
    with pytest.raises(native.IllegalArgumentException):
        FilterFactory.getHalfBand(13, WindowType.HAMMING)


@pytest.mark.parametrize("rate,count", [(2.4e6, 96), (1e7, 400), (2e7, 800), (2.88e6, 114), (1.75e6, 70)])
def test_channel_count(rate, count):
    assert ComplexPolyphaseChannelizerM2.getChannelCount(rate) == count


def test_channel_calculator_matches_oracle_sweep():
    rng = np.random.default_rng(5)
    for fs, m in [(1e7, 400), (2.4e6, 96), (2e7, 800)]:
        centre = 851000000
        mine = ChannelCalculator(fs, m, centre)
        ref = oracle.ChannelCalculator(fs, m, centre)
        freqs = list(rng.integers(int(centre - fs / 2) + 20000, int(centre + fs / 2) - 20000, 300))
        # exact bin centres and bin boundaries exercise the boundary policies
        freqs += [centre + k * 25000 for k in range(-m // 2 + 1, m // 2)]
        freqs += [centre + k * 25000 + 12500 for k in range(-m // 2 + 1, m // 2 - 1)]
        for f in freqs:
            for bw in (12500, 25000):
                try:
                    want = ref.channel_indexes(int(f), bw)
                except ValueError:
                    with pytest.raises(native.IllegalArgumentException):
                        mine.getChannelIndexes(TunerChannel(int(f), bw))
                    continue
                got = mine.getChannelIndexes(TunerChannel(int(f), bw))
                assert got == want, (fs, f, bw)
                assert mine.getCenterFrequencyForIndexes(got) == ref.center_frequency_for_indexes(want)


def test_channel_calculator_known_cases():
    c = ChannelCalculator(1e7, 400, 850000000)
    assert c.getChannelIndexes(TunerChannel(850000000, 12500)) == [0]
    assert c.getChannelIndexes(TunerChannel(850025000, 12500)) == [1]
    assert c.getChannelIndexes(TunerChannel(849975000, 12500)) == [399]
    assert c.getChannelIndexes(TunerChannel(850012500, 12500)) == [0, 1]   # straddles the bin boundary
    assert c.getCenterFrequencyForIndexes([1]) == 850025000
    assert c.getCenterFrequencyForIndexes([0, 1]) == 850012500
    with pytest.raises(native.IllegalArgumentException):
        c.getChannelIndexes(TunerChannel(860000000, 12500))               # outside the tuner bandwidth
