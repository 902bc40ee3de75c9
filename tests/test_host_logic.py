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


# ---------------------------------------------------------------------------------------------- Remez designer
DECODER_FILTERS = {
    # decoder: (sampleRate, passBandCutoff, stopBandStart, passBandRipple, stopBandRipple, oddLength, taps)
    "c4fm": (50000.0, 5100, 6500, 0.01, 0.01, None, 72),       # P25P1DecoderC4FM.java:136-148
    "hdqpsk": (50000.0, 6500, 7200, 0.005, 0.01, None, 154),   # P25P2DecoderHDQPSK.java:155-166
    "nbfm": (50000.0, 10000, 12500, 0.01, 0.005, True, 45),    # NBFMDecoder.java:306-325 (sampleRate = 2 x 25 kHz)
}


def _product_taps(fs, p, s, rp, rs, odd, order=0, density=16):
    from sdrtrunk_b200.dsp import FilterFactory, FIRFilterSpecification
    b = (FIRFilterSpecification.lowPassBuilder().sampleRate(fs).passBandCutoff(p).passBandAmplitude(1.0).passBandRipple(rp)
         .stopBandAmplitude(0.0).stopBandStart(s).stopBandRipple(rs).gridDensity(density))
    if order:
        b = b.order(order)
    if odd is not None:
        b = b.oddLength(odd)
    return FilterFactory.getTaps(b.build())


@pytest.mark.parametrize("decoder", sorted(DECODER_FILTERS))
def test_remez_decoder_filters_two_restatements_agree_bit_for_bit(decoder):
    """RemezFIRFilterDesigner (J/dsp/filter/fir/remez/RemezFIRFilterDesigner.java:52-672) restated twice -- oracle/orc_remez.c
    statement for statement, sdrtrunk_b200/csrc/remez.cpp on its own structure: the taps the decoders run must agree to
    the last float bit, have the reference's estimated lengths, and be the equiripple low-pass the specification asks for."""
    import scipy.signal as ss
    fs, p, s, rp, rs, odd, n = DECODER_FILTERS[decoder]
    want = oracle.remez_low_pass(fs, p, s, rp, rs, odd_length=odd)
    got = _product_taps(fs, p, s, rp, rs, odd)
    assert want is not None and got is not None
    assert got.size == n == want.size
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert np.array_equal(got, got[::-1])                               # linear phase
    w, h = ss.freqz(got.astype(np.float64), worN=8192, fs=fs)
    mag = np.abs(h)
    assert np.max(np.abs(mag[w <= p] - 1.0)) < 0.05 and np.max(mag[w >= s]) < 0.012     # ~ -40 dB stop band
    # equiripple: the stop-band error touches its bound at (nearly) every lobe
    lobes = mag[w >= s]
    peaks = lobes[1:-1][(lobes[1:-1] > lobes[:-2]) & (lobes[1:-1] > lobes[2:])]
    assert peaks.size > 5 and np.min(peaks) > 0.7 * np.max(peaks)
    # same design problem as the textbook algorithm: close to scipy's remez with the specification's weights
    rip = lambda db: (10 ** (db / 20) - 1) / (10 ** (db / 20) + 1)
    ref = ss.remez(n, [0, p, s, fs / 2], [1, 0], weight=[1 / rip(rp), 1 / rip(rs)], fs=fs)
    assert np.max(np.abs(got - ref)) < 1e-2                              # (the Java resamples its polynomial on a slightly different grid than it inverts on)


def test_remez_order_estimate_and_other_specifications():
    from sdrtrunk_b200.dsp import FIRFilterSpecification
    for args in ((50000.0, 5100, 6500, 0.01, 0.01), (50000.0, 6500, 7200, 0.005, 0.01), (50000.0, 10000, 12500, 0.01, 0.005),
                 (25000.0, 3000, 4000, 0.1, 0.02), (48000.0, 3000, 3400, 0.02, 0.001)):
        assert FIRFilterSpecification.estimateFilterOrder(*args) == oracle.remez_estimate_order(*args)
    assert FIRFilterSpecification.estimateFilterOrder(50000.0, 5100, 6500, 0.01, 0.01) == 71
    rng = np.random.default_rng(12)
    agreed = 0
    for _ in range(12):                                                 # random specifications, explicit orders, both parities
        fs = float(rng.choice([8000, 25000, 48000, 50000]))
        p = float(rng.uniform(0.05, 0.3) * fs)
        s = p + float(rng.uniform(0.03, 0.1) * fs)
        rp, rs = float(rng.uniform(0.005, 0.1)), float(rng.uniform(0.002, 0.05))
        order = int(rng.integers(20, 90))
        odd = [None, True, False][int(rng.integers(0, 3))]
        density = int(rng.choice([8, 16]))
        want = oracle.remez_low_pass(fs, p, s, rp, rs, order=order, odd_length=odd, grid_density=density)
        got = _product_taps(fs, p, s, rp, rs, odd, order, density)
        assert (want is None) == (got is None)                          # both fail to converge, or neither
        if want is not None:
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
            agreed += 1
    assert agreed >= 6
