"""Committed golden vectors (tests/golden/, minted by tools/make_golden.py from the oracle: the reference has no
fixtures for this path).  CPU half: the oracle still reproduces them (guards the checker).  GPU half: the CUDA path
through the C ABI reproduces them (bit-exact where the arithmetic is restated op for op, 1e-4 relative RMS after the
inverse DFT, dibits bit-exact)."""
import os

import numpy as np
import pytest

import oracle
import siggen as sg

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-4


def load(name):
    return np.load(os.path.join(G, name))


# ------------------------------------------------------------------------------------------------ oracle vs golden
def test_oracle_reproduces_channelizer_golden():
    g = load("channelizer_m96.npz")
    m = 96
    assert np.array_equal(oracle.sinc_m2_channelizer(25000.0, m, 9), g["taps"])
    assert np.array_equal(oracle.Channelizer(g["taps"], m).receive(g["x"], mode="f32"), g["results_f32"])
    assert np.array_equal(oracle.Channelizer(g["taps"], m).receive(g["x"], mode="raw"), g["accumulators"])
    assert sg.rel_rms(g["results_f32"], g["results_f64"]) < 1e-6
    one = oracle.OneChannelOutputProcessor(50000.0, 4, float(m))
    one.set_frequency_offset(700)
    assert np.array_equal(one.process(g["results_f32"]), g["bin4_offset700"])
    two = oracle.TwoChannelOutputProcessor(50000.0, 88, 89, g["synth"], float(m))
    two.set_frequency_offset(-300)
    assert np.array_equal(two.process(g["results_f32"]), g["bins88_89_offset_m300"])


def test_oracle_reproduces_filter_fm_chain_golden():
    g = load("filters.npz")
    assert np.array_equal(oracle.Decimator(8).decimate_complex(g["x"]), g["decimate8"])
    assert np.array_equal(oracle.ComplexFIR(g["fir"]).filter(g["x"]), g["fir72"])
    assert np.array_equal(np.concatenate([oracle.agc_block(g["x"][:2048]), oracle.agc_block(g["x"][2048:])]), g["agc"])
    f = load("fm.npz")
    assert np.array_equal(oracle.FMDemodulator(1.0).demodulate(f["x"]), f["fm"])
    assert np.array_equal(oracle.SquelchingFMDemodulator(0.01, -40.0, 4).demodulate(f["x"]), f["squelch_fm"])
    p = load("p25_chains.npz")
    for kind, okind in (("c4fm", oracle.C4FM), ("lsm", oracle.LSM), ("hdqpsk", oracle.HDQPSK), ("dmr", oracle.DMR)):
        taps = p[kind + "_fir"] if kind + "_fir" in p.files else None
        d, agc = oracle.P25Chain(okind, 50000.0, taps).receive(p[kind + "_x"], want_agc=True)
        assert np.array_equal(d, p[kind + "_dibits"]) and d.size > 350
        assert np.array_equal(agc, p[kind + "_agc"])
    c = load("converters.npz")
    assert np.array_equal(oracle.convert_samples(c["raw8"].tobytes(), "u8"), c["u8"])
    assert np.array_equal(oracle.convert_samples(c["raw8"].tobytes(), "s8"), c["s8"])
    assert np.array_equal(oracle.convert_samples(c["raw16"].tobytes(), "s16le"), c["s16le"])


def test_remez_designers_reproduce_golden_taps():
    """the decoders' baseband filters as RemezFIRFilterDesigner designs them: oracle and product (csrc/remez.cpp)"""
    from sdrtrunk_b200.dsp import FilterFactory, FIRFilterSpecification
    g = load("remez.npz")
    assert np.array_equal(oracle.c4fm_baseband_taps(), g["c4fm"]) and g["c4fm"].size == 72
    assert np.array_equal(oracle.hdqpsk_baseband_taps(), g["hdqpsk"]) and g["hdqpsk"].size == 154
    assert np.array_equal(oracle.nbfm_iq_taps(), g["nbfm"]) and g["nbfm"].size == 45
    b = FIRFilterSpecification.lowPassBuilder
    assert np.array_equal(FilterFactory.getTaps(b().sampleRate(50000).passBandCutoff(5100).passBandRipple(0.01).stopBandStart(6500)
                                                .stopBandRipple(0.01).build()), g["c4fm"])
    assert np.array_equal(FilterFactory.getTaps(b().sampleRate(50000.0).passBandCutoff(6500).passBandRipple(0.005).stopBandStart(7200)
                                                .stopBandRipple(0.01).build()), g["hdqpsk"])
    assert np.array_equal(FilterFactory.getTaps(b().sampleRate(25000.0 * 2).gridDensity(16).oddLength(True).passBandCutoff(10000)
                                                .passBandRipple(0.01).stopBandStart(12500).stopBandRipple(0.005).build()), g["nbfm"])


def test_oracle_reproduces_airspy_and_sync_golden():
    g = load("airspy_sync.npz")
    c = oracle.AirspySampleConverter()
    assert np.array_equal(c.convert(g["raw_unpacked"]), g["iq"])
    c = oracle.AirspySampleConverter()
    c.setSamplePacking(True)
    assert np.array_equal(c.convert(g["raw_packed"]), g["iq"])
    chain = oracle.P25Chain(oracle.C4FM, 50000.0, g["sync_fir"])
    chain.attach_sync(oracle.SYNC_P25_PHASE1, 50000.0)
    sym = chain.receive(g["sync_x"])
    assert np.array_equal(sym, g["sync_symbols"])
    events = (sym >> 2) & 7
    assert events[events > 0][0] == oracle.SYNC_EVENT_90_CW and np.count_nonzero(events == oracle.SYNC_EVENT_SYNC) >= 4


# ------------------------------------------------------------------------------------------------ CUDA vs golden
@pytest.mark.gpu
def test_cuda_airspy_and_sync_match_golden(gpu):
    from sdrtrunk_b200.dsp import AirspySampleConverter, Bank
    g = load("airspy_sync.npz")
    for packed, key in ((False, "raw_unpacked"), (True, "raw_packed")):
        c = AirspySampleConverter(maxSamples=8192)
        c.setSamplePacking(packed)
        assert np.array_equal(c.convert(g[key]), g["iq"])
    bank = Bank.preset(gpu.PRESET_P25_C4FM, 1, 50000.0, g["sync_fir"], max_samples_per_call=g["sync_x"].size // 2)
    bank.setSyncDetector(gpu.SYNC_P25_PHASE1)
    assert np.array_equal(bank.process(g["sync_x"].reshape(1, -1))[0], g["sync_symbols"])


@pytest.mark.gpu
def test_cuda_channelizer_matches_golden(gpu):
    from sdrtrunk_b200.dsp import ComplexPolyphaseChannelizerM2, FilterFactory
    g = load("channelizer_m96.npz")
    m = 96
    assert np.array_equal(FilterFactory.getSincM2Channelizer(25000.0, m, 9), g["taps"])
    got = ComplexPolyphaseChannelizerM2(g["taps"], 2400000, m).receive(g["x"])
    assert got.shape == g["results_f64"].shape
    assert sg.rel_rms(got, g["results_f64"]) < TOL and sg.rel_rms(got, g["results_f32"]) < TOL
    ch = ComplexPolyphaseChannelizerM2(g["taps"], 2400000, m)
    ch.setOutputChannels([([4], 700), ([88, 89], -300)], g["synth"])
    rows = ch.receiveChannels(g["x"])
    assert sg.rel_rms(rows[0], g["bin4_offset700"]) < TOL
    assert sg.rel_rms(rows[1], g["bins88_89_offset_m300"]) < TOL


@pytest.mark.gpu
def test_cuda_filters_fm_and_chains_match_golden(gpu):
    from sdrtrunk_b200.dsp import (Bank, ByteSampleConverter, ComplexFeedForwardGainControl, ComplexFIRFilter2,
                                   DecimationFilterFactory, FMDemodulator, Signed16BitSampleConverter,
                                   SignedByteSampleConverter, SquelchingFMDemodulator)
    g = load("filters.npz")
    assert np.array_equal(DecimationFilterFactory.getComplexDecimationFilter(8).decimateComplex(g["x"]), g["decimate8"])
    assert np.array_equal(ComplexFIRFilter2(g["fir"]).filter(g["x"]), g["fir72"])
    agc = ComplexFeedForwardGainControl(32)
    assert np.array_equal(np.concatenate([agc.filter(g["x"][:2048]), agc.filter(g["x"][2048:])]), g["agc"])
    f = load("fm.npz")
    assert np.max(np.abs(FMDemodulator(1.0).demodulate(f["x"]) - f["fm"])) < 1e-6
    got = SquelchingFMDemodulator(0.01, -40.0, 4).demodulate(f["x"])
    assert np.array_equal(got == 0.0, f["squelch_fm"] == 0.0) and np.max(np.abs(got - f["squelch_fm"])) < 1e-6
    p = load("p25_chains.npz")
    presets = {"c4fm": gpu.PRESET_P25_C4FM, "lsm": gpu.PRESET_P25_LSM, "hdqpsk": gpu.PRESET_P25_HDQPSK, "dmr": gpu.PRESET_DMR}
    for kind, preset in presets.items():
        taps = p[kind + "_fir"] if kind + "_fir" in p.files else None
        bank = Bank.preset(preset, 1, 50000.0, taps, max_samples_per_call=4096)
        dibits, agc_out = bank.process(p[kind + "_x"].reshape(1, -1), want_filtered=True)
        assert np.array_equal(dibits[0], p[kind + "_dibits"]), kind
        assert np.array_equal(agc_out[0], p[kind + "_agc"]), kind
    c = load("converters.npz")
    assert np.array_equal(ByteSampleConverter().convertSamples(c["raw8"].tobytes()), c["u8"])
    assert np.array_equal(SignedByteSampleConverter().convertSamples(c["raw8"].tobytes()), c["s8"])
    assert np.array_equal(Signed16BitSampleConverter().convertSamples(c["raw16"].tobytes()), c["s16le"])
