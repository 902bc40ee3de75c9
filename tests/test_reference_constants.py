"""Pins the data and constants the reference embeds in its own sources (the only "golden vectors" it holds for this
path) against what the oracle and the product carry.  Reads /root/reference, so it only runs in the build container
(skipped elsewhere, e.g. on the GPU box)."""
import os
import re

import numpy as np
import pytest

import oracle

REF = "/root/reference/src/main/java/io/github/dsheirer"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference sources not present")


def java(path):
    return open(os.path.join(REF, path)).read()


def test_mmse_interpolator_table_is_the_references():
    rows = []
    for line in java("dsp/filter/interpolator/Interpolator.java").splitlines():
        vals = re.findall(r"[-+]?\d\.\d+e[-+]\d+f", line)
        if len(vals) == 8:
            rows.append([np.float32(v[:-1]) for v in vals])
    ref = np.array(rows, np.float32)
    assert ref.shape == (129, 8)
    text = open(os.path.join(ROOT, "include", "sdr_mmse_taps.h")).read()
    mine = np.array([np.float32(v[:-1]) for v in re.findall(r"[-+]?\d\.\d+e[-+]\d+f", text)], np.float32).reshape(129, 8)
    assert np.array_equal(mine, ref)


def test_window_and_envelope_constants():
    w = java("dsp/filter/Window.java")
    a = [float(re.search(r"double a%d = ([0-9.]+);" % i, w).group(1)) for i in range(3)]
    n = 23
    x = np.arange(n)
    want = a[0] - a[1] * np.cos(2 * np.pi * x / (n - 1)) + a[2] * np.cos(4 * np.pi * x / (n - 1))
    assert np.allclose(oracle.window("blackman", n), want, rtol=0, atol=1e-15)
    c = java("sample/complex/Complex.java")
    assert "(0.4f * quadratureAbsolute)" in c and "1.9999f - magnitudeSquared()" in c
    assert "MINIMUM_ENVELOPE = 0.0001f" in java("dsp/gain/ComplexFeedForwardGainControl.java")
    y = oracle.agc_block(np.array([3e-5, -1e-5] * 1024, np.float32))          # envelope below the minimum: gain 1 / 1e-4
    assert np.array_equal(y[:2], np.array([3e-5, -1e-5], np.float32) * (np.float32(1.0) / np.float32(1e-4)))


def test_decoder_front_end_constants():
    assert "SAMPLE_COUNTER_GAIN = 0.3f" in java("module/decode/p25/phase1/P25P1DecoderC4FM.java")
    assert "SAMPLE_COUNTER_GAIN = 0.3f" in java("module/decode/p25/phase1/P25P1DecoderLSM.java")
    assert "SAMPLE_COUNTER_GAIN = 0.4f" in java("module/decode/dmr/DMRDecoder.java")
    p2 = java("module/decode/p25/phase2/P25P2DecoderHDQPSK.java")
    assert "SYMBOL_TIMING_GAIN = 0.1f" in p2 and "super(6000.0)" in p2 and "setPLLBandwidth(PLLBandwidth.BW_300)" in p2
    assert "setPLLBandwidth(PLLBandwidth.BW_300)" in java("module/decode/p25/phase1/P25P1DecoderC4FM.java")
    assert "setPLLBandwidth(PLLBandwidth.BW_200)" in java("module/decode/p25/phase1/P25P1DecoderLSM.java")
    assert "setPLLBandwidth(PLLBandwidth.BW_300)" in java("module/decode/dmr/DMRDecoder.java")
    bw = dict(re.findall(r"(BW_\d+)\((\d+\.\d+),", java("dsp/psk/pll/PLLBandwidth.java")))
    assert bw == {"BW_400": "400.0", "BW_300": "300.0", "BW_250": "250.0", "BW_200": "200.0"}
    assert "MAXIMUM_DEVIATION_SAMPLES_PER_SYMBOL = 0.02f" in java("dsp/psk/InterpolatingSampleBuffer.java")
    dib = dict((m[0], int(m[1])) for m in re.findall(r"(D\d\d_\w+)\((?:true|false), (?:true|false), (\d),", java("dsp/symbol/Dibit.java")))
    assert dib == {"D01_PLUS_3": 1, "D00_PLUS_1": 0, "D10_MINUS_1": 2, "D11_MINUS_3": 3}
    # the presets of the C ABI carry the same numbers
    from sdrtrunk_b200 import native
    cfg = native.BankConfig()
    for preset, rate, bw_, gain in ((native.PRESET_P25_C4FM, 4800.0, 300.0, 0.3), (native.PRESET_P25_LSM, 4800.0, 200.0, 0.3),
                                    (native.PRESET_P25_HDQPSK, 6000.0, 300.0, 0.1), (native.PRESET_DMR, 4800.0, 300.0, 0.4)):
        native.check(native.lib().sdrgpu_bank_config_preset(cfg, preset, 1, 50000.0, None, 0, 1024))
        assert (cfg.symbol_rate, cfg.pll_bandwidth) == (rate, bw_) and abs(cfg.sample_counter_gain - gain) < 1e-7
        assert cfg.block_size == 1024 and cfg.agc == 1
    assert "PROCESSED_BUFFER_SAMPLE_SIZE = 2048" in java("dsp/filter/channelizer/PolyphaseChannelSource.java")


def test_half_band_stage_plan_is_the_references():
    """ComplexDecimateX{N}Filter stage lengths / windows"""
    plan = {2: [(63, "HAMMING")], 4: [(23, "BLACKMAN"), (63, "HAMMING")], 8: [(15, "BLACKMAN"), (23, "BLACKMAN"), (63, "HAMMING")]}
    for rate, stages in plan.items():
        src = java("dsp/filter/decimate/ComplexDecimateX%dFilter.java" % rate)
        found = re.findall(r"DECIMATE_BY_\d+_FILTER_LENGTH = (\d+)", src)
        wins = re.findall(r"WINDOW_TYPE = Window\.WindowType\.(\w+)", src)
        assert [int(v) for v in found] == [stages[0][0]], (rate, found)
        assert wins == [stages[0][1]], (rate, wins)


def test_sync_patterns_and_detector_constants():
    """FrameSync patterns, match threshold, delay lengths and PLL corrections of the P25 sync detectors, as carried by
    the oracle (orc_sync.c) and the product (SyncTraits in bank.cu)."""
    fs = dict((k, int(v, 16)) for k, v in re.findall(r"(P25_PHASE\d_\w+)\(\s*0x([0-9A-Fa-f]+)l\s*\)", java("dsp/symbol/FrameSync.java")))
    assert len(fs) == 8
    for src in (open(os.path.join(ROOT, "oracle", "orc_sync.c")).read(),
                open(os.path.join(ROOT, "sdrtrunk_b200", "csrc", "bank.cu")).read()):
        mine = set(int(v, 16) for v in re.findall(r"0x([0-9A-Fa-f]{10,12})ull", src))
        assert set(fs.values()) <= mine
    p1 = java("module/decode/p25/phase1/P25P1SyncDetector.java")
    p2 = java("module/decode/p25/phase2/P25P2SyncDetector.java")
    assert "SYNC_MATCH_THRESHOLD = 4" in p1 and "SYNC_MATCH_THRESHOLD = 4" in p2
    assert "DEFAULT_SYMBOL_RATE = 4800" in p1 and "DEFAULT_SYMBOL_RATE = 6000" in p2
    assert "getMessageLength(), 48)" in p1 and "MultiSyncPatternMatcher(syncDetectListener, 1440, 40)" in p2
    assert "LOGICAL_LINK_DATA_UNIT_1(5, 1568," in java("module/decode/p25/phase1/P25P1DataUnitID.java")
    dud = java("module/decode/p25/phase1/P25P1DataUnitDetector.java")
    assert "DATA_UNIT_DIBIT_LENGTH = 57" in dud and "SYNC_DIBIT_LENGTH = 24" in dud
    assert "new DibitDelayBuffer(160)" in java("module/decode/p25/phase2/P25P2SuperFrameDetector.java")
    assert oracle.SyncDetector(oracle.SYNC_P25_PHASE1, 50000.0).delay == 57 - 24
    assert oracle.SyncDetector(oracle.SYNC_P25_PHASE2, 50000.0).delay == 160
    # the rotated patterns are the normal one with every dibit's constellation point turned by 90 / 180 degrees:
    # +1 (00) -> +3 (01) -> -3 (11) -> -1 (10) -> +1 going counter-clockwise
    ccw = {0: 1, 1: 3, 3: 2, 2: 0}
    for phase, bits in (("PHASE1", 48), ("PHASE2", 40)):
        d = [(fs["P25_%s_NORMAL" % phase] >> (bits - 2 - 2 * k)) & 3 for k in range(bits // 2)]
        def rotate(seq, quarter_turns):
            for _ in range(quarter_turns):
                seq = [ccw[v] for v in seq]
            return seq
        def pack(seq):
            v = 0
            for x in seq:
                v = (v << 2) | x
            return v
        assert pack(rotate(d, 1)) == fs["P25_%s_ERROR_90_CCW" % phase]
        assert pack(rotate(d, 2)) == fs["P25_%s_ERROR_180" % phase]
        assert pack(rotate(d, 3)) == fs["P25_%s_ERROR_90_CW" % phase]
