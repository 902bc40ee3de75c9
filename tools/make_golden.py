#!/usr/bin/env python3
"""Mints the golden vectors under tests/golden/ from the CPU oracle (the reference ships none for this path and no
JVM is available, SURVEY.md section 8c: parity stays "unpinned"; these fixtures pin the ORACLE and the CUDA path to
each other and guard both against regressions).  Seeded, small, deterministic:  python tools/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
import siggen as sg  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    rng = np.random.default_rng(20240101)
    # ---- channelizer, M = 96: tones + noise, 40 blocks
    m = 96
    taps = oracle.sinc_m2_channelizer(25000.0, m, 9)
    n = 48 * 40
    z = sg.awgn(rng, n, 1e-3) + sg.tone(2.4e6, 4 * 25000.0 + 1500.0, n, 0.3) + sg.tone(2.4e6, -7 * 25000.0 - 800.0, n, 0.2)
    x = sg.interleave(z)
    res32 = oracle.Channelizer(taps, m).receive(x, mode="f32")
    res64 = oracle.Channelizer(taps, m).receive(x, mode="f64")
    raw = oracle.Channelizer(taps, m).receive(x, mode="raw")
    one = oracle.OneChannelOutputProcessor(50000.0, 4, float(m))
    one.set_frequency_offset(700)
    synth = oracle.sinc_m2_synthesizer(50000.0, 25000.0, 2, 9)
    two = oracle.TwoChannelOutputProcessor(50000.0, 88, 89, synth, float(m))
    two.set_frequency_offset(-300)
    np.savez_compressed(os.path.join(OUT, "channelizer_m96.npz"), taps=taps, x=x, results_f32=res32, results_f64=res64,
                        accumulators=raw, bin4_offset700=one.process(res32), bins88_89_offset_m300=two.process(res32),
                        synth=synth)
    # ---- filters: half-band cascade x8, 72-tap FIR, AGC
    xf = (0.3 * rng.standard_normal(2 * 2048)).astype(np.float32)
    fir = oracle.c4fm_baseband_taps()
    np.savez_compressed(os.path.join(OUT, "filters.npz"), x=xf, fir=fir, decimate8=oracle.Decimator(8).decimate_complex(xf),
                        fir72=oracle.ComplexFIR(fir).filter(xf), agc=np.concatenate([oracle.agc_block(xf[:2048]),
                                                                                     oracle.agc_block(xf[2048:])]))
    # ---- FM
    xm = sg.interleave(sg.nbfm(25000.0, 2048, audio_hz=700.0) + sg.awgn(rng, 2048, 1e-3))
    np.savez_compressed(os.path.join(OUT, "fm.npz"), x=xm, fm=oracle.FMDemodulator(1.0).demodulate(xm),
                        squelch_fm=oracle.SquelchingFMDemodulator(0.01, -40.0, 4).demodulate(xm))
    # ---- P25 chains: 4 buffers each
    out = {}
    for kind, okind, rate in (("c4fm", oracle.C4FM, 4800.0), ("lsm", oracle.LSM, 4800.0), ("hdqpsk", oracle.HDQPSK, 6000.0),
                              ("dmr", oracle.DMR, 4800.0)):
        nn = 4 * 1024
        dib = rng.integers(0, 4, int(nn * rate / 50000) + 8)
        if kind in ("c4fm", "dmr"):
            zz = sg.c4fm(dib, carrier_offset=120.0, timing_phase=0.37, n_samples=nn)
            t = fir
        else:
            zz = sg.dqpsk(dib, symbol_rate=rate, carrier_offset=-80.0, timing_phase=0.61, n_samples=nn)
            t = None if kind == "lsm" else oracle.hdqpsk_baseband_taps()
        xx = sg.interleave(zz + sg.awgn(rng, nn, 0.03))
        d, agc = oracle.P25Chain(okind, 50000.0, t).receive(xx, want_agc=True)
        out[kind + "_x"] = xx
        out[kind + "_dibits"] = d
        out[kind + "_agc"] = agc
        if t is not None:
            out[kind + "_fir"] = t
    np.savez_compressed(os.path.join(OUT, "p25_chains.npz"), **out)
    # ---- sample converters
    raw8 = rng.integers(0, 256, 512, dtype=np.uint8)
    raw16 = rng.integers(-32768, 32768, 256).astype("<i2")
    np.savez_compressed(os.path.join(OUT, "converters.npz"), raw8=raw8, raw16=raw16, u8=oracle.convert_samples(raw8.tobytes(), "u8"),
                        s8=oracle.convert_samples(raw8.tobytes(), "s8"), s16le=oracle.convert_samples(raw16.tobytes(), "s16le"))
    # ---- Airspy native buffers and the on-device sync detector (own seed: the vectors above stay as they are)
    rng2 = np.random.default_rng(20240102)
    real = sg.airspy_real_signal(rng2, 6000, [(1.1e6, 0.3), (-2.7e6, 0.15)], dc=0.02)
    raw_u = sg.airspy_raw(real)
    raw_p = sg.airspy_raw(real, packed=True)
    conv = oracle.AirspySampleConverter()
    iq = np.concatenate([conv.convert(raw_u[:2 * 2500]), conv.convert(raw_u[2 * 2500:])])
    dib = sg.dibits_with_sync(rng2, 1200, sg.P25_PHASE1_SYNC, 48)
    nn = 12 * 1024
    xs = sg.interleave(sg.c4fm(dib, carrier_offset=-1250.0, timing_phase=0.4, n_samples=nn, amplitude=0.5) + sg.awgn(rng2, nn, 0.01))
    chain = oracle.P25Chain(oracle.C4FM, 50000.0, fir)
    chain.attach_sync(oracle.SYNC_P25_PHASE1, 50000.0)
    np.savez_compressed(os.path.join(OUT, "airspy_sync.npz"), raw_unpacked=raw_u, raw_packed=raw_p, iq=iq, sync_x=xs, sync_fir=fir,
                        sync_symbols=chain.receive(xs))
    # ---- the decoders' Remez-designed baseband filters (oracle/orc_remez.c; csrc/remez.cpp must reproduce them bit for bit)
    np.savez_compressed(os.path.join(OUT, "remez.npz"), c4fm=oracle.c4fm_baseband_taps(), hdqpsk=oracle.hdqpsk_baseband_taps(),
                        nbfm=oracle.nbfm_iq_taps())
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
