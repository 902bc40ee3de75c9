#!/usr/bin/env python3
"""Soak run on the GPU box: many calls of every entry point with ragged sizes, format / layout / detector switches and
handle churn; checks that device memory does not grow and that results stay equal to a fresh handle's.
usage: python tools/soak.py [seconds]"""
import os
import sys
import time

import numpy as np
import scipy.signal as ss

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import siggen as sg  # noqa: E402
import torch  # noqa: E402
from sdrtrunk_b200 import native  # noqa: E402
from sdrtrunk_b200.dsp import AirspySampleConverter, Bank, ComplexPolyphaseChannelizerM2, FilterFactory, Pipeline  # noqa: E402


def main(seconds):
    native.init(0)
    rng = np.random.default_rng(0)
    m, fs = 96, 2.4e6
    taps = FilterFactory.getSincM2Channelizer(25000.0, m, 9)
    fir = oracle_taps()
    n_ch = 8 * 1024
    bins = [3, 17, 40, 60]
    base = [sg.c4fm(sg.dibits_with_sync(rng, n_ch // 10 + 8, sg.P25_PHASE1_SYNC, 48), carrier_offset=off, n_samples=n_ch,
                    amplitude=0.05) for off in (0.0, 1150.0, -120.0, 2300.0)]
    x = sg.interleave(sg.multiplex(base, bins, m, n_ch) + sg.awgn(rng, n_ch * m // 2, 1e-3))
    raw = sg.airspy_raw(rng.uniform(-0.5, 0.5, 1 << 16), packed=True)

    def build():
        chan = ComplexPolyphaseChannelizerM2(taps, int(fs), m, maxInputFloats=x.size)
        chan.setChannels(bins)
        bank = Bank.preset(native.PRESET_P25_C4FM, len(bins), 50000.0, fir, max_samples_per_call=n_ch)
        bank.setSyncDetector(native.SYNC_P25_PHASE1)
        return chan, bank, Pipeline(chan, bank)

    chan, bank, pipe = build()
    want = pipe.process(x)
    del pipe, bank, chan
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    t0, calls, rebuilds = time.time(), 0, 0
    air = AirspySampleConverter(maxSamples=1 << 17)
    air.setSamplePacking(True)
    while time.time() - t0 < seconds:
        chan, bank, pipe = build()
        rebuilds += 1
        pipe.setChunks(int(rng.integers(1, 9)))
        # ragged cuts of the same stream must give the same dibits
        cuts = np.sort(rng.choice(np.arange(2, x.size // 2, 2), size=int(rng.integers(0, 6)), replace=False)) * 2
        parts, pos = [], 0
        for c in list(cuts) + [x.size]:
            parts.append(pipe.process(x[pos:c]))
            pos = c
            calls += 1
        got = [np.concatenate(p) for p in zip(*parts)]
        for a, b in zip(got, want):
            assert np.array_equal(a, b)
        air.convert(raw[:3 * int(rng.integers(1, raw.size // 3))])
        calls += 1
        pipe.dispose()
        bank.dispose()
        chan.dispose()
    torch.cuda.synchronize()
    free1 = torch.cuda.mem_get_info()[0]
    print("soak: %d calls, %d handle sets created and destroyed in %.0f s; device memory delta %.1f MB; airspy fallbacks %d" %
          (calls, rebuilds, time.time() - t0, (free0 - free1) / 1e6, air.mismatches()))
    assert free0 - free1 < 64e6


if __name__ == "__main__":
    main(float(sys.argv[1]) if len(sys.argv) > 1 else 30.0)
