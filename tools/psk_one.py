#!/usr/bin/env python3
"""One demodulator layout on one bank size, a few calls, kernel time printed: the command ncu captures psk kernels from.
usage (GPU box): python tools/psk_one.py [channels=6400] [lanes=4] [samples=24576] [preset=c4fm|hdqpsk] [calls=4]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import siggen as sg  # noqa: E402
from psk_layout_sweep import oracle_taps  # noqa: E402
from sdrtrunk_b200 import native  # noqa: E402
from sdrtrunk_b200.dsp import Bank  # noqa: E402


def main():
    c = int(sys.argv[1]) if len(sys.argv) > 1 else 6400
    lanes = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 24 * 1024
    preset = sys.argv[4] if len(sys.argv) > 4 else "c4fm"
    calls = int(sys.argv[5]) if len(sys.argv) > 5 else 4
    sync = int(sys.argv[6]) if len(sys.argv) > 6 else 0
    native.init(0)
    if len(sys.argv) > 7:
        native.check(native.lib().sdrgpu_set_tuning(0, int(sys.argv[7])))
        native.check(native.lib().sdrgpu_set_tuning(2, 1))
    rng = np.random.default_rng(0)
    base = []
    for k in range(8):
        dib = rng.integers(0, 4, int(n * 4800 / 50000) + 8)
        z = sg.c4fm(dib, carrier_offset=rng.uniform(-200, 200), timing_phase=rng.uniform(0, 1), n_samples=n)
        base.append(sg.interleave(z + sg.awgn(rng, n, 0.03)))
    x = np.tile(np.stack(base), (c // 8, 1))
    bank = Bank.preset(native.PRESET_P25_C4FM if preset == "c4fm" else native.PRESET_P25_HDQPSK, c, 50000.0, oracle_taps(),
                       max_samples_per_call=n)
    bank.setDemodulatorLanes(lanes)
    if sync:
        bank.setSyncDetector(sync)
    bank.enableTiming(True)
    for i in range(calls):
        bank.process(x)
        f, d = bank.lastKernelMs()
        print("call %d: %d channels x %d samples, lanes %d: filters %.3f ms, demodulator %.3f ms" % (i, c, n, lanes, f, d), flush=True)
    bank.dispose()


if __name__ == "__main__":
    main()
