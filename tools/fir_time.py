#!/usr/bin/env python3
"""fir_agc kernel time of a C4FM bank (72 taps + block AGC) by channel count; SDRGPU_FIR_SPLIT=0 selects the general
kernel.  usage (GPU box): python tools/fir_time.py [channels ...]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from sdrtrunk_b200 import native  # noqa: E402
from sdrtrunk_b200.dsp import Bank  # noqa: E402
from psk_layout_sweep import oracle_taps  # noqa: E402


def main():
    native.init(0)
    rng = np.random.default_rng(0)
    n = 12 * 1024
    fir = oracle_taps()
    if os.environ.get("FIR_TILES_PER_CTA"):   # launch shape: consecutive buffers per CTA (sdrgpu_set_tuning)
        native.check(native.lib().sdrgpu_set_tuning(native.TUNE_FIR_TILES_PER_CTA, int(os.environ["FIR_TILES_PER_CTA"])))
    for c in [int(a) for a in sys.argv[1:]] or [800, 6400]:
        x = rng.standard_normal((c, 2 * n), dtype=np.float32) * 0.1
        bank = Bank.preset(native.PRESET_P25_C4FM, c, 50000.0, fir, max_samples_per_call=n)
        bank.enableTiming(True)
        best = 1e9
        for _ in range(4):
            bank.process(x)
            best = min(best, bank.lastKernelMs()[0])
        bank.dispose()
        print("%5d channels x %d samples: filter stage %.4f ms (SDRGPU_FIR_SPLIT=%s, FIR_TILES_PER_CTA=%s)" %
              (c, n, best, os.environ.get("SDRGPU_FIR_SPLIT", "1"), os.environ.get("FIR_TILES_PER_CTA", "auto")),
              flush=True)


if __name__ == "__main__":
    main()
