#!/usr/bin/env python3
"""C4FM chain on 400 channels with every channel frequency-corrected (the usual case in sdrtrunk: the requested channel
frequency is rarely the centre of its polyphase bin) against the uncorrected selection: device-resident single pass and
host buffers (chunked pipeline).  usage (GPU box): python tools/offset_chain_time.py"""
import ctypes as C
import os
import sys
import time

import numpy as np
import scipy.signal as ss

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from sdrtrunk_b200 import native  # noqa: E402
from sdrtrunk_b200.dsp import Bank, ComplexPolyphaseChannelizerM2, Pipeline  # noqa: E402


def oracle_taps():
    """the C4FM decoder's own baseband filter (libsdrgpu's restatement of RemezFIRFilterDesigner)"""
    from sdrtrunk_b200.dsp import FilterFactory, FIRFilterSpecification
    return FilterFactory.getTaps(FIRFilterSpecification.lowPassBuilder().sampleRate(50000.0).passBandCutoff(5100)
                                 .passBandRipple(0.01).stopBandStart(6500).stopBandRipple(0.01).build())


def main():
    native.init(0)
    L = native.lib()
    m, fs = 400, 10e6
    n = 48 * 1024 * (m // 2)
    fir = oracle_taps()
    x_dev = torch.randn(2 * n, device="cuda", dtype=torch.float32) * 0.01
    x_host = torch.empty(2 * n, dtype=torch.float32, pin_memory=True)
    x_host.copy_(x_dev)
    stride = n // (m // 2) // 8 + 64
    sym_dev = torch.zeros((m, stride), dtype=torch.uint8, device="cuda")
    cnt_dev = torch.zeros(m, dtype=torch.int32, device="cuda")
    sym_host = torch.zeros((m, stride), dtype=torch.uint8, pin_memory=True)
    cnt_host = torch.zeros(m, dtype=torch.int32, pin_memory=True)
    for label, offsets in (("uncorrected", None), ("all 400 channels corrected", [([k], 1250 if k % 2 else -3000) for k in range(m)])):
        chan = ComplexPolyphaseChannelizerM2(fs, 9, maxInputFloats=2 * n)
        if offsets:
            chan.setOutputChannels(offsets)
        bank = Bank.preset(native.PRESET_P25_C4FM, m, 2 * fs / m, fir, max_samples_per_call=n // (m // 2))
        pipe = Pipeline(chan, bank)

        def dev():
            native.check(L.sdrgpu_pipeline_process(pipe._h, C.c_void_p(x_dev.data_ptr()), 2 * n, native.DEVICE,
                                                   C.c_void_p(sym_dev.data_ptr()), stride, None, 0, C.c_void_p(cnt_dev.data_ptr()),
                                                   native.DEVICE))

        def host():
            native.check(L.sdrgpu_pipeline_process(pipe._h, C.c_void_p(x_host.data_ptr()), 2 * n, native.HOST,
                                                   C.c_void_p(sym_host.data_ptr()), stride, None, 0, C.c_void_p(cnt_host.data_ptr()),
                                                   native.HOST))

        out = []
        for fn in (dev, host):
            for _ in range(3):
                fn()
            bank.sync()
            torch.cuda.synchronize()
            t = time.time()
            for _ in range(10):
                fn()
            bank.sync()
            torch.cuda.synchronize()
            out.append((time.time() - t) / 10 * 1e3)
        print("%-28s device-resident %.3f ms, host buffers %.3f ms per call" % (label, *out), flush=True)
        pipe.dispose()
        bank.dispose()
        chan.dispose()


if __name__ == "__main__":
    main()
