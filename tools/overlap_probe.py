#!/usr/bin/env python3
"""Headline step time (8 x 20 MS/s -> 6400 C4FM channels, device resident) against the knobs that decide how the filter
kernels of time chunk i+1 share the GPU with the demodulator of chunk i: time chunks per call and resident CTAs per SM of
fir_agc_kernel / pfb2_kernel (sdrgpu_set_tuning).  usage (GPU box): python tools/overlap_probe.py [tuners]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    from sdrtrunk_b200 import native
    tuners = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    native.init(0)
    L = native.lib()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    _, timed = bench.make_timed(torch, dist, dev, 1)
    inputs = bench.TunerInputs(torch, dev, "c4fm_20m", 0, tuners)
    w = bench.TunerWorkload("c4fm_20m", inputs, tuners, 0)
    grid = [(1, 0, 0), (2, 0, 0), (3, 0, 0), (4, 0, 0), (6, 0, 0), (8, 0, 0), (12, 0, 0), (16, 0, 0), (4, 0, 0), (8, 0, 0)]
    for chunks in ():
        for fir in (1, 2, 3, 4):
            grid.append((chunks, fir, 0))
        for fir, pfb in ((1, 1), (2, 1), (3, 1)):
            grid.append((chunks, fir, pfb))
    for chunks, fir, pfb in grid:
        native.check(L.sdrgpu_set_tuning(0, fir))
        native.check(L.sdrgpu_set_tuning(1, pfb))
        w.pipeline.setDeviceChunks(chunks)
        ms, _ = timed(w.step_device, w.stream, 6, 3)
        print("chunks %2d  fir_ctas_per_sm %d  pfb_ctas_per_sm %d : %.3f ms/step  %.2f GS/s" %
              (chunks, fir, pfb, ms / 6, w.total_complex / (ms / 6 * 1e-3) / 1e9), flush=True)


if __name__ == "__main__":
    main()
