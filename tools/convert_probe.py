#!/usr/bin/env python3
"""Channelizer step on device-resident 8-bit tuner samples (M = 800, 20 M samples): conversion fused into the filter
bank's loads (default) against convert_kernel + float filter bank (SDRGPU_FUSE_CONVERT=0), and the float-input step.
usage (GPU box): python tools/convert_probe.py"""
import ctypes as C
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
    name = sys.argv[1] if len(sys.argv) > 1 else "channelizer_20m"
    native.init(0)
    L = native.lib()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    _, timed = bench.make_timed(torch, dist, dev, 1)
    inputs = bench.TunerInputs(torch, dev, name, 0, 1)
    w = bench.TunerWorkload(name, inputs, 1, 0)
    ms, _ = timed(w.step_device, w.stream, 10, 3)
    print("%s float32 input: %.4f ms/step" % (name, ms / 10))
    x8 = torch.clamp(torch.round(inputs.x[0] * 128.0), -128, 127).to(torch.int8)
    w.chans[0].setSampleFormat("s8")

    def step():
        native.check(L.sdrgpu_chan_process(w.chans[0]._h, C.c_void_p(x8.data_ptr()), w.n_floats, native.DEVICE,
                                           C.c_void_p(w.out_dev.data_ptr()), 2 * w.n_blocks, native.DEVICE,
                                           native.LAYOUT_CHANNELS, None))

    ms, _ = timed(step, w.stream, 10, 3)
    print("%s signed 8-bit input (SDRGPU_FUSE_CONVERT=%s): %.4f ms/step" % (name, os.environ.get("SDRGPU_FUSE_CONVERT", "1"), ms / 10))


if __name__ == "__main__":
    main()
