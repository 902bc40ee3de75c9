#!/usr/bin/env python3
"""One traced device-resident step of the headline workload (SDRGPU_TRACE=1 prints the chunk timeline to stderr).
usage (GPU box): SDRGPU_TRACE=1 python tools/trace_probe.py [tuners] [device_chunks]"""
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
    chunks = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    native.init(0)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    _, timed = bench.make_timed(torch, dist, dev, 1)
    inputs = bench.TunerInputs(torch, dev, "c4fm_20m", 0, tuners)
    w = bench.TunerWorkload("c4fm_20m", inputs, tuners, 0)
    w.pipeline.setDeviceChunks(chunks)
    ms, _ = timed(w.step_device, w.stream, 3, 2)
    print("device resident, %d chunks: %.3f ms/step" % (chunks, ms / 3))


if __name__ == "__main__":
    main()
