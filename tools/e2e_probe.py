#!/usr/bin/env python3
"""Host-buffer (e2e) step of the headline workload against the number of time chunks, beside the bare H2D copy time of
the same pinned buffers: how much of the e2e step is PCIe.  usage (GPU box): python tools/e2e_probe.py [tuners]"""
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
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    _, timed = bench.make_timed(torch, dist, dev, 1)
    inputs = bench.TunerInputs(torch, dev, "c4fm_20m", 0, tuners)
    w = bench.TunerWorkload("c4fm_20m", inputs, tuners, 0)
    w.set_format("s8")
    bufs, _ = w.host_inputs("s8")
    dst = [torch.empty_like(b, device=dev) for b in bufs]
    s = torch.cuda.Stream(device=dev)

    def h2d():
        with torch.cuda.stream(s):
            for d, b in zip(dst, bufs):
                d.copy_(b, non_blocking=True)

    ms, wall = timed(h2d, s, 5, 2)
    nbytes = sum(b.numel() for b in bufs)
    print("bare H2D of %d x %.1f MB pinned int8: %.3f ms/step (wall %.3f) = %.1f GB/s" %
          (tuners, bufs[0].numel() / 1e6, ms / 5, wall / 5, nbytes / (ms / 5 * 1e-3) / 1e9), flush=True)
    for dchunks in (4,):
        for chunks in [int(a) for a in sys.argv[2:]] or [4, 8]:
            w.pipeline.setChunks(chunks)
            ms, wall = timed(w.step_host, w.stream, 5, 3)
            per = max(ms, wall) / 5
            print("host chunks %2d: e2e %.3f ms/step (events %.3f, wall %.3f) = %.2f GS/s" %
                  (chunks, per, ms / 5, wall / 5, w.total_complex / (per * 1e-3) / 1e9), flush=True)
    for chunks in [int(a) for a in sys.argv[2:]] or [4, 8]:
        w.pipeline.setChunks(chunks)
        ms, wall = timed(w.step_host_stream, w.stream, 6, 3)
        w.stream_drain()
        per = max(ms, wall) / 6
        print("host chunks %2d, two calls in flight (submit / wait): e2e %.3f ms/step (events %.3f, wall %.3f) = %.2f GS/s" %
              (chunks, per, ms / 6, wall / 6, w.total_complex / (per * 1e-3) / 1e9), flush=True)
    w.pipeline.setChunks(8)
    w.set_format("f32")
    ms, _ = timed(w.step_device, w.stream, 5, 3)
    print("device resident: %.3f ms/step" % (ms / 5))


if __name__ == "__main__":
    main()
