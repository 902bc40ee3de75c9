import sys, os
sys.path.insert(0, "."); sys.path.insert(0, "tests"); sys.path.insert(0, "tools")
import bench, torch, torch.distributed as dist
from sdrtrunk_b200 import native
native.init(0)
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
_, timed = bench.make_timed(torch, dist, dev, 1)
inputs = bench.TunerInputs(torch, dev, "c4fm_20m", 0, 8)
w = bench.TunerWorkload("c4fm_20m", inputs, 8, 0)
for ch in w.chans:
    ch.setOutputChannels([([k], 37 if k % 2 else -53) for k in range(w.m)])
ms, _ = timed(w.step_device, w.stream, 4, 3)
print("corrected: %.3f ms/step" % (ms / 4))
