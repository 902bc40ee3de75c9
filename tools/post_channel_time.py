import sys, time
import os; ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'tests'))
import numpy as np, torch, ctypes as C
from sdrtrunk_b200 import native
from sdrtrunk_b200.dsp import ComplexPolyphaseChannelizerM2
native.init(0)
m=400; fs=10e6
n=10_000_000
ch=ComplexPolyphaseChannelizerM2(fs, 9, maxInputFloats=2*n)
x=torch.randn(2*n, device='cuda', dtype=torch.float32)*0.01
nb=n//(m//2)
out=torch.empty((m, 2*nb), device='cuda', dtype=torch.float32)
L=native.lib()
def run():
    native.check(L.sdrgpu_chan_process(ch._h, C.c_void_p(x.data_ptr()), 2*n, native.DEVICE, C.c_void_p(out.data_ptr()), 2*nb, native.DEVICE, native.LAYOUT_CHANNELS, None))
for label, chans in (("identity", None), ("400 offsets", [([k], 1250 if k%2 else -3000) for k in range(m)]), ("40 offsets", [([k], 1250 if k%10==0 else 0) for k in range(m)])):
    if chans is not None: ch.setOutputChannels(chans)
    for _ in range(3): run()
    torch.cuda.synchronize(); t=time.time()
    for _ in range(10): run()
    ch.sync(); torch.cuda.synchronize()
    print(label, (time.time()-t)/10*1e3, "ms per 10M-sample call")
