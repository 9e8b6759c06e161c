import sys; sys.path.insert(0,"vlm-compression_b200")
import torch
from vlmc import native
native.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for R, C in [(4096,4096),(11008,4096),(4096,11008)]:
    ts=[]
    for seed in range(8):
        g=torch.Generator(device="cuda").manual_seed(100+seed)
        W=(torch.randn(R,C,device="cuda",generator=g)*0.02).half()
        s=torch.exp(torch.rand(C,device="cuda",generator=g)*2.77-1.386)**2*2048+0.1
        flush.zero_()
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); native.wanda_rowselect(W,s,round(C*0.6)); b.record(); torch.cuda.synchronize()
        ts.append(round(a.elapsed_time(b)*1e3))
    print(R,C,"k=0.6C us per seed:",ts, flush=True)
