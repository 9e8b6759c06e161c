import os, sys
sys.path.insert(0, "vlm-compression_b200")
import torch
from vlmc import native
native.load()
dev="cuda"; C=4096; T=128*2048
x = torch.empty(T, C, device=dev, dtype=torch.float16)
for j in range(0, T, 16384): x[j:j+16384] = torch.randn(16384, C, device=dev).half()
H = torch.zeros(C, C, device=dev)
tag = "single" if os.environ.get("VLMC_HESS_2CTA") == "0" else "pair"
for kc in (256, 512, 1024, 2048, 8192):
    native.hessian_accum(x, H, 0, 1, kc=kc); torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3): native.hessian_accum(x, H, 0, 1, kc=kc)
    b.record(); torch.cuda.synchronize()
    print(tag, "kc", kc, f"{a.elapsed_time(b)/3:.3f} ms", flush=True)
