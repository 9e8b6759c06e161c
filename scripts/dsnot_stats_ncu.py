"""vlmc_dsnot_stats at the bench shape (128 calls x 2048 fp16 tokens), timed, for ncu.  python scripts/dsnot_stats_ncu.py [C]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vlm-compression_b200"))
import torch
from vlmc import native
native.load()
C = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N, S = 128, 2048
g = torch.Generator(device="cuda").manual_seed(C)
x = torch.empty(N, S, C, device="cuda", dtype=torch.float16)
for j in range(0, N, 8):
    x[j:j + 8] = (torch.randn(8, S, C, device="cuda", generator=g) * (torch.rand(C, device="cuda", generator=g) + 0.5) + 0.3).half()
st = [torch.zeros(C, device="cuda") for _ in range(4)]
for rep in range(3):
    for t in st:
        t.zero_()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    native.dsnot_stats(x, st[0], st[1], st[2], st[3], 0, 1, 0, nseg=N)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    print(f"dsnot_stats C={C}: {ms:.3f} ms = {x.numel() * 2 / ms / 1e6:.0f} GB/s", flush=True)
