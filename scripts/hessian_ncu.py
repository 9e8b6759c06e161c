"""One vlmc_hessian_accum launch at the bench shape (T = 128 x 2048 fp16 tokens) after a warm-up one, for ncu --set full."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vlm-compression_b200"))
import torch
from vlmc import native
native.load()
C = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = 128 * 2048
g = torch.Generator(device="cuda").manual_seed(C)
x = torch.empty(T, C, device="cuda", dtype=torch.float16)
for j in range(0, T, 16384):
    x[j:j + 16384] = (torch.randn(16384, C, device="cuda", generator=g) * (torch.rand(C, device="cuda", generator=g) + 0.5)).half()
H = torch.zeros(C, C, device="cuda")
for _ in range(2):
    native.hessian_accum(x.view(128, 2048, C), H, 0, 128)
torch.cuda.synchronize()
print("done", C)
