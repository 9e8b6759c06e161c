"""One launch of each per-row select kernel at 4096 x 4096 fp16 (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vlm-compression_b200"))
import torch
from vlmc import native
native.load()
torch.manual_seed(0)
R, C = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, int(sys.argv[2]) if len(sys.argv) > 2 else 4096
W0 = (torch.randn(R, C, device="cuda") * 0.02).half()
s = torch.exp(torch.rand(C, device="cuda") * 4 - 2) * 50
keep = torch.empty(R, C, dtype=torch.bool, device="cuda")
for rep in range(2):
    W = W0.clone()
    native.wanda_rowselect(W, s, C // 2, keep_mask=keep)
torch.cuda.synchronize()
