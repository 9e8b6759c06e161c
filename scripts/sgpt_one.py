"""One SparseGPT linear end to end (Hessian -> factor -> OBS sweep) for profiling: python scripts/sgpt_one.py R C [T]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vlm-compression_b200"))
import torch
from vlmc import native

R, C = int(sys.argv[1]), int(sys.argv[2])
T = int(sys.argv[3]) if len(sys.argv) > 3 else 4 * C
torch.manual_seed(0)
x = torch.randn(T, C, device="cuda").half()
H = torch.zeros(C, C, device="cuda")
native.hessian_accum(x, H, 0, 1)
W = (torch.randn(R, C, device="cuda") * 0.02).half()
damp, dead = native.hessian_prepare(H, 0.01)
U, status = native.chol_inv_upper(H)
native.obs_sweep(W, U, 0.5, dead=dead)
torch.cuda.synchronize()
print("status", status.item(), "sparsity", (W == 0).float().mean().item())
