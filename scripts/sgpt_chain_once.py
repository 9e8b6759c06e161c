"""One SparseGPT chain of the down_proj shape (C = 11008, R = 4096): Hessian, factorisation, OBS sweep - for ncu captures.
python scripts/sgpt_chain_once.py [C R]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vlm-compression_b200"))
import torch
from vlmc import native
native.load()
torch.manual_seed(0)
C, R = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) >= 3 else (11008, 4096)
x = (torch.randn(2 * C, C, device="cuda") * (torch.rand(C, device="cuda") + 0.5)).half()
H = torch.zeros(C, C, device="cuda")
native.hessian_accum(x, H, 0, 1)
del x
damp, dead = native.hessian_prepare(H, 0.01)
os.environ["VLMC_CHOL_LOOKAHEAD"] = "0"
U, status = native.chol_inv_upper(H)
W = (torch.randn(R, C, device="cuda") * 0.02).half()
native.obs_sweep(W, U, 0.5, dead=dead)
torch.cuda.synchronize()
print("status", status.item(), float(W.float().abs().mean()))
