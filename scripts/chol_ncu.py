"""Two chol_inv_upper calls at one size (the second one is the one to read in an ncu launch list)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vlm-compression_b200"))
import torch
from vlmc import native
native.load()
torch.manual_seed(0)
C = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
x = (torch.randn(2 * C, C, device="cuda") * (torch.rand(C, device="cuda") + 0.5)).half()
H = torch.zeros(C, C, device="cuda")
native.hessian_accum(x, H, 0, 1)
del x
native.hessian_prepare(H, 0.01)
os.environ["VLMC_CHOL_LOOKAHEAD"] = "0"
U, status = native.chol_inv_upper(H)
torch.cuda.synchronize()
print("status", status.item())
