"""One DSnoT refine at 4096 x 4096 fp16, 60 % (for ncu / timing of the walk and apply kernels)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vlm-compression_b200"))
import torch
from vlmc import native
native.load()
torch.manual_seed(0)
R, C = 4096, int(sys.argv[1]) if len(sys.argv) > 1 else 4096
W0 = (torch.randn(R, C, device="cuda") * 0.02).half()
scal = torch.exp(torch.rand(C, device="cuda") * 4 - 2) * 50
summ = torch.randn(C, device="cuda") * 20
var = torch.exp(torch.rand(C, device="cuda") * 3 - 2)
for rep in range(4):
    os.environ["VLMC_DSNOT_WALK_V1"] = "1" if rep < 2 else "0"
    W = W0.clone()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    keep, ncyc = native.dsnot_refine(W, scal, summ, var, round(C * 0.6))
    b.record()
    torch.cuda.synchronize()
    fb = int(native.workspace(W, 4)[:4].view(torch.int32).item())
    print(f"dsnot_refine R={R} C={C} walk_v1={os.environ['VLMC_DSNOT_WALK_V1']}: {a.elapsed_time(b):.3f} ms, cycles {int(ncyc.item())}, "
          f"rows handed back {fb if rep >= 2 else '-'}", flush=True)
