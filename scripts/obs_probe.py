"""OBS sweep alone (U from one factorisation) at the Vicuna shapes: CUDA-event timings.  python scripts/obs_probe.py [C R]..."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vlm-compression_b200"))
import torch
from vlmc import native

native.load()
torch.manual_seed(0)
args = [int(a) for a in sys.argv[1:]] or [4096, 4096, 4096, 11008, 11008, 4096]
for C, R in zip(args[0::2], args[1::2]):
    x = (torch.randn(2 * C, C, device="cuda") * (torch.rand(C, device="cuda") + 0.5)).half()
    H = torch.zeros(C, C, device="cuda")
    native.hessian_accum(x, H, 0, 1)
    del x
    damp, dead = native.hessian_prepare(H, 0.01)
    U, status = native.chol_inv_upper(H)
    W = (torch.randn(R, C, device="cuda") * 0.02).half()
    native.obs_sweep(W.clone(), U, 0.5, dead=dead)
    torch.cuda.synchronize()
    for mode, kw in (("unstructured", dict(sparsity=0.5)), ("2:4", dict(sparsity=0.0, prune_n=2, prune_m=4))):
        W2 = W.clone()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        native.obs_sweep(W2, U, dead=dead, **kw)
        b.record()
        torch.cuda.synchronize()
        print(f"C={C} R={R} {mode}: obs_sweep {a.elapsed_time(b):.2f} ms", flush=True)
    del H, U, W, W2
