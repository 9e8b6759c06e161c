"""OBS sweep timings per super-block size (VLMC_OBS_SUPERBLOCK) at the Vicuna shapes, unstructured 50 % and 2:4."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vlm-compression_b200"))
import torch
from vlmc import native

native.load()
torch.manual_seed(0)
for R, C in ((4096, 4096), (11008, 4096), (4096, 11008)):
    x = (torch.randn(2 * C, C, device="cuda") * (torch.rand(C, device="cuda") + 0.5)).half()
    H = torch.zeros(C, C, device="cuda")
    native.hessian_accum(x, H, 0, 1)
    del x
    damp, dead = native.hessian_prepare(H, 0.01)
    U, status = native.chol_inv_upper(H)
    W = (torch.randn(R, C, device="cuda") * 0.02).half()
    ref = None
    for sb, la in (("1", "0"), ("4", "0"), ("4", "1"), ("8", "0"), ("8", "1")):
        os.environ["VLMC_OBS_SUPERBLOCK"] = sb
        os.environ["VLMC_CHOL_LOOKAHEAD"] = la
        for nm in ((0, 0), (2, 4)):
            W2 = W.clone()
            native.obs_sweep(W2, U, 0.5, *nm, dead=dead)
            torch.cuda.synchronize()
            W2 = W.clone()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            native.obs_sweep(W2, U, 0.5, *nm, dead=dead)
            b.record()
            torch.cuda.synchronize()
            if sb == "1":
                ref = ref or {}
                ref[nm] = W2
                extra = ""
            else:
                r = ref[nm].float()
                extra = (f" vs sb=1: mask agreement {float(((W2 == 0) == (ref[nm] == 0)).float().mean()):.6f}, rel Frobenius "
                         f"{float((W2.float() - r).norm() / r.norm()):.2e}")
            print(f"R={R} C={C} sb={sb} lookahead={la} {'2:4' if nm[0] else 'unstructured'}: obs_sweep {a.elapsed_time(b):.2f} ms{extra}", flush=True)
