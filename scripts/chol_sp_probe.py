"""vlmc_chol_inv_upper with the two-level (super-panel) trailing update: VLMC_CHOL_SUPERPANEL = 1 / 2 / 4 / 8.  python scripts/chol_sp_probe.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vlm-compression_b200"))
import torch
from vlmc import native
native.load()
torch.manual_seed(0)
for C in [int(a) for a in sys.argv[1:]] or [4096, 11008]:
    x = (torch.randn(2 * C, C, device="cuda") * (torch.rand(C, device="cuda") + 0.5)).half()
    H = torch.zeros(C, C, device="cuda")
    native.hessian_accum(x, H, 0, 1)
    del x
    native.hessian_prepare(H, 0.01)
    ref = None
    U = None
    for sp in ("1", "4", "2", "8", "1", "4"):
        os.environ["VLMC_CHOL_SUPERPANEL"] = sp
        U, status = native.chol_inv_upper(H, U)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            U, status = native.chol_inv_upper(H, U)
        b.record()
        torch.cuda.synchronize()
        if ref is None:
            ref = U.clone()
        err = float((U - ref).abs().max() / ref.abs().max())
        print(f"C={C} superpanel={sp}: chol_inv_upper {a.elapsed_time(b) / 3:.2f} ms  status {status.item()}  max diff vs plain loop {err:.1e}", flush=True)
    del ref, U, H
