"""One chol_inv_upper + one obs_sweep per size (for ncu launch lists) and CUDA-event timings."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vlm-compression_b200"))
import torch
from vlmc import native

native.load()
torch.manual_seed(0)
sizes = [int(a) for a in sys.argv[1:]] or [4096, 11008]
for C in sizes:
    x = (torch.randn(4 * C if C <= 4096 else 2 * C, C, device="cuda") * (torch.rand(C, device="cuda") + 0.5)).half()
    H = torch.zeros(C, C, device="cuda")
    native.hessian_accum(x, H, 0, 1)
    del x
    damp, dead = native.hessian_prepare(H, 0.01)
    U, status = native.chol_inv_upper(H)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    import time
    for la in ("0", "1", "0", "1"):
        os.environ["VLMC_CHOL_LOOKAHEAD"] = la
        U, status = native.chol_inv_upper(H, U)
        torch.cuda.synchronize()
        a.record()
        h0 = time.perf_counter()
        U, status = native.chol_inv_upper(H, U)
        h1 = time.perf_counter()
        b.record()
        torch.cuda.synchronize()
        print(f"C={C} lookahead={la}: chol_inv_upper {a.elapsed_time(b):.2f} ms (host issue {1e3 * (h1 - h0):.2f} ms) "
              f"status {status.item()}", flush=True)
    R = 4096
    W = (torch.randn(R, C, device="cuda") * 0.02).half()
    native.obs_sweep(W.clone(), U, 0.5, dead=dead)
    torch.cuda.synchronize()
    W2 = W.clone()
    a.record()
    h0 = time.perf_counter()
    native.obs_sweep(W2, U, 0.5, dead=dead)
    h1 = time.perf_counter()
    b.record()
    torch.cuda.synchronize()
    print(f"C={C} R={R}: obs_sweep {a.elapsed_time(b):.2f} ms (host issue {1e3 * (h1 - h0):.2f} ms)", flush=True)
