"""K3 CTA-pair (default) vs single-CTA (VLMC_HESS_2CTA=0) at the Vicuna shapes: run once per variant (the switch is read once per
process); prints timings and a checksum so the two runs can be compared.  python scripts/hessian_pair_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vlm-compression_b200"))
import torch
from vlmc import native

native.load()
dev = "cuda"
tag = "single" if os.environ.get("VLMC_HESS_2CTA") == "0" else "pair"
for C, T in ((512, 4096), (1408, 2048), (4096, 128 * 2048), (11008, 128 * 2048)):
    g = torch.Generator(device=dev).manual_seed(C)
    x = torch.empty(T, C, device=dev, dtype=torch.float16)
    for j in range(0, T, 16384):
        n = min(16384, T - j)
        x[j:j + n] = (torch.randn(n, C, device=dev, generator=g) * (torch.rand(C, device=dev, generator=g) + 0.5)).half()
    H = torch.zeros(C, C, device=dev)
    native.hessian_accum(x, H, 0, 1)
    torch.cuda.synchronize()
    if C <= 1408:
        ref = (2.0 * x.double().T @ x.double()).float()
        err = float((H - ref).abs().max() / ref.abs().max())
    else:
        err = float("nan")
    sym = float((H - H.T).abs().max())
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        native.hessian_accum(x, H, 0, 1)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    print(f"{tag} C={C} T={T}: {ms:.3f} ms, {2.0 * T * C * C / ms / 1e9 / 2:.0f} TF/s executed (SYRK half), "
          f"err vs fp64 {err:.2e}, asym {sym:.1e}, checksum {float(H.double().sum()):.6e} diag {float(H.diagonal().double().sum()):.6e}", flush=True)
    del x, H
