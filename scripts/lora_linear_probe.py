"""K23 (vlmc_sparselora_linear_forward) against K15 + the library GEMM at the SparseLoRA shapes.  python scripts/lora_linear_probe.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vlm-compression_b200"))
import torch
from vlmc import native

dev = "cuda"
dt = torch.bfloat16
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=10):
    fn(); fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / n


shapes = [("q/k/v/o", 4096, 4096), ("gate/up", 11008, 4096), ("down", 4096, 11008), ("vit qkv", 4224, 1408), ("vit fc1", 6144, 1408),
          ("vit fc2", 1408, 6144)]
only = sys.argv[1:] and sys.argv[1]
for name, R, C in shapes:
    W = (torch.randn(R, C, device=dev) * 0.02).to(dt)
    A = torch.randn(8, C, device=dev) * 0.1
    B = torch.randn(R, 8, device=dev) * 0.1
    mask = torch.rand(R, C, device=dev) < 0.5
    for T in (512, 1024, 2048, 4096, 8192, 16384):
        x = (torch.randn(T, C, device=dev) * 0.5).to(dt)
        y = torch.empty(T, R, device=dev, dtype=dt)
        weff = torch.empty_like(W)

        def fused():
            native.sparselora_linear_forward(x, W, A, B, 2.0, mask, True, out=y)

        def k15():
            native.sparselora_effective_weight(W, A, B, 2.0, mask, True, out=weff)

        def lib():
            torch.nn.functional.linear(x, weff, None)

        k15()
        tf, tk, tl = timeit(fused), timeit(k15), timeit(lib)
        err = float((y.float() - torch.nn.functional.linear(x, weff).float()).abs().max() / y.float().abs().max())
        fl = 2.0 * T * R * C
        print(f"{name:8s} R={R:5d} C={C:5d} T={T:5d}: fused {tf * 1e3:7.1f} us ({fl / tf / 1e9:6.0f} TF/s)   K15 {tk * 1e3:6.1f} + GEMM {tl * 1e3:7.1f} us "
              f"({fl / tl / 1e9:6.0f} TF/s) = {(tk + tl) * 1e3:7.1f} us   fused / (K15 + GEMM) = {tf / (tk + tl):.2f}   rel diff {err:.1e}", flush=True)
