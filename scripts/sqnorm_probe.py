"""vlmc_sqnorm_accum (single tensor) at the bench shapes for several VLMC_STATS_WAVES.  python scripts/sqnorm_probe.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vlm-compression_b200"))
import torch
from vlmc import native
native.load()
dev = "cuda"
T = 128 * 2048
for C in (4096, 11008):
    x = torch.empty(128, 2048, C, device=dev, dtype=torch.float16)
    for j in range(128):
        x[j] = torch.randn(2048, C, device=dev).half()
    x2 = torch.empty_like(x); x2.copy_(x)          # two buffers alternate: 2 x 2.1 GB >> L2
    ref = None
    for waves in ("1", "4", "16", "32", "1", "32"):
        os.environ["VLMC_STATS_WAVES"] = waves
        s = torch.zeros(C, device=dev)
        native.sqnorm_accum(x, s, 0, 128)
        if ref is None:
            ref = s.clone()
        err = float(((s - ref).abs() / ref.abs()).max())
        for _ in range(2):
            native.sqnorm_accum(x2, s, 128, 128)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(10):
            native.sqnorm_accum(x if i % 2 == 0 else x2, s, 128, 128)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        print(f"C={C} waves={waves}: {ms * 1e3:8.1f} us  {T * C * 2 / ms / 1e6:7.1f} GB/s  max rel diff vs one wave {err:.1e}", flush=True)
    del x, x2
