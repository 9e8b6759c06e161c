"""K1 / K2 tile-width sweep: CUDA-event timings of vlmc_sqnorm_accum / vlmc_dsnot_stats at the Vicuna shapes for
each VLMC_STATS_CX (column lanes per CTA).  python scripts/stats_probe.py"""
import os
import sys

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
    st = [torch.zeros(C, device=dev) for _ in range(4)]
    for cx in (32, 64, 128, 256):
        os.environ["VLMC_STATS_CX"] = str(cx)
        for name, fn in (("sqnorm", lambda: native.sqnorm_accum(x, st[0], 0, 128)),
                         ("dsnot", lambda: native.dsnot_stats(x, st[0], st[1], st[2], st[3], 0, 1, 0, nseg=128))):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10):
                fn()
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 10
            print(f"C={C} cx={cx:3d} {name:6s}: {ms * 1e3:8.1f} us  {T * C * 2 / ms / 1e6:7.1f} GB/s", flush=True)
    del x
