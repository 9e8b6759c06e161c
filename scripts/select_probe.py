"""Wanda selection kernels alone at Vicuna sizes (for ncu and CUDA-event timing): python scripts/select_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vlm-compression_b200"))
import torch
from vlmc import native

native.load()
torch.manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for R, C in [(4096, 4096), (11008, 4096), (4096, 11008)]:
    W0 = (torch.randn(R, C, device="cuda") * 0.02).half()
    s = torch.exp(torch.rand(C, device="cuda") * 4 - 2) * 50
    for name, fn in [("nm 2:4", lambda W: native.wanda_nm(W, s, 2, 4)),
                     ("rowselect 50%", lambda W: native.wanda_rowselect(W, s, C // 2))]:
        ts = []
        for rep in range(5):
            W = W0.clone()
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn(W)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        t = sorted(ts)[len(ts) // 2]
        print(f"{name:14s} {R}x{C}: {t * 1e3:.1f} us, {R * C * 5 / t / 1e6:.0f} GB/s", flush=True)

# K7: whole-matrix threshold (ViT unstructured path, wanda_pruner.py:682-683) at the EVA ViT-g shapes and one Vicuna shape
for R, C in [(4224, 1408), (6144, 1408), (1408, 6144), (4096, 4096)]:
    W0 = (torch.randn(R, C, device="cuda") * 0.02).half()
    s = torch.exp(torch.rand(C, device="cuda") * 4 - 2) * 50
    ts = []
    for rep in range(5):
        W = W0.clone()
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        native.wanda_threshold(W, s, int(R * C * 0.5))
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    t = sorted(ts)[len(ts) // 2]
    print(f"threshold 50%  {R}x{C}: {t * 1e3:.1f} us, {R * C * 5 / t / 1e6:.0f} GB/s", flush=True)
