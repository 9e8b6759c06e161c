"""K3 at the benched shapes: token-slab sweep at C = 11008 / 4096, T = 262,144 fp16 tokens."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vlm-compression_b200"))
import torch
from vlmc import native

native.load()
dev = "cuda"
T = 128 * 2048
for C in (11008, 4096):
    g = torch.Generator(device=dev).manual_seed(C)
    x = torch.empty(T, C, device=dev, dtype=torch.float16)
    for j in range(0, T, 16384):
        x[j:j + 16384] = (torch.randn(16384, C, device=dev, generator=g) * (torch.rand(C, device=dev, generator=g) + 0.5)).half()
    H = torch.zeros(C, C, device=dev)
    nt = (C + 255) // 256
    executed = 2.0 * T * (nt * (nt + 1) // 2) * 256.0 * 256.0
    for slab in (0, 16384, 32768, 65536, 131072, 262144):
        native.hessian_accum(x.view(128, 2048, C), H, 0, 128, slab_tokens=slab)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            native.hessian_accum(x.view(128, 2048, C), H, 0, 128, slab_tokens=slab)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 3
        print(f"C={C} T={T} slab_tokens={slab or 'default'}: {ms:.3f} ms, executed {executed / ms / 1e9:.0f} TFLOP/s, "
              f"logical {2.0 * T * C * C / ms / 1e9:.0f} TFLOP/s", flush=True)
    del x, H
