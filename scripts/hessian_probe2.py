"""kc / slab sweep: accuracy (vs fp64) and throughput of the tcgen05 Hessian kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vlm-compression_b200"))
import torch
from vlmc import native

def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(n): fn()
    t1.record(); torch.cuda.synchronize()
    return t0.elapsed_time(t1) / n

torch.manual_seed(0)
C, T = 1024, 16 * 2048
x = (torch.randn(T, C, device="cuda") * (torch.rand(C, device="cuda") * 3.75 + 0.25) + torch.randn(C, device="cuda") * 0.3).half()
ref = (x.double().T @ x.double()) * (2.0 / 16)
big = ref.abs() > 1e-3 * ref.abs().max()
for kc in (64, 128, 256, 512, 1024, 2048, 4096):
    H = torch.zeros(C, C, device="cuda")
    native.hessian_accum(x, H, 0, 16, kc=kc)
    d = (H.double() - ref).abs()
    print(f"kc={kc}: rel_inf={d.max().item() / ref.abs().max().item():.3e} rel_elem={(d[big] / ref.abs()[big]).max().item():.3e} "
          f"mean_signed={((H.double() - ref)[big] / ref[big]).mean().item():.3e}", flush=True)
for C in (4096, 11008):
    x = torch.randn(128 * 2048, C, device="cuda").half()
    H = torch.zeros(C, C, device="cuda")
    flop = 2.0 * x.shape[0] * C * C
    for kc in (128, 256, 512, 1024):
        for slab in (16384, 32768, 65536, 1 << 30):
            ms = timeit(lambda: native.hessian_accum(x, H, 0, 128, kc=kc, slab_tokens=slab))
            print(f"C={C} T={x.shape[0]} kc={kc} slab={slab}: {ms:.3f} ms executed {flop / 2 / ms / 1e9:.1f} TFLOP/s", flush=True)
    x1 = x[:2048]
    for kc in (256, 512):
        ms = timeit(lambda: native.hessian_accum(x1, H, 1, 1, kc=kc), n=10)
        print(f"C={C} per-sample call (T=2048, RMW of H) kc={kc}: {ms:.3f} ms executed {2.0 * 2048 * C * C / 2 / ms / 1e9:.1f} TFLOP/s", flush=True)
    del x, H
