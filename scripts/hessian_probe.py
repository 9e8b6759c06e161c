"""GPU probe for the tcgen05 Hessian kernel: correctness on small cases with diagnostics, then timing."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vlm-compression_b200"))
import torch
from vlmc import native

def check(T, C, dtype, kc=0, b=1):
    torch.manual_seed(T + C)
    x = (torch.randn(T, C, device="cuda") * (torch.rand(C, device="cuda") * 2 + 0.25) + 0.1).to(dtype)
    H = torch.zeros(C, C, device="cuda")
    native.hessian_accum(x, H, 0, b, kc=kc)
    torch.cuda.synchronize()
    ref = (x.double().T @ x.double()) * (2.0 / b)
    err = (H.double() - ref).abs().max().item() / ref.abs().max().item()
    sym = torch.equal(H, H.T)
    print(f"T={T} C={C} {dtype} kc={kc}: rel_inf={err:.3e} symmetric={sym}", flush=True)
    if err > 1e-4:
        d = (H.double() - ref).abs()
        bad = (d > 1e-4 * ref.abs().max()).nonzero()
        print("  bad count", bad.shape[0], "first", bad[:8].tolist(), flush=True)
        print("  H[0,:8]", H[0, :8].tolist()); print("  ref[0,:8]", ref[0, :8].tolist())
        blk = (d.view(C // 64, 64, C // 64, 64).amax((1, 3)) > 1e-4 * ref.abs().max()).int()
        print("  bad 64x64 blocks:\n", blk[:16, :16], flush=True)
    return err

if __name__ == "__main__":
    errs = []
    for T, C, dt in [(64, 256, torch.float16), (128, 256, torch.bfloat16), (2048, 512, torch.float16),
                     (2048, 4096, torch.float16), (300, 1408, torch.float16)]:
        errs.append(check(T, C, dt))
    if max(errs) > 1e-4:
        print("CORRECTNESS FAILED; skipping timing"); sys.exit(1)
    # second call (running average) 
    x = torch.randn(2048, 1024, device="cuda").half(); H = torch.zeros(1024, 1024, device="cuda")
    native.hessian_accum(x, H, 0, 1); native.hessian_accum(x, H, 1, 1); torch.cuda.synchronize()
    ref = (x.double().T @ x.double()) * 2.0
    print("running avg err", ((H.double() - ref).abs().max() / ref.abs().max()).item())
    for C in (4096, 11008):
        for nseq in (1, 16, 128):
            T = nseq * 2048
            x = torch.randn(T, C, device="cuda").half()
            H = torch.zeros(C, C, device="cuda")
            for kc in (2048, 8192):
                native.hessian_accum(x, H, 0, nseq, kc=kc); torch.cuda.synchronize()
                t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
                t0.record()
                for _ in range(3): native.hessian_accum(x, H, 0, nseq, kc=kc)
                t1.record(); torch.cuda.synchronize()
                ms = t0.elapsed_time(t1) / 3
                flop = 2.0 * T * C * C
                print(f"C={C} T={T} kc={kc}: {ms:.3f} ms  logical {flop / ms / 1e9:.1f} TFLOP/s  executed(SYRK) {flop / 2 / ms / 1e9:.1f}", flush=True)
            if C == 4096 and nseq == 16:
                t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
                xf = x.float(); torch.cuda.synchronize(); t0.record(); Hr = xf.T @ xf; t1.record(); torch.cuda.synchronize()
                print(f"  torch fp32 matmul same shape: {t0.elapsed_time(t1):.3f} ms")
            del x, H
