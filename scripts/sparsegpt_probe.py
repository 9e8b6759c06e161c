"""GPU probe: sgemm, chol_inv_upper and obs_sweep step by step against torch / the numpy oracle."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vlm-compression_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from vlmc import native
from oracle import oracle
import golden_util as gu

def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(n): fn()
    t1.record(); torch.cuda.synchronize()
    return t0.elapsed_time(t1) / n

torch.manual_seed(0)
for C, T in [(128, 512), (256, 1024), (200, 1024), (1408, 8192), (4096, 16384)]:
    x = (torch.randn(T, C, device="cuda") * (torch.rand(C, device="cuda") * 2 + 0.5)).half()
    H = torch.zeros(C, C, device="cuda"); native.hessian_accum(x, H, 0, 1)
    U, status = native.chol_inv_upper(H)
    torch.cuda.synchronize()
    Hd = H.double()
    Uref = torch.linalg.cholesky(torch.linalg.inv(Hd), upper=True)
    err = ((U.double() - Uref).abs().max() / Uref.abs().max()).item()
    recon = ((U.double().T @ U.double()) @ Hd - torch.eye(C, device="cuda", dtype=torch.float64)).abs().max().item()
    low = U.tril(-1).abs().max().item()
    print(f"chol C={C}: status={status.item()} rel err vs fp64 {err:.3e}  |U^T U H - I|max {recon:.3e}  lower-part max {low:.1e}", flush=True)
# not PD -> status 1
H = torch.randn(256, 64, device="cuda"); H = H @ H.T
print("rank-deficient status:", native.chol_inv_upper(H.contiguous())[1].item())
for C in (4096, 11008):
    x = torch.randn(2 * C, C, device="cuda").half()
    H = torch.zeros(C, C, device="cuda"); native.hessian_accum(x, H, 0, 1)
    U = torch.empty(C, C, device="cuda")
    ms = timeit(lambda: native.chol_inv_upper(H, U), n=2)
    print(f"chol_inv_upper C={C}: {ms:.2f} ms ({2/3*C**3/ms/1e9:.1f} TFLOP/s useful)", flush=True)
    Hc = H.clone()
    def ref():
        L = torch.linalg.cholesky(Hc); Hi = torch.cholesky_inverse(L); return torch.linalg.cholesky(Hi, upper=True)
    print(f"   torch (cuSOLVER) 3-step chain: {timeit(ref, n=1):.2f} ms", flush=True)

# obs sweep vs oracle on the golden cases + a bigger random one
g = gu.load("sparsegpt.npz")
for name in g["cases"]:
    sp, n, m = g[f"{name}|cfg"]; tag = str(g[f"{name}|tag"])
    H = torch.from_numpy(g[f"{name}|H"]).cuda()
    W = gu.to_torch(g[f"{name}|W_before"], tag, "cuda")
    damp, dead = native.hessian_prepare(H, 0.01)
    steps = 0
    while True:
        U, status = native.chol_inv_upper(H)
        if status.item() == 0: break
        native.hessian_add_damp(H, damp); steps += 1
    keep, score = native.obs_sweep(W, U, sp, int(n), int(m), dead=dead, want_mask=True)
    torch.cuda.synchronize()
    got = W.float().cpu().numpy(); ref = g[f"{name}|W_after"]
    fro = np.linalg.norm(got - ref) / np.linalg.norm(ref)
    agree = ((got == 0) == (ref == 0)).mean()
    # same U fed to the oracle: isolates the sweep from the factorisation
    Wo, so, _ = oracle.sparsegpt_fasterprune(g[f"{name}|W_before"], tag, None, sp, int(n), int(m), U=U.cpu().numpy(), dead=dead.cpu().numpy().astype(bool))
    fro2 = np.linalg.norm(got - Wo) / np.linalg.norm(Wo); agree2 = ((got == 0) == (Wo == 0)).mean()
    print(f"obs {name}: damp steps {steps} vs reference fro {fro:.2e} mask {agree:.5f} | vs oracle(same U) fro {fro2:.2e} mask {agree2:.5f} score {score.item():.6f} / {float(g[name + '|importance_score']):.6f} keepmask ok {np.array_equal(keep.cpu().numpy(), got != 0)}", flush=True)

for (R, C) in [(4096, 4096), (4096, 11008), (11008, 4096)]:
    x = torch.randn(4 * C, C, device="cuda").half()
    H = torch.zeros(C, C, device="cuda"); native.hessian_accum(x, H, 0, 1)
    U, _ = native.chol_inv_upper(H)
    W0 = (torch.randn(R, C, device="cuda") * 0.02).half()
    W = W0.clone()
    ms = timeit(lambda: native.obs_sweep(W.copy_(W0), U, 0.5), n=1)
    sp = (W == 0).float().mean().item()
    ms24 = timeit(lambda: native.obs_sweep(W.copy_(W0), U, 0.0, 2, 4), n=1)
    print(f"obs_sweep R={R} C={C}: unstructured {ms:.1f} ms (sparsity {sp:.4f}), 2:4 {ms24:.1f} ms", flush=True)
