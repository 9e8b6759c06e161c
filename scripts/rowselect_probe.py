"""K5 per-row select: CTA-per-row kernel (default) against the warp-per-row one (VLMC_ROWSELECT_LEGACY=1).
GPU time only: 6 launches on 6 weight copies (working set > L2) captured into one CUDA graph and replayed."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vlm-compression_b200"))
import torch
from vlmc import native

native.load()
torch.manual_seed(0)
NCOPY = 6
for R, C, dt in ((4096, 4096, torch.float16), (11008, 4096, torch.float16), (4096, 11008, torch.float16),
                 (2048, 2048, torch.bfloat16), (5120, 2048, torch.bfloat16), (2048, 5120, torch.bfloat16), (4096, 4096, torch.float32)):
    W0 = (torch.randn(R, C, device="cuda") * 0.02).to(dt)
    s = torch.exp(torch.rand(C, device="cuda") * 4 - 2) * 50
    res = {}
    for legacy in ("1", "0"):
        os.environ["VLMC_ROWSELECT_LEGACY"] = legacy
        Ws = [W0.clone() for _ in range(NCOPY)]
        keeps = [torch.empty(R, C, dtype=torch.bool, device="cuda") for _ in range(NCOPY)]
        means = [torch.empty(1, device="cuda") for _ in range(NCOPY)]
        native.wanda_rowselect(W0.clone(), s, C // 2, keep_mask=keeps[0], score_mean=means[0])     # warm: workspace, attributes
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for W, kp, m in zip(Ws, keeps, means):
                native.wanda_rowselect(W, s, C // 2, keep_mask=kp, score_mean=m)
        ts = []
        for rep in range(3):
            for W in Ws:
                W.copy_(W0)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            g.replay()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) / NCOPY)
        res[legacy] = (min(ts), Ws[0].clone(), keeps[0].clone(), means[0].item())
        del g
    same = torch.equal(res["0"][1], res["1"][1]) and torch.equal(res["0"][2], res["1"][2])
    nbytes = R * C * (2 * W0.element_size() + 1)
    print(f"R={R} C={C} {str(dt)[6:]}: legacy {res['1'][0] * 1e3:.1f} us ({nbytes / res['1'][0] / 1e6:.0f} GB/s), "
          f"cta {res['0'][0] * 1e3:.1f} us ({nbytes / res['0'][0] / 1e6:.0f} GB/s), identical {same}, "
          f"mean rel diff {abs(res['0'][3] - res['1'][3]) / abs(res['1'][3]):.1e}", flush=True)
