"""K14 / K15 over a Vicuna block: per-linear launches and the one-launch batch, GB/s of the 5 B / weight stream."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vlm-compression_b200"))
import torch
from vlmc import native
native.load()
D, FF = 4096, 11008
LIN = [(D, D)] * 4 + [(FF, D)] * 2 + [(D, FF)]
g = torch.Generator(device="cuda").manual_seed(11)
for r in (8, 4, 2):
    items = []
    for R, C in LIN:
        W = (torch.randn(R, C, device="cuda", generator=g) * 0.02).half()
        A = torch.randn(r, C, device="cuda", generator=g) * 0.1
        B = torch.randn(R, r, device="cuda", generator=g) * 0.1
        M = torch.rand(R, C, device="cuda", generator=g) < 0.5
        items.append((W, A, B, M, torch.empty_like(W)))
    nbytes = sum(W.numel() for W, *_ in items) * 5

    def timed(fn, reps=5):
        """GPU time: the launches of fn captured into a CUDA graph (no host gaps), replayed reps times."""
        fn(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        g.replay(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps): g.replay()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps
    t1 = timed(lambda: [native.sparselora_merge(W, A, B, 2.0, M, remask=True) for W, A, B, M, _ in items])
    t2 = timed(lambda: native.sparselora_merge_batch([i[0] for i in items], [i[1] for i in items], [i[2] for i in items],
                                                     [2.0] * len(items), [i[3] for i in items], remask=True))
    t3 = timed(lambda: [native.sparselora_effective_weight(W, A, B, 2.0, M, True, out=o) for W, A, B, M, o in items])
    print(f"rank {r}: merge per linear {t1:.3f} ms ({nbytes / t1 / 1e6:.0f} GB/s), merge one launch {t2:.3f} ms ({nbytes / t2 / 1e6:.0f} GB/s), "
          f"effective weight per linear {t3:.3f} ms ({nbytes / t3 / 1e6:.0f} GB/s)", flush=True)
    del items
