"""SURVEY 8f-4 kernels alone (K18-K22) at the size of one Vicuna-7B block (202 M fp32 scores in 7 tensors, 0.81 GB) and of
four blocks (3.2 GB >> L2), CUDA-event timing: python scripts/global_select_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vlm-compression_b200"))
import torch
from vlmc import native
from vlmc.compression.pruners import layer_sparsity as ls

native.load()
torch.manual_seed(0)
shapes = [(4096, 4096)] * 4 + [(11008, 4096)] * 2 + [(4096, 11008)]


def timed(fn, reps=5):
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


for blocks in ((1,) if os.environ.get("VLMC_PROBE_ONE") else (1, 4)):
    params = [(torch.randn(s, device="cuda") * 0.02).half() for _ in range(blocks) for s in shapes]
    grads = [(torch.randn_like(p.float()) * 1e-3).half() for p in params]
    acc = [torch.zeros(p.shape, device="cuda") for p in params]
    scores = [torch.empty_like(a) for a in acc]
    n = sum(p.numel() for p in params)
    seg0 = [0] * len(params)
    segs = list(range(len(params)))
    t = timed(lambda: native.importance_accum(acc, grads, "obd"))
    print(f"[{blocks} block(s), {n / 1e6:.0f} M scores] importance_accum: {t:.3f} ms, {n * 10 / t / 1e6:.0f} GB/s (10 B/elem)")
    t = timed(lambda: native.importance_finalize(acc, params, scores, "obd", 5))
    print(f"  importance_finalize: {t:.3f} ms, {n * 10 / t / 1e6:.0f} GB/s (10 B/elem)")
    t = timed(lambda: native.scores_sum(scores))
    print(f"  scores_sum: {t:.3f} ms, {n * 4 / t / 1e6:.0f} GB/s")
    k = torch.tensor([n // 2])
    t = timed(lambda: native.scores_kth(scores, seg0, [n // 2]))
    print(f"  scores_kth global (3 passes): {t:.3f} ms, {n * 12 / t / 1e6:.0f} GB/s (12 B/elem)")
    t = timed(lambda: native.scores_kth(scores, segs, [s.numel() // 5 * 4 + 1 for s in scores]))
    print(f"  scores_kth per tensor (3 passes): {t:.3f} ms, {n * 12 / t / 1e6:.0f} GB/s")
    thr = native.scores_kth(scores, seg0, [n // 2])
    masks = [torch.empty_like(s) for s in scores]
    t = timed(lambda: native.scores_mask(scores, seg0, thr, outs=masks))
    print(f"  scores_mask (fp32 mask out): {t:.3f} ms, {n * 8 / t / 1e6:.0f} GB/s (8 B/elem)")
    t = timed(lambda: native.scores_mask(scores, seg0, thr, outs=masks, params=params))
    print(f"  scores_mask + param *= mask (fp16): {t:.3f} ms, {n * 12 / t / 1e6:.0f} GB/s (12 B/elem)")
    d = dict(zip(map(str, range(len(scores))), scores))
    t = timed(lambda: ls.get_mask(d, 0.5, 0.8), reps=3)
    print(f"  LayerSparsity.get_mask (protect + global select + masks): {t:.3f} ms")
    # what the reference does with the same scores on the GPU (it does it on the CPU): cat + topk
    if blocks == 1:
        def ref():
            allv = torch.cat([s.flatten() for s in scores])
            thr = torch.topk(allv, n // 2, largest=False)[0][-1]
            return [(s > thr).float() for s in scores]
        t = timed(ref, reps=3)
        print(f"  torch cat + topk + compare on the same GPU: {t:.3f} ms")
    del params, grads, acc, scores, masks, d
    torch.cuda.empty_cache()
print("ok")
