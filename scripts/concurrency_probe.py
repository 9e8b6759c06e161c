"""The 7 factorisation chains and the 7 sweep chains of a Vicuna block: one after the other vs concurrently
(vlmc.schedule), and whether the power-capped SM clock after the tensor-core phase slows them.  python scripts/concurrency_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vlm-compression_b200"))
import torch
from vlmc import native, schedule

native.load()
torch.manual_seed(0)
dev = "cuda"
shapes = [(4096, 4096)] * 4 + [(11008, 4096)] * 2 + [(4096, 11008)]
Hs0 = {}
for C in (4096, 11008):
    x = (torch.randn(2 * C, C, device=dev) * (torch.rand(C, device=dev) + 0.5)).half()
    H = torch.zeros(C, C, device=dev)
    native.hessian_accum(x, H, 0, 1)
    Hs0[C] = H
    del x
Ws0 = [(torch.randn(R, C, device=dev) * 0.02).half() for R, C in shapes]
Us = [torch.empty(C, C, device=dev) for _, C in shapes]


def timed(fn):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    r = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b), r


def sequential():
    out = []
    for (R, C), U in zip(shapes, Us):
        H = Hs0[C].clone()
        damp, dead = native.hessian_prepare(H, 0.01)
        native.chol_inv_upper(H, U)
        out.append((U, dead))
    return out


for rep in range(2):
    ms, facs = timed(sequential)
    print(f"factor one by one: {ms:.2f} ms", flush=True)
for rep in range(2):
    Hs = [Hs0[C].clone() for _, C in shapes]
    ms, facs = timed(lambda: schedule.factor_concurrent(Hs, 0.01, Us))
print(f"factor concurrent: {ms:.2f} ms", flush=True)

for rep in range(2):
    Ws = [W.clone() for W in Ws0]
    ms, _ = timed(lambda: [native.obs_sweep(W, U, 0.5, dead=d) for W, (U, d) in zip(Ws, facs)])
    print(f"sweep one by one: {ms:.2f} ms", flush=True)
for rep in range(2):
    Ws = [W.clone() for W in Ws0]
    ms, _ = timed(lambda: schedule.sweep_concurrent([(W, U, d, 0.5, 0, 0) for W, (U, d) in zip(Ws, facs)]))
print(f"sweep concurrent: {ms:.2f} ms", flush=True)

# ---- does the tensor-core phase slow the latency-bound chains that follow it (power-capped SM clock)? ----
import threading
import time


def sm_clock_sampler(stop, out):
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(0)
        while not stop.is_set():
            out.append((time.perf_counter(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                        pynvml.nvmlDeviceGetPowerUsage(h) / 1e3))
            time.sleep(0.002)
    except Exception as e:  # noqa: BLE001
        out.append(("nvml", str(e)))


xs = {C: torch.randn(64 * 2048, C, device=dev).half() for C in (4096, 11008)}
Hh = {C: torch.zeros(C, C, device=dev) for C in (4096, 11008)}
for mode in ("cold", "hot", "hot", "cold"):
    torch.cuda.synchronize()
    time.sleep(0.5)
    stop, samples = threading.Event(), []
    th = threading.Thread(target=sm_clock_sampler, args=(stop, samples), daemon=True)
    th.start()
    t_h = 0.0
    if mode == "hot":
        t_h, _ = timed(lambda: [native.hessian_accum(xs[C].view(64, 2048, C), Hh[C], 0, 64) for C in (4096, 4096, 4096, 4096, 4096, 4096, 11008)])
    Hs = [Hs0[C].clone() for _, C in shapes]
    t0 = time.perf_counter()
    ms, facs = timed(lambda: schedule.factor_concurrent(Hs, 0.01, Us))
    t1 = time.perf_counter()
    stop.set()
    th.join()
    clk = [s[1] for s in samples if isinstance(s[0], float) and t0 <= s[0] <= t1]
    pw = [s[2] for s in samples if isinstance(s[0], float) and t0 <= s[0] <= t1]
    print(f"{mode}: hessians {t_h:.1f} ms, then factor concurrent {ms:.2f} ms; SM clock during the factorisations "
          f"min {min(clk) if clk else None} / median {sorted(clk)[len(clk) // 2] if clk else None} MHz, power max {max(pw) if pw else None} W", flush=True)
