"""The SparseGPT bench step (7 Hessians over 128 x 2048 tokens, then the pipelined chains) with a 1 ms NVML trace:
SM clock / power per phase, chains timed right after the Hessians and after an idle gap.  python scripts/sgpt_step_probe.py"""
import os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "vlm-compression_b200"))
import torch
import bench
from vlmc import native, schedule
native.load()
dev = "cuda"
torch.manual_seed(0)
inputs = bench.make_inputs(torch, dev, bench.N_SEQ, 1000)
shape = {n: (R, C, inp) for n, R, C, inp in bench.LINEARS}
H = {n: torch.zeros(shape[n][1], shape[n][1], device=dev) for n in shape}
U = {n: torch.empty(shape[n][1], shape[n][1], device=dev) for n in shape}

samples, stop = [], threading.Event()
def sampler():
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    while not stop.is_set():
        samples.append((time.perf_counter(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1e3))
        time.sleep(0.001)
threading.Thread(target=sampler, daemon=True).start()

def stats(t0, t1):
    clk = sorted(s[1] for s in samples if t0 <= s[0] <= t1)
    pw = [s[2] for s in samples if t0 <= s[0] <= t1]
    if not clk:
        return "no samples"
    return f"clock min {clk[0]} med {clk[len(clk) // 2]} max {clk[-1]} MHz, power max {max(pw):.0f} W, {len(clk)} samples"

def hessians():
    for n in shape:
        H[n].zero_()
        native.hessian_accum(inputs[shape[n][2]], H[n], 0, bench.N_SEQ)

def chains(ws):
    return schedule.sparsegpt_block([(ws[n], H[n], 0.5, 0, 0) for n in shape], 0.01, 128, [U[n] for n in shape])

for rep, gap in enumerate([0.0, 0.0, 1e-6, 1e-6, 0.02, 0.5, 0.0, 1e-6]):
    ws = bench.make_block(torch, dev, seed=rep)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    t0 = time.perf_counter(); e[0].record()
    hessians()
    e[1].record()
    if gap:
        torch.cuda.synchronize(); time.sleep(gap)
    t1 = time.perf_counter(); e[2].record()
    chains(ws)
    e[3].record(); torch.cuda.synchronize(); t2 = time.perf_counter()
    th = t0 + e[0].elapsed_time(e[1]) / 1e3
    print(f"rep {rep} gap {gap}: hessians {e[0].elapsed_time(e[1]):.1f} ms [{stats(t0, th)}]; chains {e[2].elapsed_time(e[3]):.1f} ms "
          f"(host {1e3 * (t2 - t1):.1f}) [{stats(t1 if gap else th, t2)}]", flush=True)
stop.set()
