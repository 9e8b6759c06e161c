"""Experiment: the SparseGPT block step with the chains of a linear started as soon as ITS Hessian is complete, the
remaining Hessians running on a reduced grid (VLMC_HESS_MAX_SMS) at the same time.  python scripts/sgpt_overlap_probe.py"""
import os, sys, time
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
side = torch.cuda.Stream()


def baseline(ws):
    for n in shape:
        H[n].zero_()
        native.hessian_accum(inputs[shape[n][2]], H[n], 0, bench.N_SEQ)
    schedule.sparsegpt_block([(ws[n], H[n], 0.5, 0, 0) for n in shape], 0.01, 128, [U[n] for n in shape])


def overlapped(ws, first, cap):
    """`first`: the linears whose Hessians come first at full width; their chains then run on the side stream family while
    the remaining Hessians are accumulated on `cap` SMs."""
    main = torch.cuda.current_stream()
    rest = [n for n in shape if n not in first]
    os.environ.pop("VLMC_HESS_MAX_SMS", None)
    for n in first:
        H[n].zero_()
        native.hessian_accum(inputs[shape[n][2]], H[n], 0, bench.N_SEQ)
    ev = torch.cuda.Event(); ev.record(main)
    side.wait_event(ev)
    with torch.cuda.stream(side):
        schedule.sparsegpt_block([(ws[n], H[n], 0.5, 0, 0) for n in first], 0.01, 128, [U[n] for n in first])
        done = torch.cuda.Event(); done.record(side)
    os.environ["VLMC_HESS_MAX_SMS"] = str(cap)
    for n in rest:
        H[n].zero_()
        native.hessian_accum(inputs[shape[n][2]], H[n], 0, bench.N_SEQ)
    os.environ.pop("VLMC_HESS_MAX_SMS", None)
    schedule.sparsegpt_block([(ws[n], H[n], 0.5, 0, 0) for n in rest], 0.01, 128, [U[n] for n in rest])
    main.wait_event(done)


def timed(fn, reps=3):
    out = []
    for rep in range(reps + 1):
        ws = bench.make_block(torch, dev, seed=rep)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn(ws)
        b.record()
        torch.cuda.synchronize()
        if rep:
            out.append(a.elapsed_time(b))
    return out


print("baseline", [round(t, 1) for t in timed(baseline)], flush=True)
small = ["q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj"]
for cap in (140, 132, 120):
    print(f"down_proj first, rest on {cap} SMs", [round(t, 1) for t in timed(lambda ws: overlapped(ws, ["down_proj"], cap))], flush=True)
for cap in (140, 132, 120, 100):
    print(f"six small first, down_proj Hessian on {cap} SMs", [round(t, 1) for t in timed(lambda ws: overlapped(ws, small, cap))], flush=True)
print("baseline", [round(t, 1) for t in timed(baseline)], flush=True)
