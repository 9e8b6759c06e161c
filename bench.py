#!/usr/bin/env python
"""bench.py - seconds per Vicuna-7B block pruned (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--method wanda_nm|wanda_unstructured]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...     # the reference's CPU path (torch port, oracle/cpu_port.py)

A "step" prunes ONE Vicuna-7B decoder layer (q,k,v,o 4096x4096; gate,up 11008x4096; down 4096x11008, fp16,
random init) the way the reference's per-layer wrapper API dictates: for each of the 7 linears, Wanda
statistics over its 128 x 2048 synthetic fp16 calibration tokens (WrappedGPT.add_batch), then score + mask
selection + weight zeroing.  Block forwards are excluded (activations are the synthetic inputs), SURVEY 8(d).

N > 1 (strong scaling, the plan of SURVEY 8e): the 128 calibration sequences are split across ranks, one
packed NCCL all-reduce merges the 7 scaler_row vectors, output rows are split for selection and all-gathered.

value     device-resident inputs, CUDA-event timed, max over ranks
e2e       the same step through the public wrapper API from PINNED HOST buffers (activations + weights H2D,
          pruned weights + masks D2H inside the timed region)
roofline  the dominant kernel (colstats / sqnorm_accum): algorithmic bytes = T*C*2 per launch
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "vlm-compression_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

D, FF = 4096, 11008
LINEARS = [  # name, rows, cols, input id
    ("q_proj", D, D, "attn_in"), ("k_proj", D, D, "attn_in"), ("v_proj", D, D, "attn_in"),
    ("o_proj", D, D, "attn_out"), ("gate_proj", FF, D, "mlp_in"), ("up_proj", FF, D, "mlp_in"),
    ("down_proj", D, FF, "mlp_mid"),
]
INPUT_DIMS = {"attn_in": D, "attn_out": D, "mlp_in": D, "mlp_mid": FF}
N_SEQ, SEQ_LEN = 128, 2048


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ GPU arm
def make_block(torch, dev, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    return {name: (torch.randn(r, c, device=dev, generator=g) * 0.02).half() for name, r, c, _ in LINEARS}


def make_inputs(torch, dev, n_seq, seed):
    """Synthetic calibration activations [n_seq, SEQ_LEN, C] fp16: N(0,1) * per-channel gain + offset."""
    out = {}
    for i, (name, C) in enumerate(INPUT_DIMS.items()):
        g = torch.Generator(device=dev).manual_seed(seed + i)
        gain = torch.exp(torch.rand(C, device=dev, generator=g) * 2.77 - 1.386)
        off = torch.randn(C, device=dev, generator=g) * 0.3
        x = torch.empty(n_seq, SEQ_LEN, C, device=dev, dtype=torch.float16)
        for j in range(n_seq):  # chunked: never materialise the fp32 copy of the whole set
            x[j] = (torch.randn(SEQ_LEN, C, device=dev, generator=g) * gain + off).half()
        out[name] = x
    return out


class Timers:
    def __init__(self, torch):
        self.torch, self.pairs = torch, []

    def span(self):
        a, b = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
        return a, b


def prune_block(torch, native, parallel, weights, inputs, method, calib_batch, rank, world, stat_events=None,
                n_total=N_SEQ):
    """One step.  Returns the number of vlmc kernel launches issued."""
    launches = 0
    n_local = next(iter(inputs.values())).shape[0]
    scalers = {}
    for name, R, C, inp in LINEARS:                       # phase 1: statistics (per-linear API, no de-duplication)
        x = inputs[inp]
        s = torch.zeros(C, device=x.device, dtype=torch.float32)
        n = 0
        for j in range(0, n_local, calib_batch):
            xb = x[j:j + calib_batch]
            if stat_events is not None:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
            native.sqnorm_accum(xb, s, n, xb.shape[0])
            if stat_events is not None:
                b.record()
                stat_events.append((a, b, xb.numel() * 2))
            n += xb.shape[0]
            launches += 1
        scalers[name] = s
    if world > 1:
        parallel.merge_running_means(list(scalers.values()), n_local, n_total=n_total)
    masks = {}
    for name, R, C, _ in LINEARS:                         # phase 2: score + select + apply
        W = weights[name]
        if method == "wanda_nm":
            def sel(Wr, s, keep):
                return native.wanda_nm(Wr, s, 2, 4, keep_mask=keep)[1]
        else:
            def sel(Wr, s, keep):
                return native.wanda_rowselect(Wr, s, int(C * 0.5), keep_mask=keep)[1]
        masks[name], _ = parallel.prune_linear_row_sharded(W, scalers[name], sel, rank, world)
        launches += 2
    return launches, masks


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from vlmc import native, parallel
    native.load()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    s0, s1 = parallel.sample_range(N_SEQ, rank, world)
    inputs = make_inputs(torch, dev, s1 - s0, seed=1000 + 17 * rank)
    nsets = min(args.steps + args.warmup, 12)
    wsets = [make_block(torch, dev, seed=s) for s in range(nsets)]
    pristine = None
    if args.steps + args.warmup > nsets:
        pristine = make_block(torch, dev, seed=0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step_i = 0

    def one_step(events=None):
        nonlocal step_i
        w = wsets[step_i % nsets]
        if pristine is not None and step_i >= nsets:
            for k in w:
                w[k].copy_(pristine[k])
        step_i += 1
        return prune_block(torch, native, parallel, w, inputs, args.method, args.calib_batch, rank, world, events)

    for _ in range(args.warmup):
        one_step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    stat_events = []
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    launches = 0
    for _ in range(args.steps):
        l, _ = one_step(stat_events)
        launches += l
    t1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([t0.elapsed_time(t1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms.item()) / args.steps
    stat_ms = sum(a.elapsed_time(b) for a, b, _ in stat_events)
    stat_bytes = sum(nb for _, _, nb in stat_events)

    # ---- e2e: same step through the wrapper API from pinned host buffers ---------------------------------
    e2e = run_e2e(torch, native, parallel, dev, args, rank, world, inputs)
    if world > 1:
        t = torch.tensor([e2e["ms"]], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e["ms"] = float(t.item())

    if rank == 0:
        hbm, how = peaks()
        achieved = stat_bytes / (stat_ms * 1e-3) / 1e9
        out = {
            "metric": "s_per_vicuna7b_block_pruned", "value": ms_per_step / 1e3, "unit": "s/block",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"{args.method} on one InstructBLIP-Vicuna-7B LLM block (7 linears, fp16 weights, "
                                   f"random init), {N_SEQ}x{SEQ_LEN} fp16 calibration tokens per linear",
                       "method": args.method, "calib_batch": args.calib_batch,
                       "l2": "inputs larger than L2 (12.2 GB of activations per step, fresh weight set per step)",
                       "parallelism": f"tokens/{world} + allreduce, rows/{world} + allgather" if world > 1 else "1 GPU"},
            "gpu_launches": launches,
            "clocks": clocks,
            "e2e": {"value": e2e["ms"] / 1e3, "unit": "s/block", "h2d_bytes_per_step": e2e["h2d"],
                    "d2h_bytes_per_step": e2e["d2h"]},
            "roofline": {"bound": "hbm", "kernel": "colstats_kernel (vlmc_sqnorm_accum)", "achieved": achieved,
                         "peak": hbm, "peak_source": how, "unit": "GB/s", "frac": achieved / hbm, "traffic": None,
                         "launches": len(stat_events), "avg_launch_ms": stat_ms / max(len(stat_events), 1),
                         "share_of_step": stat_ms / (ms_per_step * args.steps)},
        }
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(args, budget_s=20.0)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_e2e(torch, native, parallel, dev, args, rank, world, dev_inputs):
    """Host buffers in, host buffers out: per distinct input, pinned chunks stream H2D on a copy stream while the
    statistics kernels consume the previous chunk; weights H2D, pruned weights and masks D2H."""
    from vlmc.compression.pruners.wanda_pruner import WrappedGPT
    n_local = next(iter(dev_inputs.values())).shape[0]
    chunk = min(8, n_local)
    host_in = {}
    for k, x in dev_inputs.items():
        h = torch.empty(x.shape, dtype=x.dtype, pin_memory=True)
        h.copy_(x)
        host_in[k] = h
    torch.cuda.synchronize()
    del dev_inputs
    host_w = {name: (torch.randn(r, c) * 0.02).half().pin_memory() for name, r, c, _ in LINEARS}
    host_out_w = {name: torch.empty(r, c, dtype=torch.float16, pin_memory=True) for name, r, c, _ in LINEARS}
    host_out_m = {name: torch.empty(r, c, dtype=torch.bool, pin_memory=True) for name, r, c, _ in LINEARS}
    copy_stream = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream(dev)
    bufs = {C: [torch.empty(chunk, SEQ_LEN, C, device=dev, dtype=torch.float16) for _ in range(2)] for C in (D, FF)}
    h2d = sum(h.numel() * 2 for h in host_in.values()) + sum(w.numel() * 2 for w in host_w.values())
    d2h = sum(w.numel() * 3 for w in host_w.values())

    class Lin:  # minimal stand-in for nn.Linear: the wrapper only needs .weight (shape, device)
        def __init__(self, w):
            self.weight = w

    def step():
        dW = {}
        with torch.cuda.stream(copy_stream):
            for name in host_w:
                dW[name] = host_w[name].to(dev, non_blocking=True)
        w_ready = torch.cuda.Event(); w_ready.record(copy_stream)
        wrappers = {name: WrappedGPT(Lin(dW[name])) for name, *_ in LINEARS}
        for inp, C in INPUT_DIMS.items():
            users = [n for n, _, _, i in LINEARS if i == inp]
            free = [torch.cuda.Event(), torch.cuda.Event()]
            for ci, j in enumerate(range(0, n_local, chunk)):
                b = bufs[C][ci % 2]
                n = min(chunk, n_local - j)
                with torch.cuda.stream(copy_stream):
                    if ci >= 2:
                        copy_stream.wait_event(free[ci % 2])
                    b[:n].copy_(host_in[inp][j:j + n], non_blocking=True)
                    ready = torch.cuda.Event(); ready.record(copy_stream)
                main.wait_event(ready)
                for u in users:                      # per-linear API: each linear reads the chunk itself
                    wrappers[u].add_batch(b[:n], None)
                free[ci % 2] = torch.cuda.Event(); free[ci % 2].record(main)
        scal = [wrappers[n].scaler_row for n, *_ in LINEARS]
        if world > 1:
            parallel.merge_running_means(scal, n_local, n_total=N_SEQ)
        main.wait_event(w_ready)
        for name, R, C, _ in LINEARS:
            W = dW[name]
            if args.method == "wanda_nm":
                sel = lambda Wr, s, keep: native.wanda_nm(Wr, s, 2, 4, keep_mask=keep)[1]
            else:
                sel = lambda Wr, s, keep, C=C: native.wanda_rowselect(Wr, s, int(C * 0.5), keep_mask=keep)[1]
            keep, _ = parallel.prune_linear_row_sharded(W, wrappers[name].scaler_row, sel, rank, world)
            host_out_w[name].copy_(W, non_blocking=True)
            host_out_m[name].copy_(keep, non_blocking=True)

    step()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    k = max(2, min(args.steps, 5))
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    wall0 = time.perf_counter()
    t0.record()
    for _ in range(k):
        step()
    t1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - wall0) * 1e3
    return {"ms": max(t0.elapsed_time(t1), wall) / k, "h2d": h2d, "d2h": d2h, "steps": k}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_block_seconds(args, n_seq_sample, threads):
    """The reference's algorithm on host cores (oracle/cpu_port.py): statistics on n_seq_sample of the 128 sequences
    (scaled linearly to 128, labelled extrapolated) + full selection for the 7 linears."""
    import torch
    from oracle import cpu_port
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(0)
    t_stats = 0.0
    scal = {}
    xs = {inp: (torch.randn(SEQ_LEN, C, generator=g)).half() for inp, C in INPUT_DIMS.items()}
    for name, R, C, inp in LINEARS:
        st = cpu_port.WandaStat(C)
        t0 = time.perf_counter()
        for _ in range(n_seq_sample):
            st.add_batch(xs[inp].unsqueeze(0))
        t_stats += time.perf_counter() - t0
        scal[name] = st.scaler_row
    t_sel = 0.0
    for name, R, C, _ in LINEARS:
        W = (torch.randn(R, C, generator=g) * 0.02).half()
        t0 = time.perf_counter()
        if args.method == "wanda_nm":
            cpu_port.wanda_select(W, scal[name], 0.5, 2, 4)
        else:
            cpu_port.wanda_select(W, scal[name], 0.5)
        t_sel += time.perf_counter() - t0
    return t_stats * (N_SEQ / n_seq_sample) + t_sel, t_stats, t_sel


def cpu_baseline(args, budget_s):
    threads = os.cpu_count() or 1
    n_sample = 2
    total, t_stats, t_sel = cpu_block_seconds(args, n_sample, threads)
    return {"value": total, "unit": "s/block", "cores": threads, "kind": "port",
            "sample": f"statistics on {n_sample} of {N_SEQ} sequences per linear ({t_stats:.2f} s, scaled x{N_SEQ // n_sample}, "
                      f"extrapolated) + full {args.method} selection of the 7 linears ({t_sel:.2f} s)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    for _ in range(args.warmup):
        cpu_block_seconds(args, 1, threads)
    vals = []
    k = args.steps
    n_sample = 1
    for _ in range(k):
        total, t_stats, t_sel = cpu_block_seconds(args, n_sample, threads)
        vals.append(total)
    v = sum(vals) / len(vals)
    print(json.dumps({
        "impl": "reference", "metric": "s_per_vicuna7b_block_pruned", "value": v, "unit": "s/block",
        "n_gpus": args.gpus, "steps": k, "warmup": args.warmup, "ms_per_step": v * 1e3,
        "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.method} on one InstructBLIP-Vicuna-7B LLM block, {N_SEQ}x{SEQ_LEN} fp16 "
                               "calibration tokens per linear", "method": args.method},
        "cpu_baseline": {"value": v, "unit": "s/block", "cores": threads, "kind": "port",
                         "sample": f"each step: statistics on {n_sample} of {N_SEQ} sequences per linear scaled "
                                   f"x{N_SEQ // n_sample} (extrapolated) + full selection of the 7 linears"},
        "e2e": {"value": v, "unit": "s/block", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="vlmc", choices=["vlmc", "reference"])
    ap.add_argument("--method", default="wanda_nm", choices=["wanda_nm", "wanda_unstructured"])
    ap.add_argument("--calib-batch", type=int, default=N_SEQ,
                    help="sequences per add_batch call (reference hooks use 1; the wrapper API takes any b)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "vlmc":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
