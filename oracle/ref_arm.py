"""TEST / BASELINE INFRASTRUCTURE ONLY - times the UNMODIFIED reference on host cores (bench.py `--impl reference`
and `cpu_baseline`) and, in the live differential tests, runs it beside the CUDA path on the same tensors.

The reference is loaded by path through oracle/ref_loader.py (in the build container from /root/reference, on the GPU
box from the verbatim copy under baseline/_ref that oracle/fetch_ref.py leaves).  Nothing under vlmc/ imports this.

How a block is run without LAVIS (block forwards are excluded from the metric, activations are synthetic):
the reference's composite pruner `BLIPT5LayerWandaPruner.prune()` / `BLIPT5LayerDSnoTPruner.prune()` is driven through its
own public entry point on a stand-in model whose ONE block holds real `nn.Linear` layers of the named shapes (so
`find_layers`, the forward hooks, `WrappedGPT.add_batch` and the inline score / select / apply code at
wanda_pruner.py:287-347 all run as shipped) but whose block forward skips the GEMMs: every linear has an instance-level
`forward` that returns a preallocated dummy output, so `Module.__call__` still fires the hooks with the synthetic input.
"""
import contextlib
import time
import types

import torch
import torch.nn as nn

from oracle import ref_loader

VICUNA_BLOCK = [  # name, rows, cols, input id   (modeling_llama.py:151-153,178-181)
    ("self_attn.q_proj", 4096, 4096, "attn_in"), ("self_attn.k_proj", 4096, 4096, "attn_in"),
    ("self_attn.v_proj", 4096, 4096, "attn_in"), ("self_attn.o_proj", 4096, 4096, "attn_out"),
    ("mlp.gate_proj", 11008, 4096, "mlp_in"), ("mlp.up_proj", 11008, 4096, "mlp_in"),
    ("mlp.down_proj", 4096, 11008, "mlp_mid"),
]
T5XL_ENC_BLOCK = [  # modeling_t5.py:320-326, d_model 2048, d_ff 5120
    ("layer.0.SelfAttention.q", 2048, 2048, "attn_in"), ("layer.0.SelfAttention.k", 2048, 2048, "attn_in"),
    ("layer.0.SelfAttention.v", 2048, 2048, "attn_in"), ("layer.0.SelfAttention.o", 2048, 2048, "attn_out"),
    ("layer.1.DenseReluDense.wi_0", 5120, 2048, "mlp_in"), ("layer.1.DenseReluDense.wi_1", 5120, 2048, "mlp_in"),
    ("layer.1.DenseReluDense.wo", 2048, 5120, "mlp_mid"),
]


def synth_acts(shape, seed, dtype, device="cpu"):
    """N(0,1) * per-channel gain (LogUniform[0.25,4]) + per-channel offset N(0, 0.3^2)."""
    C = shape[-1]
    g = torch.Generator().manual_seed(seed)
    gain = torch.exp(torch.rand(C, generator=g) * 2.77 - 1.386)
    off = torch.randn(C, generator=g) * 0.3
    return (torch.randn(shape, generator=g) * gain + off).to(dtype).to(device)


class NoGemmBlock(nn.Module):
    """One transformer block's linears (real nn.Linear modules, random init N(0, 0.02^2)); forward() hands every linear
    its synthetic calibration input and returns the block input unchanged.  `distinct` resident sequences per input are
    cycled over the calibration set (the work per call is what is timed, not the values)."""

    def __init__(self, linears, seq_len, dtype, distinct=8, seed=0, device="cpu"):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.names = []
        self.acts = {}
        self.calls = 0
        dims = {}
        for name, R, C, inp in linears:
            lin = nn.Linear(C, R, bias=False)
            lin.weight.data = (torch.randn(R, C, generator=g) * 0.02).to(dtype).to(device)
            dummy = torch.zeros(1, 1, 1, dtype=dtype, device=device)
            lin.forward = types.MethodType(lambda self_, x, _d=dummy: _d, lin)     # no GEMM: hooks still fire
            parent = self
            *path, leaf = name.split(".")
            for p in path:
                if not hasattr(parent, p):
                    setattr(parent, p, nn.Module())
                parent = getattr(parent, p)
            setattr(parent, leaf, lin)
            self.names.append((name, inp))
            dims[inp] = C
        for i, (inp, C) in enumerate(dims.items()):
            if inp != "attn_in":       # attn_in is the block input the Catcher captured
                self.acts[inp] = [synth_acts((1, seq_len, C), 2000 + 31 * i + j, dtype, device) for j in range(distinct)]

    def linear(self, name):
        m = self
        for p in name.split("."):
            m = getattr(m, p)
        return m

    def forward(self, x, attention_mask=None, position_ids=None, dense=False):
        j = self.calls
        self.calls += 1
        for name, inp in self.names:
            self.linear(name)(x if inp == "attn_in" else self.acts[inp][j % len(self.acts[inp])])
        return (x,)


class StandIn(nn.Module):
    """What the composite pruners need of a LAVIS model: `llm_model.config.use_cache`,
    `llm_model.model.layers`, `maybe_autocast`, `forward(batch)`."""

    def __init__(self, block):
        super().__init__()
        self.llm_model = nn.Module()
        self.llm_model.config = types.SimpleNamespace(use_cache=True)
        self.llm_model.model = nn.Module()
        self.llm_model.model.layers = nn.ModuleList([block])

    def maybe_autocast(self, dtype=torch.float16):
        return contextlib.nullcontext()

    def forward(self, batch):
        return self.llm_model.model.layers[0](batch["x"], attention_mask=None, position_ids=None)


def calib_loader(n_seq, seq_len, C, dtype, distinct=8, device="cpu"):
    xs = [synth_acts((1, seq_len, C), 1000 + j, dtype, device) for j in range(min(distinct, n_seq))]
    return [{"text_input": ["a"], "x": xs[j % len(xs)]} for j in range(n_seq)]


def run_composite(which, linears, n_seq, seq_len, dtype, sparsity, prune_n=0, prune_m=0, threads=None, device="cpu",
                  distinct=8, record=None):
    """`BLIPT5Layer{Wanda,DSnoT}Pruner(model=stand-in, data_loader=..., **cfg).prune()`, the reference's own entry point.
    Returns (seconds inside prune(), stand-in model).  record: optional list that receives every per-layer WrappedGPT the
    reference creates (a recording subclass is swapped in for the duration; the reference file is not touched)."""
    if threads:
        torch.set_num_threads(threads)
    ref = ref_loader.load()
    block = NoGemmBlock(linears, seq_len, dtype, distinct=distinct, device=device)
    model = StandIn(block).eval()
    C_in = next(C for _, _, C, inp in linears if inp == "attn_in")
    loader = calib_loader(n_seq, seq_len, C_in, dtype, distinct, device)
    cfg = dict(t5_prune_spec=f"24-{1.0 - sparsity}-1.0-1.0", vit_prune_spec="39-1.0-1.0-1.0", t5_pruning_method="none",
               vit_pruning_method="none", t5_model_prefix="llm_model", num_samples=n_seq, sparsity_ratio_granularity=None,
               score_method="obd_avg", prune_n=prune_n, prune_m=prune_m)
    mod = {"wanda": ref.wanda, "dsnot": ref.dsnot}[which]
    cls = {"wanda": "BLIPT5LayerWandaPruner", "dsnot": "BLIPT5LayerDSnoTPruner"}[which]
    pruner = getattr(mod, cls)(model=model, data_loader=loader, **cfg)
    orig = mod.WrappedGPT
    if record is not None:
        class Recording(orig):
            def __init__(self, layer, *a, **k):
                super().__init__(layer, *a, **k)
                record.append(self)
        mod.WrappedGPT = Recording
    try:
        if device != "cpu":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(None):
            pruner.prune()
        if device != "cpu":
            torch.cuda.synchronize()
        return time.perf_counter() - t0, model
    finally:
        mod.WrappedGPT = orig


def sparsegpt_block_seconds(linears, n_seq, seq_len, dtype, sparsity=0.5, sample_seqs=8, threads=None, full=("q", "down")):
    """BASELINE.md 4 / SURVEY 8(d): the reference has no llm_model SparseGPT driver (F10), so its per-layer class is timed:
    `SparseGPT.add_batch` on `sample_seqs` of the n_seq sequences per distinct width (scaled linearly in T, stated as
    extrapolated) and `fasterprune` IN FULL for one [C_small x C_small] linear and the down projection; the remaining
    linears are scaled from the measured one of the same width by their row count."""
    if threads:
        torch.set_num_threads(threads)
    ref = ref_loader.load()
    g = torch.Generator().manual_seed(0)
    t_add, t_prune, notes = {}, {}, []
    widths = sorted({C for _, _, C, _ in linears})
    for C in widths:
        lin = nn.Linear(C, 8, bias=False)
        lin.weight.data = lin.weight.data.to(dtype)
        sg = ref.sparsegpt.SparseGPT(lin)
        xs = [synth_acts((1, seq_len, C), 3000 + j, dtype) for j in range(2)]
        t0 = time.perf_counter()
        for j in range(sample_seqs):
            sg.add_batch(xs[j % 2], None)
        t_add[C] = (time.perf_counter() - t0) / sample_seqs
    total_add = sum(t_add[C] * n_seq for _, _, C, _ in linears)
    notes.append("add_batch on %d of %d sequences per width (%s s/call), scaled x%d (extrapolated)" % (
        sample_seqs, n_seq, ", ".join(f"C={C}: {t_add[C]:.2f}" for C in widths), n_seq // sample_seqs))
    measured = {}
    for C in widths:
        name, R, _, _ = next(l for l in linears if l[2] == C and (l[1] == C or C == max(widths)))
        lin = nn.Linear(C, R, bias=False)
        lin.weight.data = (torch.randn(R, C, generator=g) * 0.02).to(dtype)
        sg = ref.sparsegpt.SparseGPT(lin)
        x = synth_acts((1, max(2 * C, seq_len), C), 3100 + C, dtype)      # T >= 2C: H positive definite, no damping
        sg.add_batch(x, None)
        t0 = time.perf_counter()
        sg.fasterprune(sparsity, prune_n=0, prune_m=0, percdamp=0.01, blocksize=128)
        measured[C] = (R, time.perf_counter() - t0)
        sg.free()
        notes.append(f"fasterprune in full on {name} [{R}x{C}]: {measured[C][1]:.1f} s")
    total_prune = 0.0
    for name, R, C, _ in linears:
        R0, t = measured[C]
        total_prune += t if R == R0 else t * (R / R0)      # sweep + sorts scale with the rows; the factorisation does not
    notes.append("other linears scaled from the measured one of the same width by rows (upper bound on the factorisation share)")
    return total_add + total_prune, "; ".join(notes)
