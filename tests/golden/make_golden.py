"""Generates the committed golden fixtures by running the UNMODIFIED reference (loaded by path through
oracle/ref_loader.py) on seeded inputs.  Needs /root/reference, so it only runs in the build container:

    python tests/golden/make_golden.py [name ...]        # rewrites tests/golden/*.npz (all, or the named generators)
    python tests/golden/make_golden.py --check [name ...] # regenerates in memory, compares with the committed files, exit 1 on a difference

The fixtures pin oracle/oracle.py (tests/test_oracle_vs_golden.py, CPU) and the CUDA kernels
(tests/test_gpu_parity.py, GPU box, where the reference itself is absent).
Half / bfloat16 tensors are stored as float32 (exact up-cast) next to a dtype tag.
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ref_loader  # noqa: E402
import toy_model  # noqa: E402

TAG = {torch.float32: "f32", torch.float16: "f16", torch.bfloat16: "bf16"}


def f32(t):
    return t.detach().float().cpu().numpy().copy()   # copy: fp32 CPU tensors would alias the live state


def packw(t):
    """Weights in their own precision: bf16 as raw uint16 bits, f16 / f32 natively (see tests/golden_util.py)."""
    t = t.detach().cpu()
    if t.dtype == torch.bfloat16:
        return t.view(torch.int16).numpy().view(np.uint16)
    return t.numpy().copy()


CHECK = False          # --check: regenerate in memory and compare with the committed files, write nothing
MISMATCH = []


def save(name, **arrs):
    path = os.path.join(HERE, name)
    if CHECK:
        old = np.load(path, allow_pickle=False)
        bad = sorted(set(old.files) ^ set(arrs))
        for k in set(old.files) & set(arrs):
            a, b = np.asarray(arrs[k]), old[k]
            if a.shape != b.shape or a.dtype != b.dtype or not np.array_equal(a, b, equal_nan=a.dtype.kind == "f"):
                bad.append(k)
        print(f"{name}: {'ok' if not bad else 'DIFFERS in %d of %d arrays: %s' % (len(bad), len(arrs), bad[:6])}")
        if bad:
            MISMATCH.append(name)
        return
    np.savez_compressed(path, **arrs)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def act(shape, seed, dtype, C):
    g = torch.Generator().manual_seed(seed)
    gain = torch.exp(torch.rand(C, generator=g) * 2.77 - 1.386)      # LogUniform[0.25, 4]
    off = torch.randn(C, generator=g) * 0.3
    return (torch.randn(shape, generator=g) * gain + off).to(dtype)


def gen_wanda_stats(ref):
    out = {}
    for tag, dtype in (("bf16", torch.bfloat16), ("f16", torch.float16), ("f32", torch.float32)):
        C = 96
        layer = nn.Linear(C, 8, bias=False)
        w = ref.wanda.WrappedGPT(layer)
        shapes = [(1, 40, C), (40, C), (3, 17, C), (1, 64, C)]   # 3-D b=1, 2-D, 3-D b=3, again b=1
        for i, shp in enumerate(shapes):
            x = act(shp, 10 + i, dtype, C)
            w.add_batch(x, None)
            out[f"{tag}_x{i}"] = f32(x)
            out[f"{tag}_scaler{i}"] = f32(w.scaler_row)
            out[f"{tag}_n{i}"] = np.int64(w.nsamples)
    save("wanda_stats.npz", **out)


def gen_dsnot_stats(ref):
    out = {}
    for tag, dtype in (("bf16", torch.bfloat16), ("f32", torch.float32)):
        C = 96
        layer = nn.Linear(C, 8, bias=False)
        w = ref.dsnot.WrappedGPT(layer)
        shapes = [(1, 40, C), (1, 40, C), (2, 24, C), (33, C)]
        for i, shp in enumerate(shapes):
            x = act(shp, 20 + i, dtype, C)
            w.add_batch(x, None)
            out[f"{tag}_x{i}"] = f32(x)
            for k in ("scaler_row", "sum_metric_row", "mean", "var"):
                out[f"{tag}_{k}{i}"] = f32(getattr(w, k)).reshape(-1)
            out[f"{tag}_n{i}"] = np.int64(w.nsamples)
            out[f"{tag}_ntok{i}"] = np.int64(w.ntokens)
    save("dsnot_stats.npz", **out)


def run_composite(ref, which, cfg, seed=0, **model_kw):
    """Runs a reference composite pruner on the toy model, recording every per-layer wrapper."""
    mod = getattr(ref, which)
    recorded = []
    orig = mod.WrappedGPT if which != "sparsegpt" else mod.SparseGPT

    class Recording(orig):
        def __init__(self, layer, *a, **k):
            super().__init__(layer, *a, **k)
            recorded.append(self)
    if which == "sparsegpt":
        mod.SparseGPT = Recording
    else:
        mod.WrappedGPT = Recording
    try:
        torch.manual_seed(seed)          # biases / embeddings come from the global generator
        model = toy_model.ToyBlip(seed=seed, **model_kw).eval()
        before = {n: p.detach().clone() for n, p in model.named_parameters()}
        cls = {"wanda": "BLIPT5LayerWandaPruner", "dsnot": "BLIPT5LayerDSnoTPruner",
               "sparsegpt": "BLIPT5LayerSparseGPTPruner"}[which]
        pruner = getattr(mod, cls)(model=model, data_loader=toy_model.toy_batches(cfg["num_samples"],
                                                                                  d_vit=model_kw.get("d_vit", 64)),
                                   **cfg)
        pruner.prune()
    finally:
        if which == "sparsegpt":
            mod.SparseGPT = orig
        else:
            mod.WrappedGPT = orig
    return model, before, recorded


def collect_layers(model, before, recorded, stat_names, store_after=False):
    out = {}
    names = []
    by_layer = {id(w.layer): w for w in recorded}
    for name, module in model.named_modules():
        if type(module) is nn.Linear and id(module) in by_layer:
            w = by_layer[id(module)]
            key = name.replace(".", "/")
            names.append(key)
            out[f"{key}|W_before"] = packw(before[name + ".weight"])
            if store_after:
                out[f"{key}|W_after"] = packw(module.weight)
            else:   # Wanda / DSnoT only zero weights: W_after == W_before * mask, checked here, not stored
                assert torch.equal(module.weight, before[name + ".weight"] * module.mask)
            out[f"{key}|tag"] = np.array(TAG[module.weight.dtype])
            out[f"{key}|mask"] = np.packbits(module.mask.cpu().numpy(), axis=1)
            for s in stat_names:
                v = getattr(w, s, None)
                if v is not None:
                    out[f"{key}|{s}"] = f32(v).reshape(-1)
            sc = getattr(module.weight, "importance_score", None)
            if sc is not None:
                out[f"{key}|importance_score"] = np.float64(sc)
    out["layers"] = np.array(names)
    return out


def gen_wanda_toy(ref):
    # unstructured: LLM per-row path at 60 % (int() count), ViT whole-matrix path at 50 %
    cfg = toy_model.pruner_cfg(0.4, 0.5)
    model, before, rec = run_composite(ref, "wanda", cfg)
    save("wanda_toy_unstructured.npz", **collect_layers(model, before, rec, ["scaler_row"]))
    for n, m in ((2, 4), (4, 8)):
        cfg = toy_model.pruner_cfg(0.5, 0.5, prune_n=n, prune_m=m)
        model, before, rec = run_composite(ref, "wanda", cfg)
        save(f"wanda_toy_{n}of{m}.npz", **collect_layers(model, before, rec, ["scaler_row"]))


DSNOT_TOY = dict(d_llm=296, ff=488, n_llm=1, n_vit=0)      # round(C*0.6) != int(C*0.6) for both widths


def _excise_fixup(text):
    """dsnot_pruner.py with the swap write-back block (:734-740 and its ViT copy) removed = upstream DSnoT."""
    lines = text.split("\n")
    out, skip = [], 0
    for ln in lines:
        if "sub_mask_prune = torch.gather(weight_mask, 1, pruning_indice)" in ln:
            skip = 7                                   # the 6 statements + the blank line between them
        if skip:
            skip -= 1
            continue
        out.append(ln)
    assert len(lines) - len(out) == 14, len(lines) - len(out)
    return "\n".join(out)


def gen_dsnot_toy(ref):
    """Reference composite DSnoT pruner on a 1-layer toy LLM (rows need > 100 kept and pruned columns).
    All cases share the model seed, so W_before and the statistics are stored once; each case adds its masks."""
    stats = ["scaler_row", "sum_metric_row", "mean", "var"]
    upstream = ref_loader.load(patch_source={"dsnot_pruner": _excise_fixup})
    cases = {
        # name: (module namespace, cfg overrides)
        "shipped_unstr60": (ref, dict()),
        "upstream_unstr60": (upstream, dict()),
        "upstream_unstr60_samesign": (upstream, dict(without_same_sign=False)),
        "upstream_unstr60_magnitude": (upstream, dict(initial_method="magnitude")),
        "shipped_2of4": (ref, dict(prune_n=2, prune_m=4)),
        "shipped_4of8": (ref, dict(prune_n=4, prune_m=8)),
        "shipped_without_dsnot": (ref, dict(without_DSnoT=True)),
    }
    out = {}
    for name, (ns, extra) in cases.items():
        cfg = toy_model.pruner_cfg(0.4, 1.0, **extra)
        model, before, rec = run_composite(ns, "dsnot", cfg, **DSNOT_TOY)
        got = collect_layers(model, before, rec, stats)
        if "layers" not in out:
            out.update({k: v for k, v in got.items() if not k.endswith("|mask")})
        else:
            for k in got:
                if k.endswith("|W_before") or k.split("|")[-1] in stats:
                    assert np.array_equal(out[k], got[k]), k
        for key in got["layers"]:
            out[f"{name}|{key}|mask"] = got[f"{key}|mask"]
    out["cases"] = np.array(list(cases))
    save("dsnot_toy.npz", **out)


def gen_lora_merge(ref):
    out = {}
    torch.manual_seed(5)
    for tag, dtype in (("bf16", torch.bfloat16), ("f16", torch.float16), ("f32", torch.float32)):
        for r in (2, 4, 8):
            lin = ref.lora.Linear(72, 40, r=r, lora_alpha=16, bias=False)
            lin.weight.data = (torch.randn(40, 72) * 0.05).to(dtype)
            lin.lora_A.weight.data = torch.randn(r, 72) * 0.1
            lin.lora_B.weight.data = torch.randn(40, r) * 0.1
            lin.mask = torch.rand(40, 72) < 0.5
            lin.sparse = True
            key = f"{tag}_r{r}"
            out[f"{key}|W_before"] = f32(lin.weight)
            out[f"{key}|A"] = f32(lin.lora_A.weight)
            out[f"{key}|B"] = f32(lin.lora_B.weight)
            out[f"{key}|mask"] = lin.mask.numpy().copy()
            out[f"{key}|scaling"] = np.float64(lin.scaling)
            lin.merge()                                       # lora.py:384-387
            out[f"{key}|W_merged"] = f32(lin.weight)
            lin.weight.data[~lin.mask] = 0                    # train.py:634-637
            out[f"{key}|W_remasked"] = f32(lin.weight)
    save("lora_merge.npz", **out)


def gen_lora_forward(ref):
    """The reference's lora.Linear forward (r > 0, not merged) and its autograd backward, both weight expressions."""
    out = {}
    torch.manual_seed(6)
    R, C, T = 40, 72, 9
    for tag, dtype in (("bf16", torch.bfloat16), ("f16", torch.float16), ("f32", torch.float32)):
        for r in (2, 8):
            for sparse in (True, False):
                lin = ref.lora.Linear(C, R, r=r, lora_alpha=16, bias=False)
                lin.weight.data = (torch.randn(R, C) * 0.05).to(dtype)
                lin.lora_A.weight.data = torch.randn(r, C) * 0.1
                lin.lora_B.weight.data = torch.randn(R, r) * 0.1
                lin.mask = torch.rand(R, C) < 0.5
                lin.sparse = sparse
                key = f"{tag}_r{r}_{'sparse' if sparse else 'dense'}"
                out[f"{key}|W"] = f32(lin.weight)
                out[f"{key}|A"] = f32(lin.lora_A.weight)
                out[f"{key}|B"] = f32(lin.lora_B.weight)
                out[f"{key}|mask"] = lin.mask.numpy().copy()
                out[f"{key}|scaling"] = np.float64(lin.scaling)
                with torch.no_grad():                          # F.linear(I, W_eff) = W_eff^T exactly
                    out[f"{key}|W_eff"] = f32(lin(torch.eye(C, dtype=dtype)).T)
                x = (torch.randn(T, C) * 0.5).to(dtype).requires_grad_(True)
                gy = (torch.randn(T, R) * 0.5).to(dtype)
                y = lin(x)
                y.backward(gy)
                out[f"{key}|x"] = f32(x)
                out[f"{key}|gy"] = f32(gy)
                out[f"{key}|y"] = f32(y)
                out[f"{key}|G"] = f32(gy.t() @ x.detach())     # dL/dW_eff as F.linear's backward forms it
                out[f"{key}|dx"] = f32(x.grad)
                out[f"{key}|dA"] = f32(lin.lora_A.weight.grad)
                out[f"{key}|dB"] = f32(lin.lora_B.weight.grad)
    save("lora_forward.npz", **out)


def gen_sparsegpt(ref):
    """The reference's SparseGPT class on one linear: add_batch x3 then fasterprune, in three regimes."""
    out = {}
    cases = {
        # name: (R, C, tokens per call, weight dtype, act dtype, sparsity, n, m, dead channel)
        "unstr_bf16": (40, 256, 384, torch.bfloat16, torch.bfloat16, 0.5, 0, 0, False),
        "nm24_f16": (40, 256, 384, torch.float16, torch.float16, 0.0, 2, 4, False),
        "damped_bf16": (24, 256, 64, torch.bfloat16, torch.bfloat16, 0.5, 0, 0, False),     # 192 tokens < C: not PD
        "dead_f16": (24, 128, 256, torch.float16, torch.float16, 0.6, 0, 0, True),
        "ragged_f32": (16, 200, 512, torch.float32, torch.float16, 0.5, 0, 0, False),        # C % 128 != 0
    }
    for name, (R, C, T, wdt, adt, sp, n, m, dead) in cases.items():
        g = torch.Generator().manual_seed(sum(map(ord, name)))
        lin = nn.Linear(C, R, bias=False)
        lin.weight.data = (torch.randn(R, C, generator=g) * 0.05).to(wdt)
        sg = ref.sparsegpt.SparseGPT(lin)
        for i in range(3):
            x = act((1, T, C), 40 + i, adt, C)
            if dead:
                x[..., 5] = 0
                x[..., 77] = 0
            sg.add_batch(x, None)
            if name == "unstr_bf16":        # activations are kept for one case only (the others pin fasterprune on H)
                out[f"{name}|x{i}"] = packw(x)
        out[f"{name}|H"] = f32(sg.H)
        out[f"{name}|W_before"] = f32(lin.weight)
        out[f"{name}|tag"] = np.array(TAG[wdt])
        out[f"{name}|atag"] = np.array(TAG[adt])
        out[f"{name}|cfg"] = np.array([sp, n, m], dtype=np.float64)
        sg.fasterprune(sp, prune_n=n, prune_m=m, percdamp=0.01, blocksize=128)
        out[f"{name}|W_after"] = f32(lin.weight)
        out[f"{name}|importance_score"] = np.float64(lin.weight.importance_score)
    out["cases"] = np.array(list(cases))
    # ---- the +-inf clamps and the second damping loop (:101-109, :133-157), forced -------------------------------------
    def planted(name, plant, R=24, C=256, T=1024, wdt=torch.bfloat16):
        g = torch.Generator().manual_seed(sum(map(ord, name)))
        lin = nn.Linear(C, R, bias=False)
        lin.weight.data = (torch.randn(R, C, generator=g) * 0.05).to(wdt)
        sg = ref.sparsegpt.SparseGPT(lin)
        sg.add_batch(act((1, T, C), 41, wdt, C), None)
        plant(sg.H)
        out[f"{name}|H"] = f32(sg.H)
        out[f"{name}|W_before"] = f32(lin.weight)
        out[f"{name}|tag"] = np.array(TAG[wdt])
        out[f"{name}|cfg"] = np.array([0.5, 0, 0], dtype=np.float64)
        return lin, sg

    def plant_inf_H(H):           # +inf on the diagonal of an uncoupled channel, a symmetric pair of -inf: first clamp
        H[7, :] = 0
        H[:, 7] = 0
        H[7, 7] = float("inf")
        H[20, 33] = H[33, 20] = float("-inf")

    def plant_tiny_diag(H):       # an uncoupled channel with a denormal diagonal: 1 / it overflows -> +inf in H^-1, second clamp
        H[9, :] = 0
        H[:, 9] = 0
        H[9, 9] = 1e-39
    for name, plant in (("inf_H_bf16", plant_inf_H), ("inf_Hinv_bf16", plant_tiny_diag)):
        lin, sg = planted(name, plant)
        sg.fasterprune(0.5, prune_n=0, prune_m=0, percdamp=0.01, blocksize=128)
        out[f"{name}|W_after"] = f32(lin.weight)
        out[f"{name}|importance_score"] = np.float64(lin.weight.importance_score)
    # second damping loop: torch.cholesky_inverse is stubbed (test instrumentation of torch, the reference file is
    # unmodified) to hand back an INDEFINITE inverse: the true one shifted down so that exactly 3 damping steps of
    # percdamp * mean|diag| are needed (2.5 steps of margin: robust against roundoff)
    name = "second_damp_bf16"
    lin, sg = planted(name, lambda H: None, C=128, T=512)
    real_inv = torch.cholesky_inverse
    injected = {}

    def fake_inverse(L, *a, **k):
        Hinv = real_inv(L, *a, **k)
        lam = torch.linalg.eigvalsh(Hinv.double()).min().item()
        m = Hinv.diag().abs().mean().item()
        shift = (lam + 0.025 * m) / 1.025
        Hinv = Hinv - shift * torch.eye(Hinv.shape[0])
        injected["Hinv"] = Hinv.clone()
        return Hinv
    torch.cholesky_inverse = fake_inverse
    try:
        sg.fasterprune(0.5, prune_n=0, prune_m=0, percdamp=0.01, blocksize=128)
    finally:
        torch.cholesky_inverse = real_inv
    out[f"{name}|Hinv_injected"] = f32(injected["Hinv"])
    out[f"{name}|W_after"] = f32(lin.weight)
    out["forced_cases"] = np.array(["inf_H_bf16", "inf_Hinv_bf16", "second_damp_bf16"])
    save("sparsegpt.npz", **out)


def gen_sparsegpt_4096(ref):
    """the reference's fasterprune at a benched shape, 4096 x 4096 bf16, 50 % unstructured (CPU,
    ~2 minutes).  The inputs are NOT stored: H is an exact integer matrix times a power of two and W comes from a numpy
    PCG64 stream, so the GPU test regenerates both bit for bit from `seed` (tests/test_gpu_parity.py::_exact_hessian,
    _golden_weights; H_sum is the check).  Stored: the full keep mask (bit-packed, 2 MiB), 96 rows of the pruned weights
    and the norm of every pruned row."""
    seed, R, C, T = 4096, 4096, 4096, 8192
    rng = np.random.default_rng(seed)
    gain = rng.integers(1, 5, size=C)
    x = (rng.integers(-2, 3, size=(T, C)) * gain + rng.integers(-1, 2, size=C)).astype(np.float32)
    xt = torch.from_numpy(x)
    H = (xt.double().t() @ xt.double()).float() * (2.0 / T)
    rng = np.random.default_rng(seed + 1)
    W = torch.from_numpy((rng.standard_normal((R, C)) * 0.02).astype(np.float32)).to(torch.bfloat16)
    lin = nn.Linear(C, R, bias=False)
    lin.weight.data = W.clone()
    sg = ref.sparsegpt.SparseGPT(lin)
    sg.H = H.clone()
    sg.nsamples = 1
    sg.fasterprune(0.5, prune_n=0, prune_m=0, percdamp=0.01, blocksize=128)
    out = lin.weight.data
    rows = np.arange(0, R, R // 96)[:96].astype(np.int32)
    save("sparsegpt_4096.npz", seed=np.int64(seed), H_sum=np.float64(H.double().sum().item()),
         mask=np.packbits((out != 0).numpy(), axis=1), rows=rows, W_rows=packw(out[rows.astype(np.int64)]),
         row_norms=out.float().norm(dim=1).numpy().astype(np.float32),
         importance_score=np.float64(lin.weight.importance_score))


SLOW = set()        # generators that only run when named on the command line (none at present)


def gen_reorder(ref):
    doc = torch.tensor([[1., -2., 3.], [-2., 2., -4.], [5., 6., -7.], [-6., -7., -4.]])
    g = torch.Generator().manual_seed(3)
    rnd = torch.randn(16, 37, generator=g)
    rnd[rnd.abs() < 0.2] = 0.0
    save("reorder.npz", doc_in=doc.numpy(), doc_out=ref.dsnot.return_reorder_indice(doc).numpy(),
         rnd_in=rnd.numpy(), rnd_out=ref.dsnot.return_reorder_indice(rnd).numpy())


class _AllocModel(nn.Module):
    """Parameters named like the reference's check() expects (wanda_pruner.py:875-885): two sub-models with blocks."""

    def __init__(self, seed=0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)

        def lin(o, i):
            l = nn.Linear(i, o, bias=False)
            l.weight.data = torch.randn(o, i, generator=g) * 0.05
            return l
        self.t5_model = nn.ModuleDict({"encoder": nn.ModuleDict({"block": nn.ModuleList(
            [nn.ModuleDict({"q": lin(24, 24), "wi": lin(60, 24), "wo": lin(24, 60)}) for _ in range(3)])})})
        self.visual_encoder = nn.ModuleDict({"blocks": nn.ModuleList(
            [nn.ModuleDict({"qkv": lin(48, 16), "fc1": lin(40, 16)}) for _ in range(2)])})

    def forward(self, samples):
        x = samples["x"]
        h = x
        for b in self.t5_model["encoder"]["block"]:
            h = h + torch.tanh(b["wo"](torch.relu(b["wi"](b["q"](h)))))
        v = samples["v"]
        acc = 0
        for b in self.visual_encoder["blocks"]:
            acc = acc + b["qkv"](v).pow(2).mean() + b["fc1"](v).abs().mean()
        return {"loss": h.pow(2).mean() + acc}


def gen_layer_sparsity(ref):
    """LayerSparsity.get_mask / get_layerwise_mask / compute_importance_scores / return_sparsity of the
    UNMODIFIED reference (layer_single_base_pruner.py:111-475) on seeded inputs."""
    LS = ref.base.LayerSparsity
    out = {}
    g = torch.Generator().manual_seed(11)
    shapes = {"a.weight": (48, 64), "b.weight": (33, 17), "c.weight": (128, 96), "d.weight": (5, 7)}

    def fresh_scores(kind):
        gg = torch.Generator().manual_seed({"obd": 21, "ties": 22, "signed": 23}[kind])
        sc = {}
        for k, shp in shapes.items():
            if kind == "obd":        # w^2 * g^2-like: non-negative, wide dynamic range, different scale per tensor
                t = (torch.randn(shp, generator=gg) ** 2) * (torch.randn(shp, generator=gg) ** 2) * (10.0 ** (len(sc) - 2))
            elif kind == "ties":     # heavily quantised scores: the threshold value is shared by many entries
                t = torch.randint(0, 6, shp, generator=gg).float() * 0.25
            else:                    # signed scores with zeros and negative zeros (blipt5_mag_pruner feeds signed weights)
                t = torch.randn(shp, generator=gg)
                t[torch.rand(shp, generator=gg) < 0.1] = 0.0
                t[torch.rand(shp, generator=gg) < 0.05] = -0.0
            sc[k] = t.float().contiguous()
        return sc

    dummy = LS(None, None, None, 1, 0.5, 0.8, "obd_avg")
    cases = []
    for kind in ("obd", "ties", "signed"):
        base = fresh_scores(kind)
        for k, t in base.items():
            out[f"scores|{kind}|{k}"] = t.numpy().copy()
        for p, ms in ((0.5, 0.8), (0.3, 1.0), (0.6, 0.7), (0.05, 0.5)):
            sc = {k: t.clone() for k, t in base.items()}
            masks = dummy.get_mask(sc, p, ms)
            tag = f"{kind}|{p}|{ms}"
            cases.append(tag)
            for k in sc:
                out[f"global|{tag}|{k}"] = np.packbits(masks[k].numpy().astype(bool).ravel())
                changed = sc[k] != base[k]                    # entries get_mask overwrote with finfo.max (:160)
                assert bool((sc[k][changed] == torch.finfo(torch.float32).max).all())
                out[f"global_protected|{tag}|{k}"] = np.packbits(changed.numpy().ravel())
        saved_cuda = torch.Tensor.cuda
        torch.Tensor.cuda = lambda self, *a, **k: self          # get_layerwise_mask calls .cuda() (:184)
        try:
            for p in (0.5, 0.25):
                sc = {k: t.clone() for k, t in base.items()}
                masks = dummy.get_layerwise_mask(sc, p)
                for k in sc:
                    out[f"layerwise|{kind}|{p}|{k}"] = np.packbits(masks[k].numpy().astype(bool).ravel())
        finally:
            torch.Tensor.cuda = saved_cuda
    out["global_cases"] = np.array(cases)
    out["names"] = np.array(list(shapes))

    # compute_importance_scores + return_sparsity through the toy model
    model = _AllocModel(seed=5)
    gd = torch.Generator().manual_seed(6)
    loader = [{"x": torch.randn(4, 24, generator=gd), "v": torch.randn(4, 16, generator=gd), "text_input": ["t"] * 4}
              for _ in range(3)]
    loss_func = lambda m, d, cuda_enabled: (m(d)["loss"], len(d["text_input"]))    # utils.py:21-31 without prepare_sample
    names = [k for k, v in model.named_parameters()]
    out["model_names"] = np.array(names)
    for k, v in model.named_parameters():
        out[f"param|{k}"] = f32(v)
    for bi, d in enumerate(loader):
        grads = torch.autograd.grad(model(d)["loss"], list(model.parameters()))
        for k, gr in zip(names, grads):
            out[f"grad|{bi}|{k}"] = f32(gr)
    alloc_cases = []
    for method in ("obd_avg", "aobd_avg", "gradient_avg", "obd_sum"):
        for gran in ("layer", "block", "model"):
            for sparsity, ms in ((0.5, 0.8), (0.6, 0.7)):
                def group_of(name):
                    if gran == "layer":
                        return name
                    if name.startswith("t5_model"):
                        return "t5_model" if gran == "model" else ".".join(name.split(".")[:4])
                    return "visual_encoder" if gran == "model" else ".".join(name.split(".")[:3])
                mapping = {k: group_of(k) for k in names}
                ls = LS(model, loader, loss_func, 12, sparsity, ms, method, 1, 1e-3, mapping)
                res = ls.return_sparsity()
                tag = f"{method}|{gran}|{sparsity}|{ms}"
                alloc_cases.append(tag)
                out[f"alloc|{tag}"] = np.array([res[k] for k in names], dtype=np.float64)
                if gran == "layer" and sparsity == 0.5:
                    for k in names:
                        out[f"importance|{method}|{k}"] = f32(ls.importance_measure[k])
    # per-model allocation (prune_per_model)
    mapping = {k: ".".join(k.split(".")[:4]) if k.startswith("t5_model") else ".".join(k.split(".")[:3]) for k in names}
    ls = LS(model, loader, loss_func, 12, 0.5, 0.8, "obd_avg", 1, 1e-3, mapping, prune_per_model=True,
            per_model_group=["t5_model", "visual_encoder"], per_model_sparsity=[0.6, 0.4])
    res = ls.return_sparsity()
    out["alloc_per_model"] = np.array([res[k] for k in names], dtype=np.float64)
    out["alloc_cases"] = np.array(alloc_cases)

    # "real" score_compute: global_iterative_pruning with 3 iterations, sparsity of every parameter afterwards
    # (return_sparsity takes this branch for score_compute "real*", :246-249, but compute_importance_scores then has no
    # rule to apply, :455 / :466; the loop is driven directly with the "obd" rule)
    ls = LS(model, loader, loss_func, 12, 0.5, 0.8, "obd_avg", 1, 1e-3, {k: k for k in names})
    real = ls.global_iterative_pruning(0.5, {k: k for k in names}, iteratation=3, max_sparsity_per_layer=1.0)
    out["real_sparsity"] = np.array([real[k] for k in names], dtype=np.float64)
    save("layer_sparsity.npz", **out)


def main():
    torch.set_num_threads(8)
    ref = ref_loader.load()
    global CHECK
    CHECK = "--check" in sys.argv[1:]
    only = set(a for a in sys.argv[1:] if not a.startswith("--"))
    gens = dict(wanda_stats=gen_wanda_stats, dsnot_stats=gen_dsnot_stats, wanda_toy=gen_wanda_toy,
                dsnot_toy=gen_dsnot_toy, lora_merge=gen_lora_merge, lora_forward=gen_lora_forward, sparsegpt=gen_sparsegpt,
                reorder=gen_reorder, layer_sparsity=gen_layer_sparsity, sparsegpt_4096=gen_sparsegpt_4096)
    for name, fn in gens.items():
        if (not only and name not in SLOW) or name in only:
            fn(ref)
    if MISMATCH:
        sys.exit(f"fixtures differ from the generator: {MISMATCH}")


if __name__ == "__main__":
    main()
