"""CPU: the pieces behind bench.py's `--impl reference` / `cpu_baseline` legs.

* oracle/cpu_port.py (the fallback when the reference files are absent) replayed against the committed fixtures the
  unmodified reference produced - the pin its docstring claims.
* baseline/_ref (oracle/fetch_ref.py) is a byte-identical copy of the reference files it names.
* oracle/ref_arm.py drives the reference's own composite pruner on the no-GEMM stand-in block and the result equals the
  numpy oracle on the same statistics.
* tests/golden/make_golden.py --check: the committed fixtures are what the committed generator produces today.
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import golden_util as gu
from oracle import cpu_port, fetch_ref, oracle, ref_loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs_ref = pytest.mark.skipif(not ref_loader.available(), reason="no reference tree (neither /root/reference nor baseline/_ref)")


def rel_inf(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("tag", ["bf16", "f16", "f32"])
def test_cpu_port_wanda_stats_vs_golden(tag):
    g = gu.load("wanda_stats.npz")
    st = cpu_port.WandaStat(96)
    for i in range(4):
        st.add_batch(gu.to_torch(g[f"{tag}_x{i}"], tag))
        assert st.nsamples == int(g[f"{tag}_n{i}"])
        assert rel_inf(st.scaler_row.numpy(), g[f"{tag}_scaler{i}"]) < 1e-6


@pytest.mark.parametrize("tag", ["bf16", "f32"])
def test_cpu_port_dsnot_stats_vs_golden(tag):
    g = gu.load("dsnot_stats.npz")
    st = cpu_port.DSnoTStat(96)
    for i in range(4):
        st.add_batch(gu.to_torch(g[f"{tag}_x{i}"], tag))
        assert st.nsamples == int(g[f"{tag}_n{i}"]) and st.ntokens == int(g[f"{tag}_ntok{i}"])
        for k in ("scaler_row", "sum_metric_row", "mean", "var"):
            assert rel_inf(getattr(st, k).reshape(-1).numpy(), g[f"{tag}_{k}{i}"]) < 2e-6, (k, i)


def test_cpu_port_wanda_select_vs_golden():
    g = gu.load("wanda_toy_unstructured.npz")
    for key in g["layers"]:
        L = gu.layer(g, key)
        W = gu.to_torch(L["W_before"], L["tag"])
        s = torch.from_numpy(L["scaler_row"])
        vit = key.startswith("visual_encoder")
        keep, score = cpu_port.wanda_select(W, s, 0.5 if vit else 1 - 0.4, whole_matrix=vit)
        assert np.array_equal(keep.numpy(), L["mask"]), key
        assert np.array_equal(W.float().numpy(), L["W_after"]), key
        assert abs(score - float(L["importance_score"])) <= 2e-6 * abs(score), key
    g = gu.load("wanda_toy_2of4.npz")
    for key in g["layers"]:
        L = gu.layer(g, key)
        keep, _ = cpu_port.wanda_select(gu.to_torch(L["W_before"], L["tag"]), torch.from_numpy(L["scaler_row"]), 0.5, 2, 4)
        assert np.array_equal(keep.numpy(), L["mask"]), key      # same torch.topk, same ties


def test_cpu_port_lora_merge_vs_oracle():
    rng = np.random.default_rng(3)
    W = torch.from_numpy(rng.standard_normal((24, 64)).astype(np.float32) * 0.05).half()
    A = torch.from_numpy(rng.standard_normal((4, 64)).astype(np.float32) * 0.1)
    B = torch.from_numpy(rng.standard_normal((24, 4)).astype(np.float32) * 0.1)
    M = torch.from_numpy(rng.random((24, 64)) < 0.5)
    want = oracle.sparselora_merge(W.float().numpy(), "f16", A.numpy(), B.numpy(), 2.0, M.numpy(), True)
    got = cpu_port.lora_merge(W.clone(), A, B, 2.0, M)
    assert np.array_equal(got.float().numpy(), want)


@pytest.mark.parametrize("case", ["unstr_bf16", "nm24_f16", "damped_bf16", "dead_f16", "ragged_f32"])
def test_cpu_port_sparsegpt_vs_golden(case):
    g = gu.load("sparsegpt.npz")
    tag = str(g[f"{case}|tag"])
    sp, n, m = g[f"{case}|cfg"]
    p = cpu_port.SparseGPTPort(gu.to_torch(g[f"{case}|W_before"], tag))
    p.H = torch.from_numpy(g[f"{case}|H"].copy())
    p.fasterprune(float(sp), int(n), int(m))
    got, want = p.W.float().numpy(), g[f"{case}|W_after"]
    assert ((got == 0) == (want == 0)).mean() >= 0.999
    assert np.linalg.norm(got - want) / np.linalg.norm(want) < 1e-3


def test_ref_copy_is_verbatim():
    """baseline/_ref (when built) holds byte-identical copies; with /root/reference present the digests are re-derived."""
    if not os.path.isfile(os.path.join(fetch_ref.DST, "MANIFEST.json")):
        pytest.skip("baseline/_ref not built (python oracle/fetch_ref.py)")
    assert fetch_ref.verify()
    if os.path.isfile(os.path.join(fetch_ref.SRC, fetch_ref.FILES[0])):
        for rel in fetch_ref.FILES:
            assert fetch_ref._sha(os.path.join(fetch_ref.SRC, rel)) == fetch_ref._sha(os.path.join(fetch_ref.DST, rel)), rel


@needs_ref
@pytest.mark.parametrize("nm", [(0, 0), (2, 4)])
def test_ref_arm_stand_in_matches_oracle(nm):
    """The no-GEMM stand-in really runs the reference's statistics + selection: masks equal the oracle's on the same data."""
    from oracle import ref_arm
    linears = [("self_attn.q_proj", 24, 64, "attn_in"), ("self_attn.o_proj", 16, 64, "attn_out"),
               ("mlp.down_proj", 24, 96, "mlp_mid")]
    n_seq, seq = 6, 40
    before = {}
    _, model = ref_arm.run_composite("wanda", linears, n_seq, seq, torch.bfloat16, 0.5, *nm, distinct=4)
    blk = model.llm_model.model.layers[0]
    ref_blk = ref_arm.NoGemmBlock(linears, seq, torch.bfloat16, distinct=4)      # same seeds: same weights and activations
    loader = ref_arm.calib_loader(n_seq, seq, 64, torch.bfloat16, 4)
    for name, R, C, inp in linears:
        W0 = ref_blk.linear(name).weight.data.float().numpy()
        s, n = np.zeros(C, np.float32), 0
        for j in range(n_seq):
            x = loader[j]["x"] if inp == "attn_in" else ref_blk.acts[inp][j % 4]
            s, n = oracle.wanda_add_batch(s, n, x[0].float().numpy(), 1)
        if nm[0]:
            keep, Wp, _ = oracle.wanda_nm(W0, s, *nm)
            S = oracle.wanda_scores(W0, s).reshape(R, -1, nm[1])
            tie_free = np.array([[len(set(r.tolist())) == nm[1] for r in grp] for grp in S])
            got = blk.linear(name).mask.numpy().reshape(S.shape)
            assert np.array_equal(got[tie_free], keep.reshape(S.shape)[tie_free]), name
        else:
            keep, Wp, _ = oracle.wanda_rowselect(W0, s, int(C * 0.5))
            assert np.array_equal(blk.linear(name).mask.numpy(), keep), name
            assert np.array_equal(blk.linear(name).weight.data.float().numpy(), Wp), name


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="the generator needs the reference tree")
def test_committed_goldens_are_reproducible():
    """`make_golden.py --check` regenerates every fixture in memory and compares it with the committed file."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "golden", "make_golden.py"), "--check"],
                       capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
