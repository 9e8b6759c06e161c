"""CPU: the reference arm of bench.py (`--impl reference`) runs without a GPU and prints the contract's JSON line; ranks
other than 0 exit 0 without work (the driver launches it under torchrun for N > 1)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra, *args):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, env=env, capture_output=True,
                          text=True, timeout=600)


import pytest


@pytest.mark.parametrize("kind", ["reference", "port"])
def test_reference_arm_prints_the_contract_line(kind):
    """kind "reference": the unmodified reference files (here /root/reference, on the GPU box baseline/_ref) run the whole
    block; kind "port": they are absent (VLMC_REFERENCE_ROOT points nowhere) and oracle/cpu_port.py times a bounded sample."""
    sys.path.insert(0, ROOT)
    from oracle import ref_loader
    if kind == "reference" and not ref_loader.available():
        pytest.skip("no reference tree")
    r = _run({} if kind == "reference" else {"VLMC_REFERENCE_ROOT": "/nonexistent"}, "--impl", "reference", "--steps", "1",
             "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["higher_is_better"] is False and line["vs_baseline"] is None
    assert line["metric"] == "s_per_vicuna7b_block_pruned" and line["unit"] == "s/block" and line["value"] > 0
    assert abs(line["ms_per_step"] - 1e3 * line["value"]) < 1e-6 * line["ms_per_step"]
    assert line["steps"] == 1 and line["warmup"] == 0 and line["n_gpus"] == 1 and line["data"] == "synthetic"
    assert "workload" in line["config"] and "model" not in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == kind and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--impl", "reference", "--gpus", "2", "--steps", "1",
             "--warmup", "0")
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_statistics_roofline_is_quoted_on_distinct_bytes_when_batched(monkeypatch):
    """SURVEY 8(d) asks to say which byte count the roofline uses: with the block's statistics as ONE launch the linears fed
    the same activations share them through L2, so `achieved` / `frac` count the four distinct tensors (12.21 of the 18.66 GB)
    and the per-linear figure sits beside them; with one launch per linear, or shared inputs, nothing is rescaled."""
    sys.path.insert(0, ROOT)
    import bench
    pk = {"hbm": 6545.0, "tensor": 1688.0, "tensor_sustained": 1408.0, "source": "test"}

    def res(world=1, calib_batch=bench.N_SEQ, shared=False):
        return {"method": "wanda_nm", "steps": 10, "eager_ms_per_step": 2.05, "shared": shared, "world": world, "calib_batch": calib_batch,
                "kernels": {"sqnorm_accum": {"ms": 18.3, "work": 18.66e9 * 10, "spans": 10},
                            "wanda_select": {"ms": 1.86, "work": 1.012e9 * 10, "spans": 10}}}
    monkeypatch.delenv("VLMC_BENCH_STATS_BATCH", raising=False)
    r = bench.roofline_of(res(), pk)
    distinct = (3 * 4096 + 11008) / (6 * 4096 + 11008)
    assert abs(r["frac_per_linear_bytes"] - 18.66e9 / 1.83e-3 / 1e9 / 6545.0) < 1e-9
    assert abs(r["frac"] - r["frac_per_linear_bytes"] * distinct) < 1e-12 and r["frac"] < 1.1 and "bytes_basis" in r
    assert bench.roofline_of(res(world=2), pk)["frac"] == r["frac"]                       # batched at every N > 1
    for plain in (res(calib_batch=16), res(shared=True)):
        q = bench.roofline_of(plain, pk)
        assert "frac_per_linear_bytes" not in q and abs(q["frac"] - r["frac_per_linear_bytes"]) < 1e-9
    monkeypatch.setenv("VLMC_BENCH_STATS_BATCH", "0")
    assert "frac_per_linear_bytes" not in bench.roofline_of(res(), pk)
