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
