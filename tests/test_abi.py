"""CPU: the C-ABI library builds (nvcc cross-compiles sm_100a), loads, and exports every symbol
include/vlmc.h declares.  No compute calls here (no GPU)."""
import ctypes
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "vlmc.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vlmc_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_hot_path():
    syms = declared_symbols()
    for s in ("vlmc_sqnorm_accum", "vlmc_dsnot_stats", "vlmc_wanda_rowselect", "vlmc_wanda_nm",
              "vlmc_wanda_threshold", "vlmc_sparselora_merge", "vlmc_version", "vlmc_workspace_bytes"):
        assert s in syms


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in include/vlmc.h but not exported: {missing}"


def test_binding_table_matches_header(built_lib):
    from vlmc import native
    assert sorted(native.SIGNATURES) == declared_symbols()
    lib = native.load()
    assert lib.vlmc_version() == 1
    assert lib.vlmc_status_string(-3).decode().startswith("pointer is not CUDA")
    assert lib.vlmc_workspace_bytes(native.OP_SQNORM, 2048, 4096, 0) > native.WS_COUNTER_BYTES


def test_library_is_sm100a_only(built_lib):
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "--list-elf", built_lib], capture_output=True, text=True)
    if out.returncode != 0:
        import pytest
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs
