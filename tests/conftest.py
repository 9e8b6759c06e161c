import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vlm-compression_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """Without a CUDA device the gpu-marked tests are skipped instead of erroring in their fixtures (a plain `pytest tests`
    is green on a CPU box).  VLMC_REQUIRE_GPU=1 keeps them strict: on the B200 box a missing device must fail loudly."""
    if os.environ.get("VLMC_REQUIRE_GPU") == "1":
        return
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (B200 box: pytest -m gpu)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def built_lib():
    """Path of libvlmc.so, building it with nvcc when it is missing or stale (cross-compiles without a GPU)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("vlmc_build", os.path.join(ROOT, "vlm-compression_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    if os.path.exists(mod.OUT) and not os.path.exists("/usr/local/cuda/bin/nvcc"):
        return mod.OUT
    return mod.build()
