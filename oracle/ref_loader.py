"""TEST INFRASTRUCTURE ONLY - loads the *unmodified* reference hot-path modules by file path.

The reference package (`lavis`) is not importable as a package in this image (omegaconf is
missing and `lavis/datasets/data_utils.py` is absent from the tree), but the seven files on
the calibration-and-masking path only need torch + transformers.Conv1D once four names are
stubbed.  This loader is used (a) by `tests/golden/make_golden.py` to generate the committed
fixtures and (b) by the CPU tests, which skip when no reference tree is found, and (c) by bench.py's
`--impl reference` / `cpu_baseline` legs and the live GPU differential tests, which read the verbatim copy
under `baseline/_ref` (oracle/fetch_ref.py) on the GPU box.  Nothing under `vlmc/` may import this file.

Stubs (and the reference line that needs each):
  lavis.common.registry.registry.register_pruner   wanda_pruner.py:84 (decorator)
  lavis.datasets.data_utils.prepare_sample         pruners/utils.py:3 (file missing upstream)
  lora.Linear8bitLt                                wanda_pruner.py:13 (only defined with bitsandbytes)
  torch.cuda.synchronize (CPU-only hosts)          sparsegpt_pruner.py:212
"""
import importlib
import importlib.util
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


def _resolve_root():
    """VLMC_REFERENCE_ROOT, else the reference tree of the build container, else the verbatim copy that
    oracle/fetch_ref.py leaves under baseline/_ref (git-ignored; it travels to the GPU box with the snapshot)."""
    env = os.environ.get("VLMC_REFERENCE_ROOT")
    if env:
        return env
    for cand in ("/root/reference", os.path.join(os.path.dirname(_HERE), "baseline", "_ref")):
        if os.path.isfile(os.path.join(cand, "lavis/compression/pruners/wanda_pruner.py")):
            return cand
    return "/root/reference"


REF_ROOT = _resolve_root()

_PRUNER_FILES = ("utils", "base_pruner", "layer_single_base_pruner",
                 "wanda_pruner", "sparsegpt_pruner", "dsnot_pruner", "global_pruner")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "lavis/compression/pruners/wanda_pruner.py"))


class _PassThroughRegistry:
    """Stands in for lavis.common.registry.registry: decorators return the class unchanged."""
    table = {}

    @classmethod
    def register_pruner(cls, name):
        def deco(klass):
            cls.table[name] = klass
            return klass
        return deco

    @classmethod
    def get_pruner_class(cls, name):
        return cls.table.get(name)


def _empty_pkg(name, path=None, **attrs):
    mod = types.ModuleType(name)
    mod.__path__ = [] if path is None else [path]
    for k, v in attrs.items():
        setattr(mod, k, v)
    sys.modules[name] = mod
    return mod


_cache = None


def load(patch_source=None):
    """Return a namespace with .wanda, .sparsegpt, .dsnot, .global_pruner, .lora, .base (reference modules).

    patch_source: optional {module_short_name: callable(str)->str} applied to the file text
    before exec (used only to excise dsnot_pruner.py:734-740 for the upstream-semantics test).
    A patched load is never cached.
    """
    global _cache
    if patch_source is None and _cache is not None:
        return _cache
    if not available():
        raise FileNotFoundError(f"reference tree not found under {REF_ROOT}")
    import torch

    saved = {k: v for k, v in sys.modules.items() if k == "lavis" or k.startswith("lavis.")}
    for k in saved:
        del sys.modules[k]
    try:
        for pkg in ("lavis", "lavis.common", "lavis.datasets", "lavis.compression",
                    "lavis.compression.pruners", "lavis.peft", "lavis.peft.src"):
            _empty_pkg(pkg)
        _empty_pkg("lavis.common.registry", registry=_PassThroughRegistry)
        _empty_pkg("lavis.datasets.data_utils", prepare_sample=lambda s, cuda_enabled=True: s)
        _empty_pkg("lavis.peft.src.peft", path=f"{REF_ROOT}/lavis/peft/src/peft")
        _empty_pkg("lavis.peft.src.peft.tuners", path=f"{REF_ROOT}/lavis/peft/src/peft/tuners")
        lora = importlib.import_module("lavis.peft.src.peft.tuners.lora")
        if not hasattr(lora, "Linear8bitLt"):
            lora.Linear8bitLt = type("Linear8bitLt", (), {})
        if not torch.cuda.is_available():
            torch.cuda.synchronize = lambda *a, **k: None

        mods = {}
        for short in _PRUNER_FILES:
            full = f"lavis.compression.pruners.{short}"
            path = f"{REF_ROOT}/lavis/compression/pruners/{short}.py"
            if patch_source and short in patch_source:
                with open(path) as f:
                    text = patch_source[short](f.read())
                mod = types.ModuleType(full)
                mod.__file__ = path
                sys.modules[full] = mod
                exec(compile(text, path, "exec"), mod.__dict__)
            else:
                spec = importlib.util.spec_from_file_location(full, path)
                mod = importlib.util.module_from_spec(spec)
                sys.modules[full] = mod
                spec.loader.exec_module(mod)
            mods[short] = mod
        ns = types.SimpleNamespace(
            wanda=mods["wanda_pruner"], sparsegpt=mods["sparsegpt_pruner"], dsnot=mods["dsnot_pruner"],
            base=mods["layer_single_base_pruner"], global_pruner=mods["global_pruner"], lora=lora,
            registry=_PassThroughRegistry)
    finally:
        # leave sys.modules without our fake 'lavis' so nothing else resolves it by accident
        for k in [k for k in sys.modules if k == "lavis" or k.startswith("lavis.")]:
            del sys.modules[k]
        sys.modules.update(saved)
    if patch_source is None:
        _cache = ns
    return ns
