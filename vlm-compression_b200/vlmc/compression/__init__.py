"""load_pruner: same contract as lavis/compression/__init__.py:29-46 (OmegaConf is optional here)."""
from vlmc.common.registry import registry
from vlmc.compression.pruners.layer_single_base_pruner import BasePruner  # noqa: F401
from vlmc.compression.pruners.wanda_pruner import BLIPT5LayerWandaPruner  # noqa: F401
from vlmc.compression.pruners.sparsegpt_pruner import BLIPT5LayerSparseGPTPruner  # noqa: F401
from vlmc.compression.pruners.dsnot_pruner import BLIPT5LayerDSnoTPruner  # noqa: F401
from vlmc.compression.pruners.global_pruner import BLIPT5MagPruner, BLIPT5RandPruner, BLIPT5AOBDPruner  # noqa: F401

__all__ = ["BasePruner", "load_pruner"]


def load_pruner(name, model, data_loader, cfg_path=None, cfg=None):
    if cfg_path is not None:
        import yaml
        with open(cfg_path) as f:
            cfg = yaml.safe_load(f)
    try:
        pruner = registry.get_pruner_class(name)(model=model, data_loader=data_loader, **(cfg or {}))
    except TypeError:
        # the reference prints and exits (lavis/compression/__init__.py:39-44)
        print(f"Pruner {name} not found. Available pruners:\n" + ", ".join(registry.list_pruners()))
        exit(1)
    return pruner
