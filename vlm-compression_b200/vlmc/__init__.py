"""vlmc - B200-native calibration-and-masking path of VLM-Compression (Wanda / SparseGPT / DSnoT / SparseLoRA).

Mirrors the reference's module layout for this path only:
    vlmc.compression.load_pruner            <- lavis/compression/__init__.py:29-46
    vlmc.compression.pruners.wanda_pruner   <- lavis/compression/pruners/wanda_pruner.py
    vlmc.compression.pruners.sparsegpt_pruner, dsnot_pruner
    vlmc.peft.lora                          <- lavis/peft/src/peft/tuners/lora.py (Linear: mask / sparse / merge)
    vlmc.common.registry                    <- lavis/common/registry.py (pruner registry only)
Every statistic, selection and update runs in libvlmc.so (include/vlmc.h); see vlmc.native.
"""
__version__ = "0.1.0"
