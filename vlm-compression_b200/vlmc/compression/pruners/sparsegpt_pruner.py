"""SparseGPT pruner: same per-layer wrapper and entry point as lavis/compression/pruners/sparsegpt_pruner.py.

  SparseGPT.add_batch    <- :68-79    H accumulation on the tensor cores (vlmc_hessian_accum, K3)
  SparseGPT.fasterprune  <- :81-215   dead channels + conditional damping loop (host, like the reference) around
                                      vlmc_chol_inv_upper (K10) and vlmc_obs_sweep (K11-K13)
  BLIPT5LayerSparseGPTPruner <- :1005-1091, registered as "blipt5_sparsegpt_pruner"; also drives
                                      llm_model.model.layers, which the reference cannot
"""
import torch
import torch.nn as nn

from vlmc import native
from vlmc.common.registry import registry
from vlmc.compression.pruners.wanda_pruner import BLIPT5LayerWandaPruner


class SparseGPT:
    def __init__(self, layer):
        self.layer = layer
        self.dev = self.layer.weight.device
        W = layer.weight.data
        if not isinstance(layer, nn.Linear):
            raise NotImplementedError("only nn.Linear (and LoRA Linear) layers are on the InstructBLIP path")
        self.rows = W.shape[0]
        self.columns = W.shape[1]
        self.H = torch.zeros((self.columns, self.columns), device=self.dev)
        self.nsamples = 0

    def add_batch(self, inp, out=None):
        if len(inp.shape) == 2:
            inp = inp.unsqueeze(0)
        b = inp.shape[0]
        native.hessian_accum(inp, self.H, self.nsamples, b)
        self.nsamples += b

    def fasterprune(self, sparsity, prune_n=0, prune_m=0, blocksize=128, percdamp=.01):
        H = self.H
        del self.H
        damp, dead = native.hessian_prepare(H, percdamp)            # :95-96, :111
        U = None
        while True:                                                 # :114-128: damp only after a failed attempt
            U, status = native.chol_inv_upper(H, U)
            if status.item() == 0:
                break
            native.hessian_add_damp(H, damp)
        _, score = native.obs_sweep(self.layer.weight.data, U, sparsity, prune_n, prune_m, dead=dead,
                                    blocksize=blocksize)
        setattr(self.layer.weight, "importance_score", score.item())
        torch.cuda.synchronize()                                    # :212

    def free(self):
        self.H = None
        torch.cuda.empty_cache()


@registry.register_pruner("blipt5_sparsegpt_pruner")
class BLIPT5LayerSparseGPTPruner(BLIPT5LayerWandaPruner):
    pruner_name = "blipt5_sparsegpt_pruner"

    def make_wrapper(self, module):
        return SparseGPT(module)

    def _prune_linear(self, vit, lora_model):
        def fn(i, name, module, wrapper, sparsity, expected_nsamples):
            assert wrapper.nsamples == expected_nsamples
            wrapper.fasterprune(sparsity, prune_n=self.prune_n, prune_m=self.prune_m, percdamp=0.01, blocksize=128)
            wrapper.free()
        return fn

    def prune(self, importance_scores=None, keep_indices_or_masks=None):
        return super().prune(importance_scores, keep_indices_or_masks, lora_model=False)
