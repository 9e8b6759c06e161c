"""SparseGPT pruner: same per-layer wrapper and entry point as lavis/compression/pruners/sparsegpt_pruner.py.

  SparseGPT.add_batch    <- :68-79    H accumulation on the tensor cores (vlmc_hessian_accum, K3)
  SparseGPT.fasterprune  <- :81-215   dead channels + conditional damping loop (host, like the reference) around
                                      vlmc_chol_inv_upper (K10) and vlmc_obs_sweep (K11-K13)
  fasterprune_block      <- :441-452  the reference's `for name in subset: fasterprune; free` loop: the chains of the
                                      block's linears are independent and run concurrently (vlmc.schedule), linears
                                      that share their input share one H and one factor
  BLIPT5LayerSparseGPTPruner <- :1005-1091, registered as "blipt5_sparsegpt_pruner"; also drives
                                      llm_model.model.layers, which the reference cannot
"""
import os

import torch
import torch.nn as nn

from vlmc import native, schedule
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
        # False: U comes from the fused factorisation (one blocked Cholesky + one triangular inverse), and the explicit
        # cholesky_inverse -> clamp -> cholesky(upper) stage of :131-157 runs only for a Hessian whose factor is flagged
        # (VLMC_HUGE_FACTOR).  True: always in the reference's three-step order.  Same U mathematically.
        self.exact_reference_order = os.environ.get("VLMC_SPARSEGPT_EXACT_ORDER") == "1"

    def add_batch(self, inp, out=None):
        if len(inp.shape) == 2:
            inp = inp.unsqueeze(0)
        b = inp.shape[0]
        N = getattr(self, "_global_n", None)
        if N is None:
            native.hessian_accum(inp, self.H, self.nsamples, b)
        else:
            # data-parallel calibration (layerwise.prune_blocks): this rank sees a share of the N samples; its partial sum
            # carries the divisor of the WHOLE set from the start - first call H = (2/N) X^T X, later calls
            # H += (2/N) X^T X - so the ranks' Hessians simply add (one all-reduce, no rescaling pass over C^2 floats)
            native.hessian_accum(inp, self.H, 0 if self.nsamples == 0 else N, N if self.nsamples == 0 else 0)
        self.nsamples += b

    def fasterprune(self, sparsity, prune_n=0, prune_m=0, blocksize=128, percdamp=.01):
        H = self.H
        del self.H
        group = getattr(self, "_shared", None)       # set when several linears adopted one H (layerwise.InputSharing)
        if group is not None and group.get("percdamp") == percdamp and group.get("H") is H:
            U, dead = group["U"], group["dead"]                     # the same H was factorised for a sibling linear
        else:
            damp, dead = native.hessian_prepare(H, percdamp)        # :95-96, :111
            U, status = native.chol_inv_upper(H)
            # :101-157: the +-inf clamps, the damp-only-after-a-failure loops (bounded, unlike the reference's `while
            # True`) and, when the factor is flagged or on request, the second stage in the reference's own order
            schedule.resolve_factor(H, U, status, damp, percdamp, exact_reference_order=self.exact_reference_order)
            if group is not None:
                group.update(H=H, U=U, dead=dead, percdamp=percdamp)
        _, score = native.obs_sweep(self.layer.weight.data, U, sparsity, prune_n, prune_m, dead=dead,
                                    blocksize=blocksize)
        setattr(self.layer.weight, "importance_score", score.item())
        torch.cuda.synchronize()                                    # :212

    def free(self):
        # :217-219.  The reference also calls torch.cuda.empty_cache() here; that hands the H / U blocks (up to 0.5 GB
        # each) back to the driver after EVERY linear, a synchronous cudaFree measured at 5-485 ms per call, only for the
        # next linear to cudaMalloc them again.  The caching allocator reuses them instead; the per-model
        # empty_cache() at the end of the block loop (layerwise.prune_blocks) is kept.
        self.H = None
        self._shared = None


def fasterprune_block(wrappers, sparsities, prune_n=0, prune_m=0, blocksize=128, percdamp=.01):
    """fasterprune for ALL linears of one block (the reference's loop at :441-452), scheduled as concurrent chains:
    every distinct H is prepared and factorised on its own stream (one host sync for all status words, damping only
    for the ones that failed, :114-128), then every OBS sweep runs on its own stream.  Per linear the kernels and their
    order are exactly those of SparseGPT.fasterprune, so the weights are bit-identical to calling it one by one."""
    items = []
    for w, sp in zip(wrappers, sparsities):
        items.append((w.layer.weight.data, w.H, sp, prune_n, prune_m))
    scores, _ = schedule.sparsegpt_block(items, percdamp, blocksize,
                                         exact_reference_order=any(getattr(w, "exact_reference_order", False) for w in wrappers))
    for w, v in zip(wrappers, scores.tolist()):
        setattr(w.layer.weight, "importance_score", v)
        del w.H
    torch.cuda.synchronize()                                        # :212


@registry.register_pruner("blipt5_sparsegpt_pruner")
class BLIPT5LayerSparseGPTPruner(BLIPT5LayerWandaPruner):
    pruner_name = "blipt5_sparsegpt_pruner"

    def make_wrapper(self, module):
        return SparseGPT(module)

    def _prune_linear(self, vit, lora_model):
        def fn(i, name, module, wrapper, sparsity, expected_nsamples):
            assert wrapper.nsamples == expected_nsamples
            # n:m runs never read the ratio (sparsegpt_pruner.py:176-187): a missing sparsity_dict entry arrives as None
            self._pending.append((wrapper, 0.0 if sparsity is None else sparsity))      # pruned together in finish_block
        return fn

    def finish_block(self, subset, wrapped):
        from vlmc import parallel
        from vlmc.compression.pruners.layerwise import _dist_info
        pending, self._pending = getattr(self, "_pending", []), []
        if pending:
            rank, world = _dist_info(self)
            if world > 1:
                # data-parallel run: every rank holds the merged Hessians; WHOLE linears are dealt to the ranks (longest
                # chain first), each rank runs its chains concurrently, the pruned weights are broadcast from their owners
                ws = [w for w, _ in pending]
                scores = torch.zeros(len(ws), dtype=torch.float32, device=ws[0].layer.weight.device)

                def run(indices):
                    fasterprune_block([ws[i] for i in indices], [pending[i][1] for i in indices], prune_n=self.prune_n,
                                      prune_m=self.prune_m, percdamp=0.01, blocksize=128)
                    for i in indices:
                        scores[i] = ws[i].layer.weight.importance_score
                parallel.prune_linears_task_parallel([w.layer.weight.data for w in ws], run, rank, world)
                parallel.allreduce_sum(scores)
                for w, v in zip(ws, scores.tolist()):
                    setattr(w.layer.weight, "importance_score", v)
                    w.H = None
            else:
                fasterprune_block([w for w, _ in pending], [s for _, s in pending], prune_n=self.prune_n,
                                  prune_m=self.prune_m, percdamp=0.01, blocksize=128)
            for w, _ in pending:
                w.free()

    def _prune(self, *args, **kwargs):
        self._pending = []
        return super()._prune(*args, **kwargs)

    def prune(self, importance_scores=None, keep_indices_or_masks=None):
        return super().prune(importance_scores, keep_indices_or_masks, lora_model=False)
