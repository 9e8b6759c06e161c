"""DSnoT pruner: same per-layer wrapper, helper and entry point as lavis/compression/pruners/dsnot_pruner.py.

  WrappedGPT                <- :53-105     scaler_row / sum_metric_row / mean / var in one pass (vlmc_dsnot_stats, K2)
  return_reorder_indice     <- :1881-1925  kept for callers; the refine kernel never materialises this ordering
  dsnot_prune_linear        <- :359-755    initial mask + prune/regrow cycles (vlmc_dsnot_refine_walk/_apply, K8+K9)
  BLIPT5LayerDSnoTPruner    <- :1599-1863  registered as "blipt5_dsnot_pruner"
"""
import os

import torch

from vlmc import native
from vlmc.common.registry import registry
from vlmc.compression.pruners.wanda_pruner import BLIPT5LayerWandaPruner


class WrappedGPT:
    """Per-linear DSnoT statistics (dsnot_pruner.py:53-105); add_batch is one kernel launch."""

    def __init__(self, layer, initial_method="wanda", layer_id=0, layer_name="none"):
        if initial_method == "sparsegpt":
            raise NotImplementedError("initial_method='sparsegpt' is dead code in the reference (add_batch never "
                                      "accumulates H, dsnot_pruner.py:66-101): SURVEY F11")
        self.layer = layer
        self.dev = self.layer.weight.device
        self.rows = layer.weight.data.shape[0]
        self.columns = layer.weight.data.shape[1]
        self.nsamples = 0
        self.initial_method = initial_method
        self.scaler_row = torch.zeros((self.columns), device=self.dev)
        self.sum_metric_row = torch.zeros((self.columns), device=self.dev)
        self.mean = torch.zeros((self.columns), device=self.dev)
        self.var = torch.zeros((self.columns), device=self.dev)
        self.ntokens = 0
        self.layer_id = layer_id
        self.layer_name = layer_name

    def add_batch(self, inp, out=None):
        if len(inp.shape) == 2:
            inp = inp.unsqueeze(0)
        b = inp.shape[0]
        ntok = inp.numel() // inp.shape[-1]
        if self.mean.dim() == 1:      # the reference's mean / var become [C, 1] after the first call (:89-93)
            self.mean = self.mean.reshape(-1, 1)
            self.var = self.var.reshape(-1, 1)
        # a chunk of K stacked calibration samples (layerwise.stack_calibration) is K reference calls: var is the
        # token-weighted mean of per-CALL biased variances (:92), so the kernel treats the chunk as K segments
        calls = int(getattr(self, "_stacked_calls", 1) or 1)
        if calls > 1 and b % calls == 0:
            native.dsnot_stats(inp, self.scaler_row, self.sum_metric_row, self.mean, self.var, self.nsamples,
                               b // calls, self.ntokens, nseg=calls)
        else:
            native.dsnot_stats(inp, self.scaler_row, self.sum_metric_row, self.mean, self.var, self.nsamples, b,
                               self.ntokens)
        self.ntokens += ntok
        self.nsamples += b

    @staticmethod
    def add_batch_many(pairs):
        """[(wrapper, inp)] of one block forward -> ONE launch per run of equal (nsamples, batch, calls) (in practice one per
        block forward; vlmc_dsnot_stats_batch).  The same update as add_batch on each pair, bit for bit."""
        groups = {}
        for w, inp in pairs:
            if len(inp.shape) == 2:
                inp = inp.unsqueeze(0)
            calls = int(getattr(w, "_stacked_calls", 1) or 1)
            if not (calls > 1 and inp.shape[0] % calls == 0):
                calls = 1
            groups.setdefault((w.nsamples, inp.shape[0], calls, inp.device), []).append((w, inp))
        for (n, b, calls, _), items in groups.items():
            if len(items) == 1:
                items[0][0].add_batch(items[0][1])
                continue
            for w, _ in items:
                if w.mean.dim() == 1:
                    w.mean = w.mean.reshape(-1, 1)
                    w.var = w.var.reshape(-1, 1)
            native.dsnot_stats_batch([x for _, x in items], [(w.scaler_row, w.sum_metric_row, w.mean, w.var) for w, _ in items],
                                     n, b // calls, [w.ntokens for w, _ in items], nseg=calls)
            for w, x in items:
                w.ntokens += x.numel() // x.shape[-1]
                w.nsamples += b

    def free(self):
        self.H = None          # dsnot_pruner.py:103-105 (its empty_cache() is left to the end of the block loop, see SparseGPT.free)


def return_reorder_indice(input_tensor):
    """dsnot_pruner.py:1881-1925: positions of the negative entries in order, then of the positive entries
    reversed (right-aligned); slots in between (zeros in the input) hold index 0."""
    R, C = input_tensor.shape
    idx = torch.arange(C, device=input_tensor.device).expand(R, C)
    neg, pos = input_tensor < 0, input_tensor > 0
    big = torch.full_like(idx, C)
    neg_sorted = torch.sort(torch.where(neg, idx, big), dim=1)[0]
    pos_sorted = torch.flip(torch.sort(torch.where(pos, idx, big), dim=1)[0], dims=[1])
    neg_sorted = torch.where(neg_sorted == C, torch.zeros_like(neg_sorted), neg_sorted)
    pos_sorted = torch.where(pos_sorted == C, torch.zeros_like(pos_sorted), pos_sorted)
    return (neg_sorted + pos_sorted).to(torch.int64)


def dsnot_prune_linear(module, wrapper, sparsity, prune_n=0, prune_m=0, lora_model=False, initial_method="wanda",
                       pow_of_var_regrowing=1.0, max_cycle_time=100, update_threshold=0.1, without_same_sign=True,
                       without_DSnoT=False, ref_fixup=True, argmin_rule=1, reduce_ncycles=None, elide_noop_swaps=True):
    """One linear (dsnot_pruner.py:359-755).  Sets module.mask (True = kept), zeroes pruned weights unless lora_model.
    Returns the executed cycle count (1-elem int tensor) or None when nothing ran.

    elide_noop_swaps: with the shipped reference semantics (ref_fixup) every unstructured swap is written back by
    :734-740, and the candidates always come from the initial kept / pruned sets, so the final mask IS the initial
    selection (tests assert the equality on every fixture and at Vicuna size).  True (default) skips the cycle
    loop and runs the initial selection only - same mask, same weights, no cycle count: nothing the loop computes is
    observable in the shipped semantics.  False executes the loop (vlmc_dsnot_refine) exactly as the reference does; the
    upstream semantics (ref_fixup=False) and the n:m branch always execute it."""
    W = module.weight.data
    C = W.shape[1]
    if prune_n == 0:
        if sparsity == 0.:
            return None                                          # :560-561 `continue`: the layer is left untouched
        k = round(C * sparsity)                                  # :562 (python round)
        if without_DSnoT or (elide_noop_swaps and ref_fixup):    # :577-578: the initial mask only
            scal = wrapper.scaler_row if initial_method == "wanda" else torch.ones_like(wrapper.scaler_row)
            keep, _ = native.wanda_rowselect(W, scal, k, zero_w=not lora_model)
            setattr(module, "mask", keep)
            return None
    else:
        k = 0
    keep, ncyc = native.dsnot_refine(W, wrapper.scaler_row, wrapper.sum_metric_row, wrapper.var, k, prune_n, prune_m,
                                     pow_of_var=pow_of_var_regrowing, max_cycle_time=int(max_cycle_time),
                                     update_threshold=update_threshold, without_same_sign=without_same_sign,
                                     initial_method=initial_method, argmin_rule=argmin_rule, ref_fixup=ref_fixup,
                                     zero_w=not lora_model, reduce_ncycles=reduce_ncycles)
    setattr(module, "mask", keep)
    return ncyc


@registry.register_pruner("blipt5_dsnot_pruner")
class BLIPT5LayerDSnoTPruner(BLIPT5LayerWandaPruner):
    pruner_name = "blipt5_dsnot_pruner"

    def __init__(self, model, data_loader, initial_method="wanda", skip_layer=None, skip_sub_layer=None,
                 pow_of_var_regrowing=1., max_cycle_time=1e2, update_threshold=0.1, without_same_sign=True,
                 without_DSnoT=False, upstream_semantics=False, elide_noop_swaps=True, **kwargs):
        super().__init__(model, data_loader, **kwargs)
        self.pow_of_var_regrowing = pow_of_var_regrowing
        self.without_same_sign = without_same_sign
        self.without_DSnoT = without_DSnoT
        self.update_threshold = update_threshold
        self.skip_layer = skip_layer
        self.skip_sub_layer = skip_sub_layer
        self.max_cycle_time = max_cycle_time
        self.initial_method = initial_method
        # False (default): bit-for-bit the shipped reference, whose write-back block (:734-740) turns the unstructured
        # swaps into no-ops.  True: the upstream DSnoT behaviour (that block removed).
        self.upstream_semantics = upstream_semantics
        self.elide_noop_swaps = elide_noop_swaps          # see dsnot_prune_linear

    def make_wrapper(self, module):
        return WrappedGPT(module, initial_method=self.initial_method)

    def _prune_linear(self, vit, lora_model):
        def fn(i, name, module, wrapper, sparsity, expected_nsamples):
            assert wrapper.nsamples == expected_nsamples          # :360
            if (self.prune_n == 0 and sparsity != 0. and (self.without_DSnoT or (self.elide_noop_swaps and not self.upstream_semantics))
                    and os.environ.get("VLMC_ROWSELECT_BATCH") != "0"):
                # the initial selection IS the result (dsnot_prune_linear, elide_noop_swaps): the block's selections go
                # through ONE batched call in finish_block, with DSnoT's rounded prune count (:562) and no importance score
                scal = wrapper.scaler_row if self.initial_method == "wanda" else torch.ones_like(wrapper.scaler_row)
                self._pending_rows.append((module, scal, sparsity, lora_model, round(module.weight.shape[1] * sparsity), False))
                return
            dsnot_prune_linear(module, wrapper, sparsity, self.prune_n, self.prune_m, lora_model=lora_model,
                               initial_method=self.initial_method, pow_of_var_regrowing=self.pow_of_var_regrowing,
                               max_cycle_time=self.max_cycle_time, update_threshold=self.update_threshold,
                               without_same_sign=self.without_same_sign, without_DSnoT=self.without_DSnoT,
                               ref_fixup=not self.upstream_semantics, elide_noop_swaps=self.elide_noop_swaps)
        return fn
