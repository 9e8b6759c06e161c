"""Wanda pruner: same entry points and per-layer wrapper as lavis/compression/pruners/wanda_pruner.py,
with the statistics and the mask selection running in sm_100a kernels (include/vlmc.h).

  WrappedGPT                  <- wanda_pruner.py:51-81   (scaler_row accumulation, K1)
  BLIPT5LayerWandaPruner      <- wanda_pruner.py:796-1044 (registered as "blipt5_wanda_pruner")
  per-linear score + select   <- wanda_pruner.py:316-341 (LLM/T5: per-row, K5/K6), :664-687 (ViT: K7/K6)
"""
import os

import torch
import torch.nn as nn

from vlmc import native
from vlmc.common.registry import registry
from vlmc.compression.pruners.layer_single_base_pruner import LayerWiseBasePruner
from vlmc.compression.pruners.layerwise import find_layers, get_module_recursive, prune_blocks  # noqa: F401
from vlmc.compression.pruners.utils import print_time


class WrappedGPT:
    """Per-linear Wanda statistics (wanda_pruner.py:51-81).

    scaler_row[c] is the running mean over samples of sum_t x[t, c]^2; `nsamples` counts the leading
    batch dimension, not tokens (:71,:77-78).  add_batch is one kernel launch (vlmc_sqnorm_accum).
    """

    def __init__(self, layer, layer_id=0, layer_name="none"):
        self.layer = layer
        self.dev = self.layer.weight.device
        self.rows = layer.weight.data.shape[0]
        self.columns = layer.weight.data.shape[1]
        self.scaler_row = torch.zeros((self.columns), device=self.dev)
        self.nsamples = 0
        self.layer_id = layer_id
        self.layer_name = layer_name

    def add_batch(self, inp, out=None):
        if len(inp.shape) == 2:
            inp = inp.unsqueeze(0)
        b = inp.shape[0]
        native.sqnorm_accum(inp, self.scaler_row, self.nsamples, b)
        self.nsamples += b

    @staticmethod
    def add_batch_many(pairs):
        """[(wrapper, inp)] of one block forward -> ONE multi-tensor launch per run of equal (nsamples, b) (in practice:
        one per block forward; vlmc_sqnorm_accum_batch).  Same update as add_batch on each pair."""
        groups = {}
        for w, inp in pairs:
            if len(inp.shape) == 2:
                inp = inp.unsqueeze(0)
            groups.setdefault((w.nsamples, inp.shape[0], inp.device), []).append((w, inp))
        for (n, b, _), items in groups.items():
            if len(items) == 1:
                items[0][0].add_batch(items[0][1])
                continue
            native.sqnorm_accum_batch([x for _, x in items], [w.scaler_row for w, _ in items], n, b)
            for w, _ in items:
                w.nsamples += b


def wanda_prune_block_nm(modules, scaler_rows, prune_n, prune_m, lora_model=False):
    """n:m score + select + apply for ALL linears of one block in one launch per dtype (vlmc_wanda_nm_batch); the
    reference's per-linear loop (wanda_pruner.py:313-347) with identical masks and weights.  Sets module.mask;
    returns the importance scores as a list of 1-element device tensors (one per module)."""
    out = [None] * len(modules)
    groups = {}
    for i, mod in enumerate(modules):
        W = mod.weight.data
        groups.setdefault((W.dtype, W.device), []).append(i)
    for idx in groups.values():
        for c0 in range(0, len(idx), 16):
            chunk = idx[c0:c0 + 16]
            keeps, means = native.wanda_nm_batch([modules[i].weight.data for i in chunk], [scaler_rows[i] for i in chunk],
                                                 prune_n, prune_m, zero_w=not lora_model)
            for j, i in enumerate(chunk):
                setattr(modules[i], "mask", keeps[j])
                out[i] = means[j:j + 1]
    return out


def wanda_prune_block_rows(modules, scaler_rows, sparsities, lora_model=False, streams=3, ks=None):
    """Per-row top-k (wanda_pruner.py:332-341) for the linears of one block: the same one launch per linear as
    wanda_prune_linear, but longest first and dealt over `streams` side streams - a warp of vlmc_wanda_rowselect owns whole
    rows, so the tail of one launch (warps that got one row fewer) is filled by the next linear's rows instead of idling
    (3.37 -> 3.16 ms per Vicuna block).  Masks are allocated on the caller's stream.  Identical masks, weights and scores.
    Sets module.mask; returns the importance scores as a list of 1-element device tensors (one per module).
    ks[i] (optional) overrides the rows' prune count int(C * sparsity) of module i (DSnoT rounds it)."""
    from vlmc.schedule import Fork

    def kof(i):
        return int(ks[i]) if ks is not None and ks[i] is not None else int(modules[i].weight.shape[1] * sparsities[i])
    out = [None] * len(modules)
    keeps = [torch.empty(m.weight.shape, dtype=torch.bool, device=m.weight.device) for m in modules]
    means = [torch.empty(1, dtype=torch.float32, device=m.weight.device) for m in modules]
    order = sorted(range(len(modules)), key=lambda i: -modules[i].weight.numel())
    by_dev = {}
    for i in order:
        by_dev.setdefault(modules[i].weight.device, []).append(i)
    for dev, idx in list(by_dev.items()):
        # ONE call per device and dtype (vlmc_wanda_rowselect_batch: the linears of equal row length share a launch whose
        # CTAs walk the concatenated rows); VLMC_ROWSELECT_BATCH=0 restores one launch per linear on side streams
        if os.environ.get("VLMC_ROWSELECT_BATCH") != "0" and len({modules[i].weight.dtype for i in idx}) == 1:
            _, mm = native.wanda_rowselect_batch([modules[i].weight.data for i in idx], [scaler_rows[i] for i in idx],
                                                 [kof(i) for i in idx],
                                                 zero_w=not lora_model, keep_masks=[keeps[i] for i in idx])
            for j, i in enumerate(idx):
                means[i] = mm[j:j + 1]
            del by_dev[dev]
    for dev, idx in by_dev.items():
        with Fork(dev, max(1, min(streams, len(idx)))) as fk:
            for slot, i in enumerate(idx):
                W = modules[i].weight.data
                with fk.stream(slot):
                    native.wanda_rowselect(W, scaler_rows[i], kof(i), zero_w=not lora_model,
                                           keep_mask=keeps[i], score_mean=means[i])
    for i, mod in enumerate(modules):
        setattr(mod, "mask", keeps[i])
        out[i] = means[i]
    return out


def wanda_prune_linear(module, scaler_row, sparsity, prune_n=0, prune_m=0, lora_model=False, whole_matrix=False):
    """Score + select + apply for one linear (wanda_pruner.py:316-341 / :664-687).

    Sets module.mask (True = kept) and zeroes pruned weights in place unless lora_model.
    Returns the device scalar mean(|W| * sqrt(scaler_row)) (the reference's importance_score).
    """
    W = module.weight.data
    if prune_n != 0:
        keep, mean = native.wanda_nm(W, scaler_row, prune_n, prune_m, zero_w=not lora_model)
    elif whole_matrix:
        keep, mean = native.wanda_threshold(W, scaler_row, int(W.numel() * sparsity), zero_w=not lora_model)
    else:
        keep, mean = native.wanda_rowselect(W, scaler_row, int(W.shape[1] * sparsity), zero_w=not lora_model)
    setattr(module, "mask", keep)
    return mean


@registry.register_pruner("blipt5_wanda_pruner")
class BLIPT5LayerWandaPruner(LayerWiseBasePruner):
    pruner_name = "blipt5_wanda_pruner"

    def __init__(self, model, data_loader, t5_prune_spec=None, vit_prune_spec=None, t5_pruning_method=None,
                 vit_pruning_method=None, t5_importance_scores_cache=None, t5_keep_indices_or_masks_cache=None,
                 vit_importance_scores_cache=None, vit_keep_indices_or_masks_cache=None,
                 importance_scores_cache=None, keep_indices_or_masks_cache=None, is_strct_pruning=False,
                 num_samples=64, is_global=False, t5_model_prefix="t5_model", vit_model_prefix="visual_encoder",
                 sparsity_ratio_granularity=None, max_sparsity_per_layer=0.8, score_method="obd_avg",
                 num_data_first_stage=128, num_noise=1, sparsity_dict=None, noise_eps=1e-3,
                 prune_per_model=False, peft_postfix="", prune_n=0, prune_m=0, share_inputs=True,
                 qformer_prune_spec=None, qformer_model_prefix="Qformer", calib_batch=16, data_parallel=False,
                 batch_statistics=True, **kwargs):
        super().__init__(model=model, data_loader=data_loader, prune_spec=None, is_strct_pruning=is_strct_pruning,
                         importance_scores_cache=importance_scores_cache,
                         keep_indices_or_masks_cache=keep_indices_or_masks_cache, is_global=is_global,
                         num_samples=num_samples, model_prefix=f"{vit_model_prefix}+{t5_model_prefix}",
                         sparsity_ratio_granularity=sparsity_ratio_granularity,
                         max_sparsity_per_layer=max_sparsity_per_layer, score_method=score_method,
                         num_data_first_stage=num_data_first_stage, num_noise=num_noise,
                         sparsity_dict=sparsity_dict, noise_eps=noise_eps, prune_per_model=prune_per_model,
                         prune_n=prune_n, prune_m=prune_m)
        self.t5_prune_spec = t5_prune_spec
        self.vit_prune_spec = vit_prune_spec
        self.peft_postfix = peft_postfix
        self.vit_dense = True
        self.llm_dense = True
        assert t5_pruning_method is not None
        assert vit_pruning_method is not None
        self.t5_model_prefix = t5_model_prefix
        self.vit_model_prefix = vit_model_prefix
        self._pending_scores = []
        self._pending_nm = []
        self._pending_rows = []
        # linears fed by the same tensor accumulate their statistics once (layerwise.InputSharing); False restores
        # the reference's one-accumulation-per-linear schedule.  The results are identical either way.
        self.share_inputs = share_inputs
        # calibration samples of equal shape are stacked into chunks of calib_batch for the block forwards, so every
        # linear gets ONE add_batch per chunk (layerwise.stack_calibration); 1 = one sample per forward like the
        # reference (wanda_pruner.py:308-311).  Statistics agree to rounding (tests: 1e-5), masks on tie-free data.
        self.calib_batch = calib_batch
        # the Wanda statistics of the linears a block forward reaches are accumulated by ONE multi-tensor launch after the
        # forward (vlmc_sqnorm_accum_batch) instead of one launch per hook; False: one launch per hook.  Statistics agree
        # to rounding (the partial sums are chunked differently), like calib_batch.
        self.batch_statistics = batch_statistics
        # True (and torch.distributed initialised with > 1 rank): the ranks split the calibration samples, merge the
        # statistics with one all-reduce per block and end up with identical pruned replicas.  False: every
        # rank prunes its replica on its own, like the reference's torchrun launch does.
        self.data_parallel = data_parallel
        # EXTENSION (no reference behaviour to match): the reference never prunes the Q-Former (its pruners
        # walk visual_encoder.blocks, t5/llm layers or OPT layers only).  With a "<layers>-<keep>-1.0-1.0" spec the 12
        # BertLayers under <prefix>.bert.encoder.layer are pruned with the same per-linear rule as the language model.
        self.qformer_prune_spec = qformer_prune_spec
        self.qformer_model_prefix = qformer_model_prefix

    # ---- hooks the shared block loop calls ------------------------------------------------------
    def forward_to_cache(self, model, batch, lora_model=False):
        if lora_model:
            return model(batch, vit_dense=self.vit_dense, llm_dense=self.llm_dense)
        return model(batch)

    def make_wrapper(self, module):
        return WrappedGPT(module)

    def _prune_linear(self, vit, lora_model):
        def fn(i, name, module, wrapper, sparsity, expected_nsamples):
            assert wrapper.nsamples == expected_nsamples
            if self.prune_n != 0:                   # n:m: the whole block in one launch (finish_block)
                self._pending_nm.append((module, wrapper.scaler_row, lora_model))
                return
            if not vit:                             # per-row top-k: the block's launches dealt over streams (finish_block)
                self._pending_rows.append((module, wrapper.scaler_row, sparsity, lora_model, None, True))
                return
            mean = wanda_prune_linear(module, wrapper.scaler_row, sparsity, self.prune_n, self.prune_m,
                                      lora_model=lora_model, whole_matrix=vit)
            self._pending_scores.append((module, mean))
        return fn

    def finish_block(self, subset, wrapped):
        if self._pending_nm:
            mods = [m for m, _, _ in self._pending_nm]
            means = wanda_prune_block_nm(mods, [s for _, s, _ in self._pending_nm], self.prune_n, self.prune_m,
                                         lora_model=self._pending_nm[0][2])
            self._pending_scores.extend(zip(mods, means))
            self._pending_nm = []
        if self._pending_rows:
            # entries: (module, scaler_row, sparsity, lora_model, k override or None, importance score wanted)
            mods = [e[0] for e in self._pending_rows]
            means = wanda_prune_block_rows(mods, [e[1] for e in self._pending_rows], [e[2] for e in self._pending_rows],
                                           lora_model=self._pending_rows[0][3], ks=[e[4] for e in self._pending_rows])
            self._pending_scores.extend((m, v) for m, v, e in zip(mods, means, self._pending_rows) if e[5])
            self._pending_rows = []
        # one host sync per block instead of the reference's full-matrix .cpu() per linear (:320)
        if self._pending_scores:
            vals = torch.cat([m for _, m in self._pending_scores]).tolist()
            for (module, _), v in zip(self._pending_scores, vals):
                setattr(module.weight, "importance_score", v)
            self._pending_scores = []

    def _prune(self, model, dataloader, model_prefix, module_to_process, n_samples, sparsity_ratio,
               lora_model=False, vit=False, replay_all_args=False):
        return prune_blocks(self, model, dataloader, model_prefix, module_to_process, n_samples, sparsity_ratio,
                            lora_model, vit, self.make_wrapper, self._prune_linear(vit, lora_model),
                            replay_all_args=replay_all_args)

    # ---- entry point (wanda_pruner.py:947-1044) -------------------------------------------------
    @print_time
    def prune(self, importance_scores=None, keep_indices_or_masks=None, lora_model=False):
        dtype_record, requires_grad_record, device = self.model_setup_and_record_attributes(self.model)
        global_sparsity_dict = None
        _, vit_keep_ratio, _, _ = self.convert_spec_to_list(self.vit_prune_spec)
        _, t5_keep_ratio, _, _ = self.convert_spec_to_list(self.t5_prune_spec)
        if self.sparsity_ratio_granularity not in [None, "none"]:
            global_sparsity_dict = self.get_sparsity(1 - t5_keep_ratio, self.sparsity_ratio_granularity)
        self.vit_dense = float(vit_keep_ratio) < 1.0
        self.llm_dense = float(t5_keep_ratio) < 1.0

        if self.vit_prune_spec is not None and float(vit_keep_ratio) < 1.0:
            sparsity = global_sparsity_dict if global_sparsity_dict not in [None, "none"] \
                else self.get_sparsity(1 - vit_keep_ratio, None)
            self.model = self._prune(self.model, self.data_loader, self.vit_model_prefix,
                                     f"{self.vit_model_prefix}.blocks", self.num_samples, sparsity,
                                     lora_model=lora_model, vit=True)

        if self.qformer_prune_spec is not None:
            _, q_keep_ratio, _, _ = self.convert_spec_to_list(self.qformer_prune_spec)
            if float(q_keep_ratio) < 1.0:
                sparsity = global_sparsity_dict if global_sparsity_dict not in [None, "none"] \
                    else self.get_sparsity(1 - q_keep_ratio, None)
                self.model = self._prune(self.model, self.data_loader, self.qformer_model_prefix,
                                         f"{self.qformer_model_prefix}.bert.encoder.layer", self.num_samples, sparsity,
                                         lora_model=lora_model, vit=False, replay_all_args=True)

        if self.t5_prune_spec is not None and float(t5_keep_ratio) < 1.0:
            sparsity = global_sparsity_dict if global_sparsity_dict is not None \
                else self.get_sparsity(1 - t5_keep_ratio, None)
            if "t5_model" in self.t5_model_prefix:
                stacks = [f"{self.t5_model_prefix}.encoder.block", f"{self.t5_model_prefix}.decoder.block"]
            else:
                stacks = [f"{self.t5_model_prefix}{self.peft_postfix}.model.layers"]
            for stack in stacks:
                self.model = self._prune(self.model, self.data_loader, self.t5_model_prefix, stack,
                                         self.num_samples, sparsity, lora_model=lora_model, vit=False)

        self.model_reset(self.model, dtype_record, requires_grad_record, device)
        return self.model, global_sparsity_dict
