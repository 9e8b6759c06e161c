"""Global (whole-model) magnitude / first-order pruners over the GPU select.

Mirrors lavis/compression/pruners/global_pruner.py: `BLIPT5GlobalPruner` (:47-236) with `get_mask`,
`get_layerwise_mask`, `global_iterative_pruning`, `prune`, and the registered `blipt5_mag_pruner` (:238-243) and
`blipt5_aobd_pruner` (:253-300).  The reference builds every score on the CPU and runs torch.topk over the concatenated
model; here scores stay in HBM and thresholds come from the exact radix select (K18-K20, layer_sparsity.py).
`blipt5_rand_pruner` (:245-250) is the same flow over random scores.  `blipt5_mezo_pruner` (one zeroth-order number per
layer, so whole layers are pruned) is not mirrored.
"""
import torch

from vlmc import native
from vlmc.common.registry import registry
from vlmc.compression.pruners import layer_sparsity
from vlmc.compression.pruners.layer_single_base_pruner import LayerWiseBasePruner
from vlmc.compression.pruners.utils import loss_vision_language, print_time


class BLIPT5GlobalPruner(LayerWiseBasePruner):
    pruner_name = "blipt5_global_pruner"

    def __init__(self, model, data_loader, t5_prune_spec=None, vit_prune_spec=None, t5_pruning_method=None,
                 vit_pruning_method=None, t5_importance_scores_cache=None, t5_keep_indices_or_masks_cache=None,
                 vit_importance_scores_cache=None, vit_keep_indices_or_masks_cache=None,
                 importance_scores_cache=None, keep_indices_or_masks_cache=None, is_strct_pruning=False,
                 num_samples=64, is_global=False, t5_model_prefix="t5_model", vit_model_prefix="visual_encoder",
                 sparsity_ratio_granularity=None, max_sparsity_per_layer=0.8, score_method="obd_avg",
                 num_data_first_stage=128, num_noise=1, sparsity_dict=None, prune_per_model=False, iteration=1,
                 **kwargs):
        super().__init__(model=model, data_loader=data_loader, prune_spec=None, is_strct_pruning=is_strct_pruning,
                         importance_scores_cache=importance_scores_cache,
                         keep_indices_or_masks_cache=keep_indices_or_masks_cache, is_global=is_global,
                         num_samples=num_samples, model_prefix="tmp",
                         sparsity_ratio_granularity=sparsity_ratio_granularity,
                         max_sparsity_per_layer=max_sparsity_per_layer, score_method=score_method,
                         num_data_first_stage=num_data_first_stage, num_noise=num_noise, sparsity_dict=sparsity_dict)
        self.t5_prune_spec = t5_prune_spec
        self.vit_prune_spec = vit_prune_spec
        self.t5_model_prefix = t5_model_prefix
        self.vit_model_prefix = vit_model_prefix
        self.prune_per_model = prune_per_model
        self.iteration = iteration

    def compute_importance_scores(self, model, data_loader=None, dict_layers_to_prune={}, loss_func=None):
        raise NotImplementedError

    def get_mask(self, importance_scores, p, max_sparsity_per_layer, params=None):
        return layer_sparsity.get_mask(importance_scores, p, max_sparsity_per_layer, params=params)

    def get_layerwise_mask(self, importance_scores, p, params=None):
        return layer_sparsity.get_layerwise_mask(importance_scores, p, params=params)

    def forward_to_cache(self, model, batch, device=None):
        return model(batch)

    # global_pruner.py:153-198
    def global_iterative_pruning(self, target_sparsity, dict_layers_to_prune, iteratation=1, max_sparsity_per_layer=1.0):
        named = dict(self.model.named_parameters())
        masks = None
        for i in range(1, iteratation + 1):
            p_i = target_sparsity ** (iteratation / i)
            measure = self.compute_importance_scores(self.model, self.data_loader, dict_layers_to_prune,
                                                     loss_vision_language)
            measure = {k: v for k, v in measure.items() if k in dict_layers_to_prune}
            if masks is not None:
                for k in measure:
                    measure[k] *= masks[k]
            params = {k: named[k] for k in measure}
            if self.is_global and not self.prune_per_model:
                print("global")
                masks = self.get_mask(measure, p_i, max_sparsity_per_layer, params=params)
            elif self.is_global and self.prune_per_model:
                print("model-level global")
                masks = {}
                for prefix in (self.vit_model_prefix, self.t5_model_prefix):
                    part = {k: v for k, v in measure.items() if k.startswith(prefix)}
                    masks.update(self.get_mask(part, p_i, max_sparsity_per_layer, params={k: params[k] for k in part}))
            else:
                print("layer-wise")
                masks = self.get_layerwise_mask(measure, p_i, params=params)
            print(f"Step {i}, target sparsity: {p_i:.4f}")
        for k, frac in zip(named, layer_sparsity.zero_fraction(list(named.values()))):
            print(k, " sparsity: ", frac)
        return self.model

    # global_pruner.py:200-236
    @print_time
    def prune(self, importance_scores=None, keep_indices_or_masks=None):
        print("In: ", self.pruner_name)
        dtype_record, requires_grad_record, device = self.model_setup_and_record_attributes(self.model)
        if self.t5_prune_spec is None or self.vit_prune_spec is None:
            return self.model, None
        _, vit_keep_ratio, _, _ = self.convert_spec_to_list(self.vit_prune_spec)
        _, t5_keep_ratio, _, _ = self.convert_spec_to_list(self.t5_prune_spec)
        vit_keep_ratio = min(t5_keep_ratio, vit_keep_ratio)
        prunable = set(self.prunable_parameter_names())
        parameters_to_prune = {k: v for k, v in self.model.named_parameters() if k in prunable}
        self.model = self.global_iterative_pruning(1 - vit_keep_ratio, parameters_to_prune, iteratation=self.iteration,
                                                   max_sparsity_per_layer=1.0)
        self.model_reset(self.model, dtype_record, requires_grad_record, device)
        return self.model, None


@registry.register_pruner("blipt5_mag_pruner")
class BLIPT5MagPruner(BLIPT5GlobalPruner):
    pruner_name = "blipt5_mag_pruner"

    def compute_importance_scores(self, model, data_loader=None, dict_layers_to_prune={}, loss_func=None):
        # as shipped (global_pruner.py:242-243) the score is the SIGNED weight, up-cast - not its magnitude.  Only the
        # tensors that will be selected are materialised (the reference copies every parameter and filters afterwards).
        return {k: v.data.float().clone() for k, v in model.named_parameters() if k in dict_layers_to_prune}


@registry.register_pruner("blipt5_rand_pruner")
class BLIPT5RandPruner(BLIPT5GlobalPruner):
    pruner_name = "blipt5_rand_pruner"

    def compute_importance_scores(self, model, data_loader=None, dict_layers_to_prune={}, loss_func=None):
        # global_pruner.py:249-250: standard-normal scores (the random baseline).  Drawn on the device that holds the
        # parameter, so the stream differs from a CPU run of the reference; only the selected tensors are materialised.
        return {k: torch.randn_like(v.data).float().contiguous() for k, v in model.named_parameters()
                if k in dict_layers_to_prune}


@registry.register_pruner("blipt5_aobd_pruner")
class BLIPT5AOBDPruner(BLIPT5GlobalPruner):
    pruner_name = "blipt5_aobd_pruner"

    # global_pruner.py:256-300: |w| * |mean over batches of |grad||
    @print_time
    def compute_importance_scores(self, model, data_loader=None, dict_layers_to_prune={}, loss_func=None):
        names, params = [], []
        for k, v in model.named_parameters():
            if k in dict_layers_to_prune:
                names.append(k)
                params.append(v)
        device = next(iter(model.parameters())).device
        acc = [torch.zeros(p.shape, dtype=torch.float32, device=p.device) for p in params]
        accum_samples = 0
        num_batches = 0
        for d in data_loader:
            if accum_samples >= self.num_samples:
                break
            loss, batch_len = loss_func(model, d, device != "cpu")
            accum_samples += batch_len
            num_batches += 1
            grads = torch.autograd.grad(loss, params)
            native.importance_accum(acc, [g.data.contiguous() for g in grads], "abs")
        scores = [torch.empty_like(a) for a in acc]
        native.importance_finalize(acc, [p.data for p in params], scores, "abs", num_batches)
        return dict(zip(names, scores))
