"""API shell of the reference's pruner base classes (boundary only, no arithmetic).

Mirrors lavis/compression/pruners/base_pruner.py:7-82 and layer_single_base_pruner.py:10-108:
constructor arguments, spec parsing, requires_grad bookkeeping around prune().  `LayerSparsity`
(layer_single_base_pruner.py:111-475) lives in layer_sparsity.py and is re-exported here, where the
reference defines it; with the scripts' `sparsity_ratio_granularity none` it degenerates to a constant lookup.
"""
import yaml

from vlmc.compression.pruners.layer_sparsity import LayerSparsity, UniformSparsity  # noqa: F401
from vlmc.compression.pruners.utils import loss_vision_language


class BasePruner:
    def __init__(self, model, data_loader, is_strct_pruning=False, keep_indices_or_masks_cache=None,
                 importance_scores_cache=None, is_global=False, num_samples=64):
        self.model = model
        self.data_loader = data_loader
        self.is_strct_pruning = is_strct_pruning
        self.is_global = is_global
        self.num_samples = num_samples
        self.keep_indices_or_masks_cache = keep_indices_or_masks_cache
        self.importance_scores_cache = importance_scores_cache

    def prune(self, importance_scores=None, keep_indices_or_masks=None):
        raise NotImplementedError


class LayerWiseBasePruner(BasePruner):
    def __init__(self, model, data_loader, prune_spec=None, importance_scores_cache=None,
                 keep_indices_or_masks_cache=None, is_strct_pruning=False, num_samples=64, is_global=False,
                 model_prefix="t5_model", sparsity_ratio_granularity=None, max_sparsity_per_layer=0.8,
                 score_method="obd_avg", num_data_first_stage=128, num_noise=1, sparsity_dict=None,
                 noise_eps=1e-3, prune_per_model=False, prune_n=0, prune_m=0, **kwargs):
        super().__init__(model=model, data_loader=data_loader, is_strct_pruning=is_strct_pruning,
                         importance_scores_cache=importance_scores_cache,
                         keep_indices_or_masks_cache=keep_indices_or_masks_cache, is_global=is_global,
                         num_samples=num_samples)
        self.sparsity_ratio_granularity = sparsity_ratio_granularity
        self.max_sparsity_per_layer = max_sparsity_per_layer
        self.score_method = score_method
        self.num_data_first_stage = num_data_first_stage
        self.num_noise = num_noise
        self.sparsity_dict = sparsity_dict
        self.noise_eps = noise_eps
        self.prune_per_model = prune_per_model
        self.prune_spec = prune_spec
        self.model_prefix = model_prefix
        self.prune_n, self.prune_m = prune_n, prune_m
        self.model_stem = getattr(self.model, model_prefix, None)

    # layer_single_base_pruner.py:72-97
    def model_setup_and_record_attributes(self, model):
        dtype_record, requires_grad_record = {}, {}
        for n, p in model.named_parameters():
            dtype_record[n] = p.data.dtype
            requires_grad_record[n] = p.requires_grad
            p.requires_grad = True
        device = next(iter(model.parameters())).device
        return dtype_record, requires_grad_record, device

    def model_reset(self, model, dtype_record, requires_grad_record, device):
        for n, p in model.named_parameters():
            p.requires_grad = requires_grad_record[n]
            if p.data.dtype != dtype_record[n]:
                p.data = p.data.type(dtype_record[n])
        model.to(device)

    # "<layers>-<keep>-<attn>-<ffn>", only field 2 is used (layer_single_base_pruner.py:99-105)
    def convert_spec_to_list(self, spec):
        num_layers, res_keep, attn_keep, ffn_keep = spec.split("-")
        return int(num_layers), float(res_keep), float(attn_keep), float(ffn_keep)

    def prunable_parameter_names(self):
        """wanda_pruner.py:875-885: 2-D parameters of the transformer blocks of either sub-model."""
        t5, vit = getattr(self, "t5_model_prefix", self.model_prefix), getattr(self, "vit_model_prefix", self.model_prefix)
        return [k for k, v in self.model.named_parameters()
                if len(v.shape) == 2 and ".block" in k and "relative_attention_bias.weight" not in k
                and (k.startswith(t5) or k.startswith(vit))]

    def layer_to_group_mapping(self, sparsity_ratio_granularity):
        """wanda_pruner.py:871-920: which parameters share one sparsity ("model" / "block" / "layer")."""
        if sparsity_ratio_granularity in (None, "none"):
            return {}
        t5, vit = getattr(self, "t5_model_prefix", self.model_prefix), getattr(self, "vit_model_prefix", self.model_prefix)
        if sparsity_ratio_granularity not in ("model", "layer", "block"):
            raise NotImplementedError

        def group_of(name):
            if sparsity_ratio_granularity == "layer":
                return name
            for prefix, fields in ((t5, 4), (vit, 3)):
                if name.startswith(prefix):
                    return prefix if sparsity_ratio_granularity == "model" else ".".join(name.split(".")[:fields])
            return "other"

        return {k: group_of(k) for k in self.prunable_parameter_names()}

    def get_sparsity(self, original_sparsity, sparsity_ratio_granularity=None):
        """wanda_pruner.py:865-939.  A sparsity_dict yaml wins; granularity none -> constant; otherwise the ECoFLaP
        allocation over first-order importance scores (LayerSparsity, scores and selection on the GPU)."""
        if self.sparsity_dict is not None:
            with open(self.sparsity_dict, "r") as f:
                return yaml.load(f, Loader=yaml.FullLoader)
        sparsity_module = LayerSparsity(
            self.model, self.data_loader, loss_vision_language, self.num_data_first_stage, original_sparsity,
            self.max_sparsity_per_layer, self.score_method, self.num_noise, self.noise_eps,
            self.layer_to_group_mapping(sparsity_ratio_granularity))
        return sparsity_module.return_sparsity()
