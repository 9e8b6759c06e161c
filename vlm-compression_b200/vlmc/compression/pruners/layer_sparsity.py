"""LayerSparsity: the ECoFLaP global sparsity allocation, scores and selection on the GPU.

Mirrors `LayerSparsity` of lavis/compression/pruners/layer_single_base_pruner.py:111-475: same constructor, same
`get_mask` / `get_layerwise_mask` / `global_iterative_pruning` / `return_sparsity` / `compute_importance_scores`.
What changes is where the work runs.  The reference moves every gradient and every score to the CPU and calls
`torch.topk` on the concatenation of the whole model; here the scores stay fp32 in HBM and

  * the gradient statistics and the score build are one fused multi-tensor pass each (K22, `native.importance_*`),
  * the protected top-(1 - max_sparsity) fraction per tensor and the global threshold come from an exact radix
    select over the tensors where they lie (K18 `native.scores_kth`, K19 `native.scores_protect`),
  * masks and the in-place `param *= mask` are one pass (K20 `native.scores_mask`),
  * group scores are fixed-order fp64 sums (K21 `native.scores_sum`).

The allocation loop itself (`compute_the_sparsity_per_group`, :304-378) works on one number per group; it is restated
below with the reference's tensor dtypes, including its quirks (the kept-parameter vector turns float32 after the first
round, and the "remove the extra parameters" branch adds them, :358).

The zeroth-order (MeZO) estimators (:477-729; the scripts' `olmezo-gradient_sum`) are kept as host loops: their cost is two
model forwards per perturbation, and the perturbation has to be torch's own `torch.normal` stream under the reference's
seeds (a custom generator could not reproduce its numbers), so there is no kernel to write for them.  They follow the
reference's order of operations exactly, including the inexact "recover the weight" step in 16-bit parameters.
"""
import numpy as np
import torch

from vlmc import native
from vlmc.compression.pruners.utils import print_time


class UniformSparsity:
    """layer_single_base_pruner.py:251-255: every key maps to the same sparsity."""

    def __init__(self, sparsity):
        self.sparsity = sparsity

    def __getitem__(self, key):
        return self.sparsity


def _ranks_for_protection(tensors, max_sparsity_per_layer):
    """(indices, ranks): tensors with int(numel * (1 - max_sparsity)) > 0 entries to protect (:153) and the rank, counted
    from the smallest, of the j-th largest score (:157-158)."""
    idx, ranks = [], []
    for i, t in enumerate(tensors):
        j = int(t.numel() * (1 - max_sparsity_per_layer))
        if j > 0:
            idx.append(i)
            ranks.append(t.numel() - j + 1)
    return idx, ranks


def get_mask(importance_scores, p, max_sparsity_per_layer, params=None):
    """`LayerSparsity.get_mask` (:149-176; identical at global_pruner.py:108-135) on device-resident scores.

    importance_scores: dict name -> contiguous fp32 CUDA tensor, modified in place like the reference (protected
    entries become finfo.max).  Returns dict name -> fp32 0/1 mask.  With `params` (dict name -> parameter tensor) the
    prune step `v.data *= mask` (:223-225) is fused into the mask pass."""
    names = list(importance_scores)
    tensors = [importance_scores[k] for k in names]
    idx, ranks = _ranks_for_protection(tensors, max_sparsity_per_layer)
    if idx:
        sub = [tensors[i] for i in idx]
        seg = list(range(len(sub)))
        native.scores_protect(sub, seg, native.scores_kth(sub, seg, ranks))
    total = sum(t.numel() for t in tensors)
    thr = native.scores_kth(tensors, [0] * len(tensors), [int(p * total)])            # :168-170
    masks = {k: torch.empty_like(t) for k, t in zip(names, tensors)}
    native.scores_mask(tensors, [0] * len(tensors), thr, outs=[masks[k] for k in names],
                       params=None if params is None else [params[k].data for k in names])
    return masks


def get_layerwise_mask(importance_scores, p, params=None):
    """`LayerSparsity.get_layerwise_mask` (:178-190; global_pruner.py:137-148): one threshold per tensor, all tensors in
    one select."""
    names = list(importance_scores)
    tensors = [importance_scores[k] for k in names]
    seg = list(range(len(tensors)))
    thr = native.scores_kth(tensors, seg, [int(p * t.numel()) for t in tensors])
    masks = {k: torch.empty_like(t) for k, t in zip(names, tensors)}
    native.scores_mask(tensors, seg, thr, outs=[masks[k] for k in names],
                       params=None if params is None else [params[k].data for k in names])
    return masks


def zero_fraction(tensors):
    """(v == 0).float().sum() / v.numel() per tensor (:228-229) from the K17 non-zero counts."""
    nnz = native.count_nonzero([t.data for t in tensors]).tolist()
    return [(t.numel() - n) / t.numel() for t, n in zip(tensors, nnz)]


def compute_the_sparsity_per_group(total_parameters_to_keep, group_scores, group_num_parameters,
                                   max_sparsity_per_layer=0.8):
    """:304-378.  group_scores / group_num_parameters: dicts with the same keys.  One value per group, so this is host
    arithmetic; the tensor dtypes (and therefore the roundings) are the reference's."""
    keys = list(group_num_parameters.keys())
    scores = torch.tensor([float(group_scores[k]) for k in keys], dtype=torch.float32)
    capacity = torch.tensor([int(group_num_parameters[k]) for k in keys], dtype=torch.int64)
    keep_fraction = 1 - max_sparsity_per_layer
    keep = torch.zeros(len(keys), dtype=torch.int64)
    keep += torch.ceil(capacity * keep_fraction).int()                                # the guaranteed part (:309)

    while keep.sum() < total_parameters_to_keep:
        missing = total_parameters_to_keep - keep.sum()
        grant = torch.ceil((scores / torch.sum(scores)) * missing)                    # :313-316
        keep = keep + grant                                                           # float32 from here on (:318)
        scores[keep >= capacity] = 0                                                  # full groups leave the auction
        keep = torch.clamp(keep, max=capacity)
        if grant.sum() == 0:                                                          # :326-342 nothing was granted
            short = total_parameters_to_keep - keep.sum()
            while short > 0:
                open_groups = torch.where(scores > 0)[0]
                if len(open_groups) == 0:
                    raise RuntimeError("sparsity allocation cannot place the remaining parameters "
                                       "(the reference loops forever here)")
                for g in open_groups:
                    take = min(short, capacity[g] - keep[g])
                    keep[g] += take
                    short -= take
                    if short == 0:
                        break
        if keep.sum() > total_parameters_to_keep:                                     # :344-364
            excess = keep.sum() - total_parameters_to_keep
            while excess > 0:
                moved = False
                for g in torch.argsort(keep, descending=True, stable=True):
                    spare = min(excess, keep[g] - (capacity[g] * keep_fraction).int())
                    moved = moved or bool(spare > 0)                                  # (`spare` may BE `excess`: test first)
                    keep[g] += spare                                                  # as shipped (:358): added, not removed
                    excess -= spare
                    if excess == 0:
                        break
                if not moved:
                    raise RuntimeError("sparsity allocation cannot remove the extra parameters "
                                       "(the reference loops forever here)")

    return {k: torch.clamp(1 - kept / cap, min=0, max=1).item() for k, kept, cap in zip(keys, keep, capacity)}


class LayerSparsity:
    def __init__(self, model, data_loader, loss_func, num_samples, original_sparsity, max_sparsity_per_layer=0.8,
                 score_method="obd_avg", num_noise=1, noise_eps=1e-3, layer_to_group_mapping={},
                 prune_per_model=False, per_model_group=["t5_model", "visual"], per_model_sparsity=[],
                 data_parallel=False):
        # data_parallel (new; the reference runs replicas): with torch.distributed initialised and a loader every rank
        # iterates identically, calibration batches are dealt round-robin to the ranks and the gradient statistics merged
        # with one all-reduce (vlmc.parallel.importance_accumulate_data_parallel); the select then runs replicated.
        self.data_parallel = data_parallel
        self.importance_measure = {}
        self.model = model
        self.data_loader = data_loader
        self.loss_func = loss_func
        self.num_samples = num_samples
        self.original_sparsity = original_sparsity
        self.layer_to_group_mapping = layer_to_group_mapping
        self.max_sparsity_per_layer = max_sparsity_per_layer
        self.num_noise = num_noise
        self.noise_eps = noise_eps
        self.prune_per_model = prune_per_model
        self.score_method = score_method
        self.per_model_group = per_model_group
        self.per_model_sparsity = per_model_sparsity
        if score_method is not None:
            self.score_compute, self.score_aggregate = score_method.split("_")
        assert self.max_sparsity_per_layer >= self.original_sparsity

    def get_mask(self, importance_scores, p, max_sparsity_per_layer, params=None):
        return get_mask(importance_scores, p, max_sparsity_per_layer, params=params)

    def get_layerwise_mask(self, importance_scores, p, params=None):
        return get_layerwise_mask(importance_scores, p, params=params)

    # :192-238
    def global_iterative_pruning(self, target_sparsity, dict_layers_to_prune, iteratation=1, max_sparsity_per_layer=1.0):
        selected = {k: v for k, v in self.model.named_parameters() if k in dict_layers_to_prune}
        saved = {k: v.data.clone() for k, v in selected.items()}      # stays in HBM (the reference parks it on the CPU)
        masks = None
        for i in range(1, iteratation + 1):
            p_i = target_sparsity ** (iteratation / i)
            measure = self.compute_importance_scores(dict_layers_to_prune)
            measure = {k: v for k, v in measure.items() if k in dict_layers_to_prune}
            if masks is not None:
                for k in measure:
                    measure[k] *= masks[k]
            print("global")
            masks = self.get_mask(measure, p_i, max_sparsity_per_layer,
                                  params={k: selected[k] for k in measure})            # mask + `v.data *= mask` in one pass
            print(f"Step {i}, target sparsity: {p_i:.4f}")
        everything = dict(self.model.named_parameters())
        sparsity_dict = dict(zip(everything, zero_fraction(list(everything.values()))))
        for k, v in selected.items():
            v.data = saved[k]
        return sparsity_dict

    # :241-420
    @print_time
    def return_sparsity(self):
        original_sparsity = self.original_sparsity
        mapping = self.layer_to_group_mapping
        print(f"layer_to_group_mapping: {mapping}")
        if self.score_compute.startswith("real"):
            return self.global_iterative_pruning(original_sparsity, mapping, iteratation=3, max_sparsity_per_layer=1.0)
        if mapping is None or len(mapping) == 0:
            return UniformSparsity(original_sparsity)
        if len(self.importance_measure) == 0:
            if self.score_compute.startswith("mezo"):                                 # :263-270, same order of tests
                self.importance_measure = self.compute_importance_scores_mezo_diff(mapping)
            elif self.score_compute.startswith("lmezo"):
                self.importance_measure = self.compute_importance_scores_mezo_layer(mapping)
            elif self.score_compute.startswith("olmezo"):
                self.importance_measure = self.compute_importance_scores_mezo_layer_one(mapping)
            else:
                self.importance_measure = self.compute_importance_scores(mapping)

        groups = {}
        for layer, group in mapping.items():
            groups.setdefault(group, []).append(layer)
        numel = {k: v.numel() for k, v in self.model.named_parameters() if k in mapping}
        total_parameters_to_keep = int(sum(numel.values()) * (1 - original_sparsity))

        layers = [l for members in groups.values() for l in members]
        sums = dict(zip(layers, native.scores_sum([self.importance_measure[l] for l in layers]).tolist()))
        group_scores, group_num_parameters = {}, {}
        for group, members in groups.items():
            # the reference adds float32 tensor sums (:296); the fp64 sums are rounded where it holds float32
            score = torch.zeros((), dtype=torch.float32)
            for l in members:
                score = score + torch.tensor(sums[l], dtype=torch.float64).float()
            n = sum(numel[l] for l in members)
            if self.score_aggregate == "avg":
                score = score / n
            group_scores[group] = score
            group_num_parameters[group] = n

        if self.prune_per_model:
            group_sparsity = {}
            for prefix, sparsity in zip(self.per_model_group, self.per_model_sparsity):
                print(prefix)
                sub_scores = {k: v for k, v in group_scores.items() if k.startswith(prefix)}
                sub_numel = {k: v for k, v in group_num_parameters.items() if k.startswith(prefix)}
                group_sparsity.update(compute_the_sparsity_per_group(
                    int(sum(sub_numel.values()) * (1 - sparsity)), sub_scores, sub_numel,
                    max_sparsity_per_layer=self.max_sparsity_per_layer))
        else:
            group_sparsity = compute_the_sparsity_per_group(
                total_parameters_to_keep, group_scores, group_num_parameters,
                max_sparsity_per_layer=self.max_sparsity_per_layer)

        kept = sum((1 - group_sparsity[k]) * group_num_parameters[k] for k in group_num_parameters)
        print(f"compute_total_keep_parameters: {kept}, total_parameters_to_keep: {total_parameters_to_keep}")
        layer_sparsity = {k: group_sparsity[v] for k, v in mapping.items()}
        print(f"layer_sparsity: {layer_sparsity}")
        return layer_sparsity

    # :422-475
    @print_time
    def compute_importance_scores(self, layer_to_group_mapping):
        names, params = [], []
        for k, v in self.model.named_parameters():
            if k in layer_to_group_mapping:
                names.append(k)
                params.append(v)
        device = next(iter(self.model.parameters())).device
        # one packed fp32 buffer, the per-parameter accumulators are views of it (one all-reduce when data-parallel)
        sizes = [p.numel() for p in params]
        offsets = [0]
        for n in sizes:
            offsets.append(offsets[-1] + (n + 3) // 4 * 4)                            # 16-byte aligned views
        packed = torch.zeros(offsets[-1], dtype=torch.float32, device=params[0].device)
        acc = [packed[o:o + n].view(p.shape) for o, n, p in zip(offsets, sizes, params)]
        accum_mode = "obd" if self.score_compute == "obd" else "abs"                  # :455-458

        def grads_of(d):
            loss, batch_len = self.loss_func(self.model, d, device != "cpu")
            grads = torch.autograd.grad(loss, params)
            assert len(grads) == len(names) == len(params)
            return grads, batch_len

        def accumulate(grads):
            native.importance_accum(acc, [g.data.contiguous() for g in grads], accum_mode)

        import torch.distributed as dist
        if self.data_parallel and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            from vlmc import parallel
            num_batches = parallel.importance_accumulate_data_parallel(
                self.data_loader, self.num_samples, grads_of, accumulate, packed, dist.get_rank(), dist.get_world_size())
        else:
            accum_samples = 0
            num_batches = 0
            for d in self.data_loader:
                if accum_samples >= self.num_samples:
                    break
                grads, batch_len = grads_of(d)
                accum_samples += batch_len
                num_batches += 1
                accumulate(grads)
        if "obd" in self.score_compute:                                               # also taken by "aobd" (:466)
            final_mode = "obd"
        elif "gradient" in self.score_compute:
            final_mode = "gradient"
        else:
            raise UnboundLocalError("importance_measure is not defined for score_compute "
                                    f"{self.score_compute!r} (layer_single_base_pruner.py:466-475)")
        scores = [torch.empty_like(a) for a in acc]
        native.importance_finalize(acc, [p.data for p in params], scores, final_mode, num_batches)
        return dict(zip(names, scores))

    # ---- zeroth-order estimators (:477-729): host loops over model forwards -------------------------------------------
    def zo_perturb_parameters(self, params, random_seed=1, scaling_factor=1, zo_eps=1e-3):
        """theta <- theta + scaling_factor * z * zo_eps with z ~ N(0, 1) drawn from torch's generator seeded with
        random_seed (:477-491); each product is a tensor op in the parameter's dtype, like the reference."""
        torch.manual_seed(random_seed)
        for p in params:
            noise = torch.normal(mean=0, std=1, size=p.data.size(), device=p.data.device, dtype=p.data.dtype)
            step = scaling_factor * noise
            step = step * zo_eps
            p.data = p.data + step

    def _selected(self, layer_to_group_mapping):
        names, params = [], []
        for k, v in self.model.named_parameters():
            if k in layer_to_group_mapping:
                names.append(k)
                params.append(v)
        return names, params

    def _projected_gradient(self, params, d, device, zo_eps):
        """One two-point estimate along a fresh direction: (L(theta + eps z) - L(theta - eps z)) / (2 eps); the parameters
        are stepped +1, -2, +1 like the reference (:632-641), so they come back only up to rounding."""
        seed = np.random.randint(1000000000)
        self.zo_perturb_parameters(params, random_seed=seed, scaling_factor=1, zo_eps=zo_eps)
        with torch.no_grad():
            loss_plus, batch_len = self.loss_func(self.model, d, device != "cpu")
        self.zo_perturb_parameters(params, random_seed=seed, scaling_factor=-2, zo_eps=zo_eps)
        with torch.no_grad():
            loss_minus, batch_len = self.loss_func(self.model, d, device != "cpu")
        self.zo_perturb_parameters(params, random_seed=seed, scaling_factor=1, zo_eps=zo_eps)
        return ((loss_plus - loss_minus) / (2 * zo_eps)).item(), batch_len, seed

    def _zeroth_order_scores(self, prefix, names, params, estimate):
        """The three score rules shared by the estimators (:565-570, :648-653, :724-729)."""
        if self.score_compute == prefix + "-gradient":
            return {k: estimate[k].abs() for k in names}
        if self.score_compute == prefix + "-aobd":
            return {k: v.data.float().abs() * estimate[k].abs() for k, v in zip(names, params)}
        if self.score_compute == prefix + "-obd":
            return {k: v.data.float() ** 2 * estimate[k] ** 2 for k, v in zip(names, params)}
        raise UnboundLocalError(f"importance_measure is not defined for score_compute {self.score_compute!r}")

    def _per_layer_estimate(self, layer_to_group_mapping, n_mezo, absolute_each):
        """:574-646 / :655-722: one parameter tensor at a time, n_mezo directions per batch; the estimate of a layer is one
        number (a 1-element float32 tensor)."""
        self.model.eval()
        names, params = self._selected(layer_to_group_mapping)
        device = next(iter(self.model.parameters())).device
        estimate = {}
        for i, (name, param) in enumerate(zip(names, params)):
            print(i, name)
            total = torch.zeros(1, dtype=torch.float32, device=param.device)
            accum_samples = 0
            for d in self.data_loader:
                if accum_samples >= self.num_samples:
                    break
                per_batch = 0
                for _ in range(n_mezo):
                    if accum_samples >= self.num_samples:
                        break
                    g, batch_len, seed = self._projected_gradient([param], d, device, self.noise_eps)
                    accum_samples += batch_len
                    torch.manual_seed(seed)                     # the reference re-seeds here (:643, :715)
                    per_batch += abs(g) if absolute_each else g
                total = total + torch.tensor([per_batch], dtype=torch.float32, device=param.device).abs()
            estimate[name] = total
        print(estimate)
        return names, params, estimate

    @print_time
    def compute_importance_scores_mezo_layer_one(self, layer_to_group_mapping):
        """"olmezo-*" (:655-729): num_noise directions per batch, |projected gradient| summed."""
        names, params, estimate = self._per_layer_estimate(layer_to_group_mapping, self.num_noise, absolute_each=True)
        return self._zeroth_order_scores("olmezo", names, params, estimate)

    @print_time
    def compute_importance_scores_mezo_layer(self, layer_to_group_mapping):
        """"lmezo-*" (:572-653): 4 directions per batch over 8 samples (the reference overwrites num_samples, :599), the
        signed projected gradients of a batch are summed before the absolute value."""
        self.num_samples = 8
        names, params, estimate = self._per_layer_estimate(layer_to_group_mapping, 4, absolute_each=False)
        return self._zeroth_order_scores("lmezo", names, params, estimate)

    @print_time
    def compute_importance_scores_mezo_diff(self, layer_to_group_mapping):
        """"mezo-*" (:493-570): all selected parameters perturbed together, one zeroth-order SGD step per batch with learning
        rate 1e-3 / #parameters; the estimate is |theta_end - theta_start| / #batches and the weights are restored."""
        self.model.eval()
        names, params = self._selected(layer_to_group_mapping)
        saved = {k: v.data.clone() for k, v in zip(names, params)}           # stays on the device
        device = next(iter(self.model.parameters())).device
        learning_rate = 1 / sum(v.numel() for v in params) * 1e-3
        accum_samples = 0
        num_batches = 0
        for d in self.data_loader:
            if accum_samples >= self.num_samples:
                break
            print(accum_samples)
            g, batch_len, seed = self._projected_gradient(params, d, device, self.noise_eps)
            accum_samples += batch_len
            num_batches += 1
            torch.manual_seed(seed)
            for p in params:
                noise = torch.normal(mean=0, std=1, size=p.data.size(), device=p.data.device, dtype=p.data.dtype)
                p.data = p.data - g * noise * learning_rate
        estimate = {}
        for k, p in zip(names, params):
            estimate[k] = (p.data - saved[k]).float().abs() / num_batches
            p.data = saved[k]
        return self._zeroth_order_scores("mezo", names, params, estimate)

