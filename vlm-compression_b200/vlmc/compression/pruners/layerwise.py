"""The per-block calibration skeleton the three composite pruners share.

Reference: T5LayerWandaPruner._prune (wanda_pruner.py:275-354), VITLayerWandaPruner._prune (:627-699)
and their SparseGPT / DSnoT copies (sparsegpt_pruner.py:405-459, dsnot_pruner.py:312-770).  The
reference repeats this loop six times; here it is one function parameterised by
  make_wrapper(module) -> object with add_batch(inp, out)   (the per-layer statistics wrapper)
  prune_linear(name, module, wrapper, sparsity)            (mask selection / OBS sweep for one linear)
Orchestration only: all arithmetic happens in the wrappers' CUDA kernels.
"""
import contextlib
import gc

import torch
import torch.nn as nn


def get_module_recursive(base, module_to_process):
    """wanda_pruner.py:16-26."""
    for part in [p for p in module_to_process.split(".") if p]:
        base = getattr(base, part)
    return base


def prunable_types():
    from vlmc.peft.lora import Linear as LoraLinear
    return [nn.Linear, LoraLinear]


def find_layers(module, layers=None, name=""):
    """wanda_pruner.py:29-48: leaves whose exact type is prunable, keyed by dotted name."""
    layers = prunable_types() if layers is None else layers
    if type(module) in layers:
        return {name: module}
    found = {}
    for child_name, child in module.named_children():
        found.update(find_layers(child, layers, f"{name}.{child_name}" if name else child_name))
    return found


_T5_KEYS = ["attention_mask", "position_bias", "encoder_attention_mask", "encoder_decoder_position_bias",
            "layer_head_mask", "cross_attn_layer_head_mask", "encoder_hidden_states"]
_OPT_KEYS = ["attention_mask", "layer_head_mask"]
_LLM_KEYS = ["attention_mask", "position_ids"]


class _StopForward(ValueError):
    pass


class InputSharing:
    """Linears of a block that are fed THE SAME tensor (q/k/v, gate/up, wi_0/wi_1, cross-attention k/v) have identical
    statistics; the reference accumulates each of them separately (one hook per linear, wanda_pruner.py:300-311).
    Here the first linear that sees a tensor (the leader) runs the kernel and the others adopt its state after the
    pass: one read of X - and, for SparseGPT, one Hessian and one factorisation - per distinct input.
    Same tensor = same storage address, shape, strides, dtype and version counter while a reference to the leader's
    tensor is held (so the allocator cannot hand the address to another activation).  Results are bit-identical to
    per-linear accumulation because the same kernel would read the same bytes."""

    def __init__(self):
        self.seen = []          # this forward pass: (tensor, version, leader name)
        self.leader = {}        # follower name -> leader name, fixed by the first pass
        self.leaders = set()

    def begin_forward(self):
        self.seen = []

    def route(self, name, x):
        """True when `name` must accumulate x itself, False when a leader already did."""
        found = None
        for t, ver, lead in self.seen:
            if (t.data_ptr() == x.data_ptr() and t.shape == x.shape and t.stride() == x.stride()
                    and t.dtype == x.dtype and t._version == ver):
                found = lead
                break
        if found is None:
            self.seen.append((x, x._version, name))
            if name in self.leader:
                raise RuntimeError(f"{name} shared its input with {self.leader[name]} on an earlier sample but not now")
            self.leaders.add(name)
            return True
        if name in self.leaders or self.leader.setdefault(name, found) != found:
            raise RuntimeError(f"the input sharing of {name} changed between calibration samples")
        return False


def adopt_statistics(follower, leader):
    """The follower wrapper takes the leader's accumulated state BY REFERENCE (same tensors)."""
    for attr in ("scaler_row", "nsamples", "sum_metric_row", "mean", "var", "ntokens", "H"):
        if hasattr(leader, attr):
            setattr(follower, attr, getattr(leader, attr))
    group = getattr(leader, "_shared", None)
    if group is None:
        group = {}
        leader._shared = group
    follower._shared = group


def _dist_info(pruner):
    """(rank, world) when the pruner runs data-parallel: torch.distributed is initialised with more than one
    rank AND the pruner asked for it (data_parallel=True; the reference's scripts launch replicas that each prune the
    whole model redundantly, so the default keeps that behaviour)."""
    import torch.distributed as dist
    if getattr(pruner, "data_parallel", False) and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def merge_statistics_across_ranks(wrappers, n_total):
    """Data-parallel calibration: every rank accumulated the running means of ITS samples; merge them into the statistics
    of the whole calibration set.  Per distinct wrapper: the small vectors (scaler_row, sum_metric_row: sample-weighted;
    mean, var: token-weighted) travel in ONE packed all-reduce per block; a Hessian was accumulated with the divisor of
    the whole set from the start (SparseGPT.add_batch under `_global_n`) and is all-reduced in place.  Afterwards the state
    on every rank is what one rank would have accumulated over all samples (to rounding: sums in another order)."""
    import torch.distributed as dist
    seen, uniq = set(), []
    for w in wrappers:
        key = id(getattr(w, "H", None)) if getattr(w, "H", None) is not None else id(getattr(w, "scaler_row", None))
        if key not in seen:
            seen.add(key)
            uniq.append(w)
    parts, meta = [], []
    for w in uniq:
        if getattr(w, "H", None) is not None:
            dist.all_reduce(w.H, op=dist.ReduceOp.SUM)
            continue
        dev = w.scaler_row.device
        for attr, weight in (("scaler_row", float(w.nsamples)), ("sum_metric_row", float(w.nsamples)),
                             ("mean", float(getattr(w, "ntokens", 0))), ("var", float(getattr(w, "ntokens", 0)))):
            t = getattr(w, attr, None)
            if t is None:
                continue
            parts.append(t.reshape(-1).float() * weight)
            parts.append(torch.full((1,), weight, device=dev))
            meta.append((w, attr, t.numel()))
        if hasattr(w, "ntokens"):
            parts.append(torch.full((1,), float(w.ntokens), device=dev))
            meta.append((w, "__ntokens__", 0))
    if parts:
        flat = torch.cat(parts)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        off = 0
        for w, attr, k in meta:
            if attr == "__ntokens__":
                w._ntokens_total = flat[off:off + 1]
                off += 1
                continue
            t = getattr(w, attr)
            t.copy_((flat[off:off + k] / flat[off + k]).reshape(t.shape))
            off += k + 1
    for w in wrappers:
        w.nsamples = n_total
        if hasattr(w, "_ntokens_total"):
            w.ntokens = int(round(float(w._ntokens_total.item())))
            del w._ntokens_total


def capture_block_inputs(pruner, model, dataloader, model_prefix, n_samples, module_to_process, lora_model, vit,
                         replay_all_args=False):
    """Swap block 0 for a catcher and run the model until n_samples inputs are recorded.

    wanda_pruner.py:213-273 (T5 / LLM keys :224-236) and :583-625 (ViT: rel_pos_bias).
    """
    layers = get_module_recursive(model, module_to_process)
    inps, caches = [], []
    if vit or replay_all_args:
        keys = None
    elif "t5_model" in pruner.model_prefix:
        keys = _T5_KEYS
    elif "opt_model" in pruner.model_prefix:
        keys = _OPT_KEYS
    else:
        keys = _LLM_KEYS

    class Catcher(nn.Module):
        def __init__(self, module):
            super().__init__()
            self.module = module

        def forward(self, inp, *args, dense=True, **kwargs):
            inp.requires_grad = False
            inps.append(inp)
            if replay_all_args:
                # Q-Former (BertEncoder calls its layers positionally): every argument after the hidden states is
                # recorded and replayed as is
                caches.append({"__args__": args, **kwargs})
                raise _StopForward
            if vit:
                rel = args[0] if args else kwargs.get("rel_pos_bias")
                cache = {"rel_pos_bias": rel}
            else:
                cache = {k: kwargs[k] for k in keys}
            if lora_model:
                cache["dense"] = dense
            caches.append(cache)
            raise _StopForward

    layers[0] = Catcher(layers[0])
    seen = 0
    try:
        for batch in dataloader:
            if seen >= n_samples:
                break
            seen += batch["image"].shape[0] if "image" in batch else len(batch["text_input"])
            try:
                pruner.forward_to_cache(model, batch, lora_model)
            except ValueError:
                pass
    finally:
        layers[0] = layers[0].module
    return inps, [None] * len(inps), caches


def _stack_values(vals, sizes):
    """One cache entry of several calibration samples -> the entry of the stacked call, or raises _NotStackable.
    Tensors whose leading dimension is the sample's batch dimension are concatenated; anything else (None, flags,
    a tensor shared by every sample) must be the same for all samples."""
    first = vals[0]
    if isinstance(first, torch.Tensor):
        if all(isinstance(v, torch.Tensor) and v.shape == first.shape and v.dtype == first.dtype for v in vals):
            if all(v is first for v in vals):
                if first.dim() == 0 or first.shape[0] == 1:
                    return first                           # one tensor for every sample: broadcasts over the batch
            if first.dim() >= 1 and all(v.shape[0] == n for v, n in zip(vals, sizes)):
                return torch.cat(vals, dim=0)
        raise _NotStackable
    if isinstance(first, (tuple, list)):
        if not all(isinstance(v, type(first)) and len(v) == len(first) for v in vals):
            raise _NotStackable
        return type(first)(_stack_values([v[i] for v in vals], sizes) for i in range(len(first)))
    if all((v is first) or (type(v) is type(first) and v == first) for v in vals):
        return first
    raise _NotStackable


class _NotStackable(Exception):
    pass


def stack_calibration(inps, caches, calib_batch):
    """Groups consecutive calibration samples of identical shape into chunks of up to `calib_batch` samples
    (batch the calibration set).  Returns (chunk inputs, chunk caches, samples per chunk).  The block then
    runs once per chunk and every hook hands its wrapper ONE [K*b, S, C] tensor: one statistics launch (and, for
    SparseGPT, one read-modify-write of H) per chunk instead of per sample.  Samples that cannot be stacked (ragged
    sequence lengths, per-sample python arguments) stay chunks of one, i.e. the reference's schedule."""
    if calib_batch <= 1:
        return list(inps), list(caches), [1] * len(inps)
    xs, cs, counts = [], [], []
    j = 0
    while j < len(inps):
        grp = [j]
        while len(grp) < calib_batch and j + len(grp) < len(inps) and inps[j + len(grp)].shape == inps[j].shape \
                and inps[j + len(grp)].dtype == inps[j].dtype:
            grp.append(j + len(grp))
        cache = None
        while len(grp) > 1:
            try:
                sizes = [inps[g].shape[0] for g in grp]
                keys = caches[grp[0]].keys()
                if any(caches[g].keys() != keys for g in grp):
                    raise _NotStackable
                cache = {k: _stack_values([caches[g][k] for g in grp], sizes) for k in keys}
                break
            except _NotStackable:
                grp = grp[:len(grp) // 2]                  # retry with a shorter run
        if len(grp) == 1:
            xs.append(inps[j])
            cs.append(caches[j])
        else:
            xs.append(torch.cat([inps[g] for g in grp], dim=0))
            cs.append(cache)
        counts.append(len(grp))
        j += len(grp)
    return xs, cs, counts


def prune_blocks(pruner, model, dataloader, model_prefix, module_to_process, n_samples, sparsity_ratio,
                 lora_model, vit, make_wrapper, prune_linear, replay_all_args=False):
    stem = getattr(model, model_prefix, None)
    cfg = getattr(stem, "config", None) if not (vit or replay_all_args) else None
    use_cache = getattr(cfg, "use_cache", None)
    if cfg is not None:
        cfg.use_cache = False
    with torch.no_grad():
        inps, outs, caches = capture_block_inputs(pruner, model, dataloader, model_prefix, n_samples,
                                                  module_to_process, lora_model, vit, replay_all_args)
    n_samples = min(n_samples, len(inps))
    layers = get_module_recursive(model, module_to_process)
    expected_nsamples = len(inps) * inps[0].shape[0]
    # data-parallel calibration: rank r keeps samples r, r + world, ...; the statistics are merged per block
    rank, world = _dist_info(pruner)
    if world > 1:
        inps, caches = inps[:n_samples][rank::world], caches[:n_samples][rank::world]
        n_samples = len(inps)
    # calibration batching: the block runs on chunks of stacked samples (calib_batch = 1 restores the reference's
    # one-sample-per-forward schedule, wanda_pruner.py:308-311)
    inps, caches, counts = stack_calibration(inps[:n_samples], caches[:n_samples],
                                             int(getattr(pruner, "calib_batch", 1) or 1))
    outs = [None] * len(inps)
    n_samples = len(inps)                                  # from here on: number of chunks

    share = InputSharing() if getattr(pruner, "share_inputs", True) else None
    current = {"calls": 1}
    # wrappers that can take the statistics of a whole block forward in one launch (WrappedGPT.add_batch_many): their hooks
    # only record (wrapper, input); the launch is issued when the forward returns
    deferred = [] if getattr(pruner, "batch_statistics", False) else None

    def run_block(layer):
        for j in range(n_samples):
            current["calls"] = counts[j]
            if share is not None:
                share.begin_forward()
            with torch.no_grad():
                if replay_all_args:          # Q-Former: runs in its own dtype, outside autocast
                    ctx = contextlib.nullcontext()
                else:
                    ctx = model.maybe_autocast() if vit else model.maybe_autocast(dtype=torch.bfloat16)
                with ctx:
                    kw = dict(caches[j])
                    out = layer(inps[j], *kw.pop("__args__", ()), **kw)
                    outs[j] = out if vit else out[0]
            if deferred:
                for w, x, ver in deferred:
                    if x._version != ver:        # the model wrote into a linear's input after the linear ran
                        raise RuntimeError("an activation was modified in place between its linear and the end of the block "
                                           "forward: run this model with batch_statistics=False")
                type(deferred[0][0]).add_batch_many([(w, x.data) for w, x, _ in deferred])
                del deferred[:]

    for i in range(len(layers)):
        layer = layers[i]
        subset = find_layers(layer)
        wrapped = {name: make_wrapper(subset[name]) for name in subset}
        if world > 1:
            for w in wrapped.values():
                w._global_n = expected_nsamples      # Hessians accumulate with the divisor of the whole set
        if share is not None:
            share.leader, share.leaders = {}, set()

        def make_hook(nm):
            def hook(_, inp, out):
                if share is None or share.route(nm, inp[0]):
                    # DSnoT's var is a mean of PER-CALL variances: its wrapper splits a stacked chunk back into the
                    # reference's calls (vlmc_dsnot_stats nseg); the other wrappers take any batch size
                    wrapped[nm]._stacked_calls = current["calls"]
                    if deferred is not None and hasattr(wrapped[nm], "add_batch_many"):
                        deferred.append((wrapped[nm], inp[0], inp[0]._version))
                    else:
                        wrapped[nm].add_batch(inp[0].data, out.data)
            return hook
        handles = [subset[name].register_forward_hook(make_hook(name)) for name in wrapped]
        run_block(layer)
        for h in handles:
            h.remove()
        if share is not None:
            share.begin_forward()                       # drop the references to the last sample's activations
            for follower, leader in share.leader.items():
                adopt_statistics(wrapped[follower], wrapped[leader])
        if world > 1:
            merge_statistics_across_ranks(list(wrapped.values()), expected_nsamples)
        for name in subset:
            key = f"{module_to_process}.{i}.{name}.weight"
            try:
                sparsity = sparsity_ratio[key]
            except KeyError:
                # the reference reads the per-linear ratio only in its unstructured branch (wanda_pruner.py:333): an
                # n:m run must not fail on a sparsity_dict that lacks this linear
                if getattr(pruner, "prune_n", 0) == 0:
                    raise
                sparsity = None
            prune_linear(i, name, subset[name], wrapped[name], sparsity, expected_nsamples=expected_nsamples)
        pruner.finish_block(subset, wrapped)
        run_block(layer)
        inps, outs = outs, inps

    if cfg is not None:
        cfg.use_cache = use_cache
    if torch.cuda.is_available():
        torch.cuda.empty_cache()
    gc.collect()
    return model
