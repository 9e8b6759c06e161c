"""CPU: host-side mirror of the reference interface (registry, load_pruner, wrappers' bookkeeping) and the
no-CPU-fallback rule."""
import pytest
import torch
import torch.nn as nn

import toy_model


def test_registry_and_load_pruner(built_lib):
    from vlmc.common.registry import registry
    import vlmc.compression as comp
    assert "blipt5_wanda_pruner" in registry.list_pruners()
    model = toy_model.ToyBlip()
    pruner = comp.load_pruner("blipt5_wanda_pruner", model, toy_model.toy_batches(2), cfg=toy_model.pruner_cfg(0.5, 1.0))
    assert pruner.convert_spec_to_list("24-0.5-1.0-1.0") == (24, 0.5, 1.0, 1.0)
    assert pruner.get_sparsity(0.5, None)["anything"] == 0.5
    with pytest.raises(SystemExit):       # reference: TypeError -> print + exit(1)
        comp.load_pruner("no_such_pruner", model, [], cfg={})


def test_wrapper_attributes_match_reference_interface(built_lib):
    from vlmc.compression.pruners.wanda_pruner import WrappedGPT
    lin = nn.Linear(64, 24, bias=False)
    w = WrappedGPT(lin, layer_id=3, layer_name="q")
    assert (w.rows, w.columns, w.nsamples) == (24, 64, 0)
    assert w.scaler_row.shape == (64,) and w.scaler_row.dtype == torch.float32
    assert w.layer is lin and w.layer_id == 3 and w.layer_name == "q"


def test_no_cpu_fallback(built_lib):
    from vlmc.compression.pruners.wanda_pruner import WrappedGPT
    from vlmc import native
    w = WrappedGPT(nn.Linear(64, 24, bias=False))
    with pytest.raises(RuntimeError, match="no CPU path"):
        w.add_batch(torch.randn(1, 8, 64), None)
    with pytest.raises(RuntimeError, match="no CPU path"):
        native.wanda_rowselect(torch.randn(8, 64), torch.ones(64), 3)


def test_find_layers_matches_exact_types(built_lib):
    from vlmc.compression.pruners.layerwise import find_layers, get_module_recursive
    from vlmc.peft.lora import Linear as LoraLinear
    blk = toy_model.ToyLlamaLayer(32, 64)
    blk.self_attn.q_proj = LoraLinear(32, 32, r=2, lora_alpha=16, bias=False)
    found = find_layers(blk)
    assert set(found) == {"self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj", "self_attn.o_proj",
                          "mlp.gate_proj", "mlp.up_proj", "mlp.down_proj"}     # lora_A / lora_B are not leaves
    m = toy_model.ToyBlip()
    assert get_module_recursive(m, "llm_model.model.layers") is m.llm_model.model.layers


def test_lora_linear_interface(built_lib):
    from vlmc.peft.lora import Linear as LoraLinear
    lin = LoraLinear(32, 16, r=4, lora_alpha=16, bias=False)
    assert lin.mask.dtype == torch.bool and lin.mask.all() and "mask" in dict(lin.named_buffers())
    assert lin.scaling == 4.0 and lin.sparse is False
    x = torch.randn(3, 32)
    assert torch.allclose(lin(x, dense=True), x @ lin.weight.T)
    lin.sparse = True
    lin.lora_B.weight.data.normal_()
    lin.mask = torch.rand(16, 32) < 0.5
    import pytest
    with pytest.raises(RuntimeError, match="no CPU path"):      # the masked forward is K15 / K16: CUDA tensors only
        lin(x)


def test_factorisation_assignment_is_balanced_and_deterministic():
    from vlmc import parallel
    cols = [4096, 4096, 4096, 4096, 4096, 4096, 11008]          # q k v o gate up down
    for world in (1, 2, 4, 8):
        owner = parallel.assign_factorisations(cols, world)
        assert owner == parallel.assign_factorisations(cols, world) and max(owner) < world
        load = [sum(c ** 3 for c, o in zip(cols, owner) if o == r) for r in range(world)]
        assert max(load) == 11008 ** 3 or world == 1           # the big one is alone as soon as there are 2 ranks
    assert parallel.row_range(10, 0, 4) == (0, 3) and parallel.row_range(10, 3, 4) == (8, 10)


def test_input_sharing_routes_by_tensor_identity():
    """layerwise.InputSharing: a linear is a follower only when it sees the very tensor its leader saw (same storage,
    shape, strides, dtype, version); equal VALUES at another address, a view with other strides or an in-place update
    in between must not share.  The grouping is fixed by the first sample and must not change afterwards."""
    import pytest
    import torch
    from vlmc.compression.pruners.layerwise import InputSharing, adopt_statistics
    sh = InputSharing()
    h = torch.randn(1, 6, 8)
    sh.begin_forward()
    assert sh.route("q", h) and not sh.route("k", h) and not sh.route("v", h)
    assert sh.route("o", h.clone())                      # equal values, other storage
    assert sh.route("t", h.transpose(1, 2))              # same storage, other shape / strides
    h.add_(1.0)                                          # in-place update bumps the version counter
    assert sh.route("gate", h) and not sh.route("up", h)
    assert sh.leader == {"k": "q", "v": "q", "up": "gate"}
    sh.begin_forward()                                   # next sample: same grouping is fine ...
    h2 = torch.randn(1, 6, 8)
    assert sh.route("q", h2) and not sh.route("k", h2)
    with pytest.raises(RuntimeError):                    # ... a follower that suddenly sees its own tensor is not
        sh.route("v", torch.randn(1, 6, 8))

    class Wrap:
        pass
    lead, fol = Wrap(), Wrap()
    lead.scaler_row, lead.nsamples = torch.ones(8), 3
    fol.scaler_row, fol.nsamples = torch.zeros(8), 0
    adopt_statistics(fol, lead)
    assert fol.scaler_row is lead.scaler_row and fol.nsamples == 3 and fol._shared is lead._shared


def test_linear_assignment_spreads_the_chains():
    """parallel.assign_linears (SparseGPT on several GPUs: whole linears per rank): deterministic, every linear owned,
    the longest chain (down_proj, C = 11008) alone on its rank once there are enough ranks."""
    from vlmc import parallel
    shapes = [(4096, 4096)] * 4 + [(11008, 4096)] * 2 + [(4096, 11008)]
    for world in (1, 2, 3, 4, 8):
        own = parallel.assign_linears(shapes, world)
        assert own == parallel.assign_linears(shapes, world)
        assert len(own) == 7 and all(0 <= o < world for o in own)
        if world >= 4:
            assert own.count(own[6]) == 1
        if world >= 7:
            assert len(set(own)) == 7
    loads = [0.0, 0.0]
    for s, o in zip(shapes, parallel.assign_linears(shapes, 2)):
        loads[o] += parallel.chain_cost_ms(*s)
    assert max(loads) / sum(loads) < 0.6


def test_pruned_checkpoint_layout_matches_the_reference_flow(tmp_path, built_lib):
    """vlmc.checkpoint writes the four artefacts of evaluate_old.py:336-378 under the same folders and names,
    and its loaders apply the reference's prefix rules (:246-290)."""
    import torch
    import yaml
    import toy_model
    from vlmc import checkpoint
    torch.manual_seed(0)
    model = toy_model.ToyBlip(n_vit=1, n_llm=1).eval()
    lin = model.llm_model.model.layers[0].mlp.down_proj
    lin.weight.data[:, ::2] = 0
    setattr(lin.weight, "importance_score", 0.25)
    paths = checkpoint.save_pruned_model(model, "job7", "blipt5_wanda_pruner", sparsity_dict={"a.weight": 0.5},
                                         start_time=None, root=str(tmp_path))
    assert paths["pruned_checkpoint"].endswith("pruned_checkpoint/V+L/blipt5_wanda_pruner/job7.pth")
    assert paths["sparsity_dict"].endswith("sparsity_dict/job7.yaml")
    assert paths["training_statistics"].endswith("training_statistics/job7.yaml")
    assert paths["importance_scores"].endswith("importance_scores/job7.pth")
    assert yaml.safe_load(open(paths["sparsity_dict"])) == {"a.weight": 0.5}
    assert set(yaml.safe_load(open(paths["training_statistics"]))) == {"memory", "time"}
    assert torch.load(paths["importance_scores"]) == {"llm_model.model.layers.0.mlp.down_proj.weight": 0.25}
    state = torch.load(paths["pruned_checkpoint"])
    assert set(state) == set(model.state_dict())
    # loaders: a fresh model takes the pruned language model (prefix stripped) and vision tower (missing keys tolerated)
    fresh = toy_model.ToyBlip(n_vit=1, n_llm=1, seed=1).eval()
    assert checkpoint.load_pruned_language_model(fresh, paths["pruned_checkpoint"]) == "llm_model"
    assert torch.equal(fresh.llm_model.model.layers[0].mlp.down_proj.weight, lin.weight)
    vis = {("visual." + k): v for k, v in model.visual_encoder.state_dict().items() if "qkv" in k}
    vis["visual.not_in_model"] = torch.zeros(1)
    torch.save(vis, tmp_path / "vit.pth")
    assert checkpoint.load_pruned_vision_model(fresh, str(tmp_path / "vit.pth")) == "visual."
    assert torch.equal(fresh.visual_encoder.blocks[0].attn.qkv.weight, model.visual_encoder.blocks[0].attn.qkv.weight)
    assert not torch.equal(fresh.visual_encoder.blocks[0].mlp.fc1.weight, model.visual_encoder.blocks[0].mlp.fc1.weight)


# ---- the allocation loop and the group mapping are host arithmetic ---------------------------------------
def test_sparsity_allocation_matches_reference_golden(built_lib):
    import numpy as np
    import golden_util as gu
    import test_oracle_vs_golden as og
    from vlmc.compression.pruners.layer_sparsity import compute_the_sparsity_per_group
    g = gu.load("layer_sparsity.npz")
    names = [str(k) for k in g["model_names"]]
    for tag in g["alloc_cases"]:
        method, gran, sparsity, ms = str(tag).split("|")
        imp = og._ls_importance(g, method)

        def group_of(name):
            if gran == "layer":
                return name
            if name.startswith("t5_model"):
                return "t5_model" if gran == "model" else ".".join(name.split(".")[:4])
            return "visual_encoder" if gran == "model" else ".".join(name.split(".")[:3])
        scores, counts = {}, {}
        for k in names:
            gk = group_of(k)
            scores[gk] = scores.get(gk, torch.zeros(())) + torch.tensor(float(imp[k].sum(dtype=np.float64))).float()
            counts[gk] = counts.get(gk, 0) + imp[k].size
        if method.endswith("avg"):
            scores = {k: v / counts[k] for k, v in scores.items()}
        total_keep = int(sum(counts.values()) * (1 - float(sparsity)))
        got = compute_the_sparsity_per_group(total_keep, scores, counts, max_sparsity_per_layer=float(ms))
        want = g[f"alloc|{tag}"]
        assert max(abs(got[group_of(k)] - w) for k, w in zip(names, want)) < 1e-9, tag


def test_layer_to_group_mapping_and_registered_global_pruners(built_lib):
    from vlmc.common.registry import registry
    import vlmc.compression  # noqa: F401
    from vlmc.compression.pruners.layer_single_base_pruner import LayerWiseBasePruner, LayerSparsity, UniformSparsity
    assert {"blipt5_mag_pruner", "blipt5_rand_pruner", "blipt5_aobd_pruner"} <= set(registry.list_pruners())

    class M(nn.Module):
        def __init__(self):
            super().__init__()
            self.t5_model = nn.ModuleDict({"encoder": nn.ModuleDict({"block": nn.ModuleList(
                [nn.ModuleDict({"q": nn.Linear(4, 4), "relative_attention_bias": nn.Embedding(3, 4)})])})})
            self.visual_encoder = nn.ModuleDict({"blocks": nn.ModuleList([nn.ModuleDict({"fc1": nn.Linear(4, 8)})])})
            self.head = nn.Linear(4, 2)
    p = LayerWiseBasePruner(M(), [], model_prefix="t5_model")
    p.t5_model_prefix, p.vit_model_prefix = "t5_model", "visual_encoder"
    assert p.prunable_parameter_names() == ["t5_model.encoder.block.0.q.weight", "visual_encoder.blocks.0.fc1.weight"]
    assert p.layer_to_group_mapping("none") == {}
    assert p.layer_to_group_mapping("model") == {"t5_model.encoder.block.0.q.weight": "t5_model",
                                                  "visual_encoder.blocks.0.fc1.weight": "visual_encoder"}
    assert p.layer_to_group_mapping("block") == {"t5_model.encoder.block.0.q.weight": "t5_model.encoder.block.0",
                                                  "visual_encoder.blocks.0.fc1.weight": "visual_encoder.blocks.0"}
    assert list(p.layer_to_group_mapping("layer").values()) == p.prunable_parameter_names()
    with pytest.raises(NotImplementedError):
        p.layer_to_group_mapping("channel")
    ls = LayerSparsity(None, None, None, 1, 0.5, 0.8, "obd_avg", layer_to_group_mapping={})
    assert isinstance(ls.return_sparsity(), UniformSparsity) and ls.return_sparsity()["x"] == 0.5
    with pytest.raises(AssertionError):            # max_sparsity_per_layer < original_sparsity (:146)
        LayerSparsity(None, None, None, 1, 0.9, 0.8, "obd_avg")


def test_global_allocation_has_no_cpu_path(built_lib):
    from vlmc import native
    from vlmc.compression.pruners import layer_sparsity as ls
    scores = {"a": torch.rand(8, 8), "b": torch.rand(4, 4)}
    with pytest.raises(RuntimeError, match="no CPU path"):
        ls.get_mask(scores, 0.5, 1.0)
    with pytest.raises(RuntimeError, match="no CPU path"):
        ls.get_layerwise_mask(scores, 0.5)
    with pytest.raises(RuntimeError, match="no CPU path"):
        native.scores_sum(list(scores.values()))
    with pytest.raises(RuntimeError, match="no CPU path"):
        native.importance_accum([torch.zeros(4)], [torch.ones(4)], "obd")
    with pytest.raises(IndexError):                       # topk(k = 0)[0][-1] on the reference side
        ls.get_mask(scores, 0.0, 1.0)


def test_stack_calibration_groups_equal_shapes(built_lib):
    """layerwise.stack_calibration: consecutive samples of equal shape become one chunk, per-sample tensors
    are concatenated, shared / constant arguments pass through, ragged samples keep the reference's one-by-one schedule."""
    import torch
    from vlmc.compression.pruners.layerwise import stack_calibration
    shared = torch.zeros(1, 4, 8, 8)
    inps = [torch.full((1, 8, 16), float(j)) for j in range(5)] + [torch.ones(1, 6, 16)]
    caches = [{"attention_mask": torch.full((1, 1, 8, 8), float(j)), "position_ids": torch.arange(8)[None],
               "bias": shared, "flag": False, "__args__": (None, torch.full((1, 3), float(j)), 7)} for j in range(5)]
    caches.append({"attention_mask": torch.zeros(1, 1, 6, 6), "position_ids": torch.arange(6)[None], "bias": shared,
                   "flag": False, "__args__": (None, torch.zeros(1, 3), 7)})
    xs, cs, counts = stack_calibration(inps, caches, 4)
    assert counts == [4, 1, 1] and [tuple(x.shape) for x in xs] == [(4, 8, 16), (1, 8, 16), (1, 6, 16)]
    assert torch.equal(xs[0][:, 0, 0], torch.arange(4.0))
    assert cs[0]["attention_mask"].shape == (4, 1, 8, 8) and torch.equal(cs[0]["attention_mask"][:, 0, 0, 0], torch.arange(4.0))
    assert cs[0]["position_ids"].shape == (4, 8) and cs[0]["bias"] is shared and cs[0]["flag"] is False
    assert cs[0]["__args__"][0] is None and cs[0]["__args__"][1].shape == (4, 3) and cs[0]["__args__"][2] == 7
    assert cs[1] is caches[4] and cs[2] is caches[5]
    # calib_batch = 1: untouched
    xs1, cs1, counts1 = stack_calibration(inps, caches, 1)
    assert counts1 == [1] * 6 and all(a is b for a, b in zip(xs1, inps))
    # a per-sample python argument that differs cannot be stacked: the run is halved until it can
    caches[1]["flag"] = True
    _, _, counts2 = stack_calibration(inps, caches, 4)
    assert counts2[0] == 1 and sum(counts2) == 6
