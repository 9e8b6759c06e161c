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
    w = (lin.weight + (lin.lora_B.weight @ lin.lora_A.weight) * lin.scaling) * lin.mask
    assert torch.allclose(lin(x), x @ w.T, atol=1e-5)


def test_factorisation_assignment_is_balanced_and_deterministic():
    from vlmc import parallel
    cols = [4096, 4096, 4096, 4096, 4096, 4096, 11008]          # q k v o gate up down
    for world in (1, 2, 4, 8):
        owner = parallel.assign_factorisations(cols, world)
        assert owner == parallel.assign_factorisations(cols, world) and max(owner) < world
        load = [sum(c ** 3 for c, o in zip(cols, owner) if o == r) for r in range(world)]
        assert max(load) == 11008 ** 3 or world == 1           # the big one is alone as soon as there are 2 ranks
    assert parallel.row_range(10, 0, 4) == (0, 3) and parallel.row_range(10, 3, 4) == (8, 10)
