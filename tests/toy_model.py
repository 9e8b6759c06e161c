"""Toy InstructBLIP-shaped model that both the reference's composite pruners and vlmc's accept.

Structure the pruners rely on:
  model.visual_encoder.blocks      ModuleList, block(x, rel_pos_bias) -> tensor      (EVA-ViT style)
  model.llm_model.model.layers     ModuleList, layer(x, attention_mask=, position_ids=) -> (tensor,)
  model.llm_model.config.use_cache
  model.maybe_autocast(dtype=None) context manager
  model(batch[, vit_dense=, llm_dense=])
Linear names follow the real models (qkv/proj/fc1/fc2; q_proj..down_proj).
"""
import contextlib
import types

import torch
import torch.nn as nn
import torch.nn.functional as F


class ToyViTBlock(nn.Module):
    def __init__(self, d, hidden):
        super().__init__()
        self.norm1 = nn.LayerNorm(d)
        self.attn = nn.Module()
        self.attn.qkv = nn.Linear(d, 3 * d, bias=False)
        self.attn.proj = nn.Linear(d, d)
        self.norm2 = nn.LayerNorm(d)
        self.mlp = nn.Module()
        self.mlp.fc1 = nn.Linear(d, hidden)
        self.mlp.fc2 = nn.Linear(hidden, d)

    def forward(self, x, rel_pos_bias=None, dense=False):
        h = self.norm1(x)
        q, k, v = self.attn.qkv(h).chunk(3, dim=-1)
        a = torch.softmax(q @ k.transpose(-1, -2) / q.shape[-1] ** 0.5, dim=-1) @ v
        x = x + self.attn.proj(a)
        x = x + self.mlp.fc2(F.gelu(self.mlp.fc1(self.norm2(x))))
        return x


class ToyRMSNorm(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(d))

    def forward(self, x):
        v = x.float().pow(2).mean(-1, keepdim=True)
        return (self.weight * (x.float() * torch.rsqrt(v + 1e-6))).to(x.dtype)


class ToyLlamaLayer(nn.Module):
    def __init__(self, d, ff):
        super().__init__()
        self.input_layernorm = ToyRMSNorm(d)
        self.self_attn = nn.Module()
        for n in ("q_proj", "k_proj", "v_proj", "o_proj"):
            setattr(self.self_attn, n, nn.Linear(d, d, bias=False))
        self.post_attention_layernorm = ToyRMSNorm(d)
        self.mlp = nn.Module()
        self.mlp.gate_proj = nn.Linear(d, ff, bias=False)
        self.mlp.up_proj = nn.Linear(d, ff, bias=False)
        self.mlp.down_proj = nn.Linear(ff, d, bias=False)

    def forward(self, x, attention_mask=None, position_ids=None, dense=False):
        h = self.input_layernorm(x)
        q, k, v = self.self_attn.q_proj(h), self.self_attn.k_proj(h), self.self_attn.v_proj(h)
        a = torch.softmax((q @ k.transpose(-1, -2)).float() / q.shape[-1] ** 0.5, dim=-1).to(v.dtype) @ v
        x = x + self.self_attn.o_proj(a)
        h = self.post_attention_layernorm(x)
        x = x + self.mlp.down_proj(F.silu(self.mlp.gate_proj(h)) * self.mlp.up_proj(h))
        return (x,)


class ToyBertLayer(nn.Module):
    """Q-Former style layer (lavis/models/blip2_models/Qformer.py BertLayer): self-attention over the queries, cross
    attention to the image embeddings, feed forward; called POSITIONALLY by its encoder and returns a tuple."""

    def __init__(self, d, d_enc, hidden):
        super().__init__()
        self.attention = nn.Module()
        self.attention.self = nn.Module()
        for n in ("query", "key", "value"):
            setattr(self.attention.self, n, nn.Linear(d, d))
        self.attention.output = nn.Module()
        self.attention.output.dense = nn.Linear(d, d)
        self.crossattention = nn.Module()
        self.crossattention.self = nn.Module()
        self.crossattention.self.query = nn.Linear(d, d)
        self.crossattention.self.key = nn.Linear(d_enc, d)
        self.crossattention.self.value = nn.Linear(d_enc, d)
        self.crossattention.output = nn.Module()
        self.crossattention.output.dense = nn.Linear(d, d)
        self.intermediate_query = nn.Module()
        self.intermediate_query.dense = nn.Linear(d, hidden)
        self.output_query = nn.Module()
        self.output_query.dense = nn.Linear(hidden, d)

    @staticmethod
    def _attend(q, k, v):
        return torch.softmax(q @ k.transpose(-1, -2) / q.shape[-1] ** 0.5, dim=-1) @ v

    def forward(self, hidden_states, attention_mask=None, head_mask=None, encoder_hidden_states=None,
                encoder_attention_mask=None, past_key_value=None, output_attentions=False, query_length=0):
        a = self.attention.self
        h = hidden_states + self.attention.output.dense(
            self._attend(a.query(hidden_states), a.key(hidden_states), a.value(hidden_states)))
        c = self.crossattention.self
        h = h + self.crossattention.output.dense(
            self._attend(c.query(h), c.key(encoder_hidden_states), c.value(encoder_hidden_states)))
        h = h + self.output_query.dense(F.gelu(self.intermediate_query.dense(h)))
        return (h,)


class ToyBlip(nn.Module):
    def __init__(self, d_vit=64, vit_hidden=128, n_vit=2, d_llm=64, ff=176, n_llm=2, llm_dtype=torch.bfloat16,
                 seed=0, n_qformer=0, d_q=48, n_query=8):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.visual_encoder = nn.Module()
        self.visual_encoder.blocks = nn.ModuleList([ToyViTBlock(d_vit, vit_hidden) for _ in range(n_vit)])
        self.n_qformer = n_qformer
        if n_qformer:
            self.Qformer = nn.Module()
            self.Qformer.bert = nn.Module()
            self.Qformer.bert.encoder = nn.Module()
            self.Qformer.bert.encoder.layer = nn.ModuleList([ToyBertLayer(d_q, d_vit, 2 * d_q) for _ in range(n_qformer)])
            self.query_tokens = nn.Parameter(torch.zeros(1, n_query, d_q))
        self.bridge = nn.Linear(d_q if n_qformer else d_vit, d_llm)
        self.llm_model = nn.Module()
        self.llm_model.config = types.SimpleNamespace(use_cache=True)
        self.llm_model.model = nn.Module()
        self.llm_model.model.layers = nn.ModuleList([ToyLlamaLayer(d_llm, ff) for _ in range(n_llm)])
        self.embed = nn.Embedding(32, d_llm)
        for p in self.parameters():
            if p.dim() >= 2:
                p.data = torch.randn(p.shape, generator=g) * (1.5 / p.shape[-1] ** 0.5)
        self.llm_model.to(llm_dtype)
        self.embed.to(llm_dtype)
        self.llm_dtype = llm_dtype

    def maybe_autocast(self, dtype=torch.float16):
        dev = next(self.parameters()).device
        if dev.type == "cuda":
            return torch.autocast("cuda", dtype=dtype)
        return contextlib.nullcontext()

    def forward(self, batch, vit_dense=False, llm_dense=False):
        x = batch["image"]
        for blk in self.visual_encoder.blocks:
            x = blk(x, None)
        if self.n_qformer:
            q = self.query_tokens.expand(x.shape[0], -1, -1)
            for layer in self.Qformer.bert.encoder.layer:      # positional call, like BertEncoder.forward
                q = layer(q, None, None, x, None, None, False, q.shape[1])[0]
            x = q
        vis = self.bridge(x).to(self.llm_dtype)
        txt = self.embed(batch["text_ids"])
        h = torch.cat([vis, txt], dim=1)
        for layer in self.llm_model.model.layers:
            h = layer(h, attention_mask=None, position_ids=None)[0]
        return h


def toy_batches(n, d_vit=64, n_img_tok=12, n_txt=20, seed=100, device="cpu"):
    out = []
    for j in range(n):
        g = torch.Generator().manual_seed(seed + j)
        gain = torch.exp(torch.rand(d_vit, generator=g) * 2.0 - 1.0)
        img = torch.randn(1, n_img_tok, d_vit, generator=g) * gain + 0.2
        ids = torch.randint(0, 32, (1, n_txt), generator=g)
        out.append({"image": img.to(device), "text_ids": ids.to(device), "text_input": ["a"]})
    return out


def pruner_cfg(t5_keep, vit_keep, prune_n=0, prune_m=0, num_samples=8, **extra):
    """extra: e.g. qformer_prune_spec="12-0.5-1.0-1.0" (vlmc extension), share_inputs=False."""
    cfg = dict(t5_prune_spec=f"24-{t5_keep}-1.0-1.0", vit_prune_spec=f"39-{vit_keep}-1.0-1.0",
               t5_pruning_method="none", vit_pruning_method="none", t5_model_prefix="llm_model",
               vit_model_prefix="visual_encoder", num_samples=num_samples, sparsity_ratio_granularity=None,
               score_method="obd_avg", prune_n=prune_n, prune_m=prune_m)
    cfg.update(extra)
    return cfg
