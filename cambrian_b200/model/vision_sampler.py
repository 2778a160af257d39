"""Spatial Vision Aggregator — B200-native mirror of the reference's `cambrian/model/vision_sampler.py`.

Same class names, constructor signatures, attribute / state-dict key names and call conventions as the reference
(`VisionTokenSampler` vision_sampler.py:407-419, `VisionCrossAttentionLayer` :248-327, `MultiKVCrossAttention`
:155-234, `MLP` :237-245), so released checkpoints load unchanged and `cambrian_arch` / `cambrian_llama` call sites work
as they are.  The arithmetic runs in `cambrian_b200.autograd.SVALayerFn` on the hand-written sm_100a kernels
(tcgen05 GEMMs for the LayerNorm+Linear projections, the fused window-attention kernel for the softmax).

Only the `joint` layer type used by the released models is implemented; `sep` raises NotImplementedError.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ..autograd import SVALayerFn


class MultiKVCrossAttention(nn.Module):
    """Parameter container with the reference's names (q_proj / k_proj_i / v_proj_i = LayerNorm+Linear, o_proj)."""

    def __init__(self, q_dim, kv_dim_list, hidden_dim, num_heads, attention_bias=False):
        super().__init__()
        self.hidden_dim = hidden_dim
        self.num_heads = num_heads
        self.head_dim = hidden_dim // num_heads
        if self.head_dim * num_heads != hidden_dim:
            raise ValueError(
                f"hidden_dim must be divisible by num_heads (got `hidden_dim`: {hidden_dim} and `num_heads`: {num_heads}).")
        if attention_bias:
            raise NotImplementedError("attention_bias=True is never used by the reference models")
        self.q_proj = nn.Sequential(nn.LayerNorm(q_dim), nn.Linear(q_dim, hidden_dim, bias=False))
        self.num_of_kvs = len(kv_dim_list)
        for i, kv_dim in enumerate(kv_dim_list):
            setattr(self, f"k_proj_{i}", nn.Sequential(nn.LayerNorm(kv_dim), nn.Linear(kv_dim, hidden_dim, bias=False)))
            setattr(self, f"v_proj_{i}", nn.Sequential(nn.LayerNorm(kv_dim), nn.Linear(kv_dim, hidden_dim, bias=False)))
        self.o_proj = nn.Linear(hidden_dim, q_dim, bias=False)


class MLP(nn.Module):
    def __init__(self, d_in, d_hidden, d_out):
        super().__init__()
        self.linear_1 = nn.Linear(d_in, d_hidden, bias=False)
        self.act = nn.GELU()
        self.linear_2 = nn.Linear(d_hidden, d_out, bias=False)


class VisionCrossAttentionLayer(nn.Module):
    def __init__(self, q_dim, context_dim, kv_dim_list, kv_size_list, hidden_dim=1024, layer_idx=0):
        super().__init__()
        num_heads = 16
        if hidden_dim != 1024 or any(d != hidden_dim for d in kv_dim_list):
            raise NotImplementedError("the sm_100a SVA kernels are specialised for vision_hidden_size = 1024 (16 x 64)")
        self.num_of_kvs = len(kv_dim_list)
        self.q_dim = q_dim
        self.proj_context = nn.Linear(context_dim, hidden_dim, bias=False)
        self.proj_in = nn.Linear(q_dim + hidden_dim, hidden_dim, bias=False)
        self.proj_out = MLP(hidden_dim, hidden_dim, q_dim)
        self.norm = nn.LayerNorm(hidden_dim)
        self.cross_attn = MultiKVCrossAttention(hidden_dim, kv_dim_list, hidden_dim, num_heads)
        self.kv_size_list = list(kv_size_list)
        for i, kv_size in enumerate(kv_size_list):
            if kv_size > 1:
                setattr(self, f"pos_embed_{i}", nn.Parameter(torch.randn(kv_size ** 2, hidden_dim)))

    def _named_params(self):
        ca = self.cross_attn
        items = [("proj_context", self.proj_context.weight), ("proj_in", self.proj_in.weight),
                 ("out1_w", self.proj_out.linear_1.weight), ("out2_w", self.proj_out.linear_2.weight),
                 ("norm_w", self.norm.weight), ("norm_b", self.norm.bias),
                 ("q_ln_w", ca.q_proj[0].weight), ("q_ln_b", ca.q_proj[0].bias), ("q_w", ca.q_proj[1].weight),
                 ("o_w", ca.o_proj.weight)]
        for i in range(self.num_of_kvs):
            kp, vp = getattr(ca, f"k_proj_{i}"), getattr(ca, f"v_proj_{i}")
            items += [(f"k_ln_w_{i}", kp[0].weight), (f"k_ln_b_{i}", kp[0].bias), (f"k_w_{i}", kp[1].weight),
                      (f"v_ln_w_{i}", vp[0].weight), (f"v_ln_b_{i}", vp[0].bias), (f"v_w_{i}", vp[1].weight)]
            if self.kv_size_list[i] > 1:
                items.append((f"pos_embed_{i}", getattr(self, f"pos_embed_{i}")))
        return items

    def forward(self, queries, context_feature, *vision_latents_attention_mask_list, natural_layout=None):
        """Reference call convention (vision_sampler.py:270-275): queries [N,1,q_dim], context [N,1,ctx_dim], then T
        window-rearranged latents [N, r_i^2, 1024] followed by T bool masks [N, r_i^2].

        natural_layout=(B, q_side): fast path used by our cambrian_arch — latents are the un-rearranged grids
        [B, (r_i q_side)^2, 1024] and the window gather happens inside the kernels (no permute/contiguous copies)."""
        T = self.num_of_kvs
        latents = list(vision_latents_attention_mask_list[:T])
        masks = list(vision_latents_attention_mask_list[T:])
        n = queries.shape[0]
        for i, m in enumerate(masks):
            if m is not None and m.numel() != n * self.kv_size_list[i] ** 2:
                # same failure class as the reference's mask-shape check (vision_sampler.py:202-206)
                raise ValueError(f"Attention mask should be of size {(n, 1, 1, self.kv_size_list[i] ** 2)}, "
                                 f"but is {tuple(m.shape)}")
        masks = [None if m is None else m.reshape(n, -1) for m in masks] if masks else None
        if masks is not None and all(m is None for m in masks):
            masks = None
        items = self._named_params()
        names = [k for k, _ in items]
        params = [p for _, p in items]
        if any(p.dtype != torch.bfloat16 or not p.is_cuda for p in params):
            raise RuntimeError("cambrian_b200 SVA layers run in bf16 on CUDA only (no CPU / fp32 fallback): "
                               "call .to(device='cuda', dtype=torch.bfloat16)")
        meta = dict(T=T, rs=self.kv_size_list, masks=masks, natural=natural_layout, names=names, params=params,
                    feat_shapes=[t.shape for t in latents])
        q2 = queries.reshape(n, -1)
        c2 = context_feature.reshape(n, -1)
        latents = [t.to(torch.bfloat16) for t in latents]
        out = SVALayerFn.apply(meta, q2.contiguous(), c2.contiguous(), *[t.contiguous() for t in latents], *params)
        return out.view(queries.shape)


class VisionTokenSampler(nn.Module):
    def __init__(self, q_dim, context_dim, kv_dim_list, kv_size_list, vision_hidden_size, num_of_layers=1,
                 layer_type="joint"):
        super().__init__()
        assert layer_type in ["joint", "sep"]
        if layer_type != "joint":
            raise NotImplementedError("layer_type='sep' (VisionAggregationLayer) is unused by the released Cambrian models")
        self.layers = nn.ModuleList([
            VisionCrossAttentionLayer(q_dim, context_dim, kv_dim_list, kv_size_list, vision_hidden_size, idx)
            for idx in range(num_of_layers)])

    def forward(self, queries, context_feature, *vision_latents_attention_mask_list, natural_layout=None):
        for layer in self.layers:
            queries = layer(queries, context_feature, *vision_latents_attention_mask_list, natural_layout=natural_layout)
        return queries
