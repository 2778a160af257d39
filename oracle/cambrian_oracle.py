"""TEST INFRASTRUCTURE ONLY — CPU oracle for the Cambrian-1 image->text hot path.

A functional, fp32, plain-torch restatement of what the reference computes on the path SURVEY.md §8a lists
(A1-A11).  It is the checker for the CUDA path: only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
cpu_baseline / `--impl reference` leg may import it; the product (`cambrian_b200/`) never does.

Every function takes tensors plus a flat state dict that uses the REFERENCE's parameter names (SURVEY.md §8b
"State-dict keys"), so the same weights can be loaded into the reference modules, this oracle and the CUDA
modules.  Pinning status (see tests/test_oracle_pin.py, tests/golden/make_golden.py):
  * SVA layer / sampler, window rearrange (train and inference/unpad variants, unmask_attention_mask, unpad_image),
    projectors, connector + splice — pinned against the reference's own
    modules imported through oracle/ref_shim.py, and against committed golden fixtures generated from them.
  * CLIP ViT, DINOv2 ViT, LLaMA decoder layer — pinned against the installed `transformers` implementations
    the reference delegates to (clip_encoder.py:47, dino_encoder.py:81, cambrian_llama.py:23-24).
  * SigLIP ViT and ConvNeXt-XXL trunks — the reference delegates to timm 0.9.16 via open_clip, which is NOT
    installed and cannot be fetched.  Restated from the published timm definitions and pinned against the
    independent `transformers` implementations of the same architectures (SiglipVisionModel, ConvNextModel) with
    weights mapped name by name (tests/test_oracle_pin.py).  What stays unpinned against the reference's own
    dependency is one choice only: the GELU flavour timm 0.9.16 applies in the SigLIP MLP (a config field here).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

IGNORE_INDEX = -100
IMAGE_TOKEN_INDEX = -200


# ------------------------------------------------------------------------------------------------
# numerics modes
# ------------------------------------------------------------------------------------------------
# Every function below is dtype- and device-agnostic plain torch, so the same restatement serves as
#   * the fp32 oracle ("truth"): fp32 state dict + fp32 inputs, and
#   * the EAGER-bf16 oracle: `eager_bf16(sd)` + bf16 inputs.  torch then rounds to bf16 after every op exactly as the
#     reference does when it runs its eager PyTorch path in bf16 (fp32 accumulation inside each matmul, fp32 islands
#     where the reference upcasts: RMSNorm statistics, softmax of the decoder attention, token-grid interpolation,
#     the logits before the loss).  This is the reference's ACTUAL arithmetic on a GPU, and therefore the yardstick for
#     the CUDA path:  err(CUDA, fp32 oracle)  <=  slack * err(eager-bf16 oracle, fp32 oracle)   per tensor.
# Both modes may run on any torch device (the tests use the GPU box's device for the full-size cases so the oracle
# finishes in seconds; that is the checker running on torch/cuBLAS, never the product).
def eager_bf16(sd):
    """bf16 copy of a (fp32) state dict: the weights the reference holds when it runs in bf16."""
    return {k: (v.detach().to(torch.bfloat16) if v.is_floating_point() else v.detach()) for k, v in sd.items()}


def to_device(sd, device):
    return {k: v.to(device) for k, v in sd.items()}


# ------------------------------------------------------------------------------------------------
# small building blocks
# ------------------------------------------------------------------------------------------------
def _lin(sd, name, x, bias=True):
    b = sd.get(name + ".bias") if bias else None
    return F.linear(x, sd[name + ".weight"], b)


def _ln(sd, name, x, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd.get(name + ".bias"), eps)


def quick_gelu(x):
    return x * torch.sigmoid(1.702 * x)


def rms_norm(x, w, eps):
    """train_fsdp.py:1429-1435 (training-time patch): fp32 normalise, multiply by weight, then cast."""
    xf = x.float()
    return (w.float() * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps))).to(x.dtype)


# ------------------------------------------------------------------------------------------------
# A6/A7 — Spatial Vision Aggregator
# ------------------------------------------------------------------------------------------------
def window_rearrange(feat, q_side):
    """cambrian_arch.py:271-287: [B, (q r)^2, C] -> [B q^2, r^2, C]."""
    B, n, C = feat.shape
    side = int(round(n ** 0.5))
    assert (side // q_side) * q_side == side
    r = side // q_side
    t = feat.view(B, q_side, r, q_side, r, C).permute(0, 1, 3, 2, 4, 5)
    return t.reshape(B * q_side * q_side, r * r, C)


def sva_layer(sd, p, queries, ctx, feats, masks, heads=16):
    """VisionCrossAttentionLayer.forward, vision_sampler.py:270-327 (+ MultiKVCrossAttention :177-234).

    queries [N,1,Dq]; ctx [N,1,Dc]; feats[i] [N, r_i^2, Dkv]; masks[i] [N, r_i^2] bool (or None)."""
    n = queries.shape[0]
    residual = queries
    c = F.linear(ctx, sd[p + "proj_context.weight"])                        # :279
    q = F.linear(torch.cat([queries, c], -1), sd[p + "proj_in.weight"])     # :281,:292
    ks, vs, ms = [], [], []
    for i, f in enumerate(feats):
        if f.shape[1] > 1:                                                  # :304-309
            f = f + sd[p + f"pos_embed_{i}"][None].to(f.dtype)
        ks.append(_lin(sd, p + f"cross_attn.k_proj_{i}.1", _ln(sd, p + f"cross_attn.k_proj_{i}.0", f), bias=False))
        vs.append(_lin(sd, p + f"cross_attn.v_proj_{i}.1", _ln(sd, p + f"cross_attn.v_proj_{i}.0", f), bias=False))
        m = masks[i] if masks is not None and masks[i] is not None else None
        ms.append(torch.ones(n, f.shape[1], dtype=torch.bool, device=f.device) if m is None
                  else m.view(n, -1).bool().to(f.device))
    hidden = q.shape[-1]
    hd = hidden // heads
    Q = _lin(sd, p + "cross_attn.q_proj.1", _ln(sd, p + "cross_attn.q_proj.0", q), bias=False)
    Q = Q.view(n, 1, heads, hd).transpose(1, 2)
    K = torch.cat(ks, 1).view(n, -1, heads, hd).transpose(1, 2)
    V = torch.cat(vs, 1).view(n, -1, heads, hd).transpose(1, 2)
    mask = torch.cat(ms, -1)[:, None, None, :]                              # :200
    s = (Q @ K.transpose(-1, -2)) / math.sqrt(hd)
    s = s.masked_fill(~mask, float("-inf"))
    a = (torch.softmax(s, -1) @ V).transpose(1, 2).reshape(n, 1, hidden)    # :215-230
    a = F.linear(a, sd[p + "cross_attn.o_proj.weight"])                     # :232
    q = _ln(sd, p + "norm", q + a)                                          # :319-321
    q = F.linear(F.gelu(F.linear(q, sd[p + "proj_out.linear_1.weight"])), sd[p + "proj_out.linear_2.weight"])
    return q + residual                                                     # :325


def sva_agg_layer(sd, p, queries, ctx, feats, masks, heads=16):
    """VisionAggregationLayer.forward (layer_type="sep"), vision_sampler.py:330-405: one aggregate per tower —
    CrossAttention (:55-121) over the r_i^2 window when r_i > 1, a two-layer MLP of the single latent otherwise
    (AggregationBlock :124-153) — mixed with a per-query softmax over towers (weight_mlp, :369-371) and added to the
    query stream (:396-398)."""
    n = queries.shape[0]
    T = len(feats)
    residual = queries
    c = F.linear(ctx, sd[p + "proj_context.weight"])                                    # :360
    cat = torch.cat([queries, c], -1)                                                    # :362
    if T > 1:                                                                            # :364-368
        w = F.linear(F.gelu(F.linear(cat, sd[p + "weight_mlp.linear_1.weight"])), sd[p + "weight_mlp.linear_2.weight"])
        w = w.softmax(-1).unsqueeze(-1)                                                  # [N, 1, T, 1]
    else:
        w = 1
    q = F.linear(cat, sd[p + "proj_in.weight"])                                          # :370
    hidden = q.shape[-1]
    hd = hidden // heads
    aggs = []
    for i, f in enumerate(feats):
        a = p + f"aggregate_{i}.attention_layer."
        if f.shape[1] > 1:                                                               # :382-387, CrossAttention :79-121
            f = f + sd[p + f"pos_embed_{i}"][None].to(f.dtype)
            Q = _lin(sd, a + "q_proj.1", _ln(sd, a + "q_proj.0", q), bias=False).view(n, 1, heads, hd).transpose(1, 2)
            K = _lin(sd, a + "k_proj.1", _ln(sd, a + "k_proj.0", f), bias=False).view(n, -1, heads, hd).transpose(1, 2)
            V = _lin(sd, a + "v_proj.1", _ln(sd, a + "v_proj.0", f), bias=False).view(n, -1, heads, hd).transpose(1, 2)
            s_ = (Q @ K.transpose(-1, -2)) / math.sqrt(hd)
            m = masks[i] if masks is not None and masks[i] is not None else None
            if m is not None:
                s_ = s_.masked_fill(~m.view(n, 1, 1, -1).bool().to(f.device), float("-inf"))
            o = (torch.softmax(s_, -1) @ V).transpose(1, 2).reshape(n, 1, hidden)
            aggs.append(F.linear(o, sd[a + "o_proj.weight"]))
        else:                                                                            # MLP(kv_dim, q_dim, q_dim) :140
            aggs.append(F.linear(F.gelu(F.linear(f, sd[a + "linear_1.weight"])), sd[a + "linear_2.weight"]))
    agg = torch.stack(aggs, 2)                                                           # :394  [N, 1, T, hidden]
    q = q + (agg * w).sum(2)                                                             # :396
    q = _ln(sd, p + "norm", q)                                                           # :398
    q = F.linear(F.gelu(F.linear(q, sd[p + "proj_out.linear_1.weight"])), sd[p + "proj_out.linear_2.weight"])
    return q + residual                                                                  # :402


def sva_sampler(sd, p, queries, ctx, feats, masks, num_layers, layer_type="joint"):
    """VisionTokenSampler.forward, vision_sampler.py:407-419."""
    layer = sva_layer if layer_type == "joint" else sva_agg_layer
    for l in range(num_layers):
        queries = layer(sd, f"{p}layers.{l}.", queries, ctx, feats, masks)
    return queries


# ------------------------------------------------------------------------------------------------
# A5/A8 — projectors, connector, splice (static-shape branch, cambrian_arch.py:366-490)
# ------------------------------------------------------------------------------------------------
def mm_projector_aux(sd, p, x):
    """nn.Sequential(Linear, GELU, Linear, LayerNorm)  cambrian_arch.py:56."""
    return _ln(sd, p + "3", _lin(sd, p + "2", F.gelu(_lin(sd, p + "0", x))))


def mlp2x_gelu(sd, p, x):
    """nn.Sequential(Linear, GELU, Linear): cambrian_arch.py:49 / multimodal_projector/builder.py:60-67."""
    return _lin(sd, p + "2", F.gelu(_lin(sd, p + "0", x)))


def connector(sd, cfg, tower_feats, aux_masks):
    """cambrian_arch.py:366-420 (mm_projector_type == 'sva', one query group, static branch).

    tower_feats[i] [B, N_i, C_i]; aux_masks[i] [B*q^2, r_i^2] bool.  Returns image_features [B, q*(q+1), H],
    rearranged aux features / ctx for the in-LLM SVA layers."""
    B = tower_feats[0].shape[0]
    q_num = cfg["image_token_len"]
    q_side = int(q_num ** 0.5)
    aux = [mm_projector_aux(sd, f"model.mm_projector_aux_{i}.", f) for i, f in enumerate(tower_feats)]
    ctx = aux[0].mean(1).view(B, 1, 1, -1)                                   # :377
    feats_w = [window_rearrange(a, q_side) for a in aux]
    ctx_q = ctx.expand(-1, q_num, 1, -1).flatten(0, 1)
    groups = []
    for g, qn in enumerate(cfg.get("query_num_list", [q_num])):              # :382-402, one sampler per query group
        qs = int(qn ** 0.5)
        queries = sd["model.vision_query"][g].view(1, 1, 1, -1).expand(B, qn, -1, -1).flatten(0, 1)
        ctx_g = ctx.expand(-1, qn, 1, -1).flatten(0, 1)
        fw = feats_w if qs == q_side else [window_rearrange(a, qs) for a in aux]
        # a group with another side reuses the collator's masks through a raw reshape, exactly as
        # rearrange_vision_tower_features_train does (cambrian_arch.py:284: `.view(bs * q * q, r * r)`)
        gm = aux_masks if (qs == q_side or aux_masks is None) else [m.reshape(B * qn, -1) for m in aux_masks]
        qf = sva_sampler(sd, f"model.vision_sampler_{g}.", queries, ctx_g, fw, gm, cfg["connector_depth"]).view(B, qn, -1)
        if qs != q_side:                                                     # :394-401 bilinear resize of the query grid
            t = qf.permute(0, 2, 1).contiguous().view(B, -1, qs, qs)
            t = F.interpolate(t.float(), size=(q_side, q_side), mode="bilinear", align_corners=False).to(qf.dtype)
            qf = t.permute(0, 2, 3, 1).contiguous().flatten(1, 2)
        groups.append(qf)
    qf = torch.cat(groups, -1)
    img = mlp2x_gelu(sd, "model.mm_projector.", qf.view(B, q_num, -1))       # :410-411
    img = img.view(B, q_side, q_side, -1)
    nl = sd["model.image_newline"][None, None, None, :].expand(B, q_side, 1, -1)
    img = torch.cat([img, nl], 2).flatten(1, 2)                              # :413-420
    return img, feats_w, ctx_q


def splice(sd, input_ids, image_features):
    """cambrian_arch.py:457-490 static branch, one image per sample: the <image> indicator and the 599 pad ids
    that follow it are replaced by the 600 image embeddings."""
    ids = torch.where(input_ids == IMAGE_TOKEN_INDEX, 0, input_ids)
    emb = F.embedding(ids, sd["model.embed_tokens.weight"]).clone()
    for b in range(input_ids.shape[0]):
        pos = torch.where(input_ids[b] == IMAGE_TOKEN_INDEX)[0]
        if len(pos) == 0:
            continue
        s = int(pos[0])
        L = image_features.shape[1]
        emb[b, s:s + L] = image_features[b].to(emb.dtype)
    return emb


# ------------------------------------------------------------------------------------------------
# A9-A11 — LLaMA decoder with SVA insertion, lm_head, loss
# ------------------------------------------------------------------------------------------------
def rope_cos_sin(position_ids, head_dim, theta):
    inv = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.float32, device=position_ids.device) / head_dim))
    fr = position_ids[..., None].float() * inv                               # [B,S,hd/2]
    emb = torch.cat([fr, fr], -1)
    return emb.cos(), emb.sin()


def apply_rope(x, cos, sin):
    """x [B,h,S,hd]; HF rotate_half convention."""
    x1, x2 = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
    cos, sin = cos.to(x.dtype), sin.to(x.dtype)     # HF LlamaRotaryEmbedding returns cos/sin in the activation dtype
    return x * cos[:, None] + torch.cat([-x2, x1], -1) * sin[:, None]


def llama_layer(sd, p, x, cos, sin, attn_mask_2d, cfg):
    """HF LlamaDecoderLayer (transformers, called from cambrian_llama.py:142-166)."""
    B, S, H = x.shape
    nh, nkv = cfg["num_attention_heads"], cfg["num_key_value_heads"]
    hd = H // nh
    h = rms_norm(x, sd[p + "input_layernorm.weight"], cfg["rms_norm_eps"])
    q = F.linear(h, sd[p + "self_attn.q_proj.weight"]).view(B, S, nh, hd).transpose(1, 2)
    k = F.linear(h, sd[p + "self_attn.k_proj.weight"]).view(B, S, nkv, hd).transpose(1, 2)
    v = F.linear(h, sd[p + "self_attn.v_proj.weight"]).view(B, S, nkv, hd).transpose(1, 2)
    q, k = apply_rope(q, cos, sin), apply_rope(k, cos, sin)
    k = k.repeat_interleave(nh // nkv, 1)
    v = v.repeat_interleave(nh // nkv, 1)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
    causal = torch.ones(S, S, dtype=torch.bool, device=x.device).tril()
    allow = causal[None, None]
    if attn_mask_2d is not None:
        allow = allow & attn_mask_2d.bool().to(x.device)[:, None, None, :]
    s = s.masked_fill(~allow, torch.finfo(s.dtype).min)                      # HF additive-min mask semantics
    a = (torch.softmax(s.float(), -1).to(s.dtype) @ v).transpose(1, 2).reshape(B, S, H)
    x = x + F.linear(a, sd[p + "self_attn.o_proj.weight"])
    h = rms_norm(x, sd[p + "post_attention_layernorm.weight"], cfg["rms_norm_eps"])
    g = F.silu(F.linear(h, sd[p + "mlp.gate_proj.weight"])) * F.linear(h, sd[p + "mlp.up_proj.weight"])
    return x + F.linear(g, sd[p + "mlp.down_proj.weight"])


def decoder(sd, cfg, inputs_embeds, position_ids, attn_mask_2d, feats_w=None, aux_masks=None, ctx_q=None):
    """CambrianLlamaModel.forward, cambrian_llama.py:98-277, static SVA-insertion branch :168-207."""
    x = inputs_embeds
    B = x.shape[0]
    hd = cfg["hidden_size"] // cfg["num_attention_heads"]
    cos, sin = rope_cos_sin(position_ids, hd, cfg["rope_theta"])
    n_sva = 0 if cfg.get("connector_only", True) else cfg["num_of_vision_sampler_layers"]
    sites = [cfg["start_of_vision_sampler_layers"] + i * cfg["stride_of_vision_sampler_layers"]
             for i in range(n_sva)]
    q_num = cfg.get("image_token_len", 576)
    side = int(q_num ** 0.5)
    for i in range(cfg["num_hidden_layers"]):
        x = llama_layer(sd, f"model.layers.{i}.", x, cos, sin, attn_mask_2d, cfg)
        if feats_w is not None and i in sites:
            s0 = cfg["image_position"]
            blk = x[:, s0:s0 + q_num + side].clone().view(B, side, side + 1, -1)
            lq, nl = blk[:, :, :-1], blk[:, :, -1:]
            lq = lq.reshape(B * q_num, 1, -1)
            lq = sva_sampler(sd, f"model.vision_sampler_layers.{sites.index(i)}.", lq, ctx_q,
                             [f.to(lq.dtype) for f in feats_w], aux_masks, 1)
            blk = torch.cat([lq.view(B, side, side, -1), nl], 2).flatten(1, 2)
            x = x.clone()
            x[:, s0:s0 + q_num + side] = blk
    return rms_norm(x, sd["model.norm.weight"], cfg["rms_norm_eps"])


def lm_loss(sd, hidden, labels):
    """cambrian_llama.py:402-422: lm_head -> fp32 logits -> shifted CE (ignore_index -100, mean)."""
    logits = F.linear(hidden, sd["lm_head.weight"]).float()
    loss = None
    if labels is not None:
        loss = F.cross_entropy(logits[:, :-1].reshape(-1, logits.shape[-1]), labels[:, 1:].reshape(-1),
                               ignore_index=IGNORE_INDEX)
    return logits, loss


# ------------------------------------------------------------------------------------------------
# dynamic-shape (non-XLA / inference) branch: per-sample unpadded query grids
# ------------------------------------------------------------------------------------------------
def unmask_attention_mask(mask, original_size):
    """cambrian_arch.py:203-227."""
    ow, oh = original_size
    ch, cw = mask.shape[1:3]
    if ow / oh > cw / ch:
        pad = (ch - int(oh * (cw / ow))) // 2
        if pad > 0:
            mask[:, :pad, :] = 0
            mask[:, -pad:, :] = 0
    else:
        pad = (cw - int(ow * (ch / oh))) // 2
        if pad > 0:
            mask[:, :, :pad] = 0
            mask[:, :, -pad:] = 0
    return mask


def unpad_image(t, original_size):
    """cambrian_arch.py:230-256 (crop of dims 1, 2)."""
    ow, oh = original_size
    ch, cw = t.shape[1:3]
    if ow / oh > cw / ch:
        pad = (ch - int(oh * (cw / ow))) // 2
        return t[:, pad:ch - pad, :]
    pad = (cw - int(ow * (ch / oh))) // 2
    return t[:, :, pad:cw - pad]


def rearrange_inference(feats, q_side, image_sizes, unpad=False):
    """rearrange_vision_tower_features_inference, cambrian_arch.py:289-330."""
    out_f, out_m = [], []
    for f in feats:
        B, N, C = f.shape
        side = int(N ** 0.5)
        r = side // q_side
        fl, ml = [], []
        for b in range(B):
            w = f[b].view(1, q_side, r, q_side, r, C).permute(0, 1, 3, 2, 4, 5).contiguous()
            m = unmask_attention_mask(torch.ones((1, side, side), dtype=torch.bool), image_sizes[b])
            m = m.view(1, q_side, r, q_side, r).permute(0, 1, 3, 2, 4).contiguous()
            if unpad:
                w, m = unpad_image(w, image_sizes[b]), unpad_image(m, image_sizes[b])
            w = w.flatten(0, 2).flatten(1, 2)
            m = m.flatten(0, 2).flatten(1, 2).clone()
            m[m.sum(-1) == 0] = True
            fl.append(w)
            ml.append(m)
        out_f.append(torch.cat(fl, 0))
        out_m.append(torch.cat(ml, 0))
    return out_f, out_m


def prepare_dynamic(sd, cfg, tower_feats, input_ids, attention_mask, labels, image_sizes):
    """prepare_inputs_labels_for_multimodal off-XLA (cambrian_arch.py:387-389, :422-451, :493-609), SVA projector, one
    query group, right padding.  Returns (inputs_embeds, labels, attention_mask, position_ids, feats_final, masks_final,
    final_size, ctx_final)."""
    B = tower_feats[0].shape[0]
    q_num = cfg["image_token_len"]
    side = int(q_num ** 0.5)
    aux = [mm_projector_aux(sd, f"model.mm_projector_aux_{i}.", f) for i, f in enumerate(tower_feats)]
    ctx = aux[0].mean(1).view(B, 1, 1, -1)
    f_w, m_w = rearrange_inference(aux, side, image_sizes)
    queries = sd["model.vision_query"][0].view(1, 1, 1, -1).expand(B, q_num, -1, -1).flatten(0, 1)
    ctx_q = ctx.expand(-1, q_num, 1, -1).flatten(0, 1)
    qf = sva_sampler(sd, "model.vision_sampler_0.", queries, ctx_q, f_w, m_w, cfg["connector_depth"])
    img = mlp2x_gelu(sd, "model.mm_projector.", qf.view(B, q_num, -1)).view(B, side, side, -1)
    feats_final, masks_final = rearrange_inference(aux, side, image_sizes, unpad=True)
    per_sample, final_size, ctx_final = [], [], []
    for b in range(B):
        cur = unpad_image(img[b].unsqueeze(0), image_sizes[b])
        h, w = cur.shape[1:3]
        final_size.append((h, w))
        nl = sd["model.image_newline"].view(1, 1, 1, -1).expand(1, h, 1, -1)
        per_sample.append(torch.cat([cur, nl], 2).flatten(1, 2).squeeze(0))
        ctx_final.append(ctx[b].expand(h * w, 1, -1))
    ctx_final = torch.cat(ctx_final, 0)
    am = torch.ones_like(input_ids, dtype=torch.bool) if attention_mask is None else attention_mask.bool()
    lab = torch.full_like(input_ids, IGNORE_INDEX) if labels is None else labels
    embeds, new_labels = [], []
    for b in range(B):
        ids, lb = input_ids[b][am[b]], lab[b][am[b]]
        pos = torch.where(ids == IMAGE_TOKEN_INDEX)[0].tolist()
        if not pos:
            embeds.append(F.embedding(ids, sd["model.embed_tokens.weight"]))
            new_labels.append(lb)
            continue
        p0 = pos[0]
        e = F.embedding(torch.cat([ids[:p0], ids[p0 + 1:]]), sd["model.embed_tokens.weight"])
        embeds.append(torch.cat([e[:p0], per_sample[b].to(e.dtype), e[p0:]], 0))
        new_labels.append(torch.cat([lb[:p0], torch.full((per_sample[b].shape[0],), IGNORE_INDEX, dtype=lb.dtype),
                                     lb[p0 + 1:]]))
    L = max(e.shape[0] for e in embeds)
    H = embeds[0].shape[1]
    out_e = torch.zeros(B, L, H, dtype=embeds[0].dtype)
    out_l = torch.full((B, L), IGNORE_INDEX, dtype=lab.dtype)
    out_m = torch.zeros(B, L, dtype=torch.bool)
    out_p = torch.zeros(B, L, dtype=torch.long)
    for b, (e, l) in enumerate(zip(embeds, new_labels)):
        n = e.shape[0]
        out_e[b, :n], out_l[b, :n], out_m[b, :n] = e, l, True
        out_p[b, :n] = torch.arange(n)
    return out_e, out_l, out_m, out_p, feats_final, masks_final, final_size, ctx_final


def decoder_dynamic(sd, cfg, inputs_embeds, position_ids, attn_mask_2d, feats_w, aux_masks, ctx_q, final_size):
    """CambrianLlamaModel.forward with the per-sample SVA-insertion branch, cambrian_llama.py:208-253."""
    x = inputs_embeds
    B = x.shape[0]
    hd = cfg["hidden_size"] // cfg["num_attention_heads"]
    cos, sin = rope_cos_sin(position_ids, hd, cfg["rope_theta"])
    sites = [cfg["start_of_vision_sampler_layers"] + i * cfg["stride_of_vision_sampler_layers"]
             for i in range(cfg["num_of_vision_sampler_layers"])]
    s0 = cfg["image_position"]
    for i in range(cfg["num_hidden_layers"]):
        x = llama_layer(sd, f"model.layers.{i}.", x, cos, sin, attn_mask_2d, cfg)
        if i in sites:
            lqs, nls = [], []
            for b, (h, w) in enumerate(final_size):
                blk = x[b:b + 1, s0:s0 + h * (w + 1)].clone().view(1, h, w + 1, -1)
                lqs.append(blk[:, :, :-1].contiguous().view(h * w, 1, -1))
                nls.append(blk[:, :, -1:])
            lq = sva_sampler(sd, f"model.vision_sampler_layers.{sites.index(i)}.", torch.cat(lqs, 0), ctx_q,
                             [f.to(x.dtype) for f in feats_w], aux_masks, 1)
            x = x.clone()
            o = 0
            for b, (h, w) in enumerate(final_size):
                cur = lq[o:o + h * w].view(1, h, w, -1)
                o += h * w
                x[b:b + 1, s0:s0 + h * (w + 1)] = torch.cat([cur, nls[b]], 2).flatten(1, 2)
    return rms_norm(x, sd["model.norm.weight"], cfg["rms_norm_eps"])


# ------------------------------------------------------------------------------------------------
# A1-A4 — vision towers
# ------------------------------------------------------------------------------------------------
def _mha(x, wq, bq, wk, bk, wv, bv, wo, bo, heads):
    B, S, D = x.shape
    hd = D // heads
    q = F.linear(x, wq, bq).view(B, S, heads, hd).transpose(1, 2)
    k = F.linear(x, wk, bk).view(B, S, heads, hd).transpose(1, 2)
    v = F.linear(x, wv, bv).view(B, S, heads, hd).transpose(1, 2)
    a = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(hd), -1) @ v
    return F.linear(a.transpose(1, 2).reshape(B, S, D), wo, bo)


def bilinear_tokens(feat, target_tokens):
    """clip_encoder.py:70-96 / siglip_encoder.py:67-93 / dino_encoder.py:128-154: fp32 bilinear resize of the
    token grid (align_corners=False)."""
    B, n, C = feat.shape
    if n == target_tokens:
        return feat
    h, t = int(n ** 0.5), int(target_tokens ** 0.5)
    g = feat.view(B, h, h, C).permute(0, 3, 1, 2).float()
    g = F.interpolate(g, size=(t, t), mode="bilinear", align_corners=False).to(feat.dtype)
    return g.permute(0, 2, 3, 1).flatten(1, 2)


def clip_vit(sd, cfg, images, p="vision_model."):
    """ClipVisionTower._forward, clip_encoder.py:98-107 -> HF CLIPVisionModel, hidden_states[select_layer]
    (default -2), CLS dropped (clip_encoder.py:55-68)."""
    L = cfg["num_hidden_layers"]
    sel = cfg.get("select_layer", -2)
    n_run = L + 1 + sel if sel < 0 else sel           # hidden_states has L+1 entries; index -2 => L-1 layers
    x = F.conv2d(images, sd[p + "embeddings.patch_embedding.weight"], stride=cfg["patch_size"])
    x = x.flatten(2).transpose(1, 2)
    cls = sd[p + "embeddings.class_embedding"].expand(x.shape[0], 1, -1)
    x = torch.cat([cls, x], 1) + sd[p + "embeddings.position_embedding.weight"][None]
    x = _ln(sd, p + "pre_layrnorm", x)
    for i in range(n_run):
        q = f"{p}encoder.layers.{i}."
        h = _ln(sd, q + "layer_norm1", x)
        x = x + _mha(h, sd[q + "self_attn.q_proj.weight"], sd[q + "self_attn.q_proj.bias"],
                     sd[q + "self_attn.k_proj.weight"], sd[q + "self_attn.k_proj.bias"],
                     sd[q + "self_attn.v_proj.weight"], sd[q + "self_attn.v_proj.bias"],
                     sd[q + "self_attn.out_proj.weight"], sd[q + "self_attn.out_proj.bias"],
                     cfg["num_attention_heads"])
        h = _ln(sd, q + "layer_norm2", x)
        x = x + _lin(sd, q + "mlp.fc2", quick_gelu(_lin(sd, q + "mlp.fc1", h)))
    return bilinear_tokens(x[:, 1:], cfg.get("interp", x.shape[1] - 1))


def dinov2_pos_embed(pos, grid):
    """HF Dinov2Embeddings.interpolate_pos_encoding (installed transformers): bicubic, size-based."""
    n = pos.shape[1] - 1
    s = int(n ** 0.5)
    if s == grid:
        return pos
    pp = pos[:, 1:].reshape(1, s, s, -1).permute(0, 3, 1, 2).float()
    pp = F.interpolate(pp, size=(grid, grid), mode="bicubic", align_corners=False).to(pos.dtype)
    return torch.cat([pos[:, :1], pp.permute(0, 2, 3, 1).reshape(1, grid * grid, -1)], 1)


def dinov2_vit(sd, cfg, images):
    """DinoVisionTower._forward, dino_encoder.py:156-165 -> HF Dinov2Model.last_hidden_state[:, 1:]."""
    ps = cfg["patch_size"]
    x = F.conv2d(images, sd["embeddings.patch_embeddings.projection.weight"],
                 sd["embeddings.patch_embeddings.projection.bias"], stride=ps)
    grid = x.shape[-1]
    x = x.flatten(2).transpose(1, 2)
    x = torch.cat([sd["embeddings.cls_token"].expand(x.shape[0], -1, -1), x], 1)
    x = x + dinov2_pos_embed(sd["embeddings.position_embeddings"], grid)
    eps = cfg.get("layer_norm_eps", 1e-6)
    for i in range(cfg["num_hidden_layers"]):
        q = f"encoder.layer.{i}."
        h = _ln(sd, q + "norm1", x, eps)
        a = _mha(h, sd[q + "attention.attention.query.weight"], sd[q + "attention.attention.query.bias"],
                 sd[q + "attention.attention.key.weight"], sd[q + "attention.attention.key.bias"],
                 sd[q + "attention.attention.value.weight"], sd[q + "attention.attention.value.bias"],
                 sd[q + "attention.output.dense.weight"], sd[q + "attention.output.dense.bias"],
                 cfg["num_attention_heads"])
        x = x + a * sd[q + "layer_scale1.lambda1"]
        h = _ln(sd, q + "norm2", x, eps)
        if (q + "mlp.weights_in.weight") in sd:     # dinov2-giant: HF Dinov2SwiGLUFFN
            x1, x2 = _lin(sd, q + "mlp.weights_in", h).chunk(2, -1)
            m = _lin(sd, q + "mlp.weights_out", F.silu(x1) * x2)
        else:
            m = _lin(sd, q + "mlp.fc2", F.gelu(_lin(sd, q + "mlp.fc1", h)))
        x = x + m * sd[q + "layer_scale2.lambda1"]
    x = _ln(sd, "layernorm", x, eps)
    return bilinear_tokens(x[:, 1:], cfg.get("interp", x.shape[1] - 1))


def siglip_vit(sd, cfg, images):
    """SiglipVisionTower._forward, siglip_encoder.py:95-99 -> timm VisionTransformer.forward_features of
    vit_so400m_patch14_siglip_384 (class_token=False, learned pos, fused qkv with bias, erf-GELU, final norm,
    LN eps 1e-6).  timm parameter names.  Pinned against transformers' SiglipVisionModel (timm itself is not installed)."""
    x = F.conv2d(images, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=cfg["patch_size"])
    x = x.flatten(2).transpose(1, 2) + sd["pos_embed"]
    heads = cfg["num_attention_heads"]
    act = F.gelu if cfg.get("act", "gelu") == "gelu" else (lambda t: F.gelu(t, approximate="tanh"))
    for i in range(cfg["num_hidden_layers"]):
        q = f"blocks.{i}."
        h = _ln(sd, q + "norm1", x, 1e-6)
        wq, wk, wv = sd[q + "attn.qkv.weight"].chunk(3, 0)
        bq, bk, bv = sd[q + "attn.qkv.bias"].chunk(3, 0)
        x = x + _mha(h, wq, bq, wk, bk, wv, bv, sd[q + "attn.proj.weight"], sd[q + "attn.proj.bias"], heads)
        h = _ln(sd, q + "norm2", x, 1e-6)
        x = x + _lin(sd, q + "mlp.fc2", act(_lin(sd, q + "mlp.fc1", h)))
    x = _ln(sd, "norm", x, 1e-6)
    return bilinear_tokens(x, cfg.get("interp", x.shape[1]))


def _ln2d(sd, name, x, eps=1e-6):
    return _ln(sd, name, x.permute(0, 2, 3, 1), eps).permute(0, 3, 1, 2)


def convnext_trunk(sd, cfg, images):
    """CLIPConvNextTower._forward, clip_convnext_encoder.py:121-144 -> timm ConvNeXt stem + stages
    (convnext_xxlarge: depths 3-4-30-3, dims 384-768-1536-3072; block = dwconv7x7 -> LN -> fc1 -> GELU -> fc2
    -> gamma -> +residual; downsample = LN2d + conv2x2/2).  Returns the last stage (or all 4, multi-stage)
    bilinearly resized to the interp grid (:99-119) as [B, N, C].  Pinned against transformers' ConvNextModel (timm itself
    is not installed)."""
    x = F.conv2d(images, sd["stem.0.weight"], sd["stem.0.bias"], stride=4)
    x = _ln2d(sd, "stem.1", x)
    outs = []
    for s, depth in enumerate(cfg["depths"]):
        p = f"stages.{s}."
        if s > 0:
            x = _ln2d(sd, p + "downsample.0", x)
            x = F.conv2d(x, sd[p + "downsample.1.weight"], sd[p + "downsample.1.bias"], stride=2)
        for b in range(depth):
            q = f"{p}blocks.{b}."
            h = F.conv2d(x, sd[q + "conv_dw.weight"], sd[q + "conv_dw.bias"], padding=3, groups=x.shape[1])
            h = h.permute(0, 2, 3, 1)
            h = _ln(sd, q + "norm", h, 1e-6)
            h = _lin(sd, q + "mlp.fc2", F.gelu(_lin(sd, q + "mlp.fc1", h)))
            if (q + "gamma") in sd:
                h = h * sd[q + "gamma"]
            x = x + h.permute(0, 3, 1, 2)
        outs.append(x)
    feats = outs if cfg.get("multi_stage", False) else outs[-1:]
    t = int(cfg["interp"] ** 0.5)
    res = []
    for f in feats:
        g = F.interpolate(f.float(), size=(t, t), mode="bilinear", align_corners=False).to(f.dtype)
        res.append(g.flatten(2).transpose(1, 2))
    return torch.cat(res, -1)
