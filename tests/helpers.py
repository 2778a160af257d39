"""Shared test utilities: tolerances, tiny configurations, state-dict plumbing between the CUDA modules and the oracle."""
from __future__ import annotations

import types

import torch

# bf16 has 8 mantissa bits (eps = 2^-8 = 3.9e-3).  north_star's rtol=1e-3 / atol=1e-5 applies where the kernel can
# emit fp32 (GEMM with fp32 output, statistics, losses); bf16 tensors produced through chains of bf16-rounded stages
# are compared with a scaled max-error and a cosine criterion against the fp32 oracle.
BF16_TOL = 2.5e-2
FP32_RTOL, FP32_ATOL = 1e-3, 1e-5


def rel_err(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()


def cosine(a, b):
    a, b = a.detach().float().cpu().flatten(), b.detach().float().cpu().flatten()
    return (torch.dot(a, b) / (a.norm() * b.norm() + 1e-30)).item()


def assert_close_bf16(got, ref, what, tol=BF16_TOL, cos=0.999):
    e, c = rel_err(got, ref), cosine(got, ref)
    assert e < tol and c > cos, f"{what}: scaled max err {e:.3e} (tol {tol}), cosine {c:.6f}"


def sd_cpu32(module, prefix=""):
    return {prefix + k: v.detach().float().cpu() for k, v in module.state_dict().items()}


def tiny_cambrian_config(connector_only=False, sva=True, kv_last=2):
    """A small Cambrian config with the real structure: 4 towers (tiny), SVA connector depth 2, 2 in-LLM SVA sites."""
    from cambrian_b200.model.language_model.cambrian_llama import CambrianConfig
    cfg = CambrianConfig(hidden_size=256, intermediate_size=512, num_hidden_layers=4, num_attention_heads=4,
                         num_key_value_heads=2, vocab_size=1024, max_position_embeddings=512, rope_theta=500000.0,
                         rms_norm_eps=1e-5)
    q = 4  # 4 x 4 = 16 queries
    cfg.image_token_len = q * q
    cfg.mm_vision_tower_aux_list = ["siglip/CLIP-ViT-SO400M-14-384", "openai/clip-vit-large-patch14-336",
                                    "facebook/dinov2-large-res56", "clip-convnext-XXL-multi-stage-res128"]
    cfg.mm_vision_tower_aux_token_len_list = [q * q, q * q, q * q, (q * kv_last) ** 2]
    cfg.siglip_config_overrides = dict(hidden_size=288, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4,
                                       image_size=56)
    cfg.clip_config_overrides = dict(hidden_size=256, intermediate_size=512, num_hidden_layers=3, num_attention_heads=4,
                                     image_size=56)
    cfg.convnext_config_overrides = dict(depths=(1, 1, 2, 1), dims=(64, 128, 256, 512), image_size=128)
    cfg.mm_projector_type = "sva" if sva else "mlp2x_gelu"
    cfg.vision_hidden_size = 1024
    cfg.num_query_group = 1
    cfg.query_num_list = [q * q]
    cfg.connector_depth = 2
    cfg.connector_only = connector_only
    cfg.num_of_vision_sampler_layers = 2
    cfg.start_of_vision_sampler_layers = 0
    cfg.stride_of_vision_sampler_layers = 2
    cfg.image_position = 5
    cfg.fused_lm_loss = False
    return cfg


def tower_image_sizes(cfg):
    return [56, 56, 56, 128]


def oracle_cfg(cfg):
    return dict(hidden_size=cfg.hidden_size, num_attention_heads=cfg.num_attention_heads,
                num_key_value_heads=cfg.num_key_value_heads, num_hidden_layers=cfg.num_hidden_layers,
                rms_norm_eps=cfg.rms_norm_eps, rope_theta=rope_theta(cfg), image_token_len=cfg.image_token_len,
                connector_depth=cfg.connector_depth, connector_only=cfg.connector_only,
                num_of_vision_sampler_layers=cfg.num_of_vision_sampler_layers,
                start_of_vision_sampler_layers=cfg.start_of_vision_sampler_layers,
                stride_of_vision_sampler_layers=cfg.stride_of_vision_sampler_layers, image_position=cfg.image_position)


def rope_theta(cfg):
    t = getattr(cfg, "rope_theta", None)
    return t if t is not None else (getattr(cfg, "rope_parameters", None) or {}).get("rope_theta", 10000.0)


def ns(**kw):
    return types.SimpleNamespace(**kw)
