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


# --------------------------------------------------------------------------------------------------------------------
# Parity criterion (VERDICT r1 "weak #1"): the reference runs its eager PyTorch path in bf16, so its own results differ
# from exact arithmetic by some err(eager-bf16, fp32).  A CUDA tensor passes when its error against the fp32 oracle is
# no larger than SLACK x that yardstick, per tensor, in two norms:
#     fro  = ||got - ref||_F / ||ref||_F          (robust; the asserted headline, slack 1.5)
#     maxs = max|got - ref| / max|ref|            (ONE outlier element of up to 1e8 decides it: slack 3.0; measured
#                                                  worst 2.1 in 1023 tensor comparisons, profiles/r02_parity_*.md)
# The yardstick is floored at the error of ONE bf16 rounding of an exact result (fro 2^-9/sqrt(3) ~ 1.1e-3, max 2^-9):
# tensors the eager path happens to produce exactly must not demand more than bf16 storage can give.
# Tensors the kernels emit in fp32 from identical inputs (losses, fp32 GEMM outputs, statistics) are held to
# north_star's rtol 1e-3 / atol 1e-5 directly (FP32_RTOL / FP32_ATOL below).
# Every comparison is appended to gpurun_out/parity_report.jsonl (copied to profiles/ per round).
# --------------------------------------------------------------------------------------------------------------------
FRO_FLOOR, MAX_FLOOR = 1.5e-3, 4e-3
FRO_SLACK, MAX_SLACK = 1.5, 3.0
_REPORT = None


def fro_err(a, b):
    a, b = a.detach().double().flatten().cpu(), b.detach().double().flatten().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def _report(rec):
    global _REPORT
    import json
    import os
    if _REPORT is None:
        d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
        try:
            os.makedirs(d, exist_ok=True)
            _REPORT = open(os.path.join(d, "parity_report.jsonl"), "a")
        except OSError:
            _REPORT = False
    if _REPORT:
        _REPORT.write(json.dumps(rec) + "\n")
        _REPORT.flush()


def parity(got, ref32, eager, what, fro_slack=FRO_SLACK, max_slack=MAX_SLACK):
    """Returns None when `got` (CUDA path) is as close to the fp32 oracle as the eager-bf16 oracle is (see above), else a
    message.  `eager` may be None for tensors without an eager counterpart: the floors alone apply (x slack)."""
    ec_f, ec_m = fro_err(got, ref32), rel_err(got, ref32)
    ee_f, ee_m = (fro_err(eager, ref32), rel_err(eager, ref32)) if eager is not None else (0.0, 0.0)
    lim_f, lim_m = fro_slack * max(ee_f, FRO_FLOOR), max_slack * max(ee_m, MAX_FLOOR)
    ok = ec_f <= lim_f and ec_m <= lim_m
    _report(dict(what=what, cuda_fro=ec_f, eager_fro=ee_f, cuda_max=ec_m, eager_max=ee_m, limit_fro=lim_f,
                 limit_max=lim_m, ok=bool(ok), numel=int(ref32.numel())))
    if ok:
        return None
    return (f"{what}: cuda fro {ec_f:.3e} (eager {ee_f:.3e}, limit {lim_f:.3e}), cuda max {ec_m:.3e} "
            f"(eager {ee_m:.3e}, limit {lim_m:.3e})")


def assert_parity(got, ref32, eager, what, **kw):
    msg = parity(got, ref32, eager, what, **kw)
    assert msg is None, msg


class ParityCollector:
    """Collect every tensor comparison of a test and fail once at the end with all offenders (a GPU call is expensive:
    one run must show every mismatch, not the first)."""

    def __init__(self):
        self.bad = []

    def check(self, got, ref32, eager, what, **kw):
        msg = parity(got, ref32, eager, what, **kw)
        if msg is not None:
            self.bad.append(msg)

    def done(self):
        assert not self.bad, "parity failures:\n  " + "\n  ".join(self.bad)


def assert_same_training(master, master_ref, lr, steps, what):
    """Two runs of the same training recipe (different optimizer schedules) on the fp32 master weights.  Gradients are not
    bit-reproducible from run to run (dQ partials are reduce-added by TMA in arrival order), and an AdamW step moves every
    element by ~lr * sign(g): an element whose gradient sits at the rounding-noise level can land one step apart.  So: no
    element further apart than the 2 * lr * steps an optimizer can move it, and the tensors equal in the Frobenius sense —
    a schedule bug (a bucket updated twice, not at all, or from stale gradients) moves whole buckets by ~lr per element,
    i.e. >= 5e-2 relative on the 0.02-rms weights."""
    d = (master.float() - master_ref.float()).abs()
    fro = (d.double().norm() / master_ref.double().norm()).item()
    _report(dict(what=f"{what}: fp32 master", max_abs=d.max().item(), fro=fro,
                 frac_gt_0p1lr=(d > 0.1 * lr).float().mean().item(), bound_abs=2 * lr * steps))
    assert d.max().item() <= 2.05 * lr * steps, (what, d.max().item())
    assert fro < 1e-2, (what, fro)


def oracle_device():
    """Device the (torch) oracle runs on in `-m gpu` tests: the GPU, so full-size fp32 / eager-bf16 oracles take
    seconds.  TF32 is disabled so 'fp32' means fp32."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return torch.device("cuda" if torch.cuda.is_available() else "cpu")


def both_modes(fn, sd32, *tensors, device=None):
    """Run an oracle function `fn(sd, *tensors)` twice on `device`: fp32 (truth) and eager bf16 (the reference's numerics).
    Floating tensors are cast to the mode's dtype, everything else is only moved."""
    from oracle import cambrian_oracle as O
    device = device or oracle_device()
    outs = []
    for dt in (torch.float32, torch.bfloat16):
        sd = O.to_device(sd32 if dt == torch.float32 else O.eager_bf16(sd32), device)
        ts = [t.to(device=device, dtype=dt) if torch.is_tensor(t) and t.is_floating_point() else
              (t.to(device) if torch.is_tensor(t) else t) for t in tensors]
        outs.append(fn(sd, *ts))
    return outs


def bf(t):
    """round to bf16, keep fp32 storage: inputs every arm (CUDA, fp32 oracle, eager oracle) can represent exactly"""
    return t.bfloat16().float()


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
                stride_of_vision_sampler_layers=cfg.stride_of_vision_sampler_layers, image_position=cfg.image_position,
                query_num_list=list(getattr(cfg, "query_num_list", [cfg.image_token_len])))


def rope_theta(cfg):
    t = getattr(cfg, "rope_theta", None)
    return t if t is not None else (getattr(cfg, "rope_parameters", None) or {}).get("rope_theta", 10000.0)


def ns(**kw):
    return types.SimpleNamespace(**kw)
