"""TEST INFRASTRUCTURE ONLY — plain-torch stand-ins for a subset of `cambrian_b200.ops`, installed by monkeypatching inside
a CPU test process (tests/test_autograd_blocks_cpu.py) so that the HOST logic of the autograd blocks — argument order,
saved tensors, gradient routing, main_grad accumulation, layout conventions — runs in the `-m "not gpu"` suite.

This is NOT a fallback: nothing under `cambrian_b200/` imports this file, the product still raises without the CUDA
library / CUDA tensors (tests/test_abi_cpu.py), and the numerical parity claims rest on the `-m gpu` tests alone.  Each
stand-in computes in fp32 and rounds once to bf16, like the kernels it mirrors (include/cambrian_b200.h documents each).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

_ACT = {"gelu": F.gelu, "quick_gelu": lambda x: x * torch.sigmoid(1.702 * x), "silu": F.silu,
        "gelu_tanh": lambda x: F.gelu(x, approximate="tanh")}


def gemm(a, b, *, a_mn=False, b_mn=False, bias=None, colscale=None, residual=None, out=None, out_dtype=torch.bfloat16,
         accumulate=False, alpha=1.0, act=None, force_bn=0):
    A = a.float().transpose(-1, -2) if a_mn else a.float()
    B = b.float() if b_mn else b.float().transpose(-1, -2)
    y = (A @ B) * alpha
    if bias is not None:
        y = y + bias.float()
    if act not in (None, "none"):
        y = _ACT[act](y)
    if colscale is not None:
        y = y * colscale.float()
    if residual is not None:
        assert residual.shape == y.shape, (residual.shape, y.shape)
        y = y + residual.float()
    if out is not None:
        assert out.shape == y.shape, (out.shape, y.shape)
        if accumulate:
            y = y + out.float()
        out.copy_(y.to(out.dtype))
        return out
    assert not accumulate, "accumulate=True needs an explicit `out`"
    return y.to(out_dtype)


def linear(x, weight, bias=None, **kw):
    lead = x.shape[:-1]
    res = kw.pop("residual", None)
    if res is not None:
        res = res.reshape(-1, weight.shape[0])
    return gemm(x.reshape(-1, x.shape[-1]), weight, bias=bias, residual=res, **kw).view(*lead, weight.shape[0])


def f32_to_bf16(src, dst, scale=1.0, cols=None, out_ld=None):
    dst.copy_((src.float() * scale).reshape(dst.shape).to(dst.dtype))
    return dst


def _window_pos(pos, rows, side, r):
    """pos_embed row added to each latent: natural layout (row = (b, y, x) of a side x side grid) -> window position
    (y % r) * r + (x % r); window-rearranged layout (side == 0) -> row % r^2."""
    if pos is None:
        return 0
    idx = torch.arange(rows)
    if side == 0:
        w = idx % (r * r)
    else:
        w = ((idx // side) % side % r) * r + (idx % side) % r
    return pos.float()[w]


def layernorm_fwd(x, gamma, beta, eps=1e-5, pos=None, side=0, r=0, save_stats=False, out=None):
    C = x.shape[-1]
    xp = x.reshape(-1, C).float() + _window_pos(pos, x.numel() // C, side, r)
    mean, var = xp.mean(-1), xp.var(-1, unbiased=False)
    rstd = torch.rsqrt(var + eps)
    y = ((xp - mean[:, None]) * rstd[:, None] * gamma.float() + (beta.float() if beta is not None else 0)).to(torch.bfloat16)
    if out is not None:
        out.copy_(y)
        y = out
    y = y.view(x.shape)
    return (y, mean, rstd) if save_stats else y


def layernorm_bwd(dy, x, gamma, mean, rstd, pos=None, side=0, r=0, has_beta=True, dres=None):
    C = x.shape[-1]
    dy2 = dy.reshape(-1, C).float()
    xp = x.reshape(-1, C).float() + _window_pos(pos, x.numel() // C, side, r)
    xh = (xp - mean[:, None]) * rstd[:, None]
    g = dy2 * gamma.float()
    dx = rstd[:, None] * (g - g.mean(-1, keepdim=True) - xh * (g * xh).mean(-1, keepdim=True))
    if dres is not None:
        dx = dx + dres.reshape(-1, C).float()
    return (dx.to(torch.bfloat16).view(x.shape), (dy2 * xh).sum(0).to(torch.bfloat16),
            dy2.sum(0).to(torch.bfloat16) if has_beta else None)


def _win(t, batch, q_side, r, windowed, n):
    if windowed:
        return t.reshape(n, r * r, -1)
    return t.reshape(batch, q_side, r, q_side, r, -1).permute(0, 1, 3, 2, 4, 5).reshape(n, r * r, -1)


def _sva(q, ks, vs, masks, rs, batch, q_side, windowed):
    n = q.shape[0]
    K = torch.cat([_win(k, batch, q_side, r, windowed, n) for k, r in zip(ks, rs)], 1).view(n, -1, 16, 64).transpose(1, 2)
    V = torch.cat([_win(v, batch, q_side, r, windowed, n) for v, r in zip(vs, rs)], 1).view(n, -1, 16, 64).transpose(1, 2)
    ms = [torch.ones(n, r * r, dtype=torch.bool) if masks is None or masks[i] is None else masks[i].reshape(n, -1).bool()
          for i, r in enumerate(rs)]
    s = (q.view(n, 1, 16, 64).transpose(1, 2) @ K.transpose(-1, -2)) / 8.0
    s = s.masked_fill(~torch.cat(ms, 1)[:, None, None, :], float("-inf"))
    return (torch.softmax(s, -1) @ V).transpose(1, 2).reshape(n, 1024)


def sva_window_attn_fwd(q, ks, vs, masks, rs, batch, q_side, need_lse=True, windowed=False):
    out = _sva(q.float(), [k.float() for k in ks], [v.float() for v in vs], masks, rs, batch, q_side, windowed)
    return out.to(torch.bfloat16), torch.zeros(q.shape[0], 16)


def sva_window_attn_bwd(q, out, dout, lse, ks, vs, masks, rs, batch, q_side, windowed=False, dks=None, dvs=None):
    qf = q.float().requires_grad_()
    kf = [k.float().requires_grad_() for k in ks]
    vf = [v.float().requires_grad_() for v in vs]
    with torch.enable_grad():
        o = _sva(qf, kf, vf, masks, rs, batch, q_side, windowed)
    gs = [g.to(torch.bfloat16) for g in torch.autograd.grad(o, [qf] + kf + vf, dout.float())]
    T = len(ks)
    gk, gv = gs[1:1 + T], gs[1 + T:]
    if dks is not None:
        for d, g in zip(list(dks) + list(dvs), gk + gv):
            d.copy_(g)
        return gs[0], dks, dvs
    return gs[0], gk, gv


def act_fwd(x, act):
    return _ACT[act](x.float()).to(torch.bfloat16)


def act_bwd(dy, x, act):
    xf = x.float().requires_grad_()
    with torch.enable_grad():
        y = _ACT[act](xf)
    return torch.autograd.grad(y, xf, dy.float())[0].to(torch.bfloat16)


def tower_combine_fwd(logits, aggs, q_in):
    T = len(aggs)
    w = torch.softmax(logits.float()[:, :T], -1)
    return (q_in.float() + sum(w[:, t:t + 1] * aggs[t].float() for t in range(T))).to(torch.bfloat16)


def tower_combine_bwd(logits, aggs, dout):
    T = len(aggs)
    w, d = torch.softmax(logits.float()[:, :T], -1), dout.float()
    g = torch.stack([(d * a.float()).sum(-1) for a in aggs], 1)
    dl = torch.zeros(logits.shape, dtype=torch.float32)
    dl[:, :T] = w * (g - (w * g).sum(-1, keepdim=True))
    return [(w[:, t:t + 1] * d).to(torch.bfloat16) for t in range(T)], dl.to(torch.bfloat16)


def pos_grad(dx, B, side, r, out=None, accumulate=False):
    C = dx.shape[-1]
    return dx.float().reshape(B, side // r, r, side // r, r, C).sum((0, 1, 3)).reshape(r * r, C).to(torch.bfloat16)


def bilinear(x, h, w, th, tw, *, in_bs=None, out=None, out_ld=None, out_col0=0):
    B, C = x.shape[0], x.shape[-1]
    y = F.interpolate(x[:, :h * w].float().reshape(B, h, w, C).permute(0, 3, 1, 2), size=(th, tw), mode="bilinear",
                      align_corners=False)
    return y.permute(0, 2, 3, 1).reshape(B, th * tw, C).to(torch.bfloat16)


def bilinear_bwd(dout, h, w, th, tw):
    B, C = dout.shape[0], dout.shape[-1]
    x = torch.zeros(B, C, h, w, requires_grad=True)
    with torch.enable_grad():
        y = F.interpolate(x, size=(th, tw), mode="bilinear", align_corners=False)
    g = torch.autograd.grad(y, x, dout.float().reshape(B, th, tw, C).permute(0, 3, 1, 2))[0]
    return g.permute(0, 2, 3, 1).reshape(B, h * w, C).to(torch.bfloat16)


def group_colsum(x, groups, scale=1.0, out=None, accumulate=False, fp32=False):
    y = x.float().reshape(groups, -1, x.shape[-1]).sum(1) * scale
    return y if fp32 else y.to(torch.bfloat16)


def group_broadcast(dmean, rows_per_group, scale, out=None, accumulate=False):
    G, C = dmean.shape
    return (dmean.float() * scale)[:, None, :].expand(G, rows_per_group, C).reshape(G * rows_per_group, C).to(torch.bfloat16)


def add_(dst, src):
    dst.add_(src)
    return dst


def adamw(p32, m, v, g16, p16, lr, beta1, beta2, eps, wd, step, grad_scale=1.0, clip_coef=None, background=False):
    """adamw_kernel (elementwise.cu): torch.optim.AdamW arithmetic on fp32 master / moments, bf16 gradients in, bf16 copy out."""
    gs = float(clip_coef[0]) if clip_coef is not None else grad_scale
    g = g16.float() * gs
    bc1, bc2 = 1.0 - beta1 ** step, 1.0 - beta2 ** step
    p32.mul_(1.0 - lr * wd)
    m.mul_(beta1).add_(g, alpha=1.0 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1.0 - beta2)
    p32.sub_((lr / bc1) * m / (v.sqrt() / bc2 ** 0.5 + eps))
    p16.copy_(p32.to(torch.bfloat16))


def sumsq_accumulate(g16, acc, ws, background=True):
    acc[0] += g16.float().pow(2).sum()


def clip_coef(sumsq, max_norm, inv_world, coef):
    norm = sumsq[0].sqrt() * inv_world
    coef[0] = inv_world * min(1.0, max_norm / (float(norm) + 1e-6))
    coef[1] = norm
    sumsq[0] = 0.0


def require_cuda_bf16_params(params, what):
    if any(p.dtype != torch.bfloat16 for p in params):
        raise RuntimeError(f"cambrian_b200 {what} run in bf16")


_NAMES = ("gemm", "linear", "f32_to_bf16", "layernorm_fwd", "layernorm_bwd", "sva_window_attn_fwd", "sva_window_attn_bwd",
          "act_fwd", "act_bwd", "tower_combine_fwd", "tower_combine_bwd", "pos_grad", "bilinear", "bilinear_bwd",
          "group_colsum", "group_broadcast", "add_", "require_cuda_bf16_params", "adamw", "sumsq_accumulate", "clip_coef")


class _Setter:
    """monkeypatch look-alike for spawned worker processes (the process ends with the test)."""

    @staticmethod
    def setattr(obj, name, value):
        setattr(obj, name, value)


def install(monkeypatch=None):
    """Replace the emulated entry points of `cambrian_b200.ops` for the duration of one test (pytest's monkeypatch undoes
    it; a spawned worker passes nothing); every other op keeps raising without the CUDA library."""
    monkeypatch = monkeypatch or _Setter
    from cambrian_b200 import ops
    for n in _NAMES:
        monkeypatch.setattr(ops, n, globals()[n])
    monkeypatch.setattr(ops, "_require_cuda_bf16", lambda *a: None)
