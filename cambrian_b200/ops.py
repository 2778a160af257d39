"""Raw (non-autograd) Python wrappers over the C ABI: one function per entry point of
include/cambrian_b200.h.  Tensors are torch CUDA tensors used purely as device buffers; every call
launches the hand-written sm_100a kernel on the current torch stream.  No fallbacks."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check, int_array, ptr, ptr_array, stream

ACT = {None: 0, "none": 0, "gelu": 1, "gelu_erf": 1, "gelu_tanh": 2, "gelu_pytorch_tanh": 2,
       "quick_gelu": 3, "silu": 4}

_ws_cache: dict = {}


def _require_cuda_bf16(*ts):
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise _lib.CambrianB200Error("cambrian_b200 kernels need CUDA tensors (no CPU fallback)")
        if t.dtype != torch.bfloat16:
            raise ValueError(f"expected bf16 tensor, got {t.dtype}")


def workspace(nfloats: int, device) -> torch.Tensor:
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    w = _ws_cache.get(key)
    if w is None or w.numel() < nfloats:
        w = torch.empty(max(nfloats, 1 << 20), dtype=torch.float32, device=device)
        _ws_cache[key] = w
    return w


def gemm(a: torch.Tensor, b: torch.Tensor, *, a_mn: bool = False, b_mn: bool = False, bias=None,
         colscale=None, residual=None, act=None, alpha: float = 1.0, out: torch.Tensor | None = None,
         out_dtype=torch.bfloat16, accumulate: bool = False, force_bn: int = 0) -> torch.Tensor:
    """C = epi(alpha * opA(a) @ opB(b)).  2-D or batched 3-D (leading batch dim on every operand).

    a: [M, K] (a_mn=False) or [K, M] (a_mn=True);  b: [N, K] (b_mn=False, nn.Linear layout) or [K, N].
    Row strides may exceed the logical width (ld), the last dim must be contiguous.
    """
    _require_cuda_bf16(a, b, bias, colscale, residual)
    batched = a.dim() == 3
    if batched != (b.dim() == 3):
        raise ValueError("gemm: a and b must both be 2-D or both 3-D")

    def dims(t):
        if t.stride(-1) != 1:
            raise ValueError("gemm: innermost dimension must be contiguous")
        if batched:
            return t.shape[0], t.shape[1], t.shape[2], t.stride(1), t.stride(0)
        return 1, t.shape[0], t.shape[1], t.stride(0), 0

    ba, ar, ac, lda, bsa = dims(a)
    bb, br, bc, ldb, bsb = dims(b)
    M, K = (ac, ar) if a_mn else (ar, ac)
    N, Kb = (bc, br) if b_mn else (br, bc)
    if K != Kb or ba != bb:
        raise ValueError(f"gemm: shape mismatch a={tuple(a.shape)} b={tuple(b.shape)} a_mn={a_mn} b_mn={b_mn}")
    if out is None:
        if accumulate:
            raise ValueError("gemm: accumulate=True needs an explicit `out`")
        shape = (ba, M, N) if batched else (M, N)
        out = torch.empty(shape, dtype=out_dtype, device=a.device)
    if out.dtype not in (torch.bfloat16, torch.float32) or out.stride(-1) != 1:
        raise ValueError("gemm: out must be bf16/fp32 with contiguous last dim")
    ldc = out.stride(-2)
    bsc = out.stride(0) if batched else 0
    ldr = bsr = 0
    if residual is not None:
        if residual.shape != out.shape or residual.stride(-1) != 1:
            raise ValueError("gemm: residual must match the output shape")
        ldr = residual.stride(-2)
        bsr = residual.stride(0) if batched else 0
    rc = _lib.load().cb_gemm_bf16(ptr(a), ptr(b), ptr(out), M, N, K, ba, lda, ldb, ldc, bsa, bsb, bsc,
                                  int(a_mn), int(b_mn), ptr(bias), ptr(colscale), ptr(residual), ldr, bsr,
                                  float(alpha), ACT[act], int(out.dtype == torch.float32), int(accumulate),
                                  int(force_bn), stream())
    check(rc, "cb_gemm_bf16")
    return out


def linear(x: torch.Tensor, weight: torch.Tensor, bias=None, **kw) -> torch.Tensor:
    """y = x @ weight.T + bias for x [..., K], weight [N, K] (nn.Linear)."""
    lead = x.shape[:-1]
    x2 = x.reshape(-1, x.shape[-1])
    res = kw.pop("residual", None)
    if res is not None:
        res = res.reshape(-1, weight.shape[0])
    y = gemm(x2, weight, bias=bias, residual=res, **kw)
    return y.view(*lead, weight.shape[0])


def sva_window_attn_fwd(q, ks, vs, masks, rs, batch: int, q_side: int, need_lse: bool = True):
    _require_cuda_bf16(q, *ks, *vs)
    n, hidden = q.shape
    out = torch.empty_like(q)
    lse = torch.empty((n, 16), dtype=torch.float32, device=q.device) if need_lse else None
    mk = None
    if masks is not None:
        masks = [None if m is None else m.contiguous().view(torch.uint8) if m.dtype == torch.bool else m
                 for m in masks]
        mk = ptr_array(masks)
    rc = _lib.load().cb_sva_window_attn_fwd(ptr(q), ptr(out), ptr(lse), len(ks), ptr_array(ks), ptr_array(vs),
                                            mk, int_array(rs), batch, q_side, hidden, stream())
    check(rc, "cb_sva_window_attn_fwd")
    return out, lse


def sva_window_attn_bwd(q, out, dout, lse, ks, vs, masks, rs, batch: int, q_side: int):
    _require_cuda_bf16(q, out, dout, *ks, *vs)
    dq = torch.empty_like(q)
    dks = [torch.empty_like(k) for k in ks]
    dvs = [torch.empty_like(v) for v in vs]
    mk = None
    if masks is not None:
        masks = [None if m is None else m.contiguous().view(torch.uint8) if m.dtype == torch.bool else m
                 for m in masks]
        mk = ptr_array(masks)
    rc = _lib.load().cb_sva_window_attn_bwd(ptr(q), ptr(out), ptr(dout), ptr(lse), ptr(dq), len(ks),
                                            ptr_array(ks), ptr_array(vs), mk, ptr_array(dks), ptr_array(dvs),
                                            int_array(rs), batch, q_side, q.shape[1], stream())
    check(rc, "cb_sva_window_attn_bwd")
    return dq, dks, dvs


def layernorm_fwd(x, gamma, beta, eps: float = 1e-5, pos=None, side: int = 0, r: int = 0, save_stats=False):
    _require_cuda_bf16(x, gamma, beta, pos)
    C_ = x.shape[-1]
    x2 = x.reshape(-1, C_)
    rows = x2.shape[0]
    y = torch.empty_like(x2)
    mean = rstd = None
    if save_stats:
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
    rc = _lib.load().cb_layernorm_fwd(ptr(x2), ptr(gamma), ptr(beta), ptr(y), ptr(mean), ptr(rstd), rows, C_,
                                      float(eps), ptr(pos), side, r, stream())
    check(rc, "cb_layernorm_fwd")
    y = y.view(x.shape)
    return (y, mean, rstd) if save_stats else y


def layernorm_bwd(dy, x, gamma, mean, rstd, pos=None, side: int = 0, r: int = 0, has_beta: bool = True):
    _require_cuda_bf16(dy, x, gamma, pos)
    C_ = x.shape[-1]
    x2, dy2 = x.reshape(-1, C_), dy.reshape(-1, C_)
    rows = x2.shape[0]
    dx = torch.empty_like(x2)
    dgamma = torch.empty_like(gamma)
    dbeta = torch.empty_like(gamma) if has_beta else None
    nws = _lib.load().cb_norm_bwd_workspace_floats(rows, C_)
    ws = workspace(nws, x.device)
    rc = _lib.load().cb_layernorm_bwd(ptr(dy2), ptr(x2), ptr(gamma), ptr(mean), ptr(rstd), ptr(dx), ptr(dgamma),
                                      ptr(dbeta), ptr(ws), ws.numel(), rows, C_, ptr(pos), side, r, stream())
    check(rc, "cb_layernorm_bwd")
    return dx.view(x.shape), dgamma, dbeta


def rmsnorm_fwd(x, gamma, eps: float = 1e-6, hf_cast: bool = False, save_stats=False):
    _require_cuda_bf16(x, gamma)
    C_ = x.shape[-1]
    x2 = x.reshape(-1, C_)
    rows = x2.shape[0]
    y = torch.empty_like(x2)
    rstd = torch.empty(rows, dtype=torch.float32, device=x.device) if save_stats else None
    rc = _lib.load().cb_rmsnorm_fwd(ptr(x2), ptr(gamma), ptr(y), ptr(rstd), rows, C_, float(eps), int(hf_cast),
                                    stream())
    check(rc, "cb_rmsnorm_fwd")
    y = y.view(x.shape)
    return (y, rstd) if save_stats else y


def rmsnorm_bwd(dy, x, gamma, rstd):
    _require_cuda_bf16(dy, x, gamma)
    C_ = x.shape[-1]
    x2, dy2 = x.reshape(-1, C_), dy.reshape(-1, C_)
    rows = x2.shape[0]
    dx = torch.empty_like(x2)
    dgamma = torch.empty_like(gamma)
    nws = _lib.load().cb_norm_bwd_workspace_floats(rows, C_)
    ws = workspace(nws, x.device)
    rc = _lib.load().cb_rmsnorm_bwd(ptr(dy2), ptr(x2), ptr(gamma), ptr(rstd), ptr(dx), ptr(dgamma), ptr(ws),
                                    ws.numel(), rows, C_, stream())
    check(rc, "cb_rmsnorm_bwd")
    return dx.view(x.shape), dgamma
