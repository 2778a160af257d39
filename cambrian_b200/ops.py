"""Raw (non-autograd) Python wrappers over the C ABI: one function per entry point of
include/cambrian_b200.h.  Tensors are torch CUDA tensors used purely as device buffers; every call
launches the hand-written sm_100a kernel on the current torch stream.  No fallbacks."""
from __future__ import annotations

import ctypes
import os

import torch

from . import _lib
from ._lib import check, int_array, ptr, ptr_array, stream

ACT = {None: 0, "none": 0, "gelu": 1, "gelu_erf": 1, "gelu_tanh": 2, "gelu_pytorch_tanh": 2,
       "quick_gelu": 3, "silu": 4}

_ws_cache: dict = {}

# NVTX ranges around every block of the hot path (decoder layer fwd / bwd, SVA layer fwd / bwd, each tower, fused loss,
# optimizer): CB_NVTX=1 turns them on for nsys / ncu --nvtx captures; off, `nvtx()` is a shared no-op context.
_NVTX = os.environ.get("CB_NVTX", "0") != "0"


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


_NULL = _Null()


class _Range:
    __slots__ = ("name",)

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        torch.cuda.nvtx.range_push(self.name)
        return self

    def __exit__(self, *a):
        torch.cuda.nvtx.range_pop()
        return False


def nvtx(name: str):
    return _Range(name) if _NVTX else _NULL


def _require_cuda_bf16(*ts):
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise _lib.CambrianB200Error("cambrian_b200 kernels need CUDA tensors (no CPU fallback)")
        if t.dtype != torch.bfloat16:
            raise ValueError(f"expected bf16 tensor, got {t.dtype}")


def require_cuda_bf16_params(params, what: str):
    """Module-level guard: every parameter must already live on the GPU in bf16 (there is no CPU / fp32 fallback)."""
    if any(p.dtype != torch.bfloat16 or not p.is_cuda for p in params):
        raise RuntimeError(f"cambrian_b200 {what} run in bf16 on CUDA only (no CPU / fp32 fallback): "
                           "call .to(device='cuda', dtype=torch.bfloat16)")


def _chk(t, name):
    _require_cuda_bf16(t)
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")


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
    # decode step: weight streaming.  The CUDA-core GEMV is FMA-bound at 6-8 rows (2.0-2.6 TB/s at M = 8): for wide outputs
    # the 128-row tcgen05 tile streams the weights faster there (measured, profiles/r02_probe_gemv.log: N = 28672: 61 vs
    # 97 us, N = 128256: 257 vs 396 us at M = 8; N <= 6144: GEMV wins at every M), so those keep the tensor-core kernel.
    if (_GEMV and M <= 8 and not (M >= 6 and N >= 16384) and not batched and not a_mn and not b_mn
            and act in (None, "none") and colscale is None
            and not accumulate and alpha == 1.0 and force_bn == 0 and K % 8 == 0 and lda % 8 == 0 and ldb % 8 == 0
            and a.data_ptr() % 16 == 0 and b.data_ptr() % 16 == 0):
        return gemv(a, b, bias=bias, residual=residual, out=out, out_dtype=out_dtype)
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


_GEMV = os.environ.get("CB_GEMV", "1") != "0"


def gemv(x: torch.Tensor, w: torch.Tensor, bias=None, residual=None, out=None, out_dtype=torch.bfloat16) -> torch.Tensor:
    """y[M, N] = x[M, K] @ w[N, K]^T (+ bias) (+ residual) for M <= 8 (the KV-cache decode step)."""
    _require_cuda_bf16(x, w, bias, residual)
    M, K = x.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=x.device)
    if out.dtype not in (torch.bfloat16, torch.float32) or out.stride(-1) != 1 or out.shape != (M, N):
        raise ValueError("gemv: out must be a bf16 / fp32 [M, N] tensor with contiguous rows")
    if residual is not None and (residual.shape != out.shape or residual.stride(-1) != 1):
        raise ValueError("gemv: residual must match the output shape")
    check(_lib.load().cb_gemv_bf16(ptr(x), ptr(w), ptr(out), M, N, K, x.stride(0), w.stride(0), out.stride(0), ptr(bias),
                                   ptr(residual), residual.stride(0) if residual is not None else 0,
                                   int(out.dtype == torch.float32), stream()), "cb_gemv_bf16")
    return out


def gemm_set_dynamic_scheduling(on: bool) -> bool:
    """Cluster-Launch-Control tile scheduling for the persistent GEMMs (default on); returns the previous setting."""
    return bool(_lib.load().cb_gemm_set_dynamic_scheduling(int(on)))


def linear(x: torch.Tensor, weight: torch.Tensor, bias=None, **kw) -> torch.Tensor:
    """y = x @ weight.T + bias for x [..., K], weight [N, K] (nn.Linear)."""
    lead = x.shape[:-1]
    x2 = x.reshape(-1, x.shape[-1])
    res = kw.pop("residual", None)
    if res is not None:
        res = res.reshape(-1, weight.shape[0])
    y = gemm(x2, weight, bias=bias, residual=res, **kw)
    return y.view(*lead, weight.shape[0])


def sva_window_attn_fwd(q, ks, vs, masks, rs, batch: int, q_side: int, need_lse: bool = True, windowed: bool = False):
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
                                            mk, int_array(rs), batch, q_side, hidden, int(windowed), stream())
    check(rc, "cb_sva_window_attn_fwd")
    return out, lse


def sva_window_attn_bwd(q, out, dout, lse, ks, vs, masks, rs, batch: int, q_side: int, windowed: bool = False,
                        dks=None, dvs=None):
    """dks / dvs: optional preallocated contiguous destinations (slabs of a batched-GEMM operand)."""
    _require_cuda_bf16(q, out, dout, *ks, *vs)
    dq = torch.empty_like(q)
    dks = [torch.empty_like(k) for k in ks] if dks is None else dks
    dvs = [torch.empty_like(v) for v in vs] if dvs is None else dvs
    for d, k in zip(list(dks) + list(dvs), list(ks) + list(vs)):
        if d.shape != k.shape or not d.is_contiguous():
            raise ValueError("sva_window_attn_bwd: dk / dv destinations must be contiguous and shaped like k / v")
    mk = None
    if masks is not None:
        masks = [None if m is None else m.contiguous().view(torch.uint8) if m.dtype == torch.bool else m
                 for m in masks]
        mk = ptr_array(masks)
    rc = _lib.load().cb_sva_window_attn_bwd(ptr(q), ptr(out), ptr(dout), ptr(lse), ptr(dq), len(ks),
                                            ptr_array(ks), ptr_array(vs), mk, ptr_array(dks), ptr_array(dvs),
                                            int_array(rs), batch, q_side, q.shape[1], int(windowed), stream())
    check(rc, "cb_sva_window_attn_bwd")
    return dq, dks, dvs


def layernorm_fwd(x, gamma, beta, eps: float = 1e-5, pos=None, side: int = 0, r: int = 0, save_stats=False, out=None):
    """out: optional preallocated contiguous [rows, C] bf16 destination (e.g. one slab of a batched-GEMM operand)."""
    _require_cuda_bf16(x, gamma, beta, pos)
    C_ = x.shape[-1]
    x2 = x.reshape(-1, C_)
    rows = x2.shape[0]
    if out is not None:
        if out.shape != x2.shape or not out.is_contiguous() or out.dtype != torch.bfloat16:
            raise ValueError("layernorm_fwd: `out` must be a contiguous bf16 [rows, C] tensor")
        y = out
    else:
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


def layernorm_bwd(dy, x, gamma, mean, rstd, pos=None, side: int = 0, r: int = 0, has_beta: bool = True, dres=None):
    _require_cuda_bf16(dy, x, gamma, pos)
    C_ = x.shape[-1]
    x2, dy2 = x.reshape(-1, C_), dy.reshape(-1, C_)
    rows = x2.shape[0]
    dx = torch.empty_like(x2)
    dgamma = torch.empty_like(gamma)
    dbeta = torch.empty_like(gamma) if has_beta else None
    nws = _lib.load().cb_norm_bwd_workspace_floats(rows, C_)
    ws = workspace(nws, x.device)
    rc = _lib.load().cb_layernorm_bwd(ptr(dy2), ptr(x2), ptr(gamma), ptr(mean), ptr(rstd), ptr(dx), ptr(dres),
                                      ptr(dgamma), ptr(dbeta), ptr(ws), ws.numel(), rows, C_, ptr(pos), side, r,
                                      stream())
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


def rmsnorm_bwd(dy, x, gamma, rstd, dres=None):
    _require_cuda_bf16(dy, x, gamma)
    C_ = x.shape[-1]
    x2, dy2 = x.reshape(-1, C_), dy.reshape(-1, C_)
    rows = x2.shape[0]
    dx = torch.empty_like(x2)
    dgamma = torch.empty_like(gamma)
    nws = _lib.load().cb_norm_bwd_workspace_floats(rows, C_)
    ws = workspace(nws, x.device)
    rc = _lib.load().cb_rmsnorm_bwd(ptr(dy2), ptr(x2), ptr(gamma), ptr(rstd), ptr(dx), ptr(dres), ptr(dgamma), ptr(ws),
                                    ws.numel(), rows, C_, stream())
    check(rc, "cb_rmsnorm_bwd")
    return dx.view(x.shape), dgamma


# ------------------------------------------------------------------------------------------------
# attention
# ------------------------------------------------------------------------------------------------
def _bshd_strides(t, hd):
    """t is a [B, S, heads, hd] view (possibly a slice of a packed QKV buffer) with unit stride on hd and
    heads packed at stride hd."""
    if t.dim() != 4 or t.stride(3) != 1 or (t.shape[2] > 1 and t.stride(2) != hd):
        raise ValueError(f"attention operand must be [B,S,heads,hd] with packed heads, got strides {t.stride()}")
    return t.stride(0), t.stride(1)


def attn_fwd(q, k, v, *, causal: bool, kmask=None, scale: float | None = None, need_lse: bool = False, out=None):
    """q [B,Sq,nh,hd], k/v [B,Skv,nkv,hd] (views are fine) -> o [B,Sq,nh,hd] contiguous (+ lse [B,nh,Sq])."""
    _require_cuda_bf16(q, k, v)
    B, Sq, nh, hd = q.shape
    Skv, nkv = k.shape[1], k.shape[2]
    if scale is None:
        scale = hd ** -0.5
    o = out if out is not None else torch.empty((B, Sq, nh, hd), dtype=torch.bfloat16, device=q.device)
    lse = torch.empty((B, nh, Sq), dtype=torch.float32, device=q.device) if need_lse else None
    qb, qs = _bshd_strides(q, hd)
    kb, ks = _bshd_strides(k, hd)
    vb, vs_ = _bshd_strides(v, hd)
    ob, os_ = _bshd_strides(o, hd)
    if kmask is not None:
        kmask = kmask.contiguous()
        kmask = kmask.view(torch.uint8) if kmask.dtype == torch.bool else kmask.to(torch.uint8)
    rc = _lib.load().cb_attn_fwd(ptr(q), ptr(k), ptr(v), ptr(o), ptr(lse), ptr(kmask), B, nh, nkv, Sq, Skv, hd,
                                 qb, qs, kb, ks, vb, vs_, ob, os_, float(scale), int(causal), stream())
    check(rc, "cb_attn_fwd")
    return (o, lse) if need_lse else o


def attn_bwd(q, k, v, o, do, lse, *, causal: bool, kmask=None, scale: float | None = None, dq=None, dk=None, dv=None):
    """Returns dq [B,Sq,nh,hd], dk, dv [B,Skv,nkv,hd] (bf16).  dq/dk/dv may be preallocated (strided views ok)."""
    _require_cuda_bf16(q, k, v, o, do)
    B, Sq, nh, hd = q.shape
    Skv, nkv = k.shape[1], k.shape[2]
    if scale is None:
        scale = hd ** -0.5
    dev = q.device
    delta = torch.empty((B, nh, Sq), dtype=torch.float32, device=dev)
    dq_acc = torch.zeros((B, Sq, nh, hd), dtype=torch.float32, device=dev)
    dk = dk if dk is not None else torch.empty((B, Skv, nkv, hd), dtype=torch.bfloat16, device=dev)
    dv = dv if dv is not None else torch.empty((B, Skv, nkv, hd), dtype=torch.bfloat16, device=dev)
    qb, qs = _bshd_strides(q, hd)
    kb, ks = _bshd_strides(k, hd)
    vb, vs_ = _bshd_strides(v, hd)
    ob, os_ = _bshd_strides(o, hd)
    dob, dos = _bshd_strides(do, hd)
    dkb, dks = _bshd_strides(dk, hd)
    dvb, dvs = _bshd_strides(dv, hd)
    if kmask is not None:
        kmask = kmask.contiguous()
        kmask = kmask.view(torch.uint8) if kmask.dtype == torch.bool else kmask.to(torch.uint8)
    rc = _lib.load().cb_attn_bwd(ptr(q), ptr(k), ptr(v), ptr(o), ptr(do), ptr(lse), ptr(delta), ptr(dq_acc), ptr(dk),
                                 ptr(dv), ptr(kmask), B, nh, nkv, Sq, Skv, hd, qb, qs, kb, ks, vb, vs_, ob, os_, dob, dos,
                                 dkb, dks, dvb, dvs, float(scale), int(causal), stream())
    check(rc, "cb_attn_bwd")
    if dq is None:
        dq = torch.empty((B, Sq, nh, hd), dtype=torch.bfloat16, device=dev)
    # dq may be a [B,Sq,nh,hd] view into a packed dQKV buffer: rows (b,s) at stride dq.stride(1)
    assert dq.stride(3) == 1 and dq.stride(2) == hd and dq.stride(0) == Sq * dq.stride(1)
    f32_to_bf16(dq_acc, dq, scale, cols=nh * hd, out_ld=dq.stride(1))
    return dq, dk, dv


# ------------------------------------------------------------------------------------------------
# elementwise / gather / reductions
# ------------------------------------------------------------------------------------------------
def act_fwd(x, act):
    _require_cuda_bf16(x)
    x = x.contiguous()
    y = torch.empty_like(x)
    check(_lib.load().cb_act_fwd(ptr(x), ptr(y), x.numel(), ACT[act], stream()), "cb_act_fwd")
    return y


def act_bwd(dy, x, act):
    _require_cuda_bf16(dy, x)
    dy, x = dy.contiguous(), x.contiguous()
    dx = torch.empty_like(x)
    check(_lib.load().cb_act_bwd(ptr(dy), ptr(x), ptr(dx), x.numel(), ACT[act], stream()), "cb_act_bwd")
    return dx


def swiglu_fwd(gate, up):
    """gate/up: [rows, I] views with a common row stride (e.g. the two halves of a fused [rows, 2I] buffer)."""
    _require_cuda_bf16(gate, up)
    rows, I = gate.shape
    assert gate.stride(1) == 1 and up.stride(1) == 1 and gate.stride(0) == up.stride(0)
    out = torch.empty((rows, I), dtype=torch.bfloat16, device=gate.device)
    check(_lib.load().cb_swiglu_fwd(ptr(gate), ptr(up), ptr(out), rows, I, gate.stride(0), I, stream()), "cb_swiglu_fwd")
    return out


def swiglu_bwd(dout, gate, up, dgate, dup):
    _require_cuda_bf16(dout, gate, up, dgate, dup)
    rows, I = gate.shape
    assert dgate.stride(0) == dup.stride(0) and dout.stride(1) == 1
    check(_lib.load().cb_swiglu_bwd(ptr(dout), ptr(gate), ptr(up), ptr(dgate), ptr(dup), rows, I, gate.stride(0),
                                    dout.stride(0), dgate.stride(0), stream()), "cb_swiglu_bwd")


def rope_(buf, pos, cos_t, sin_t, n_heads: int, hd: int, inverse: bool = False):
    """In-place RoPE on the first n_heads heads of each row of buf [rows, ld]."""
    _require_cuda_bf16(buf)
    assert buf.dim() == 2 and buf.stride(1) == 1 and pos.dtype == torch.int64 and pos.is_contiguous()
    check(_lib.load().cb_rope(ptr(buf), ptr(pos), ptr(cos_t), ptr(sin_t), buf.shape[0], n_heads, hd, buf.stride(0),
                              cos_t.shape[0], int(inverse), stream()), "cb_rope")
    return buf


def embed_splice(ids, img_start, embed, img, newline, q_side: int):
    B, S = ids.shape
    H = embed.shape[1]
    out = torch.empty((B, S, H), dtype=torch.bfloat16, device=embed.device)
    check(_lib.load().cb_embed_splice(ptr(ids), ptr(img_start), ptr(embed), ptr(img), ptr(newline), ptr(out), B, S, H,
                                      q_side, embed.shape[0], stream()), "cb_embed_splice")
    return out


def embed_splice_bwd(dout, ids, img_start, d_embed, q_side: int, has_img: bool):
    B, S, H = dout.shape
    dev = dout.device
    d_img = torch.empty((B, q_side * q_side, H), dtype=torch.bfloat16, device=dev) if has_img else None
    d_nl = torch.empty((B * q_side, H), dtype=torch.bfloat16, device=dev) if has_img else None
    vocab = d_embed.shape[0] if d_embed is not None else 0
    check(_lib.load().cb_embed_splice_bwd(ptr(dout), ptr(ids), ptr(img_start), ptr(d_embed), ptr(d_img), ptr(d_nl), B, S,
                                          H, q_side, vocab, stream()), "cb_embed_splice_bwd")
    return d_img, d_nl


def embed_grad_sorted(dout, ids, img_start, d_embed, q_side: int):
    """Deterministic embedding-row gradient: d_embed[id] = sum over the text positions holding `id` of dout rows, summed in
    position order (d_embed must be zero on entry).  The sort of 8 K token ids is device-side index plumbing (torch)."""
    B, S, H = dout.shape
    vocab = d_embed.shape[0]
    keys = ids.reshape(B, S).clamp(min=0)
    keys = torch.where(keys >= vocab, torch.zeros_like(keys), keys)
    if img_start is not None:
        span = q_side * (q_side + 1)
        pos = torch.arange(S, device=ids.device)[None]
        st = img_start.to(torch.long)[:, None]
        keys = torch.where((st >= 0) & (pos >= st) & (pos < st + span), torch.full_like(keys, vocab), keys)
    keys = keys.reshape(-1).contiguous()
    order = torch.argsort(keys, stable=True).to(torch.int32)
    check(_lib.load().cb_embed_grad_sorted(ptr(dout), ptr(keys), ptr(order), ptr(d_embed), B * S, H, vocab, stream()),
          "cb_embed_grad_sorted")


def add_pos_tokens(patch, cls, pos):
    B, N, C_ = patch.shape
    T = N + (1 if cls is not None else 0)
    out = torch.empty((B, T, C_), dtype=torch.bfloat16, device=patch.device)
    check(_lib.load().cb_add_pos_tokens(ptr(patch), ptr(cls), ptr(pos), ptr(out), B, N, C_, stream()), "cb_add_pos_tokens")
    return out


def bilinear(x, h: int, w: int, th: int, tw: int, *, in_bs=None, out=None, out_ld=None, out_col0: int = 0):
    """x: [B, >=h*w, C] token grid (first h*w rows of each batch used, e.g. after skipping CLS via a view)."""
    B, C_ = x.shape[0], x.shape[-1]
    in_bs = x.stride(0) if in_bs is None else in_bs
    if out is None:
        out = torch.empty((B, th * tw, C_), dtype=torch.bfloat16, device=x.device)
    out_ld = out.stride(1) if out_ld is None else out_ld
    check(_lib.load().cb_bilinear(ptr(x), ptr(out), B, h, w, th, tw, C_, in_bs, out.stride(0), out_ld, out_col0, stream()),
          "cb_bilinear")
    return out


def bilinear_bwd(dout, h: int, w: int, th: int, tw: int):
    """Adjoint of `bilinear` on contiguous grids: dout [B, th*tw, C] -> din [B, h*w, C]."""
    _require_cuda_bf16(dout)
    dout = dout.contiguous()
    B, C_ = dout.shape[0], dout.shape[-1]
    if dout.shape[1] != th * tw:
        raise ValueError(f"bilinear_bwd: gradient has {dout.shape[1]} tokens, expected {th} x {tw}")
    din = torch.empty((B, h * w, C_), dtype=torch.bfloat16, device=dout.device)
    check(_lib.load().cb_bilinear_bwd(ptr(dout), ptr(din), B, h, w, th, tw, C_, stream()), "cb_bilinear_bwd")
    return din


def tower_combine_fwd(logits, aggs, q_in):
    """out = q_in + sum_t softmax(logits[:, :T])[:, t] * aggs[t]  (vision_sampler.py:369-371, :396-398).
    logits [N, >= T] bf16 (extra columns are padding), aggs: T tensors [N, C], q_in [N, C]."""
    _require_cuda_bf16(logits, q_in, *aggs)
    N, C_ = q_in.shape
    if logits.shape[0] != N or logits.stride(1) != 1 or any(a.shape != q_in.shape or not a.is_contiguous() for a in aggs):
        raise ValueError("tower_combine_fwd: logits [N, Tpad] and T contiguous aggregates shaped like q_in are required")
    out = torch.empty_like(q_in)
    check(_lib.load().cb_tower_combine_fwd(ptr(logits), logits.stride(0), ptr_array(aggs), ptr(q_in), ptr(out), N, C_,
                                           len(aggs), stream()), "cb_tower_combine_fwd")
    return out


def tower_combine_bwd(logits, aggs, dout):
    """Returns (daggs: list of T [N, C], dlogits [N, Tpad] with zero padding columns); d q_in is dout itself."""
    _require_cuda_bf16(logits, dout, *aggs)
    N, C_ = dout.shape
    if not logits.is_contiguous() or not dout.is_contiguous():
        raise ValueError("tower_combine_bwd: contiguous logits / dout are required")
    daggs = [torch.empty_like(a) for a in aggs]
    dlogits = torch.empty_like(logits)
    check(_lib.load().cb_tower_combine_bwd(ptr(logits), logits.stride(0), ptr_array(aggs), ptr(dout), ptr_array(daggs),
                                           ptr(dlogits), N, C_, len(aggs), stream()), "cb_tower_combine_bwd")
    return daggs, dlogits


def patchify_nchw(img, p: int):
    B, Cin, R, _ = img.shape
    img = img.contiguous()
    K = Cin * p * p
    Kpad = (K + 7) // 8 * 8
    g = R // p
    out = torch.empty((B * g * g, Kpad), dtype=torch.bfloat16, device=img.device)
    check(_lib.load().cb_patchify_nchw(ptr(img), ptr(out), B, Cin, R, p, Kpad, stream()), "cb_patchify_nchw")
    return out


def patchify_nhwc(x, p: int):
    B, H, W, C_ = x.shape
    out = torch.empty((B * (H // p) * (W // p), p * p * C_), dtype=torch.bfloat16, device=x.device)
    check(_lib.load().cb_patchify_nhwc(ptr(x), ptr(out), B, H, W, C_, p, stream()), "cb_patchify_nhwc")
    return out


def dwconv7(x, w, bias):
    B, H, W, C_ = x.shape
    out = torch.empty_like(x)
    check(_lib.load().cb_dwconv7(ptr(x), ptr(w), ptr(bias), ptr(out), B, H, W, C_, stream()), "cb_dwconv7")
    return out


def add_(dst, src):
    assert dst.is_contiguous() and src.is_contiguous() and dst.numel() == src.numel()
    check(_lib.load().cb_add_inplace(ptr(dst), ptr(src), dst.numel(), stream()), "cb_add_inplace")
    return dst


def group_colsum(x, groups: int, scale: float = 1.0, out=None, accumulate: bool = False, fp32: bool = False):
    C_ = x.shape[-1]
    x2 = x.reshape(-1, C_)
    rpg = x2.shape[0] // groups
    if out is None:
        out = torch.empty((groups, C_), dtype=torch.float32 if fp32 else torch.bfloat16, device=x.device)
    ob, of = (None, out) if out.dtype == torch.float32 else (out, None)
    check(_lib.load().cb_group_colsum(ptr(x2), ptr(ob), ptr(of), groups, rpg, C_, float(scale), int(accumulate), stream()),
          "cb_group_colsum")
    return out


def group_broadcast(dmean, rows_per_group: int, scale: float, out=None, accumulate: bool = False):
    groups, C_ = dmean.shape
    if out is None:
        out = torch.empty((groups * rows_per_group, C_), dtype=torch.bfloat16, device=dmean.device)
    check(_lib.load().cb_group_broadcast(ptr(dmean), ptr(out), groups, rows_per_group, C_, float(scale), int(accumulate),
                                         stream()), "cb_group_broadcast")
    return out


def pos_grad(dx, B: int, side: int, r: int, out=None, accumulate: bool = False):
    C_ = dx.shape[-1]
    if out is None:
        out = torch.empty((r * r, C_), dtype=torch.bfloat16, device=dx.device)
    check(_lib.load().cb_pos_grad(ptr(dx), ptr(out), B, side, r, C_, int(accumulate), stream()), "cb_pos_grad")
    return out


def f32_to_bf16(src, dst, scale: float = 1.0, cols: int | None = None, out_ld: int | None = None):
    """src fp32 contiguous viewed as [rows, cols]; dst bf16 rows at stride out_ld."""
    cols = src.shape[-1] if cols is None else cols
    rows = src.numel() // cols
    out_ld = cols if out_ld is None else out_ld
    check(_lib.load().cb_f32_to_bf16(ptr(src), ptr(dst), rows, cols, out_ld, float(scale), stream()), "cb_f32_to_bf16")
    return dst


def cross_entropy(logits, labels, loss_rows, loss_acc, grad_scale: float, write_grad: bool, ignore_index: int = -100,
                  scale_dev=None):
    """scale_dev: optional fp32 device tensor; its element 0 multiplies grad_scale on the device (1 / #valid labels)."""
    rows, V = logits.shape
    check(_lib.load().cb_cross_entropy_ex(ptr(logits), ptr(labels), ptr(loss_rows), ptr(loss_acc), rows, V,
                                          logits.stride(0), float(grad_scale), ptr(scale_dev), int(write_grad), ignore_index,
                                          stream()), "cb_cross_entropy_ex")


def adamw(p32, m, v, g16, p16, lr, beta1, beta2, eps, wd, step: int, grad_scale: float = 1.0, clip_coef=None,
          background: bool = False):
    """clip_coef: optional fp32 DEVICE tensor whose element 0 replaces grad_scale (written by `clip_coef`);
    background: one small block per SM so the update co-resides with persistent GEMM CTAs."""
    check(_lib.load().cb_adamw_ex(ptr(p32), ptr(m), ptr(v), ptr(g16), ptr(p16), p32.numel(), float(lr), float(beta1),
                                  float(beta2), float(eps), float(wd), int(step), float(grad_scale), ptr(clip_coef),
                                  int(background), stream()), "cb_adamw_ex")


def sumsq_accumulate(g16, acc, ws, background: bool = True):
    """acc[0] += sum(g16^2) (deterministic); g16 bf16 contiguous with numel % 8 == 0; ws fp32 scratch (>= 4096)."""
    _require_cuda_bf16(g16)
    check(_lib.load().cb_sumsq_bf16(ptr(g16), g16.numel(), ptr(acc), ptr(ws), ws.numel(), int(background), stream()),
          "cb_sumsq_bf16")


def clip_coef(sumsq, max_norm: float, inv_world: float, coef):
    """coef[0] = inv_world * min(1, max_norm / (norm + 1e-6)), coef[1] = norm of the averaged gradient; resets sumsq."""
    check(_lib.load().cb_clip_coef(ptr(sumsq), float(max_norm), float(inv_world), ptr(coef), stream()), "cb_clip_coef")


def span_gather(hidden, start: int, q_side: int):
    B, S, H = hidden.shape
    lat = torch.empty((B * q_side * q_side, H), dtype=torch.bfloat16, device=hidden.device)
    check(_lib.load().cb_span_gather(ptr(hidden), ptr(lat), B, S, H, start, q_side, stream()), "cb_span_gather")
    return lat


def span_scatter_(hidden, lat, start: int, q_side: int):
    B, S, H = hidden.shape
    check(_lib.load().cb_span_scatter(ptr(hidden), ptr(lat), B, S, H, start, q_side, stream()), "cb_span_scatter")
    return hidden


def span_gather_hw(hidden, start: int, q_h: int, q_w: int):
    """Dynamic branch (cambrian_llama.py:208-253): q_h rows of (q_w latent queries + 1 newline) -> [B*q_h*q_w, H]."""
    B, S, H = hidden.shape
    _chk(hidden, "hidden")
    lat = torch.empty((B * q_h * q_w, H), dtype=torch.bfloat16, device=hidden.device)
    check(_lib.load().cb_span_gather_hw(ptr(hidden), ptr(lat), B, S, H, start, q_h, q_w, stream()), "cb_span_gather_hw")
    return lat


def span_scatter_hw_(hidden, lat, start: int, q_h: int, q_w: int):
    B, S, H = hidden.shape
    _chk(hidden, "hidden")
    _chk(lat, "lat")
    if lat.numel() != B * q_h * q_w * H:
        raise ValueError("span_scatter_hw_: lat has the wrong number of rows")
    check(_lib.load().cb_span_scatter_hw(ptr(hidden), ptr(lat), B, S, H, start, q_h, q_w, stream()), "cb_span_scatter_hw")
    return hidden


def window_gather(feat, q_side: int, crop=None):
    """feat [B, (q r)^2, C] (natural row-major grid) -> [B*h*w, r*r, C] windows of the query rows/cols in
    crop = (y0, y1, x0, x1) (default: the whole q x q grid) — cambrian_arch.py:271-330."""
    _chk(feat, "feat")
    B, N, Cc = feat.shape
    side = int(round(N ** 0.5))
    if side * side != N or side % q_side != 0:
        raise AssertionError("window_gather: token grid is not a square multiple of the query grid")   # :277
    r = side // q_side
    y0, y1, x0, x1 = crop if crop is not None else (0, q_side, 0, q_side)
    out = torch.empty((B * (y1 - y0) * (x1 - x0), r * r, Cc), dtype=torch.bfloat16, device=feat.device)
    check(_lib.load().cb_window_gather(ptr(feat), ptr(out), B, q_side, r, Cc, y0, y1, x0, x1, stream()), "cb_window_gather")
    return out


def embed_splice_ragged(embed_w, img, newline, src, batch: int, max_len: int):
    """src: int32 [batch*max_len] row map (>=0 token id, -1 zeros, INT32_MIN newline, <=-2 image row -2-src)."""
    _chk(embed_w, "embed_w")
    H = embed_w.shape[1]
    if src.dtype != torch.int32 or src.numel() != batch * max_len or not src.is_contiguous():
        raise ValueError("embed_splice_ragged: src must be contiguous int32 [batch*max_len]")
    if img is not None:
        _chk(img, "img")
    if newline is not None:
        _chk(newline, "newline")
    out = torch.empty((batch, max_len, H), dtype=torch.bfloat16, device=embed_w.device)
    check(_lib.load().cb_embed_splice_ragged(ptr(out), ptr(embed_w), ptr(img) if img is not None else None,
                                             ptr(newline) if newline is not None else None, ptr(src), batch * max_len, H,
                                             stream()), "cb_embed_splice_ragged")
    return out


def resample_coeffs(in_size: int, out_size: int):
    """Host-side Pillow-compatible bicubic coefficient tables: (bounds int32 [out, 2], kk int32 [out, ksize])."""
    lib = _lib.load()
    ks = lib.cb_resample_ksize(in_size, out_size)
    bounds = torch.empty((out_size, 2), dtype=torch.int32)
    kk = torch.empty((out_size, ks), dtype=torch.int32)
    check(lib.cb_resample_coeffs(in_size, out_size, bounds.data_ptr(), kk.data_ptr()), "cb_resample_coeffs")
    return bounds, kk


def preprocess_image(img_u8, size: int, pad_rgb, mean, std, return_u8: bool = False):
    """img_u8: CUDA uint8 [H, W, 3] RGB -> bf16 [3, size, size] = normalise(resize(expand2square(img))) with Pillow's
    exact uint8 bicubic arithmetic (mm_utils.py:186-201).  Optionally also returns the resized uint8 image."""
    if not img_u8.is_cuda or img_u8.dtype != torch.uint8 or img_u8.dim() != 3 or img_u8.shape[2] != 3:
        raise ValueError("preprocess_image: expected a CUDA uint8 [H, W, 3] tensor")
    img_u8 = img_u8.contiguous()
    H, W = int(img_u8.shape[0]), int(img_u8.shape[1])
    lib = _lib.load()
    nbytes = lib.cb_preprocess_workspace_bytes(H, W, size)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=img_u8.device)
    out = torch.empty((3, size, size), dtype=torch.bfloat16, device=img_u8.device)
    u8 = torch.empty((size, size, 3), dtype=torch.uint8, device=img_u8.device) if return_u8 else None
    pad = (ctypes.c_int32 * 3)(*[int(v) for v in pad_rgb])
    m = (ctypes.c_float * 3)(*[float(v) for v in mean])
    sd = (ctypes.c_float * 3)(*[float(v) for v in std])
    check(lib.cb_preprocess_image(ptr(img_u8), H, W, size, ctypes.addressof(pad), ctypes.addressof(m), ctypes.addressof(sd),
                                  ptr(out), ptr(u8), ptr(ws), nbytes, stream()), "cb_preprocess_image")
    return (out, u8) if return_u8 else out


def gemm_swiglu(x2d, w_gu, gu_out=None, act_out=None):
    """(gu, act) = fused gate/up projection + SwiGLU: gu = x @ [gate; up]^T [M, 2F], act = silu(gu[:, :F]) * gu[:, F:]."""
    _require_cuda_bf16(x2d, w_gu)
    M, K = x2d.shape
    F2 = w_gu.shape[0]
    if F2 % 256 or x2d.stride(1) != 1 or w_gu.stride(1) != 1:
        raise ValueError("gemm_swiglu: need contiguous rows and F % 128 == 0")
    F = F2 // 2
    gu = gu_out if gu_out is not None else torch.empty((M, F2), dtype=torch.bfloat16, device=x2d.device)
    act = act_out if act_out is not None else torch.empty((M, F), dtype=torch.bfloat16, device=x2d.device)
    check(_lib.load().cb_gemm_swiglu_bf16(ptr(x2d), ptr(w_gu), ptr(gu), ptr(act), M, F, K, x2d.stride(0), w_gu.stride(0),
                                          gu.stride(0), act.stride(0), stream()), "cb_gemm_swiglu_bf16")
    return gu, act


# CB_FUSED_SWIGLU=0 falls back to GEMM + stand-alone SwiGLU kernel (A/B switch used for profiles/, both are CUDA paths)
_FUSED_SWIGLU = os.environ.get("CB_FUSED_SWIGLU", "1") != "0"


def mlp_gate_up(h2d, w_gu):
    """gate/up projection + SwiGLU: the fused CTA-pair kernel when the problem fills the SM pairs for >= 3 waves (the
    same rule cb_gemm_bf16 uses to pick that kernel) and F % 128 == 0, otherwise GEMM followed by the SwiGLU kernel."""
    M = h2d.shape[0]
    F2 = w_gu.shape[0]
    I = F2 // 2
    pairs = max(_lib.load().cb_sm_count() // 2, 1)
    if (_FUSED_SWIGLU and F2 % 256 == 0 and ((M + 255) // 256) * (F2 // 256) >= 3 * pairs and h2d.is_contiguous()
            and w_gu.is_contiguous()):
        return gemm_swiglu(h2d, w_gu)
    gu = gemm(h2d, w_gu)
    return gu, swiglu_fwd(gu[:, :I], gu[:, I:])
