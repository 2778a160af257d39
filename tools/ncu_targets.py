"""Launches the dominant kernels once each at benchmark shapes, for `ncu --set full` captures (profiles/README.md):
  1. gemm_bf16_tcgen05_2cta  — decoder gate/up projection  M=8192 N=28672 K=4096 (B=4 x S=2048 rows)
  2. sva_window_attn_fwd     — BASELINE grids [576]x4 and release grids [576,576,576,9216], batch 32 (inputs > L2)
  3. attn_fwd / attn_bwd     — Llama-3-8B shape, B=4 S=2048 nh=32 nkv=8 hd=128, causal
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cambrian_b200 import ops  # noqa: E402

dev = "cuda"
torch.manual_seed(0)
a = torch.randn(8192, 4096, device=dev).bfloat16()
w = torch.randn(28672, 4096, device=dev).bfloat16()
out = torch.empty(8192, 28672, device=dev, dtype=torch.bfloat16)
act = torch.empty(8192, 14336, device=dev, dtype=torch.bfloat16)
for _ in range(2):
    ops.gemm(a, w, out=out)                       # heuristic -> CTA-pair kernel
    ops.gemm(a, w, out=out, force_bn=256)         # single-CTA kernel for comparison
    ops.gemm_swiglu(a, w, out, act)               # fused gate/up + SwiGLU (what the decoder layer launches)
for rs in ([1, 1, 1, 1], [1, 1, 1, 4]):
    B, q = 32, 24
    qq = torch.randn(B * q * q, 1024, device=dev).bfloat16()
    ks = [torch.randn(B, (r * q) ** 2, 1024, device=dev).bfloat16() for r in rs]
    vs = [torch.randn(B, (r * q) ** 2, 1024, device=dev).bfloat16() for r in rs]
    for _ in range(2):
        o, lse = ops.sva_window_attn_fwd(qq, ks, vs, None, rs, B, q)
    ops.sva_window_attn_bwd(qq, o, o, lse, ks, vs, None, rs, B, q)
B, S, nh, nkv, hd = 4, 2048, 32, 8, 128
qkv = torch.randn(B, S, (nh + 2 * nkv) * hd, device=dev).bfloat16()
qv = qkv[..., : nh * hd].view(B, S, nh, hd)
kv = qkv[..., nh * hd:(nh + nkv) * hd].view(B, S, nkv, hd)
vv = qkv[..., (nh + nkv) * hd:].view(B, S, nkv, hd)
for _ in range(2):
    o, lse = ops.attn_fwd(qv, kv, vv, causal=True, need_lse=True)
    ops.attn_bwd(qv, kv, vv, o, o, lse, causal=True)
torch.cuda.synchronize()
print("done")
