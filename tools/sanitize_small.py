"""Small-shape pass through every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool memcheck  --error-exitcode 1 python tools/sanitize_small.py
    compute-sanitizer --tool racecheck --error-exitcode 1 python tools/sanitize_small.py

Shapes are tiny so the 10-50x sanitizer slowdown stays within a minute; every launch still goes through the same code paths
(TMA producer / tcgen05 issuer / TMEM epilogue roles, CTA-pair cluster, Cluster Launch Control scheduling, mbarrier rings)."""
import sys

import torch

sys.path.insert(0, ".")
from cambrian_b200 import ops  # noqa: E402

dev = "cuda"
torch.manual_seed(0)


def r(*s):
    return torch.randn(*s, device=dev).bfloat16()


def main():
    n = 0
    # GEMM: single-CTA tiles, CTA pair, every major, epilogues, batched, accumulate, CLC on and off
    for clc in (False, True):
        ops.gemm_set_dynamic_scheduling(clc)
        for (M, N, K, a_mn, b_mn, bn) in [(256, 128, 128, False, False, 64), (384, 256, 192, False, True, 128),
                                          (256, 512, 128, True, True, 256), (512, 512, 256, False, False, 512),
                                          (20 * 128, 1024, 128, False, False, 64), (2560, 2560, 128, True, False, 512)]:
            a = r(K, M) if a_mn else r(M, K)
            b = r(K, N) if b_mn else r(N, K)
            ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, force_bn=bn)
            ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, force_bn=bn, out_dtype=torch.float32)
            ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, force_bn=bn, out=r(M, N), accumulate=True)
            n += 3
        ops.gemm(r(256, 128), r(256, 128), bias=r(256), act="gelu", residual=None)
        ops.gemm(r(3, 256, 128), r(3, 256, 128))
        ops.gemm_swiglu(r(512, 128), r(2 * 256, 128))
        n += 3
    # attention fwd / bwd (hd 64 / 72->80 / 128, causal + GQA + key mask)
    for (B, S, nh, nkv, hd, causal) in [(1, 200, 4, 4, 64, False), (1, 130, 2, 2, 72, False), (2, 300, 4, 2, 128, True)]:
        q, k, v = r(B, S, nh, hd), r(B, S, nkv, hd), r(B, S, nkv, hd)
        km = torch.ones(B, S, dtype=torch.bool, device=dev)
        km[:, S - 7:] = False
        o, lse = ops.attn_fwd(q, k, v, causal=causal, kmask=km, need_lse=True)
        n += 1
        if hd in (64, 128):
            ops.attn_bwd(q, k, v, o, r(B, S, nh, hd), lse, causal=causal, kmask=km)
            n += 1
    # SVA window attention fwd / bwd (natural layout, r > 1, masks)
    Bq, qs, rs = 2, 4, [1, 2, 1]
    qq = r(Bq * qs * qs, 1024)
    ks = [r(Bq, (x * qs) ** 2, 1024) for x in rs]
    vs = [r(Bq, (x * qs) ** 2, 1024) for x in rs]
    masks = [torch.rand(Bq * qs * qs, x * x, device=dev) > 0.2 for x in rs]
    for m in masks:
        m[m.sum(1) == 0] = True
    out, lse = ops.sva_window_attn_fwd(qq, ks, vs, masks, rs, Bq, qs)
    ops.sva_window_attn_bwd(qq, out, r(*out.shape), lse, ks, vs, masks, rs, Bq, qs)
    n += 2
    # norms
    x = r(300, 1024)
    y, mean, rstd = ops.layernorm_fwd(x, r(1024), r(1024), save_stats=True)
    ops.layernorm_bwd(r(300, 1024), x, r(1024), mean, rstd)
    y, rstd = ops.rmsnorm_fwd(x, r(1024), save_stats=True)
    ops.rmsnorm_bwd(r(300, 1024), x, r(1024), rstd)
    n += 4
    # elementwise / optimizer
    ops.act_fwd(x, "gelu")
    ops.swiglu_fwd(r(64, 256), r(64, 256))
    p, m, v = torch.randn(4096, device=dev), torch.zeros(4096, device=dev), torch.zeros(4096, device=dev)
    coef = torch.ones(2, device=dev)
    acc, ws = torch.zeros(1, device=dev), torch.empty(4096, device=dev)
    g = r(4096)
    ops.sumsq_accumulate(g, acc, ws)
    ops.clip_coef(acc, 1.0, 1.0, coef)
    ops.adamw(p, m, v, g, torch.empty(4096, device=dev, dtype=torch.bfloat16), 1e-3, 0.9, 0.999, 1e-8, 0.0, 1, clip_coef=coef,
              background=True)
    n += 5
    torch.cuda.synchronize()
    print(f"sanitize_small: {n} launches groups done")


if __name__ == "__main__":
    main()
