"""GPU probe 2 (development tool): attention fwd/bwd + elementwise kernels against torch references."""
import math
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from cambrian_b200 import ops  # noqa: E402

dev = "cuda"
torch.manual_seed(0)


def rel_err(a, b):
    a, b = a.float(), b.float()
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def attn_ref(q, k, v, causal, kmask):
    B, Sq, nh, hd = q.shape
    Skv, nkv = k.shape[1], k.shape[2]
    Q = q.float().transpose(1, 2)
    K = k.float().transpose(1, 2).repeat_interleave(nh // nkv, 1)
    V = v.float().transpose(1, 2).repeat_interleave(nh // nkv, 1)
    s = Q @ K.transpose(-1, -2) / math.sqrt(hd)
    allow = torch.ones(Sq, Skv, dtype=torch.bool, device=q.device)
    if causal:
        allow = allow.tril(Skv - Sq)
    allow = allow[None, None]
    if kmask is not None:
        allow = allow & kmask[:, None, None, :]
    s = s.masked_fill(~allow, float("-inf"))
    p = torch.softmax(s, -1)
    p = torch.nan_to_num(p, 0.0)
    return (p @ V).transpose(1, 2)


def attn_case(B, Sq, Skv, nh, nkv, hd, causal, use_mask, bwd, packed=False):
    if packed:
        W = (nh + 2 * nkv) * hd
        qkv = torch.randn(B, Sq, W, device=dev).bfloat16()
        q = qkv[..., : nh * hd].view(B, Sq, nh, hd)
        k = qkv[..., nh * hd:(nh + nkv) * hd].view(B, Sq, nkv, hd)
        v = qkv[..., (nh + nkv) * hd:].view(B, Sq, nkv, hd)
    else:
        q = torch.randn(B, Sq, nh, hd, device=dev).bfloat16()
        k = torch.randn(B, Skv, nkv, hd, device=dev).bfloat16()
        v = torch.randn(B, Skv, nkv, hd, device=dev).bfloat16()
    kmask = None
    if use_mask:
        kmask = torch.rand(B, Skv, device=dev) > 0.2
        kmask[:, 0] = True
    qf, kf, vf = (t.detach().float().requires_grad_() for t in (q, k, v))
    ref = attn_ref(qf, kf, vf, causal, kmask)
    o, lse = ops.attn_fwd(q, k, v, causal=causal, kmask=kmask, need_lse=True)
    torch.cuda.synchronize()
    errs = [rel_err(o, ref)]
    if bwd:
        do = torch.randn(ref.shape, device=dev).bfloat16()
        ref.backward(do.float())
        dq, dk, dv = ops.attn_bwd(q, k, v, o, do, lse, causal=causal, kmask=kmask)
        torch.cuda.synchronize()
        errs += [rel_err(dq, qf.grad), rel_err(dk, kf.grad), rel_err(dv, vf.grad)]
    ok = all(e < 2e-2 for e in errs)
    print(f"attn B={B} Sq={Sq} Skv={Skv} nh={nh} nkv={nkv} hd={hd} causal={causal} mask={use_mask} packed={packed} "
          f"errs={[f'{e:.2e}' for e in errs]} {'OK' if ok else 'FAIL'}", flush=True)


def group_attn_fwd():
    attn_case(1, 128, 128, 1, 1, 64, False, False, False)
    attn_case(1, 128, 128, 1, 1, 128, False, False, False)
    attn_case(2, 256, 256, 2, 2, 64, False, False, False)
    attn_case(2, 577, 577, 16, 16, 64, False, False, False)
    attn_case(1, 729, 729, 16, 16, 72, False, False, False)
    attn_case(1, 512, 512, 4, 4, 128, True, False, False)
    attn_case(2, 1024, 1024, 8, 2, 128, True, True, False)
    attn_case(1, 300, 300, 4, 1, 128, True, True, False, packed=True)
    attn_case(1, 2048, 2048, 32, 8, 128, True, False, False, packed=True)


def group_attn_bwd():
    attn_case(1, 128, 128, 1, 1, 128, False, False, True)
    attn_case(1, 128, 128, 1, 1, 64, False, False, True)
    attn_case(1, 256, 256, 2, 1, 128, True, False, True)
    attn_case(2, 512, 512, 4, 4, 128, True, True, True)
    attn_case(1, 300, 300, 4, 2, 128, True, True, True, packed=True)
    attn_case(1, 1024, 1024, 8, 2, 128, True, False, True, packed=True)
    attn_case(1, 2048, 2048, 32, 8, 128, True, True, True, packed=True)


def group_attn_perf():
    for (B, S, nh, nkv, hd, causal) in [(4, 2048, 32, 8, 128, True), (8, 577, 16, 16, 64, False),
                                        (8, 729, 16, 16, 72, False), (4, 2048, 32, 32, 128, True)]:
        q = torch.randn(B, S, nh, hd, device=dev).bfloat16()
        k = torch.randn(B, S, nkv, hd, device=dev).bfloat16()
        v = torch.randn(B, S, nkv, hd, device=dev).bfloat16()
        fl = 4.0 * B * nh * S * S * hd * (0.5 if causal else 1.0)
        ms = timeit(lambda: ops.attn_fwd(q, k, v, causal=causal))
        print(f"perf attn fwd B={B} S={S} nh={nh} nkv={nkv} hd={hd} causal={causal}: {ms:.3f} ms "
              f"{fl / ms / 1e9:.1f} TFLOP/s", flush=True)
        ms2 = timeit(lambda: F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2).repeat_interleave(nh // nkv, 1),
                                                            v.transpose(1, 2).repeat_interleave(nh // nkv, 1), is_causal=causal))
        print(f"     torch sdpa: {ms2:.3f} ms {fl / ms2 / 1e9:.1f} TFLOP/s", flush=True)
        if hd in (64, 128):
            o, lse = ops.attn_fwd(q, k, v, causal=causal, need_lse=True)
            ms = timeit(lambda: ops.attn_bwd(q, k, v, o, o, lse, causal=causal), iters=10)
            print(f"perf attn bwd: {ms:.3f} ms {2.5 * fl / ms / 1e9:.1f} TFLOP/s", flush=True)


def group_elem():
    # activations
    for act, f in [("gelu", F.gelu), ("quick_gelu", lambda x: x * torch.sigmoid(1.702 * x)), ("silu", F.silu),
                   ("gelu_tanh", lambda x: F.gelu(x, approximate="tanh"))]:
        x = torch.randn(1000, 1024, device=dev).bfloat16()
        dy = torch.randn_like(x)
        xf = x.float().requires_grad_()
        ref = f(xf)
        ref.backward(dy.float())
        e1, e2 = rel_err(ops.act_fwd(x, act), ref), rel_err(ops.act_bwd(dy, x, act), xf.grad)
        print(f"act {act} errs={e1:.2e},{e2:.2e} {'OK' if max(e1, e2) < 1e-2 else 'FAIL'}", flush=True)
    # swiglu on a fused buffer
    rows, I = 777, 1024
    gu = torch.randn(rows, 2 * I, device=dev).bfloat16()
    g, u = gu[:, :I], gu[:, I:]
    gf, uf = g.float().requires_grad_(), u.float().requires_grad_()
    ref = F.silu(gf) * uf
    dout = torch.randn(rows, I, device=dev).bfloat16()
    ref.backward(dout.float())
    out = ops.swiglu_fwd(g, u)
    dgu = torch.empty_like(gu)
    ops.swiglu_bwd(dout, g, u, dgu[:, :I], dgu[:, I:])
    errs = [rel_err(out, ref), rel_err(dgu[:, :I], gf.grad), rel_err(dgu[:, I:], uf.grad)]
    print(f"swiglu errs={[f'{e:.2e}' for e in errs]} {'OK' if max(errs) < 1.5e-2 else 'FAIL'}", flush=True)
    # rope
    B, S, nh, nkv, hd = 2, 300, 8, 2, 128
    W = (nh + 2 * nkv) * hd
    qkv = torch.randn(B * S, W, device=dev).bfloat16()
    pos = torch.randint(0, 2048, (B * S,), device=dev)
    inv = 1.0 / (500000.0 ** (torch.arange(0, hd, 2, device=dev).float() / hd))
    ang = torch.arange(2048, device=dev).float()[:, None] * inv[None]
    cos_t, sin_t = ang.cos().contiguous(), ang.sin().contiguous()
    x = qkv[:, : (nh + nkv) * hd].float().view(B * S, nh + nkv, hd)
    c = torch.cat([cos_t[pos], cos_t[pos]], -1)[:, None].bfloat16().float()
    s = torch.cat([sin_t[pos], sin_t[pos]], -1)[:, None].bfloat16().float()
    rot = torch.cat([-x[..., hd // 2:], x[..., : hd // 2]], -1)
    ref = x * c + rot * s
    got = qkv.clone()
    ops.rope_(got, pos, cos_t, sin_t, nh + nkv, hd)
    e1 = rel_err(got[:, : (nh + nkv) * hd].view(B * S, nh + nkv, hd), ref)
    e0 = rel_err(got[:, (nh + nkv) * hd:], qkv[:, (nh + nkv) * hd:])
    back = got.clone()
    ops.rope_(back, pos, cos_t, sin_t, nh + nkv, hd, inverse=True)
    e2 = rel_err(back, qkv)
    print(f"rope fwd={e1:.2e} v-untouched={e0:.2e} inverse-roundtrip={e2:.2e} "
          f"{'OK' if e1 < 1e-2 and e0 == 0 and e2 < 3e-2 else 'FAIL'}", flush=True)
    # embed splice
    B, S, H, q, V = 2, 100, 256, 4, 500
    ids = torch.randint(3, V, (B, S), device=dev)
    start = torch.tensor([7, 30], device=dev, dtype=torch.int32)
    for b in range(B):
        ids[b, start[b]] = -200
        ids[b, start[b] + 1: start[b] + q * (q + 1)] = 0
    embed = torch.randn(V, H, device=dev).bfloat16()
    img = torch.randn(B, q * q, H, device=dev).bfloat16()
    nl = torch.randn(H, device=dev).bfloat16()
    out = ops.embed_splice(ids, start, embed, img, nl, q)
    ref = embed[torch.where(ids < 0, 0, ids)].clone()
    for b in range(B):
        blk = torch.cat([img[b].view(q, q, H), nl.view(1, 1, H).expand(q, 1, H)], 1).flatten(0, 1)
        ref[b, start[b]: start[b] + q * (q + 1)] = blk
    e = rel_err(out, ref)
    dout = torch.randn(B, S, H, device=dev).bfloat16()
    d_embed = torch.zeros(V, H, device=dev).bfloat16()
    d_img, d_nl = ops.embed_splice_bwd(dout, ids, start, d_embed, q, True)
    ref_de = torch.zeros(V, H, device=dev)
    ref_dimg = torch.zeros(B, q * q, H, device=dev)
    ref_dnl = torch.zeros(H, device=dev)
    for b in range(B):
        for s_ in range(S):
            st = int(start[b])
            if st <= s_ < st + q * (q + 1):
                kk = s_ - st
                r_, c_ = divmod(kk, q + 1)
                if c_ == q:
                    ref_dnl += dout[b, s_].float()
                else:
                    ref_dimg[b, r_ * q + c_] = dout[b, s_].float()
            else:
                ref_de[max(int(ids[b, s_]), 0)] += dout[b, s_].float()
    dnl = ops.group_colsum(d_nl, 1)
    errs = [e, rel_err(d_embed, ref_de), rel_err(d_img, ref_dimg), rel_err(dnl[0], ref_dnl)]
    print(f"embed_splice errs={[f'{x:.2e}' for x in errs]} {'OK' if max(errs) < 2e-2 else 'FAIL'}", flush=True)
    # add_pos_tokens
    patch = torch.randn(3, 576, 1024, device=dev).bfloat16()
    cls = torch.randn(1024, device=dev).bfloat16()
    pos = torch.randn(577, 1024, device=dev).bfloat16()
    ref = torch.cat([cls.float().expand(3, 1, -1), patch.float()], 1) + pos.float()[None]
    e1 = rel_err(ops.add_pos_tokens(patch, cls, pos), ref)
    e2 = rel_err(ops.add_pos_tokens(patch, None, pos[:576]), patch.float() + pos[:576].float()[None])
    print(f"add_pos_tokens errs={e1:.2e},{e2:.2e} {'OK' if max(e1, e2) < 1e-2 else 'FAIL'}", flush=True)
    # bilinear 27->24 and 37->24 (with CLS skip via view)
    for h, t in [(27, 24), (37, 24), (32, 24), (10, 24)]:
        x = torch.randn(2, h * h + 1, 256, device=dev).bfloat16()
        ref = F.interpolate(x[:, 1:].float().view(2, h, h, 256).permute(0, 3, 1, 2), size=(t, t), mode="bilinear",
                            align_corners=False).permute(0, 2, 3, 1).flatten(1, 2)
        got = ops.bilinear(x[:, 1:], h, h, t, t)
        print(f"bilinear {h}->{t} err={rel_err(got, ref):.2e} {'OK' if rel_err(got, ref) < 1e-2 else 'FAIL'}", flush=True)
    # patchify nchw (conv14 equivalence through the GEMM) and nhwc
    img = torch.randn(2, 3, 56, 56, device=dev).bfloat16()
    w = torch.randn(64, 3, 14, 14, device=dev).bfloat16()
    ref = F.conv2d(img.float(), w.float(), stride=14).flatten(2).transpose(1, 2).reshape(-1, 64)
    pt = ops.patchify_nchw(img, 14)
    wf = torch.zeros(64, pt.shape[1], device=dev).bfloat16()
    wf[:, :588] = w.view(64, -1)
    got = ops.gemm(pt, wf, out_dtype=torch.float32)
    e1 = rel_err(got, ref)
    x = torch.randn(2, 8, 8, 64, device=dev).bfloat16()
    w2 = torch.randn(128, 64, 2, 2, device=dev).bfloat16()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w2.float(), stride=2).permute(0, 2, 3, 1).reshape(-1, 128)
    got = ops.gemm(ops.patchify_nhwc(x, 2), w2.permute(0, 2, 3, 1).reshape(128, -1).contiguous(), out_dtype=torch.float32)
    e2 = rel_err(got, ref)
    print(f"patchify nchw={e1:.2e} nhwc={e2:.2e} {'OK' if max(e1, e2) < 1e-3 else 'FAIL'}", flush=True)
    # dwconv7 (ragged H / W, partial channel chunk, multi-step column strips)
    for (bb, hh, ww, cc) in ((2, 12, 12, 384), (1, 13, 11, 200), (2, 40, 24, 128), (1, 64, 64, 1536)):
        x = torch.randn(bb, hh, ww, cc, device=dev).bfloat16()
        w = torch.randn(cc, 1, 7, 7, device=dev).bfloat16()
        b = torch.randn(cc, device=dev).bfloat16()
        ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b.float(), padding=3, groups=cc).permute(0, 2, 3, 1)
        got = ops.dwconv7(x, w.view(cc, 49).t().contiguous().view(7, 7, cc), b)
        print(f"dwconv7 {bb}x{hh}x{ww}x{cc} err={rel_err(got, ref):.2e} {'OK' if rel_err(got, ref) < 1e-2 else 'FAIL'}", flush=True)
    x = torch.randn(4, 64, 64, 1536, device=dev).bfloat16()
    w = torch.randn(7, 7, 1536, device=dev).bfloat16()
    b = torch.randn(1536, device=dev).bfloat16()
    for _ in range(3):
        ops.dwconv7(x, w, b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.dwconv7(x, w, b)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"perf dwconv7 4x64x64x1536: {ms * 1e3:.1f} us  {2 * x.numel() * 2 / ms / 1e6:.0f} GB/s  {x.numel() * 49 * 2 / ms / 1e9:.2f} TFLOP/s fp32", flush=True)
    # dynamic-shape helpers: window gather (+ crop), rectangular span gather / scatter, ragged embed splice
    f = torch.randn(2, 144, 64, device=dev).bfloat16()                 # 12 x 12 grid, q = 4 -> r = 3
    ref = f.view(2, 4, 3, 4, 3, 64).permute(0, 1, 3, 2, 4, 5).contiguous()
    e1 = (ops.window_gather(f, 4).float() - ref.flatten(0, 2).flatten(1, 2).float()).abs().max().item()
    e2 = (ops.window_gather(f[1:2].contiguous(), 4, (1, 3, 0, 4)).float()
          - ref[1:2, 1:3].flatten(0, 2).flatten(1, 2).float()).abs().max().item()
    print(f"window_gather full={e1:.1e} crop={e2:.1e} {'OK' if max(e1, e2) == 0 else 'FAIL'}", flush=True)
    hid = torch.randn(2, 50, 64, device=dev).bfloat16()
    lat = ops.span_gather_hw(hid, 5, 3, 4)                             # 3 rows of (4 queries + newline)
    blk = hid[:, 5:5 + 15].view(2, 3, 5, 64)
    e1 = (lat.float() - blk[:, :, :4].reshape(-1, 64).float()).abs().max().item()
    h2 = hid.clone()
    new = torch.randn_like(lat)
    ops.span_scatter_hw_(h2, new, 5, 3, 4)
    want = hid.clone()
    want[:, 5:20] = torch.cat([new.view(2, 3, 4, 64), blk[:, :, 4:]], 2).flatten(1, 2)
    e2 = (h2.float() - want.float()).abs().max().item()
    print(f"span_hw gather={e1:.1e} scatter={e2:.1e} {'OK' if max(e1, e2) == 0 else 'FAIL'}", flush=True)
    emb = torch.randn(100, 64, device=dev).bfloat16()
    img = torch.randn(10, 64, device=dev).bfloat16()
    nl = torch.randn(64, device=dev).bfloat16()
    src = torch.tensor([3, 99, -2, -11, -(2 ** 31), -1, 0, -5], dtype=torch.int32, device=dev)
    got = ops.embed_splice_ragged(emb, img, nl, src, 2, 4).view(8, 64)
    want = torch.stack([emb[3], emb[99], img[0], img[9], nl, torch.zeros_like(nl), emb[0], img[3]])
    e1 = (got.float() - want.float()).abs().max().item()
    print(f"embed_splice_ragged err={e1:.1e} {'OK' if e1 == 0 else 'FAIL'}", flush=True)
    # reductions
    x = torch.randn(4 * 576, 1024, device=dev).bfloat16()
    e1 = rel_err(ops.group_colsum(x, 4, 1.0 / 576), x.float().view(4, 576, 1024).mean(1))
    e2 = rel_err(ops.group_colsum(x, 1, fp32=True), x.float().sum(0, keepdim=True))
    dm = torch.randn(4, 1024, device=dev).bfloat16()
    e3 = rel_err(ops.group_broadcast(dm, 576, 1.0 / 576), (dm.float() / 576)[:, None].expand(4, 576, 1024).reshape(-1, 1024))
    a, bb = torch.randn(4096, device=dev).bfloat16(), torch.randn(4096, device=dev).bfloat16()
    ref = a.float() + bb.float()
    e4 = rel_err(ops.add_(a, bb), ref)
    side, r, C, B = 8, 4, 256, 3
    dx = torch.randn(B * side * side, C, device=dev).bfloat16()
    idx = torch.arange(side * side, device=dev)
    pidx = ((idx // side) % r) * r + (idx % side) % r
    ref = torch.zeros(r * r, C, device=dev).index_add_(0, pidx.repeat(B), dx.float())
    e5 = rel_err(ops.pos_grad(dx, B, side, r), ref)
    print(f"reductions errs={e1:.2e},{e2:.2e},{e3:.2e},{e4:.2e},{e5:.2e} "
          f"{'OK' if max(e1, e2, e3, e4, e5) < 1e-2 else 'FAIL'}", flush=True)
    # cross entropy
    for rows, V in ((300, 32000), (192, 1024), (64, 128256)):  # incl. V/8 < blockDim (idle threads) and the Llama-3 vocab
        logits = (3 * torch.randn(rows, V, device=dev)).bfloat16()
        labels = torch.randint(0, V, (rows,), device=dev)
        labels[::7] = -100
        lf = logits.float().requires_grad_()
        ref = F.cross_entropy(lf, labels, ignore_index=-100, reduction="sum")
        ref.backward()
        loss_rows = torch.empty(rows, device=dev)
        acc = torch.zeros(2, device=dev)
        buf = logits.clone()
        ops.cross_entropy(buf, labels, loss_rows, acc, 1.0, True)
        torch.cuda.synchronize()
        e1 = abs(acc[0].item() - ref.item()) / ref.item()
        e2 = rel_err(buf, lf.grad)
        cnt = int((labels != -100).sum())
        print(f"cross_entropy loss_err={e1:.2e} grad_err={e2:.2e} count={acc[1].item()}/{cnt} "
              f"{'OK' if e1 < 1e-4 and e2 < 1e-2 and acc[1].item() == cnt else 'FAIL'}", flush=True)
    # adamw vs torch
    n = 1 << 16
    p = torch.randn(n, device=dev)
    g = torch.randn(n, device=dev).bfloat16()
    tp = torch.nn.Parameter(p.clone())
    opt = torch.optim.AdamW([tp], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)
    m, v, p16 = torch.zeros(n, device=dev), torch.zeros(n, device=dev), torch.empty(n, device=dev, dtype=torch.bfloat16)
    for step in (1, 2, 3):
        tp.grad = g.float()
        opt.step()
        ops.adamw(p, m, v, g, p16, 1e-3, 0.9, 0.999, 1e-8, 0.01, step)
    torch.cuda.synchronize()
    e1 = rel_err(p, tp.data)
    print(f"adamw err={e1:.2e} p16_err={rel_err(p16, p):.2e} {'OK' if e1 < 1e-5 else 'FAIL'}", flush=True)
    x = torch.randn(8 * 2048, 14336 * 2, device=dev).bfloat16()
    ms = timeit(lambda: ops.swiglu_fwd(x[:, :14336], x[:, 14336:]))
    print(f"perf swiglu fwd: {ms * 1e3:.0f} us {3 * 16384 * 14336 * 2 / ms / 1e6:.0f} GB/s", flush=True)


def group_attn_prof():
    B, S, nh, nkv, hd = 1, 2048, 32, 8, 128
    q = torch.randn(B, S, nh, hd, device=dev).bfloat16()
    k = torch.randn(B, S, nkv, hd, device=dev).bfloat16()
    v = torch.randn(B, S, nkv, hd, device=dev).bfloat16()
    for _ in range(2):
        o, lse = ops.attn_fwd(q, k, v, causal=True, need_lse=True)
        ops.attn_bwd(q, k, v, o, o, lse, causal=True)
    torch.cuda.synchronize()


if __name__ == "__main__":
    t0 = time.time()
    globals()["group_" + sys.argv[1]]()
    print(f"group {sys.argv[1]} done in {time.time() - t0:.1f}s", flush=True)
