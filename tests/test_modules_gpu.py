"""GPU parity tests of the drop-in modules against the CPU oracle (oracle/cambrian_oracle.py) on identical weights and
seeded inputs: SVA sampler (both layouts, fwd + bwd), the four towers, the LLaMA decoder layer (fwd + bwd), and the full
tiny Cambrian model (loss + parameter gradients + greedy generate token ids)."""
import pytest
import torch

from helpers import (FP32_RTOL, ParityCollector, assert_same_training, bf, both_modes, ns, oracle_cfg, rel_err, sd_cpu32,
                     tiny_cambrian_config, tower_image_sizes)

pytestmark = pytest.mark.gpu
dev = "cuda"

from cambrian_b200 import ops  # noqa: E402


def _cuda_bf16(m):
    return m.to(device=dev, dtype=torch.bfloat16)


@pytest.mark.parametrize("rs,use_mask", [([1, 1, 1, 1], False), ([1, 1, 1, 2], True), ([2, 1, 3], True)])
def test_sva_sampler_matches_oracle(rs, use_mask):
    from cambrian_b200.model.vision_sampler import VisionTokenSampler
    from oracle import cambrian_oracle as O
    torch.manual_seed(0)
    B, q, D, depth = 2, 4, 256, 2
    T = len(rs)
    m = _cuda_bf16(VisionTokenSampler(D, 1024, [1024] * T, rs, 1024, depth))
    for layer in m.layers:  # give pos_embed a realistic scale
        for i, r in enumerate(rs):
            if r > 1:
                getattr(layer, f"pos_embed_{i}").data.mul_(0.1)
    sd = sd_cpu32(m)
    n = B * q * q
    feats = [torch.randn(B, (r * q) ** 2, 1024) for r in rs]
    queries = torch.randn(n, 1, D)
    ctx = torch.randn(B, 1, 1024).expand(B, q * q, 1024).reshape(n, 1, 1024)
    masks = None
    if use_mask:
        masks = []
        for r in rs:
            mk = torch.rand(n, r * r) > 0.3
            mk[mk.sum(1) == 0] = True
            masks.append(mk)
    names = list(sd.keys())
    dout = torch.randn(n, 1, D)

    def run(sdd, qq, cx, dy, *fm):
        sdd = {k: v.detach().requires_grad_() for k, v in sdd.items()}
        qq = qq.detach().requires_grad_()
        fs = [f.detach().requires_grad_() for f in fm[:T]]
        out = O.sva_sampler(sdd, "", qq, cx, [O.window_rearrange(f, q) for f in fs], None if masks is None else list(fm[T:]),
                            depth)
        return out.detach(), torch.autograd.grad(out, [qq, *fs, *[sdd[k] for k in names]], dy)

    (ref, gref), (eag, geag) = both_modes(run, sd, bf(queries), bf(ctx), bf(dout), *[bf(f) for f in feats], *(masks or []))
    # ---- CUDA, natural layout (fast path)
    qg = queries.to(dev).bfloat16().requires_grad_()
    fg = [f.to(dev).bfloat16().requires_grad_() for f in feats]
    mg = [None] * T if masks is None else [mk.to(dev) for mk in masks]
    out = m(qg, ctx.to(dev).bfloat16(), *fg, *mg, natural_layout=(B, q))
    out.backward(dout.to(dev).bfloat16())
    pc = ParityCollector()
    tag = f"sva rs={rs}"
    pc.check(out, ref, eag, f"{tag}: forward (natural)")
    pc.check(qg.grad, gref[0], geag[0], f"{tag}: dqueries")
    for i in range(T):
        pc.check(fg[i].grad, gref[1 + i], geag[1 + i], f"{tag}: dfeats[{i}]")
    params = dict(m.named_parameters())
    for j, k in enumerate(names):
        pc.check(params[k].grad, gref[1 + T + j], geag[1 + T + j], f"{tag}: grad {k}")
    # ---- CUDA, reference call convention (window-rearranged latents): same numbers
    m.zero_grad()
    fw = [O.window_rearrange(bf(f), q).to(dev).bfloat16() for f in feats]
    out_w = m(queries.to(dev).bfloat16(), ctx.to(dev).bfloat16(), *fw, *mg)
    pc.check(out_w, ref, eag, f"{tag}: forward (window-rearranged API)")
    pc.done()
    assert rel_err(out_w, out) < 1e-2


@pytest.mark.parametrize("rs,use_mask,natural", [([2, 1, 3], True, True), ([1, 1, 2], False, False), ([2], True, True)])
def test_sva_sep_sampler_matches_oracle(rs, use_mask, natural):
    """layer_type="sep" (VisionAggregationLayer, vision_sampler.py:330-405): output and every gradient vs the oracle
    restatement (itself pinned to the live reference, tests/test_oracle_pin.py), natural-layout and reference (windowed)
    call conventions, with and without masks, T = 1 (no weight_mlp) included."""
    from cambrian_b200.model.vision_sampler import VisionTokenSampler
    from oracle import cambrian_oracle as O
    torch.manual_seed(1)
    B, q, D, depth = 2, 4, 256, 2
    T = len(rs)
    m = _cuda_bf16(VisionTokenSampler(D, 1024, [1024] * T, rs, 1024, depth, layer_type="sep"))
    for layer in m.layers:
        for i, r in enumerate(rs):
            if r > 1:
                getattr(layer, f"pos_embed_{i}").data.mul_(0.1)
    sd = sd_cpu32(m)
    n = B * q * q
    feats = [torch.randn(B, (r * q) ** 2, 1024) for r in rs]
    queries = torch.randn(n, 1, D)
    ctx = torch.randn(B, 1, 1024).expand(B, q * q, 1024).reshape(n, 1, 1024)
    masks = None
    if use_mask:
        masks = []
        for r in rs:
            mk = torch.rand(n, r * r) > 0.3
            mk[mk.sum(1) == 0] = True
            masks.append(mk)
    names = list(sd.keys())
    dout = torch.randn(n, 1, D)

    def run(sdd, qq, cx, dy, *fm):
        sdd = {k: v.detach().requires_grad_() for k, v in sdd.items()}
        qq = qq.detach().requires_grad_()
        cx = cx.detach().requires_grad_()
        fs = [f.detach().requires_grad_() for f in fm[:T]]
        out = O.sva_sampler(sdd, "", qq, cx, [O.window_rearrange(f, q) for f in fs], None if masks is None else list(fm[T:]),
                            depth, layer_type="sep")
        return out.detach(), torch.autograd.grad(out, [qq, cx, *fs, *[sdd[k] for k in names]], dy)

    (ref, gref), (eag, geag) = both_modes(run, sd, bf(queries), bf(ctx), bf(dout), *[bf(f) for f in feats], *(masks or []))
    qg = queries.to(dev).bfloat16().requires_grad_()
    cg = ctx.to(dev).bfloat16().requires_grad_()
    if natural:
        fg = [f.to(dev).bfloat16().requires_grad_() for f in feats]
    else:
        fg = [O.window_rearrange(bf(f), q).to(dev).bfloat16().requires_grad_() for f in feats]
    mg = [None] * T if masks is None else [mk.to(dev) for mk in masks]
    out = m(qg, cg, *fg, *mg, natural_layout=(B, q) if natural else None)
    out.backward(dout.to(dev).bfloat16())
    pc = ParityCollector()
    tag = f"sva sep rs={rs} natural={natural}"
    pc.check(out, ref, eag, f"{tag}: forward")
    pc.check(qg.grad, gref[0], geag[0], f"{tag}: dqueries")
    pc.check(cg.grad, gref[1], geag[1], f"{tag}: dcontext")
    for i in range(T):
        g = fg[i].grad if natural else fg[i].grad.float().cpu()
        gr, ge = gref[2 + i], geag[2 + i]
        if not natural:
            gr, ge = O.window_rearrange(gr, q), O.window_rearrange(ge, q)
        pc.check(g, gr, ge, f"{tag}: dfeats[{i}]")
    params = dict(m.named_parameters())
    for j, k in enumerate(names):
        gr, ge = gref[2 + T + j], geag[2 + T + j]
        if "k_proj.0.bias" in k:
            # one softmax per tower with a single query: a constant added to every key cancels — the gradient is
            # structurally zero (rounding noise only) in all three arms
            scale = float(params[k.replace("0.bias", "0.weight")].grad.float().abs().max())     # the same LayerNorm's gamma
            assert float(params[k].grad.float().abs().max()) <= 0.1 * scale + 1e-3, (k, scale)
            continue
        pc.check(params[k].grad, gr, ge, f"{tag}: grad {k}")
    pc.done()


def test_sva_mask_shape_error():
    from cambrian_b200.model.vision_sampler import VisionTokenSampler
    m = _cuda_bf16(VisionTokenSampler(256, 1024, [1024], [1], 1024, 1))
    q = torch.zeros(16, 1, 256, device=dev, dtype=torch.bfloat16)
    c = torch.zeros(16, 1, 1024, device=dev, dtype=torch.bfloat16)
    f = torch.zeros(16, 1, 1024, device=dev, dtype=torch.bfloat16)
    with pytest.raises(ValueError):
        m(q, c, f, torch.ones(16, 3, dtype=torch.bool, device=dev))


def _tower_pair(kind):
    from cambrian_b200.model.multimodal_encoder.builder import build_vision_tower_aux_list
    from oracle import cambrian_oracle as O
    cfg = tiny_cambrian_config()
    cfg.dino_config_overrides = dict(hidden_size=384, num_hidden_layers=2, num_attention_heads=6, mlp_ratio=4)
    idx = {"siglip": 0, "clip": 1, "dino": 2, "convnext": 3}[kind]
    args = ns(**{k: getattr(cfg, k) for k in ("siglip_config_overrides", "clip_config_overrides",
                                               "convnext_config_overrides", "dino_config_overrides")},
              mm_vision_tower_aux_list=[cfg.mm_vision_tower_aux_list[idx]],
              mm_vision_tower_aux_token_len_list=[576 if kind != "convnext" else 64])
    tower = build_vision_tower_aux_list(args)[0]
    return tower, O, tower_image_sizes(cfg)[idx]


@pytest.mark.parametrize("kind", ["clip", "dino", "siglip", "convnext"])
def test_tower_matches_oracle(kind):
    torch.manual_seed(1)
    tower, O, R = _tower_pair(kind)
    with torch.no_grad():  # non-trivial LayerScale / gamma / norm affine so every op is visible
        for n_, p in tower.named_parameters():
            if n_.endswith("lambda1") or n_.endswith("gamma"):
                p.copy_(0.5 + 0.5 * torch.rand_like(p))
            elif "norm" in n_ and n_.endswith("weight"):
                p.copy_(1 + 0.2 * torch.randn_like(p))
    tower = _cuda_bf16(tower)
    sd = sd_cpu32(tower.vision_tower)
    img = torch.randn(2, 3, R, R)
    c = tower.cfg
    interp = tower._interp_size
    common = dict(num_hidden_layers=c.get("num_hidden_layers"), patch_size=14,
                  num_attention_heads=c.get("num_attention_heads"), interp=interp)
    fn = {"clip": lambda s_, im: O.clip_vit(s_, dict(common, select_layer=-2), im),
          "dino": lambda s_, im: O.dinov2_vit(s_, common, im),
          "siglip": lambda s_, im: O.siglip_vit(s_, common, im),
          "convnext": lambda s_, im: O.convnext_trunk(s_, dict(depths=c["depths"], interp=interp, multi_stage=True), im)}[kind]
    with torch.no_grad():
        ref, eag = both_modes(fn, sd, bf(img))
        got = tower(img.to(dev).bfloat16())
    assert got.shape == ref.shape, (got.shape, ref.shape)
    pc = ParityCollector()
    pc.check(got, ref, eag, f"{kind} tower (small)")
    pc.done()


def test_decoder_layer_matches_oracle():
    from cambrian_b200.model.language_model.cambrian_llama import CBLlamaDecoderLayer, rope_tables
    from oracle import cambrian_oracle as O
    torch.manual_seed(2)
    cfg = tiny_cambrian_config()
    layer = CBLlamaDecoderLayer(cfg, 0)
    with torch.no_grad():
        for p in layer.parameters():
            if p.dim() == 2:
                p.normal_(0, 0.05)
            else:
                p.copy_(1 + 0.1 * torch.randn_like(p))
    layer = _cuda_bf16(layer)
    sd = sd_cpu32(layer, "model.layers.0.")
    B, S, H = 2, 300, cfg.hidden_size
    x = torch.randn(B, S, H)
    pos = torch.stack([torch.arange(S), torch.arange(S).clamp(max=250)])
    kmask = torch.ones(B, S, dtype=torch.bool)
    kmask[1, 270:] = False
    kmask[0, 40:50] = False
    ocfg = oracle_cfg(cfg)
    dout = torch.randn(B, S, H)
    dout[~kmask] = 0  # padded query rows carry no loss
    names = list(sd.keys())
    hd = H // cfg.num_attention_heads

    def run(sdd, xx, dy, pp, km):
        sdd = {k: v.detach().requires_grad_() for k, v in sdd.items()}
        xx = xx.detach().requires_grad_()
        cos, sin = O.rope_cos_sin(pp, hd, ocfg["rope_theta"])
        out = O.llama_layer(sdd, "model.layers.0.", xx, cos, sin, km, ocfg)
        return out.detach(), torch.autograd.grad(out, [xx, *[sdd[k] for k in names]], dy)

    (ref, gref), (eag, geag) = both_modes(run, sd, bf(x), bf(dout), pos, kmask)
    cos_t, sin_t = rope_tables(cfg, torch.device(dev))
    pc = ParityCollector()
    valid = kmask.to(dev)
    params = dict(layer.named_parameters())
    for recompute in (False, True):
        layer.zero_grad()
        rt = dict(pos=pos.to(dev).reshape(-1).contiguous(), cos=cos_t, sin=sin_t, kmask=valid, hf_cast=False,
                  recompute=recompute)
        xg = x.to(dev).bfloat16().requires_grad_()
        out = layer(xg, rt)
        out.backward(dout.to(dev).bfloat16())
        tag = f"decoder layer (recompute={recompute})"
        pc.check(out[valid], ref[valid], eag[valid], f"{tag}: fwd")
        pc.check(xg.grad[valid], gref[0][valid], geag[0][valid], f"{tag}: dx")
        for j, k in enumerate(names):
            pc.check(params[k[len("model.layers.0."):]].grad, gref[1 + j], geag[1 + j], f"{tag}: grad {k}")
    pc.done()


def _build_tiny_model(cfg):
    from cambrian_b200.model.language_model.cambrian_llama import CambrianLlamaForCausalLM
    torch.manual_seed(3)
    cfg.dino_config_overrides = dict(hidden_size=384, num_hidden_layers=2, num_attention_heads=6, mlp_ratio=4)
    model = CambrianLlamaForCausalLM(cfg)
    for t in model.get_model().vision_tower_aux_list:
        t.load_model()
    with torch.no_grad():
        for n_, p in model.named_parameters():
            if "pos_embed" in n_:
                p.mul_(0.1)
    model = _cuda_bf16(model)
    for t in model.get_model().vision_tower_aux_list:
        t.to(device=dev, dtype=torch.bfloat16)
    return model


def _tiny_batch(cfg, B=2, S=96):
    g = torch.Generator().manual_seed(7)
    q = int(cfg.image_token_len ** 0.5)
    span = q * (q + 1)
    ids = torch.randint(3, cfg.vocab_size, (B, S), generator=g)
    p0 = cfg.image_position
    ids[:, p0] = -200
    ids[:, p0 + 1:p0 + span] = 0
    labels = ids.clone()
    labels[:, p0:p0 + span] = -100
    labels[:, :p0] = -100
    attn = torch.ones(B, S, dtype=torch.bool)
    attn[1, S - 10:] = False
    labels[1, S - 10:] = -100
    pos = (attn.long().cumsum(1) - 1).clamp(min=0)
    images = [torch.randn(B, 3, r, r, generator=g) for r in tower_image_sizes(cfg)]
    masks = []
    for L in cfg.mm_vision_tower_aux_token_len_list:
        r = int(L ** 0.5) // q
        mk = torch.rand(B * q * q, r * r, generator=g) > 0.2
        mk[mk.sum(1) == 0] = True
        masks.append(mk)
    return ids, labels, attn, pos, images, masks


TOWER_KINDS = ("siglip", "clip", "dino", "convnext")


def _tower_fn(kind, t):
    from oracle import cambrian_oracle as O
    c = t.cfg
    common = dict(num_hidden_layers=c.get("num_hidden_layers"), patch_size=14,
                  num_attention_heads=c.get("num_attention_heads"), interp=t._interp_size)
    if kind == "siglip":
        return lambda sd, im: O.siglip_vit(sd, common, im)
    if kind == "clip":
        return lambda sd, im: O.clip_vit(sd, dict(common, select_layer=-2), im)
    if kind == "dino":
        return lambda sd, im: O.dinov2_vit(sd, common, im)
    return lambda sd, im: O.convnext_trunk(sd, dict(depths=c["depths"], interp=t._interp_size, multi_stage=True), im)


def _oracle_tower_feats(model, images):
    """fp32 oracle towers on the CPU; features rounded to bf16 (the CUDA towers hand bf16 features to the trainable part)."""
    towers = model.get_model().vision_tower_aux_list
    with torch.no_grad():
        return [bf(_tower_fn(kind, t)(sd_cpu32(t.vision_tower), bf(img))) for kind, t, img in zip(TOWER_KINDS, towers, images)]


def _oracle_forward(model, cfg, ids, labels, attn, pos, images, masks):
    """fp32 oracle, CPU (used by __graft_entry__.smoke and the engine tests)."""
    from oracle import cambrian_oracle as O
    sd = sd_cpu32(model)
    feats = _oracle_tower_feats(model, images)
    ocfg = oracle_cfg(cfg)
    sdg = {k: v.clone().requires_grad_() for k, v in sd.items()}
    img, feats_w, ctx_q = O.connector(sdg, ocfg, feats, masks)
    emb = O.splice(sdg, ids, img)
    hid = O.decoder(sdg, ocfg, emb, pos, attn, feats_w, masks, ctx_q)
    logits, loss = O.lm_loss(sdg, hid, labels)
    return logits, loss, sdg


def oracle_full_model_both(model, cfg, ids, labels, attn, pos, images, masks):
    """The whole hot path (towers -> connector -> splice -> decoder with SVA sites -> loss, + every parameter gradient)
    through the oracle in fp32 and in eager bf16.  Returns ((logits, loss, {name: grad}), same for eager)."""
    from oracle import cambrian_oracle as O
    towers = model.get_model().vision_tower_aux_list
    sd = sd_cpu32(model)
    names = list(sd.keys())
    for i, t in enumerate(towers):
        sd.update({f"tower{i}." + k: v for k, v in sd_cpu32(t.vision_tower).items()})
    fns = [_tower_fn(kind, t) for kind, t in zip(TOWER_KINDS, towers)]
    ocfg = oracle_cfg(cfg)

    def run(s, ii, ll, aa, pp, *im_masks):
        ims, mks = im_masks[:len(towers)], (list(im_masks[len(towers):]) or None)
        with torch.no_grad():
            feats = []
            for i, (f, im) in enumerate(zip(fns, ims)):
                tsd = {k[len(f"tower{i}."):]: v for k, v in s.items() if k.startswith(f"tower{i}.")}
                o = f(tsd, im)
                feats.append(o.to(torch.bfloat16).to(o.dtype))     # towers hand bf16 features on (frozen, no grad)
        s = {k: (v.detach().requires_grad_() if k in names else v) for k, v in s.items()}
        img, feats_w, ctx_q = O.connector(s, ocfg, feats, mks)
        emb = O.splice(s, ii, img)
        hid = O.decoder(s, ocfg, emb, pp, aa, feats_w, mks, ctx_q)
        logits, loss = O.lm_loss(s, hid, ll)
        grads = torch.autograd.grad(loss, [s[k] for k in names], allow_unused=True)
        return logits.detach(), loss.detach(), dict(zip(names, grads))

    return both_modes(run, sd, ids, labels, attn, pos, *[bf(i) for i in images], *(masks or []))


@pytest.mark.parametrize("fused_loss", [False, True])
def test_full_model_loss_and_grads_match_oracle(fused_loss):
    cfg = tiny_cambrian_config()
    cfg.fused_lm_loss = fused_loss
    model = _build_tiny_model(cfg)
    model.train()
    ids, labels, attn, pos, images, masks = _tiny_batch(cfg)
    (ref_logits, ref_loss, gref), (eag_logits, eag_loss, geag) = oracle_full_model_both(
        model, cfg, ids, labels, attn, pos, images, masks)
    out = model(input_ids=ids.to(dev), labels=labels.to(dev), attention_mask=attn.to(dev), position_ids=pos.to(dev),
                images=[i.to(dev).bfloat16() for i in images], image_aux_attention_masks_list=[m.to(dev) for m in masks])
    out.loss.backward()
    # loss is an fp32 scalar: north_star's rtol, or the eager path's own deviation when that is larger
    lim = max(FP32_RTOL * abs(ref_loss.item()), 1.5 * abs(eag_loss.item() - ref_loss.item()))
    assert abs(out.loss.item() - ref_loss.item()) <= lim, (out.loss.item(), ref_loss.item(), eag_loss.item())
    pc = ParityCollector()
    tag = f"tiny model (fused_loss={fused_loss})"
    if not fused_loss:
        valid = attn.to(dev)
        pc.check(out.logits[valid], ref_logits[valid], eag_logits[valid], f"{tag}: logits")
    for k, p in model.named_parameters():
        g = gref[k]
        if g is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
            continue
        assert p.grad is not None, f"missing grad for {k}"
        pc.check(p.grad, g, geag[k], f"{tag}: grad {k}")
    pc.done()


def test_engine_overlapped_optimizer_matches_serial():
    """Per-bucket AdamW on the side stream (overlapped with backward) must give exactly the parameters of the serial
    whole-buffer update: same kernels on the same values, only the schedule differs."""
    from cambrian_b200.engine import TrainEngine
    results = []
    for overlap in (True, False, "defer"):
        cfg = tiny_cambrian_config()
        cfg.fused_lm_loss = True
        model = _build_tiny_model(cfg)
        model.train()
        ids, labels, attn, pos, images, masks = _tiny_batch(cfg)
        batch = dict(input_ids=ids.to(dev), labels=labels.to(dev), attention_mask=attn.to(dev), position_ids=pos.to(dev),
                     images=[i.to(dev).bfloat16() for i in images],
                     image_aux_attention_masks_list=[m.to(dev) for m in masks])
        eng = TrainEngine(model, lr=1e-3, bucket_mb=8.0, overlap=bool(overlap))
        eng.defer_param_sync = overlap == "defer"    # towers of the next step overlap the optimizer tail
        losses = []
        for _ in range(3):
            eng.zero_grad()
            loss = model(**batch).loss
            loss.backward()
            eng.step()
            losses.append(float(loss))
        torch.cuda.synchronize()
        results.append((eng.flat_p.clone(), eng.master.clone(), losses, len(eng.buckets)))
    assert results[0][3] > 3                                      # several buckets -> the overlapped path is exercised
    assert results[0][2][0] == results[1][2][0] and results[0][2][2] < results[0][2][0]   # same start, loss goes down
    # embedding-row gradients use bf16 atomics (order-dependent rounding), so allow last-bit differences there
    assert rel_err(results[0][0], results[1][0]) < 2e-2
    assert_same_training(results[0][1], results[1][1], 1e-3, 3, "overlapped vs serial optimizer")
    assert rel_err(results[2][0], results[1][0]) < 2e-2
    assert_same_training(results[2][1], results[1][1], 1e-3, 3, "deferred vs serial optimizer")
    assert results[2][2][0] == results[1][2][0]


def test_engine_step_reproduces_autograd_gradients():
    """TrainEngine (flat buffers, main_grad accumulation, fused AdamW) must reproduce plain-autograd gradients."""
    from cambrian_b200.engine import TrainEngine
    from oracle import cambrian_oracle as O
    cfg = tiny_cambrian_config()
    cfg.fused_lm_loss = True
    model = _build_tiny_model(cfg)
    model.train()
    ids, labels, attn, pos, images, masks = _tiny_batch(cfg)
    batch = dict(input_ids=ids.to(dev), labels=labels.to(dev), attention_mask=attn.to(dev), position_ids=pos.to(dev),
                 images=[i.to(dev).bfloat16() for i in images], image_aux_attention_masks_list=[m.to(dev) for m in masks])
    out = model(**batch)
    out.loss.backward()
    ref_grads = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
    model.zero_grad(set_to_none=True)
    eng = TrainEngine(model, lr=1e-3)
    before = eng.flat_p.clone()
    eng.zero_grad()
    loss2 = model(**batch).loss
    loss2.backward()
    eng._finalize_unwritten()  # folds in the few gradients that arrive through plain autograd (vision_query slice)
    for k, p in model.named_parameters():
        if k in ref_grads:
            assert rel_err(p.main_grad, ref_grads[k]) < 2e-2, k
    eng.step()
    assert float((eng.flat_p.float() - before.float()).abs().max()) > 0
    assert torch.isfinite(eng.master).all()
    # greedy decode parity: tests/test_parity_gpu.py::test_greedy_generate_token_exact_32_tokens


# ---------------------------------------------------------------------------------------------------------------
# dynamic-shape branch (non-square images, per-sample unpadded query grids) — SURVEY.md §8f rank 3
# ---------------------------------------------------------------------------------------------------------------
def _dynamic_batch(cfg, B=2, L=40):
    g = torch.Generator().manual_seed(17)
    ids = torch.randint(3, cfg.vocab_size, (B, L), generator=g)
    ids[:, cfg.image_position] = -200                      # bare <image> indicator: the branch expands it per sample
    attn = torch.ones(B, L, dtype=torch.bool)
    if B > 1:
        attn[1, L - 7:] = False
    images = [torch.randn(B, 3, r, r, generator=g) for r in tower_image_sizes(cfg)]
    return ids, attn, images


def test_dynamic_branch_logits_match_oracle():
    from oracle import cambrian_oracle as O
    cfg = tiny_cambrian_config()
    model = _build_tiny_model(cfg).eval()
    ids, attn, images = _dynamic_batch(cfg)
    sizes = [(800, 400), (200, 500)]                       # 4x4 query grid -> (2, 4) and (4, 2) after unpadding
    with torch.no_grad():
        out = model(input_ids=ids.to(dev), attention_mask=attn.to(dev), images=[i.to(dev).bfloat16() for i in images],
                    image_sizes=sizes)
        sd = sd_cpu32(model)
        ocfg = oracle_cfg(cfg)
        feats = _oracle_tower_feats(model, images)
        emb, _, am, pos, ff, mf, fs, ctx = O.prepare_dynamic(sd, ocfg, feats, ids, attn, None, sizes)
        assert fs == [(2, 4), (4, 2)]
        hid = O.decoder_dynamic(sd, ocfg, emb, pos, am, ff, mf, ctx, fs)
        ref_logits, _ = O.lm_loss(sd, hid, None)
        sdb = O.eager_bf16(sd)
        feats_b = [f.bfloat16() for f in feats]
        emb_b, _, _, _, ff_b, mf_b, _, ctx_b = O.prepare_dynamic(sdb, ocfg, feats_b, ids, attn, None, sizes)
        eag_logits, _ = O.lm_loss(sdb, O.decoder_dynamic(sdb, ocfg, emb_b, pos, am, ff_b, mf_b, ctx_b, fs), None)
    assert out.logits.shape == ref_logits.shape
    pc = ParityCollector()
    pc.check(out.logits.cpu()[am], ref_logits[am], eag_logits[am], "dynamic-branch logits")
    pc.done()


def test_dynamic_branch_equals_static_branch_for_square_images():
    cfg = tiny_cambrian_config()
    model = _build_tiny_model(cfg).eval()
    ids, attn, images = _dynamic_batch(cfg)
    imgs = [i.to(dev).bfloat16() for i in images]
    B = ids.shape[0]
    with torch.no_grad():
        static = model(input_ids=ids.to(dev), attention_mask=attn.to(dev), images=imgs, image_sizes=[(336, 336)] * B)
        args = model._prepare_dynamic(ids.to(dev), None, attn.to(dev), None, None, imgs, [(336, 336)] * B)
        (_, pos, am, _, emb, _, ff, mf, fs, ctx) = args
        dyn = model.model(inputs_embeds=emb, attention_mask=am, position_ids=pos, vision_tower_aux_feature_list=ff,
                          vision_tower_aux_attention_masks_list=mf, final_vision_feature_size=fs,
                          global_context_feature=ctx)
        dyn_logits = ops.gemm(dyn.last_hidden_state.reshape(-1, cfg.hidden_size).contiguous(), model.lm_head.weight,
                              out_dtype=torch.float32).view(B, -1, cfg.vocab_size)
    # same arithmetic, different gather path (materialised windows vs in-kernel index arithmetic): bf16-rounding level
    assert static.logits.shape == dyn_logits.shape
    L = dyn_logits.shape[1]
    m = am[:, :L].bool()
    assert rel_err(static.logits[:, :L][m], dyn_logits[:, :L][m]) < 2e-2


def test_dynamic_branch_greedy_generate_runs():
    cfg = tiny_cambrian_config()
    model = _build_tiny_model(cfg).eval()
    ids, attn, images = _dynamic_batch(cfg, B=1, L=20)
    toks = model.generate(ids.to(dev), images=[i.to(dev).bfloat16() for i in images], image_sizes=[(200, 500)],
                          max_new_tokens=4)
    assert toks.shape == (1, 4) and int(toks.min()) >= 0 and int(toks.max()) < cfg.vocab_size


def test_zero3_sharded_generate_is_token_identical():
    """ZeRO-3 inference sharding (config 5) at world size 1: the gather pipeline (staging slots, wrap-around prefetch,
    parameter re-pointing) must not change a single generated token."""
    from cambrian_b200.sharded import Zero3Inference
    cfg = tiny_cambrian_config()
    model = _build_tiny_model(cfg).eval()
    ids, attn, images = _dynamic_batch(cfg, B=2, L=24)
    imgs = [i.to(dev).bfloat16() for i in images]
    kw = dict(images=imgs, image_sizes=[(336, 336)] * 2, attention_mask=attn.to(dev), max_new_tokens=6)
    ref = model.generate(ids.to(dev), **kw)
    z = Zero3Inference(model)
    got = model.generate(ids.to(dev), **kw)
    assert torch.equal(ref, got)
    assert z.gathers >= 6 * cfg.num_hidden_layers


def test_fused_loss_with_label_ranges_matches_full_rows():
    """The collator's label-range hint only removes rows that cannot contribute: loss and every gradient must agree
    with the all-rows fused loss."""
    from cambrian_b200.train.collator import valid_label_ranges
    cfg = tiny_cambrian_config()
    cfg.fused_lm_loss = True
    model = _build_tiny_model(cfg)
    model.train()
    ids, labels, attn, pos, images, masks = _tiny_batch(cfg)
    ranges, n_valid = valid_label_ranges(labels)
    assert 0 < sum(b - a for a, b in ranges) < labels.numel() and n_valid == sum(b - a for a, b in ranges)
    kw = dict(input_ids=ids.to(dev), labels=labels.to(dev), attention_mask=attn.to(dev), position_ids=pos.to(dev),
              images=[i.to(dev).bfloat16() for i in images], image_aux_attention_masks_list=[m.to(dev) for m in masks])
    res = []
    for lr in (None, ranges):
        model.zero_grad(set_to_none=True)
        out = model(**kw, label_ranges=lr, num_valid_labels=n_valid)
        out.loss.backward()
        res.append((out.loss.item(), {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}))
    assert abs(res[0][0] - res[1][0]) < 1e-4 * abs(res[0][0])
    assert res[0][1].keys() == res[1][1].keys()
    for k in res[0][1]:
        assert rel_err(res[1][1][k], res[0][1][k]) < 2e-2, k


def test_generate_sampling_stopping_criteria_and_streamer():
    """A12: the keywords the reference's callers pass (model_worker.py:177-187: do_sample, temperature, top_p,
    max_new_tokens, streamer, stopping_criteria, use_cache; inference.py:77-85: num_beams=1) are honoured."""
    cfg = tiny_cambrian_config()
    model = _build_tiny_model(cfg).eval()
    ids, attn, images = _dynamic_batch(cfg, B=2, L=24)
    imgs = [i.to(dev).bfloat16() for i in images]
    common = dict(images=imgs, image_sizes=[(336, 336)] * 2, attention_mask=attn.to(dev))

    class Streamer:
        def __init__(self):
            self.chunks, self.ended = [], False

        def put(self, t):
            self.chunks.append(t.clone())

        def end(self):
            self.ended = True

    greedy = model.generate(ids.to(dev), max_new_tokens=6, do_sample=False, temperature=0, num_beams=1, use_cache=True, **common)
    # seeded sampling is reproducible, differs from greedy somewhere, and stays inside the top-k set
    gens = []
    for _ in range(2):
        g = torch.Generator(device=dev).manual_seed(123)
        gens.append(model.generate(ids.to(dev), max_new_tokens=6, do_sample=True, temperature=1.5, top_p=0.9, generator=g, **common))
    assert torch.equal(gens[0], gens[1]) and gens[0].shape == (2, 6)
    # temperature -> 0 sampling collapses to greedy
    g = torch.Generator(device=dev).manual_seed(5)
    cold = model.generate(ids.to(dev), max_new_tokens=6, do_sample=True, temperature=1e-4, top_k=0, generator=g, **common)
    assert torch.equal(cold, greedy)
    # stopping criteria (HF signature: (generated_ids, scores) -> bool / [B] bool) and streamer
    st = Streamer()
    seen = []

    def stop_after_three(gen_ids, scores):
        seen.append(tuple(gen_ids.shape))
        return gen_ids.shape[1] >= 3
    out = model.generate(ids.to(dev), max_new_tokens=6, stopping_criteria=[stop_after_three], streamer=st, **common)
    assert out.shape == (2, 3) and torch.equal(out, greedy[:, :3])
    assert seen == [(2, 1), (2, 2), (2, 3)]
    assert st.ended and st.chunks[0].shape == (2, 0) and torch.equal(torch.stack(st.chunks[1:], 1), out.cpu())
    # EOS: the sequence that emits it is padded afterwards, the other one keeps going
    eos = int(greedy[0, 1])
    out = model.generate(ids.to(dev), max_new_tokens=6, eos_token_id=eos, pad_token_id=0, **common)
    row = out[0].tolist()
    k = row.index(eos)
    assert all(t == 0 for t in row[k + 1:])
    with pytest.raises(NotImplementedError):
        model.generate(ids.to(dev), max_new_tokens=2, num_beams=3, **common)
    with pytest.raises(TypeError):
        model.generate(ids.to(dev), max_new_tokens=2, not_a_generate_kwarg=True, **common)


def test_towers_on_separate_streams_give_identical_features():
    """encode_images (cambrian_arch.py:332-338) runs each frozen tower on its own stream; the features must be the ones the
    sequential schedule produces, bit for bit, call after call."""
    import cambrian_b200.model.cambrian_arch as A
    cfg = tiny_cambrian_config()
    model = _build_tiny_model(cfg).eval()
    _, _, _, _, images, _ = _tiny_batch(cfg)
    imgs = [i.to(dev).bfloat16() for i in images]
    old = A._TOWER_STREAMS
    try:
        A._TOWER_STREAMS = False
        seq = model.encode_images(imgs)
        A._TOWER_STREAMS = True
        for _ in range(3):
            par = model.encode_images(imgs)
            torch.cuda.synchronize()
            assert all(torch.equal(a, b) for a, b in zip(seq, par))
    finally:
        A._TOWER_STREAMS = old
