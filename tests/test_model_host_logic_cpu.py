"""CPU coverage of the HOST logic of the whole trainable path (cambrian_arch -> cambrian_llama -> autograd blocks ->
TrainEngine): `CambrianLlamaForCausalLM.forward` / backward and the engine's step run on the CPU with the kernels replaced
by the plain-torch stand-ins of tests/ops_emulation.py (test infrastructure, monkeypatched per test) and the frozen towers
replaced by seeded feature tensors (the towers are forward-only and covered at full depth under `-m gpu`).  What is
exercised is the product's Python: image-span location / expansion, projector and SVA wiring, splice, the decoder loop
with its in-LLM SVA sites, fused and unfused loss, every block's hand-written backward, gradient routing into the
engine's flat buffers, bucket collectives launched from inside backward on two gloo ranks.  Checked against the oracle.
Kernel numerics are NOT covered here — `-m gpu` does that."""
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ops_emulation  # noqa: E402
from helpers import oracle_cfg, tiny_cambrian_config  # noqa: E402
from oracle import cambrian_oracle as O  # noqa: E402

pytestmark = pytest.mark.skipif(torch.cuda.is_available(), reason="kernel stand-ins are for GPU-less machines only; on a GPU "
                                                                  "box the same logic is covered by the -m gpu suite")


def _fro(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-30)).item()


def _build(cfg, seed=3):
    from cambrian_b200.model.language_model.cambrian_llama import CambrianLlamaForCausalLM
    torch.manual_seed(seed)
    cfg.dino_config_overrides = dict(hidden_size=384, num_hidden_layers=2, num_attention_heads=6, mlp_ratio=4)
    model = CambrianLlamaForCausalLM(cfg)
    for t in model.get_model().vision_tower_aux_list:
        t.load_model()
    with torch.no_grad():
        for n_, p in model.named_parameters():
            if "pos_embed" in n_:
                p.mul_(0.1)
    return model.to(torch.bfloat16)


def _batch(cfg, B=2, S=64, seed=7):
    g = torch.Generator().manual_seed(seed)
    q = int(cfg.image_token_len ** 0.5)
    span = q * (q + 1)
    ids = torch.randint(3, cfg.vocab_size, (B, S), generator=g)
    p0 = cfg.image_position
    ids[:, p0] = -200
    ids[:, p0 + 1:p0 + span] = 0
    labels = ids.clone()
    labels[:, :p0 + span] = -100
    attn = torch.ones(B, S, dtype=torch.bool)
    attn[1, S - 6:] = False
    labels[1, S - 6:] = -100
    pos = (attn.long().cumsum(1) - 1).clamp(min=0)
    masks = []
    for L in cfg.mm_vision_tower_aux_token_len_list:
        r = int(L ** 0.5) // q
        mk = torch.rand(B * q * q, r * r, generator=g) > 0.2
        mk[mk.sum(1) == 0] = True
        masks.append(mk)
    return ids, labels, attn, pos, masks


def _tower_feats(model, cfg, B, seed):
    """Stand-in for the frozen towers' outputs: [B, N_i, C_i] bf16 with the token counts / widths the towers declare."""
    g = torch.Generator().manual_seed(seed)
    towers = model.get_model().vision_tower_aux_list
    return [torch.randn(B, t.num_patches, t.hidden_size, generator=g).bfloat16() for t in towers]


def _oracle(model, cfg, ids, labels, attn, pos, feats, masks):
    sd = {k: v.detach().float().clone().requires_grad_() for k, v in model.state_dict().items()}
    ocfg = oracle_cfg(cfg)
    if cfg.mm_projector_type == "sva":
        img, feats_w, ctx_q = O.connector(sd, ocfg, [f.float() for f in feats], masks)
    else:
        img, feats_w, ctx_q = O.mlp2x_gelu(sd, "model.mm_projector.", torch.cat([f.float() for f in feats], -1)), None, None
        B, q = img.shape[0], int(cfg.image_token_len ** 0.5)
        nl = sd["model.image_newline"][None, None, None, :].expand(B, q, 1, -1)
        img = torch.cat([img.view(B, q, q, -1), nl], 2).flatten(1, 2)
    hid = O.decoder(sd, ocfg, O.splice(sd, ids, img), pos, attn, feats_w, masks, ctx_q)
    logits, loss = O.lm_loss(sd, hid, labels)
    return logits, loss, sd


@pytest.mark.parametrize("fused,sva", [(False, True), (True, True), (True, False)])
def test_full_trainable_path_host_logic(monkeypatch, fused, sva):
    ops_emulation.install(monkeypatch)
    cfg = tiny_cambrian_config(sva=sva, connector_only=not sva)
    if not sva:
        cfg.mm_vision_tower_aux_list = cfg.mm_vision_tower_aux_list[1:2]         # config 2: one CLIP tower + mlp2x_gelu
        cfg.mm_vision_tower_aux_token_len_list = [cfg.image_token_len]
    cfg.fused_lm_loss = fused
    model = _build(cfg)
    model.train()
    ids, labels, attn, pos, masks = _batch(cfg)
    if not sva:
        masks = None
    feats = _tower_feats(model, cfg, ids.shape[0], 11)
    monkeypatch.setattr(type(model), "encode_images", lambda self, imgs: feats)
    images = [torch.zeros(ids.shape[0], 3, 8, 8, dtype=torch.bfloat16) for _ in feats]   # only their batch size is read
    out = model(input_ids=ids, labels=labels, attention_mask=attn, position_ids=pos, images=images,
                image_aux_attention_masks_list=masks)
    out.loss.backward()
    ref_logits, ref_loss, sd = _oracle(model, cfg, ids, labels, attn, pos, feats, masks)
    ref_loss.backward()
    assert abs(out.loss.item() - ref_loss.item()) <= 2e-2 * abs(ref_loss.item()), (out.loss.item(), ref_loss.item())
    if not fused:
        assert _fro(out.logits[attn], ref_logits[attn]) < 3e-2
    worst = ("", 0.0)
    checked = 0
    for k, p in model.named_parameters():
        g = sd[k].grad
        if g is None or float(g.abs().max()) == 0.0:
            assert p.grad is None or float(p.grad.float().abs().max()) == 0.0, k
            continue
        assert p.grad is not None, f"missing grad for {k}"
        e = _fro(p.grad, g)
        checked += 1
        if e > worst[1]:
            worst = (k, e)
    assert checked > 30 and worst[1] < 5e-2, worst


def test_decoder_layer_recompute_gives_the_same_gradients(monkeypatch):
    """Per-layer activation recompute (gradient_checkpointing) re-runs the block's forward inside backward: identical
    gradients to the stored-activation path (same stand-in kernels, so bit-identical here)."""
    ops_emulation.install(monkeypatch)
    grads = []
    for ckpt in (False, True):
        cfg = tiny_cambrian_config()
        cfg.fused_lm_loss = True
        model = _build(cfg)
        model.train()
        model.get_model().gradient_checkpointing = ckpt
        ids, labels, attn, pos, masks = _batch(cfg)
        feats = _tower_feats(model, cfg, ids.shape[0], 11)
        monkeypatch.setattr(type(model), "encode_images", lambda self, imgs: feats)
        images = [torch.zeros(ids.shape[0], 3, 8, 8, dtype=torch.bfloat16) for _ in feats]
        model(input_ids=ids, labels=labels, attention_mask=attn, position_ids=pos, images=images,
              image_aux_attention_masks_list=masks).loss.backward()
        grads.append({k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None})
    assert grads[0].keys() == grads[1].keys()
    assert all(torch.equal(grads[0][k], grads[1][k]) for k in grads[0])


def test_dynamic_shape_branch_host_logic(monkeypatch):
    """The reference's per-sample (non-XLA) branch for non-square images (cambrian_arch.py:289-330, :422-451, :493-609;
    cambrian_llama.py:208-253): unpadded query grids, ragged splice, ragged in-LLM SVA sites — logits vs the oracle."""
    ops_emulation.install(monkeypatch)
    cfg = tiny_cambrian_config()
    model = _build(cfg).eval()
    g = torch.Generator().manual_seed(17)
    B, L = 2, 40
    ids = torch.randint(3, cfg.vocab_size, (B, L), generator=g)
    ids[:, cfg.image_position] = -200                        # bare <image> indicator: the branch expands it per sample
    attn = torch.ones(B, L, dtype=torch.bool)
    attn[1, L - 7:] = False
    feats = _tower_feats(model, cfg, B, 23)
    monkeypatch.setattr(type(model), "encode_images", lambda self, imgs: feats)
    images = [torch.zeros(B, 3, 8, 8, dtype=torch.bfloat16) for _ in feats]
    sizes = [(800, 400), (200, 500)]                         # 4x4 query grid -> (2, 4) and (4, 2) after unpadding
    with torch.no_grad():
        out = model(input_ids=ids, attention_mask=attn, images=images, image_sizes=sizes)
        sd = {k: v.detach().float() for k, v in model.state_dict().items()}
        ocfg = oracle_cfg(cfg)
        emb, _, am, pos, ff, mf, fs, ctx = O.prepare_dynamic(sd, ocfg, [f.float() for f in feats], ids, attn, None, sizes)
        assert fs == [(2, 4), (4, 2)]
        ref_logits, _ = O.lm_loss(sd, O.decoder_dynamic(sd, ocfg, emb, pos, am, ff, mf, ctx, fs), None)
    assert out.logits.shape == ref_logits.shape
    assert _fro(out.logits[am], ref_logits[am]) < 3e-2


def test_greedy_generate_host_logic(monkeypatch):
    """generate(): multimodal prefill + KV-cache decode loop (the eager per-token loop; the CUDA-graph replay needs a GPU),
    token-exact against the oracle's greedy decode on the peaked test model of tests/test_parity_gpu.py."""
    from test_parity_gpu import _oracle_greedy
    ops_emulation.install(monkeypatch)
    cfg = tiny_cambrian_config()
    cfg.fused_lm_loss = True
    model = _build(cfg)
    with torch.no_grad():
        emb = model.get_model().embed_tokens.weight
        perm = torch.randperm(emb.shape[0], generator=torch.Generator().manual_seed(9))
        model.lm_head.weight.copy_(emb[perm] * 24.0)
        for n_, p in model.named_parameters():
            if ((n_.endswith("o_proj.weight") and "layers." in n_ and "vision_sampler" not in n_)
                    or n_.endswith("down_proj.weight")
                    or ("vision_sampler_layers" in n_ and n_.endswith("proj_out.linear_2.weight"))):
                p.mul_(0.4)
    model.eval()
    ids, labels, attn, pos, masks = _batch(cfg, S=96)
    S0, n_new = 40, 12
    gen_ids = ids[:1, :S0].clone()
    feats = [f[:1] for f in _tower_feats(model, cfg, 2, 31)]
    monkeypatch.setattr(type(model), "encode_images", lambda self, imgs: feats)
    images = [torch.zeros(1, 3, 8, 8, dtype=torch.bfloat16) for _ in feats]
    new = model.generate(gen_ids, images=images, image_sizes=[(56, 56)], max_new_tokens=n_new, do_sample=False)
    sd = {k: v.detach().float() for k, v in model.state_dict().items()}
    want, margins = _oracle_greedy(sd, cfg, oracle_cfg(cfg), [f.float() for f in feats], gen_ids, n_new, torch.float32,
                                   torch.device("cpu"))
    assert new[0].tolist() == want, (new[0].tolist(), want, margins)
    assert len(set(want)) >= 4


# ------------------------------------------------------------------------------------------------ N > 1 on gloo
def _train_worker(rank, world, port, q, zero, vary_labels=False):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(max(1, (os.cpu_count() or 2) // (2 * world)))   # two workers share the host: no oversubscription
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ops_emulation.install()
        from cambrian_b200.engine import TrainEngine
        cfg = tiny_cambrian_config()
        cfg.fused_lm_loss = True
        model = _build(cfg)                                    # same seed on every rank: identical replicas
        model.train()
        ids, labels, attn, pos, masks = _batch(cfg, seed=7 + rank)          # every rank its own data, fixed over the steps
        feats = _tower_feats(model, cfg, ids.shape[0], 100 + rank)
        type(model).encode_images = lambda self, imgs: feats
        images = [torch.zeros(ids.shape[0], 3, 8, 8, dtype=torch.bfloat16) for _ in feats]
        calls = {"n": 0, "bwd": False}
        ar0 = dist.all_reduce

        def counting(*a, **k):
            calls["n"] += int(calls["bwd"])
            return ar0(*a, **k)
        dist.all_reduce = counting
        eng = TrainEngine(model, lr=2e-3, bucket_mb=1.0, zero_stage=zero, max_grad_norm=1.0)
        losses = []
        from cambrian_b200.train.collator import valid_label_ranges
        for step in range(3):
            lab, hints = labels, {}
            if vary_labels:
                # the label layout differs per rank AND per step: the lm_head's chunk count and the rows it
                # processes are data-dependent, the engine's collective schedule must not be
                gl = torch.Generator().manual_seed(100 * step + rank)
                lab = labels.clone()
                lab[torch.rand(lab.shape, generator=gl) < 0.4] = -100
                ranges, nv = valid_label_ranges(lab)
                hints = dict(label_ranges=ranges, num_valid_labels=nv)
            eng.zero_grad()
            loss = model(input_ids=ids, labels=lab, attention_mask=attn, position_ids=pos, images=images,
                         image_aux_attention_masks_list=masks, **hints).loss
            calls["bwd"] = True
            loss.backward()
            calls["bwd"] = False
            eng.step()
            losses.append(float(loss.detach()))
        t = eng.flat_p.float().clone()
        lo, hi = t.clone(), t.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        ok = bool(torch.equal(lo, hi))                                      # replicas bit-identical after 3 steps
        ok &= eng._overlap_ok and len(eng.buckets) >= 4
        ok &= calls["n"] >= len(eng.buckets)                               # collectives were issued from inside backward
        ok &= vary_labels or losses[-1] < losses[0]                         # and the job trains
        if zero == 2:
            ok &= eng.master.numel() * world == eng.total
        q.put((rank, bool(ok), f"losses {losses} buckets {len(eng.buckets)} in-backward collectives {calls['n']} "
                                f"overlap_ok {eng._overlap_ok}"))
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, False, traceback.format_exc()[-2500:]))
    finally:
        dist.destroy_process_group()


def _zero3_generate_worker(rank, world, port, q):
    import torch.distributed as dist
    torch.set_num_threads(max(1, (os.cpu_count() or 2) // (2 * world)))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ops_emulation.install()
        from cambrian_b200.sharded import Zero3Inference
        cfg = tiny_cambrian_config()
        model = _build(cfg).eval()                             # identical replicas
        g = torch.Generator().manual_seed(40 + rank)           # every rank decodes its own sequence
        L = 24 + 3 * rank                                      # different prompt lengths: ranks stay in lock-step anyway
        ids = torch.randint(3, cfg.vocab_size, (1, L), generator=g)
        ids[:, cfg.image_position] = -200
        feats = _tower_feats(model, cfg, 1, 60 + rank)
        type(model).encode_images = lambda self, imgs: feats
        images = [torch.zeros(1, 3, 8, 8, dtype=torch.bfloat16) for _ in feats]
        kw = dict(images=images, image_sizes=[(336, 336)], max_new_tokens=5)
        ref = model.generate(ids, **kw)                        # unsharded
        z = Zero3Inference(model)
        shard_elems = sum(s_.numel() for s_ in z.shards)
        got = model.generate(ids, **kw)                        # decoder layers sharded 1/world, gathered a layer ahead
        ok = torch.equal(ref, got) and z.gathers >= 5 * cfg.num_hidden_layers
        q.put((rank, bool(ok), f"ref {ref.tolist()} got {got.tolist()} gathers {z.gathers} shard elems {shard_elems}"))
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, False, traceback.format_exc()[-2500:]))
    finally:
        dist.destroy_process_group()


def test_zero3_sharded_generate_two_ranks_gloo():
    """BASELINE config 5's scheme on CPU: decoder layers sharded over two gloo ranks and all-gathered one layer ahead while
    every rank runs `generate()` on its own sequence — token-identical to the unsharded model."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 37500 + os.getpid() % 2000
    procs = [ctx.Process(target=_zero3_generate_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=600) for _ in procs), key=lambda r: r[0])
    for p in procs:
        p.join(60)
    assert all(r[1] for r in res), [r[2] for r in res]


@pytest.mark.parametrize("zero,vary_labels", [(0, False), (2, False), (2, True)])
def test_whole_model_data_parallel_two_ranks_gloo(zero, vary_labels):
    """SURVEY §8e on CPU: the full tiny Cambrian model (connector + decoder with in-LLM SVA sites + fused loss) under
    TrainEngine on two gloo ranks — DDP all-reduce buckets and the ZeRO-2 sharded optimizer, gradient clipping on; one
    variant with label layouts (and the collator's label-range hints) that differ per rank and per step."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 35500 + (os.getpid() + 29 * zero + 7 * int(vary_labels)) % 2000
    procs = [ctx.Process(target=_train_worker, args=(r, 2, port, q, zero, vary_labels)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=600) for _ in procs), key=lambda r: r[0])
    for p in procs:
        p.join(60)
    assert all(r[1] for r in res), [r[2] for r in res]
