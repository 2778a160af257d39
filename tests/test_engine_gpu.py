"""GPU tests of the TrainEngine features added in round 2: device-side gradient clipping, parameter groups, background
AdamW launch shape, per-bucket parameter-ready events (deferred sync), gradient-accumulation loss scale."""
import pytest
import torch

from helpers import assert_same_training, rel_err, tiny_cambrian_config

pytestmark = pytest.mark.gpu
dev = "cuda"


def _setup(fused=True):
    import test_modules_gpu as T
    cfg = tiny_cambrian_config()
    cfg.fused_lm_loss = fused
    model = T._build_tiny_model(cfg)
    model.train()
    ids, labels, attn, pos, images, masks = T._tiny_batch(cfg)
    batch = dict(input_ids=ids.to(dev), labels=labels.to(dev), attention_mask=attn.to(dev), position_ids=pos.to(dev),
                 images=[i.to(dev).bfloat16() for i in images], image_aux_attention_masks_list=[m.to(dev) for m in masks])
    return cfg, model, batch


def test_adamw_device_coefficient_and_background_shape_match_plain_launch():
    from cambrian_b200 import ops
    torch.manual_seed(0)
    n = 8 * 4099
    p0, g = torch.randn(n, device=dev), torch.randn(n, device=dev).bfloat16()
    res = []
    for coef, bg in ((None, False), (torch.tensor([0.37, 0.0], device=dev), False), (torch.tensor([0.37, 0.0], device=dev), True)):
        p, m, v = p0.clone(), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
        p16 = torch.empty(n, device=dev, dtype=torch.bfloat16)
        for step in (1, 2, 3):
            ops.adamw(p, m, v, g, p16, 1e-2, 0.9, 0.999, 1e-8, 0.1, step, grad_scale=0.37, clip_coef=coef, background=bg)
        res.append((p, m, v, p16))
    for a, b in zip(res[0], res[1]):
        assert torch.equal(a, b)
    for a, b in zip(res[0], res[2]):
        assert torch.equal(a, b)
    # against torch.optim.AdamW on the scaled gradient
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([ref], lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.1)
    for _ in range(3):
        ref.grad = g.float() * 0.37
        opt.step()
    torch.testing.assert_close(res[0][0], ref.detach(), rtol=1e-5, atol=1e-6)


def test_sumsq_and_clip_coefficient():
    from cambrian_b200 import ops
    torch.manual_seed(1)
    g = torch.randn(8 * 100003, device=dev).bfloat16()
    acc = torch.zeros(1, device=dev)
    ws = torch.empty(4096, device=dev)
    ops.sumsq_accumulate(g[: 8 * 50000], acc, ws, background=True)
    ops.sumsq_accumulate(g[8 * 50000:], acc, ws, background=False)
    want = g.double().pow(2).sum().item()
    assert abs(acc.item() - want) <= 1e-5 * want
    coef = torch.zeros(2, device=dev)
    ops.clip_coef(acc, 1.0, 0.5, coef)
    norm = want ** 0.5 * 0.5
    assert abs(coef[1].item() - norm) <= 1e-5 * norm and abs(coef[0].item() - 0.5 * min(1.0, 1.0 / (norm + 1e-6))) < 1e-7
    assert acc.item() == 0.0            # reset for the next step


@pytest.mark.parametrize("clip", [None, 0.05])
def test_engine_schedules_give_identical_parameters(clip):
    """Overlapped (per bucket, side stream, background grid), serial and deferred-sync schedules are the same arithmetic."""
    from cambrian_b200.engine import TrainEngine
    results = []
    for overlap, defer, bg in ((True, False, True), (False, False, False), (True, True, True)):
        cfg, model, batch = _setup()
        eng = TrainEngine(model, lr=1e-3, bucket_mb=8.0, overlap=overlap, max_grad_norm=clip, background_optimizer=bg)
        eng.defer_param_sync = defer
        losses = []
        for _ in range(3):
            eng.zero_grad()
            loss = model(**batch).loss
            loss.backward()
            eng.step()
            losses.append(float(loss.detach()))
        eng.wait_for_params()
        torch.cuda.synchronize()
        results.append((eng.flat_p.clone(), eng.master.clone(), losses, len(eng.buckets), eng.grad_norm() if clip else None))
    assert results[0][3] > 3
    assert results[0][2][0] == results[1][2][0] and results[0][2][2] < results[0][2][0]
    for tag, r in (("overlapped", results[0]), ("deferred", results[2])):
        assert_same_training(r[1], results[1][1], 1e-3, 3, f"engine schedules ({tag} vs serial, clip={clip})")
        assert rel_err(r[0], results[1][0]) < 2e-2
    if clip:
        assert results[0][4] > clip           # the clip was active
        assert abs(results[0][4] - results[1][4]) < 2e-3 * results[1][4]


def test_engine_grad_norm_matches_torch_and_clip_scales_the_update():
    from cambrian_b200.engine import TrainEngine
    cfg, model, batch = _setup()
    model(**batch).loss.backward()
    want = torch.linalg.vector_norm(torch.stack([p.grad.float().norm() for p in model.parameters() if p.grad is not None])).item()
    model.zero_grad(set_to_none=True)
    eng = TrainEngine(model, lr=1e-3, max_grad_norm=1e9)
    eng.zero_grad()
    model(**batch).loss.backward()
    eng.step()
    assert abs(eng.grad_norm() - want) < 2e-2 * want, (eng.grad_norm(), want)
    assert abs(eng._coef[0].item() - 1.0) < 1e-6


def test_engine_parameter_groups():
    """mm_vision_sampler_lr (cambrian_trainer.py:285-312): group lr for 'vision_sampler' / 'vision_query' parameters; no weight
    decay on norm and bias parameters."""
    from cambrian_b200.engine import TrainEngine
    cfg, model, batch = _setup()
    eng = TrainEngine(model, lr=1e-3, weight_decay=0.1, mm_vision_sampler_lr=0.0, bucket_mb=8.0)
    hp = dict(zip(eng.names, eng.hparams))
    assert hp["model.vision_sampler_0.layers.0.proj_in.weight"] == (0.0, 0.1)
    assert hp["model.vision_query"] == (0.0, 0.1)
    assert hp["model.layers.0.self_attn.q_proj.weight"] == (1e-3, 0.1)
    assert hp["model.layers.0.input_layernorm.weight"] == (1e-3, 0.0)
    assert hp["model.mm_projector.0.bias"] == (1e-3, 0.0)
    assert hp["model.vision_sampler_0.layers.0.norm.weight"] == (0.0, 0.0)
    before = {n: p.detach().clone() for n, p in zip(eng.names, eng.params)}
    eng.zero_grad()
    model(**batch).loss.backward()
    eng.step()
    torch.cuda.synchronize()
    for n, p in zip(eng.names, eng.params):
        moved = not torch.equal(p.detach(), before[n])
        if "vision_sampler" in n or "vision_query" in n:
            assert not moved, n          # lr 0 (and decoupled decay = lr * wd = 0)
        elif n.endswith("q_proj.weight") or "mm_projector.0.weight" in n:
            assert moved, n


def test_loss_scale_for_gradient_accumulation():
    from cambrian_b200.engine import TrainEngine
    cfg, model, batch = _setup()
    eng = TrainEngine(model, lr=0.0, loss_scale=0.5)
    eng.zero_grad()
    for _ in range(2):                       # two micro-batches, each contributing half
        model(**batch).loss.backward()
    eng._finalize_unwritten()
    acc = {n: p.main_grad.clone() for n, p in zip(eng.names, eng.params)}
    cfg2, model2, _ = _setup()
    eng2 = TrainEngine(model2, lr=0.0)
    eng2.zero_grad()
    model2(**batch).loss.backward()
    eng2._finalize_unwritten()
    for n, p in zip(eng2.names, eng2.params):
        assert rel_err(acc[n], p.main_grad) < 2e-2, n


def test_eval_forward_with_labels_does_not_touch_main_grad():
    from cambrian_b200.engine import TrainEngine
    cfg, model, batch = _setup()
    eng = TrainEngine(model, lr=1e-3)
    eng.zero_grad()
    model.eval()
    before = eng.flat_g.clone()
    writes = list(eng._writes)
    out = model(**batch)                     # grad mode on, eval mode: a validation pass
    assert torch.isfinite(out.loss)
    assert torch.equal(eng.flat_g, before) and eng._writes == writes


def test_embedding_gradient_is_deterministic_and_matches_index_add():
    """Backward of the embedding rows: per-token-id sums in position order (no atomics): bit-identical run to run, equal to
    an fp32 index_add reference, image-span positions excluded, accumulation across micro-batches supported."""
    from cambrian_b200 import ops
    torch.manual_seed(3)
    B, S, H, V, q = 2, 300, 256, 50, 4
    ids = torch.randint(0, V, (B, S), device=dev)
    ids[:, 7] = -200
    img_start = torch.tensor([7, 7], dtype=torch.int32, device=dev)
    dout = torch.randn(B, S, H, device=dev).bfloat16()
    span = q * (q + 1)
    keep = torch.ones(B, S, dtype=torch.bool, device=dev)
    keep[:, 7:7 + span] = False
    ref = torch.zeros(V, H, device=dev)
    ref.index_add_(0, ids.clamp(min=0)[keep], dout.float()[keep])
    outs = []
    for _ in range(3):
        d = torch.zeros(V, H, device=dev, dtype=torch.bfloat16)
        ops.embed_grad_sorted(dout, ids, img_start, d, q)
        outs.append(d)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    assert rel_err(outs[0], ref) < 4e-3                      # one bf16 rounding of an fp32 sum
    ops.embed_grad_sorted(dout, ids, img_start, outs[0], q)  # second micro-batch accumulates
    assert rel_err(outs[0], 2 * ref) < 8e-3
