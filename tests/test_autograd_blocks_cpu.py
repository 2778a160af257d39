"""CPU coverage of the HOST logic of the SVA autograd blocks (cambrian_b200/autograd.py, model/vision_sampler.py): the
kernels are replaced by plain-torch stand-ins (tests/ops_emulation.py — test infrastructure, monkeypatched for one test
at a time), the block code itself — argument order, saved tensors, gradient routing, layout conventions, main_grad
accumulation — is the product's.  Checked against the oracle in fp32 (itself pinned to the reference).  The kernels'
numerics are NOT covered here: that is what `-m gpu` does."""
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ops_emulation  # noqa: E402
from oracle import cambrian_oracle as O  # noqa: E402

pytestmark = pytest.mark.skipif(torch.cuda.is_available(), reason="kernel stand-ins are for GPU-less machines only; on a GPU "
                                                                  "box the same logic is covered by the -m gpu suite")


def _fro(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-30)).item()


def _case(layer_type, q_dim, rs, layers, use_mask, natural, main_grad=False, seed=0):
    from cambrian_b200.model.vision_sampler import VisionTokenSampler
    torch.manual_seed(seed)
    T = len(rs)
    m = VisionTokenSampler(q_dim, 1024, [1024] * T, rs, 1024, layers, layer_type=layer_type)
    with torch.no_grad():
        for n_, p in m.named_parameters():
            if "pos_embed" in n_:
                p.mul_(0.1)
    sd32 = {k: v.detach().bfloat16().float() for k, v in m.state_dict().items()}
    m = m.to(torch.bfloat16)
    if main_grad:      # what TrainEngine attaches: gradients accumulate into preallocated buffers, autograd sees None
        for p in m.parameters():
            p.main_grad = torch.zeros_like(p)
            p._cb_fresh = set()
    B, qs = 2, 3
    n = B * qs * qs
    q = torch.randn(n, 1, q_dim).bfloat16()
    c = torch.randn(n, 1, 1024).bfloat16()
    feats_nat = [torch.randn(B, (r * qs) ** 2, 1024).bfloat16() for r in rs]
    feats_win = [O.window_rearrange(f.float(), qs) for f in feats_nat]
    masks = []
    for r in rs:
        mk = torch.rand(n, r * r) > 0.3 if use_mask else torch.ones(n, r * r, dtype=torch.bool)
        mk[mk.sum(1) == 0] = True
        masks.append(mk)
    qg, cg = q.clone().requires_grad_(), c.clone().requires_grad_()
    fin = [f.clone().requires_grad_() for f in (feats_nat if natural else [w.bfloat16() for w in feats_win])]
    out = m(qg, cg, *fin, *masks, natural_layout=(B, qs) if natural else None)
    do = torch.randn_like(out)
    out.backward(do)
    sdg = {k: v.clone().requires_grad_() for k, v in sd32.items()}
    q32, c32 = q.float().requires_grad_(), c.float().requires_grad_()
    f32 = [w.clone().requires_grad_() for w in feats_win]
    ref = O.sva_sampler(sdg, "", q32, c32, f32, masks, layers, layer_type=layer_type)
    ref.backward(do.float())
    errs = {"out": _fro(out, ref), "dq": _fro(qg.grad, q32.grad), "dc": _fro(cg.grad, c32.grad)}
    for i, (f, fr) in enumerate(zip(fin, f32)):
        g = f.grad.float()
        errs[f"dfeat{i}"] = _fro(O.window_rearrange(g, qs) if natural else g, fr.grad)
    for k, p in m.named_parameters():
        if layer_type == "sep" and "k_proj.0.bias" in k:
            continue        # structurally zero: one softmax per tower with a single query cancels a constant key offset
        got = p.main_grad if main_grad else p.grad
        if main_grad:
            assert p.grad is None, f"{k}: autograd received a gradient although main_grad is attached"
        errs["d" + k] = _fro(got, sdg[k].grad)
    worst = max(errs.items(), key=lambda kv: kv[1])
    assert worst[1] < 3e-2, (worst, errs)


@pytest.mark.parametrize("q_dim,rs,layers,use_mask,natural", [
    (256, [2, 1, 3], 2, True, True), (1024, [1, 1, 2], 1, False, False), (512, [2], 1, True, True),
    (256, [1], 1, False, False), (256, [1, 2, 1, 1], 1, True, False)])
def test_sep_layer_host_logic(monkeypatch, q_dim, rs, layers, use_mask, natural):
    ops_emulation.install(monkeypatch)
    _case("sep", q_dim, rs, layers, use_mask, natural)


@pytest.mark.parametrize("q_dim,rs,layers,use_mask,natural", [
    (256, [1, 1, 1, 2], 2, True, True), (1024, [1, 1, 1, 1], 1, False, True), (256, [2, 1, 3], 1, True, False)])
def test_joint_layer_host_logic(monkeypatch, q_dim, rs, layers, use_mask, natural):
    ops_emulation.install(monkeypatch)
    _case("joint", q_dim, rs, layers, use_mask, natural)


@pytest.mark.parametrize("layer_type", ["joint", "sep"])
def test_gradients_accumulate_into_main_grad(monkeypatch, layer_type):
    """With TrainEngine-style `main_grad` buffers every weight gradient lands in the buffer (first write of the step
    overwrites, later ones accumulate) and autograd receives None."""
    ops_emulation.install(monkeypatch)
    _case(layer_type, 256, [2, 1, 1], 2, True, True, main_grad=True)


def test_query_grid_resize_backward_host_logic(monkeypatch):
    """ResizeTokenGridFn (cambrian_arch.py:394-401) routes the gradient through the adjoint with the right grid sides."""
    import torch.nn.functional as F
    from cambrian_b200.autograd import ResizeTokenGridFn
    ops_emulation.install(monkeypatch)
    torch.manual_seed(0)
    x = torch.randn(2, 4, 64).bfloat16().requires_grad_()
    y = ResizeTokenGridFn.apply(x, 2, 4)
    g = torch.randn_like(y)
    y.backward(g)
    xf = x.detach().float().view(2, 2, 2, 64).permute(0, 3, 1, 2).requires_grad_()
    yf = F.interpolate(xf, size=(4, 4), mode="bilinear", align_corners=False)
    (gx,) = torch.autograd.grad(yf, xf, g.float().view(2, 4, 4, 64).permute(0, 3, 1, 2))
    assert _fro(y, yf.permute(0, 2, 3, 1).reshape(2, 16, 64)) < 1e-2
    assert _fro(x.grad, gx.permute(0, 2, 3, 1).reshape(2, 4, 64)) < 1e-2


# ------------------------------------------------------------------------------------------------ N > 1 on gloo
def _engine_worker(rank, world, port, q, zero, clip):
    """One data-parallel rank: a 1-layer SVA sampler under TrainEngine, real backward (the blocks' weight-gradient sites
    write into the flat gradient buffer and notify the engine), three steps on rank- and step-dependent data."""
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(max(1, (os.cpu_count() or 2) // (2 * world)))   # two workers share the host: no oversubscription
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ops_emulation.install()
        from cambrian_b200.engine import TrainEngine
        from cambrian_b200.model.vision_sampler import VisionTokenSampler
        torch.manual_seed(0)                                   # identical replicas
        rs = [1, 2]
        m = VisionTokenSampler(256, 1024, [1024] * 2, rs, 1024, 1).to(torch.bfloat16)
        calls = {"during_backward": 0, "in_backward": False}
        ar0 = dist.all_reduce

        def counting_all_reduce(*a, **k):
            calls["during_backward"] += int(calls["in_backward"])
            return ar0(*a, **k)
        dist.all_reduce = counting_all_reduce
        eng = TrainEngine(m, lr=1e-3, bucket_mb=2.0, zero_stage=zero, max_grad_norm=clip)
        grads = []                                             # this rank's bf16 gradients per step (before the reduction)
        B, qs = 2, 2
        n = B * qs * qs
        for step in range(3):
            g = torch.Generator().manual_seed(1000 * step + rank)
            q_ = torch.randn(n, 1, 256, generator=g).bfloat16()
            c_ = torch.randn(n, 1, 1024, generator=g).bfloat16()
            feats = [torch.randn(B, (r * qs) ** 2, 1024, generator=g).bfloat16() for r in rs]
            eng.zero_grad()
            out = m(q_, c_, *feats, natural_layout=(B, qs))
            loss = out.float().pow(2).mean()
            calls["in_backward"] = True
            loss.backward()
            calls["in_backward"] = False
            eng.step()
        ok = True
        msg = f"buckets {len(eng.buckets)} overlap_ok {eng._overlap_ok} collectives during backward {calls['during_backward']}"
        ok &= len(eng.buckets) >= 3 and eng._overlap_ok
        # steps 2 and 3 launch every bucket's collective from inside backward (step 1 learns the contribution counts)
        ok &= calls["during_backward"] >= 2 * len(eng.buckets)
        # replicas agree bit for bit
        t = eng.flat_p.float().clone()
        lo, hi = t.clone(), t.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        ok &= bool(torch.equal(lo, hi))
        if zero == 2:
            ok &= eng.master.numel() * world == eng.total
        q.put((rank, bool(ok), msg, t.numpy() if rank == 0 else None))   # numpy: pickled by value
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, False, traceback.format_exc()[-2000:], None))
    finally:
        dist.destroy_process_group()


def _single_process_reference(clip, world=2):
    """The same three steps in ONE process: per-rank gradients through the same blocks (plain autograd path, no engine),
    summed like the collective does (bf16), then torch-free AdamW in fp32 — what the data-parallel job must reproduce."""
    from cambrian_b200.model.vision_sampler import VisionTokenSampler
    torch.manual_seed(0)
    rs = [1, 2]
    m = VisionTokenSampler(256, 1024, [1024] * 2, rs, 1024, 1).to(torch.bfloat16)
    params = [p for _, p in m.named_parameters()]
    master = [p.detach().float().clone() for p in params]
    m1 = [torch.zeros_like(x) for x in master]
    v1 = [torch.zeros_like(x) for x in master]
    B, qs = 2, 2
    n = B * qs * qs
    for step in range(3):
        tot = [torch.zeros_like(p) for p in params]
        for rank in range(world):
            g = torch.Generator().manual_seed(1000 * step + rank)
            q_ = torch.randn(n, 1, 256, generator=g).bfloat16()
            c_ = torch.randn(n, 1, 1024, generator=g).bfloat16()
            feats = [torch.randn(B, (r * qs) ** 2, 1024, generator=g).bfloat16() for r in rs]
            m.zero_grad(set_to_none=True)
            m(q_, c_, *feats, natural_layout=(B, qs)).float().pow(2).mean().backward()
            tot = [t + p.grad for t, p in zip(tot, params)]                    # bf16 sum, as the all-reduce of bf16 buckets
        scale = 1.0 / world
        if clip:
            norm = torch.sqrt(sum(t.float().pow(2).sum() for t in tot)) * scale
            scale *= min(1.0, clip / (float(norm) + 1e-6))
        for i, p in enumerate(params):
            ops_emulation.adamw(master[i], m1[i], v1[i], tot[i], p.data, 1e-3, 0.9, 0.999, 1e-8, 0.0, step + 1, grad_scale=scale)
    return torch.cat([(x.reshape(-1)) for x in master]), [p.numel() for p in params]


@pytest.mark.parametrize("zero,clip", [(0, None), (0, 0.05), (2, 0.05)])
def test_data_parallel_engine_two_ranks_gloo_real_backward(monkeypatch, zero, clip):
    """SURVEY §8e on CPU: two gloo ranks, TrainEngine (DDP all-reduce buckets / ZeRO-2 sharded optimizer) driven by the REAL
    backward of the SVA blocks — per-bucket collectives are launched from inside backward once the contribution counts are
    learned, replicas stay bit-identical, and the result equals a single-process run over both ranks' data."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() + 17 * zero + (3 if clip else 0)) % 2000
    procs = [ctx.Process(target=_engine_worker, args=(r, 2, port, q, zero, clip)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=300) for _ in procs), key=lambda r: r[0])
    for p in procs:
        p.join(60)
    assert all(r[1] for r in res), [r[2] for r in res]
    ops_emulation.install(monkeypatch)
    ref, sizes = _single_process_reference(clip)
    got = torch.from_numpy(res[0][3])
    # the engine's flat buffer pads every parameter to 8 elements: compare parameter by parameter
    o = o_ref = 0
    worst = 0.0
    for nel in sizes:
        a, b = got[o:o + nel], ref[o_ref:o_ref + nel].bfloat16().float()
        worst = max(worst, (a - b).abs().max().item())
        o += (nel + 7) // 8 * 8
        o_ref += nel
    # bf16 parameters one Adam trajectory apart by at most rounding: a sign flip of a noise-level gradient moves an
    # element by 2 * lr per step, bf16 storage adds one ulp (0.8 % of a value of order one)
    assert worst <= 2.05 * 1e-3 * 3 + 8e-3, worst


def test_emulation_is_test_only():
    """The stand-ins live in tests/ and nothing under cambrian_b200/ refers to them."""
    root = os.path.join(os.path.dirname(HERE), "cambrian_b200")
    for d, _, files in os.walk(root):
        for f in files:
            if f.endswith(".py"):
                assert "ops_emulation" not in open(os.path.join(d, f)).read(), os.path.join(d, f)
