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


def test_emulation_is_test_only():
    """The stand-ins live in tests/ and nothing under cambrian_b200/ refers to them."""
    root = os.path.join(os.path.dirname(HERE), "cambrian_b200")
    for d, _, files in os.walk(root):
        for f in files:
            if f.endswith(".py"):
                assert "ops_emulation" not in open(os.path.join(d, f)).read(), os.path.join(d, f)
