"""Multi-GPU (NCCL) tests of the data path: need >= 2 visible CUDA devices (`gpurun --gpus 2`), skipped otherwise.
ZeRO-2 must reach the NCCL reduce_scatter_tensor / in-place all_gather_into_tensor path and reproduce the DDP engine's
parameters; ranks with different label layouts must issue identical collective sequences."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q, mode):
    import sys
    import torch.distributed as dist
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    sys.path.insert(0, os.path.dirname(here))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import test_modules_gpu as T
        from helpers import rel_err, tiny_cambrian_config
        from cambrian_b200.engine import TrainEngine
        dev = f"cuda:{rank}"
        T.dev = dev
        out = {}
        calls = {"rs": 0, "ag": 0}
        rs0, ag0 = dist.reduce_scatter_tensor, dist.all_gather_into_tensor

        def rs(*a, **k):
            calls["rs"] += 1
            return rs0(*a, **k)

        def ag(*a, **k):
            calls["ag"] += 1
            return ag0(*a, **k)
        dist.reduce_scatter_tensor, dist.all_gather_into_tensor = rs, ag
        for zero, clip in ((0, None), (2, None), (0, 0.05), (2, 0.05)):
            cfg = tiny_cambrian_config()
            cfg.fused_lm_loss = True
            model = T._build_tiny_model(cfg)          # same seed on every rank -> identical replicas
            model.train()
            eng = TrainEngine(model, lr=1e-3, bucket_mb=8.0, zero_stage=zero, max_grad_norm=clip)
            eng.defer_param_sync = True
            losses = []
            for step in range(3):
                ids, labels, attn, pos, images, masks = T._tiny_batch(cfg)
                g = torch.Generator().manual_seed(100 * step + rank)
                labels = labels.clone()                                     # label layout differs per rank AND per step
                drop = torch.rand(labels.shape, generator=g) < 0.3
                labels[drop] = -100
                images = [i + 0.1 * rank for i in images]
                batch = dict(input_ids=ids.to(dev), labels=labels.to(dev), attention_mask=attn.to(dev),
                             position_ids=pos.to(dev), images=[i.to(dev).bfloat16() for i in images],
                             image_aux_attention_masks_list=[m.to(dev) for m in masks])
                from cambrian_b200.train.collator import valid_label_ranges
                ranges, nv = valid_label_ranges(labels)
                eng.zero_grad()
                loss = model(**batch, label_ranges=ranges, num_valid_labels=nv).loss
                loss.backward()
                eng.step()
                losses.append(float(loss.detach()))
            eng.wait_for_params()
            torch.cuda.synchronize()
            out[(zero, clip)] = (eng.flat_p[: eng.offsets[-1] + eng.params[-1].numel()].float().cpu(), losses, eng._overlap_ok)
            if zero == 2:
                assert eng.master.numel() * world == eng.total        # optimizer state really is sharded
        ok = calls["rs"] > 0 and calls["ag"] > 0                       # the NCCL ZeRO-2 path was reached
        msg = f"rs={calls['rs']} ag={calls['ag']}"
        for clip in (None, 0.05):
            a, b = out[(0, clip)], out[(2, clip)]
            n = min(a[0].numel(), b[0].numel())
            e = rel_err(b[0][:n], a[0][:n])
            ok &= e < 2e-2 and all(abs(x - y) < 2e-2 * abs(x) for x, y in zip(a[1], b[1])) and a[2] and b[2]
            msg += f" clip={clip}: zero2-vs-ddp rel_err {e:.2e} losses {a[1]} {b[1]} overlap_ok {a[2]} {b[2]};"
        # replicas must agree across ranks
        t = out[(2, 0.05)][0].to(dev)
        lo, hi = t.clone(), t.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        ok &= bool(torch.equal(lo, hi))
        q.put((rank, bool(ok), msg))
    except Exception as ex:  # noqa: BLE001
        import traceback
        q.put((rank, False, traceback.format_exc()[-1500:]))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_zero2_nccl_matches_ddp_with_rank_varying_labels():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 23000 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, "zero2")) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(60)
    assert all(r[1] for r in res), res


def _symm_worker(rank, world, port, q):
    import sys
    import torch.distributed as dist
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    sys.path.insert(0, os.path.dirname(here))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from cambrian_b200.comm import SymmetricAllReduce
        dev = torch.device("cuda", rank)
        msgs, ok = [], True
        for use_mc in (True, False):
            n = 8 * world * 4099 * 3
            ar = SymmetricAllReduce(n, dev, ctas=8 if use_mc else 0, use_multicast=use_mc)
            mode = "multimem" if ar.multicast else "p2p"
            g = torch.Generator(device="cpu").manual_seed(7)
            base = torch.randint(-64, 64, (world, n), generator=g).to(torch.bfloat16)     # exactly representable, exact sums
            want = base.float().sum(0)
            for it in range(3):                                                           # back-to-back launches, sub-ranges
                ar.buf.copy_(base[rank].to(dev))
                torch.cuda.synchronize()
                dist.barrier()
                lo, hi = (0, n) if it == 0 else (8 * world * 100 * it, n - 8 * world * 50 * it)
                ar.all_reduce_(lo, hi)
                torch.cuda.synchronize()
                got = ar.buf.float().cpu()
                exp = base[rank].float().clone()
                exp[lo:hi] = want[lo:hi]
                good = torch.equal(got, exp)
                ok &= good
                msgs.append(f"{mode} it{it} {'ok' if good else 'MISMATCH max ' + str((got - exp).abs().max().item())}")
            dist.barrier()
        q.put((rank, bool(ok), "; ".join(msgs)))
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, False, traceback.format_exc()[-1500:]))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_symmetric_allreduce_kernel_multimem_and_p2p():
    """cb_allreduce_symm_bf16 on 2 GPUs: NVLS multimem path (when the fabric offers multicast) and the peer load/store path,
    whole buffer and sub-ranges, back-to-back launches: exact sums, untouched elements outside the range."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 25000 + os.getpid() % 2000
    procs = [ctx.Process(target=_symm_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(60)
    print(res)
    assert all(r[1] for r in res), res


def _engine_symm_worker(rank, world, port, q):
    import sys
    import torch.distributed as dist
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    sys.path.insert(0, os.path.dirname(here))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import test_modules_gpu as T
        from helpers import rel_err, tiny_cambrian_config
        from cambrian_b200.engine import TrainEngine
        dev = f"cuda:{rank}"
        T.dev = dev
        out = {}
        for coll in ("nccl", "multimem"):
            cfg = tiny_cambrian_config()
            cfg.fused_lm_loss = True
            model = T._build_tiny_model(cfg)
            model.train()
            eng = TrainEngine(model, lr=1e-3, bucket_mb=8.0, max_grad_norm=0.05, collective=coll)
            eng.defer_param_sync = True
            losses = []
            for step in range(3):
                ids, labels, attn, pos, images, masks = T._tiny_batch(cfg)
                images = [i + 0.1 * rank for i in images]
                batch = dict(input_ids=ids.to(dev), labels=labels.to(dev), attention_mask=attn.to(dev),
                             position_ids=pos.to(dev), images=[i.to(dev).bfloat16() for i in images],
                             image_aux_attention_masks_list=[m.to(dev) for m in masks])
                eng.zero_grad()
                loss = model(**batch).loss
                loss.backward()
                eng.step()
                losses.append(float(loss.detach()))
            eng.wait_for_params()
            torch.cuda.synchronize()
            out[coll] = (eng.flat_p[: eng.offsets[-1] + eng.params[-1].numel()].float().cpu(), losses, eng.grad_norm())
        n = min(out["nccl"][0].numel(), out["multimem"][0].numel())
        e = rel_err(out["multimem"][0][:n], out["nccl"][0][:n])
        ok = e < 2e-2 and abs(out["nccl"][2] - out["multimem"][2]) < 1e-2 * out["nccl"][2]
        q.put((rank, bool(ok), f"multimem-vs-nccl rel_err {e:.2e} losses {out['nccl'][1]} {out['multimem'][1]} norms "
                               f"{out['nccl'][2]:.4f} {out['multimem'][2]:.4f}"))
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, False, traceback.format_exc()[-1500:]))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_engine_with_in_switch_allreduce_matches_nccl():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 27000 + os.getpid() % 2000
    procs = [ctx.Process(target=_engine_symm_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(60)
    print(res)
    assert all(r[1] for r in res), res
