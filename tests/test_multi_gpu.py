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
