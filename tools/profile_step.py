"""Runs the bench workload for a few warm-up steps, then brackets ONE training step with cudaProfilerStart/Stop so that
`ncu --profile-from-start off ...` captures exactly that step (see profiles/README.md for the commands)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--micro-batch", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--small", action="store_true")
    ap.add_argument("--recompute", type=int, default=0)
    a = ap.parse_args()
    args = argparse.Namespace(small=a.small)
    from cambrian_b200.engine import TrainEngine
    from cambrian_b200.model.language_model.cambrian_llama import CambrianLlamaForCausalLM
    dev = torch.device("cuda", 0)
    cfg = bench.cambrian_8b_config(args)
    res = bench.TOWER_RES if not a.small else [384, 336, 336, 256]
    torch.manual_seed(1234)
    torch.set_default_dtype(torch.bfloat16)
    with torch.device(dev):
        model = CambrianLlamaForCausalLM(cfg)
        for t in model.get_model().vision_tower_aux_list:
            t.load_model()
    torch.set_default_dtype(torch.float32)
    model.train()
    model.get_model().gradient_checkpointing = bool(a.recompute)
    eng = TrainEngine(model)
    hb, (nv, lr) = bench.make_host_batch(cfg, a.micro_batch, 2048, 0, res)
    db, _ = bench.to_device(hb, dev)
    pos = [cfg.image_position] * a.micro_batch

    def step():
        eng.zero_grad()
        out = model(**db, num_valid_labels=nv, image_positions=pos, label_ranges=lr)
        out.loss.backward()
        eng.step()

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
