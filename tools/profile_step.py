"""Runs the bench workload for a few warm-up steps, then brackets ONE training step with cudaProfilerStart/Stop so that
`ncu --profile-from-start off ...` captures exactly that step (see profiles/README.md for the commands).  With CB_NVTX=1
every block of the step carries an NVTX range (decoder_layer.fwd / .bwd, sva_layer.*, tower.*, lm_head_loss.*,
optimizer.bucketN)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="8b-ddp")
    ap.add_argument("--micro-batch", type=int, default=0)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--small", action="store_true")
    ap.add_argument("--recompute", type=int, default=0)
    a = ap.parse_args()
    from cambrian_b200.engine import TrainEngine
    from cambrian_b200.model.language_model.cambrian_llama import CambrianLlamaForCausalLM
    dev = torch.device("cuda", 0)
    C = bench.CONFIGS[a.config]
    cfg = bench.build_config(a.config, a.small)
    res = C["res"] if not a.small else [r if r < 1024 else 256 for r in C["res"]]
    torch.manual_seed(1234)
    torch.set_default_dtype(torch.bfloat16)
    with torch.device(dev):
        model = CambrianLlamaForCausalLM(cfg)
        for t in model.get_model().vision_tower_aux_list:
            t.load_model()
    torch.set_default_dtype(torch.float32)
    model.train()
    model.get_model().gradient_checkpointing = bool(a.recompute)
    eng = TrainEngine(model, max_grad_norm=1.0)
    eng.defer_param_sync = True
    B = a.micro_batch or C["micro_batch"]
    S = C["seq"] if not a.small else 1024
    db, _ = bench.to_device(bench.make_host_batch(cfg, B, S, 0, res, True), dev)

    def step():
        eng.zero_grad()
        out = model(**db)
        out.loss.backward()
        eng.step()

    for _ in range(a.warmup):
        step()
    eng.wait_for_params()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    step()
    eng.wait_for_params()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
