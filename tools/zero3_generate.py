"""BASELINE config 5 — Cambrian-34B-shaped `generate` with ZeRO-3 style parameter sharding (cambrian_b200/sharded.py).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/zero3_generate.py [--layers 60] [--new-tokens 32] [--check]

Every rank builds the same random-init model (seeded), shards the decoder layers 1/N, and greedily decodes its own
sample (batch split 1 per GPU).  Prints prefill / per-token latency (CUDA events, max over ranks), gather count and the
per-GPU weight footprint.  --check first generates with the unsharded model and asserts token-identical output."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def yi34b_config(layers):
    from cambrian_b200.model.language_model.cambrian_llama import CambrianConfig
    cfg = CambrianConfig(hidden_size=7168, intermediate_size=20480, num_hidden_layers=layers, num_attention_heads=56,
                         num_key_value_heads=8, vocab_size=64000, max_position_embeddings=4096, rope_theta=5000000.0,
                         rms_norm_eps=1e-5)
    cfg.mm_vision_tower_aux_list = ["siglip/CLIP-ViT-SO400M-14-384", "openai/clip-vit-large-patch14-336",
                                    "facebook/dinov2-large-res336", "clip-convnext-XXL"]
    cfg.mm_vision_tower_aux_token_len_list = [576, 576, 576, 576]
    cfg.image_token_len = 576
    cfg.mm_projector_type = "sva"
    cfg.vision_hidden_size = 1024
    cfg.num_query_group = 1
    cfg.query_num_list = [576]
    cfg.connector_depth = 3
    cfg.connector_only = False
    cfg.num_of_vision_sampler_layers = min(9, max(1, layers // 7))      # 34B: 9 SVA sites, stride 7
    cfg.start_of_vision_sampler_layers = 0
    cfg.stride_of_vision_sampler_layers = 7 if layers >= 14 else 1
    cfg.image_position = 87
    return cfg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=60)
    ap.add_argument("--new-tokens", type=int, default=32)
    ap.add_argument("--prompt", type=int, default=128)
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    lrank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lrank)
    dev = torch.device("cuda", lrank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from cambrian_b200.model.language_model.cambrian_llama import CambrianLlamaForCausalLM
    from cambrian_b200.sharded import Zero3Inference
    cfg = yi34b_config(a.layers)
    torch.manual_seed(1234)
    torch.set_default_dtype(torch.bfloat16)
    with torch.device(dev):
        model = CambrianLlamaForCausalLM(cfg)
        for t in model.get_model().vision_tower_aux_list:
            t.load_model()
    torch.set_default_dtype(torch.float32)
    model.eval()
    n_layer_params = sum(p.numel() for p in model.get_model().layers.parameters())
    n_params = sum(p.numel() for p in model.parameters())
    g = torch.Generator().manual_seed(100 + rank)
    ids = torch.randint(3, cfg.vocab_size, (1, a.prompt), generator=g)
    ids[0, cfg.image_position] = -200
    images = [torch.randn(1, 3, r, r, generator=g).bfloat16().to(dev) for r in bench.CONFIGS["8b-ddp"]["res"]]
    ids = ids.to(dev)

    def run(n_new):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        toks = model.generate(ids, images=images, image_sizes=[(336, 336)], max_new_tokens=n_new)
        e1.record()
        torch.cuda.synchronize()
        return toks, e0.elapsed_time(e1)

    ref = None
    if a.check:
        ref, _ = run(a.new_tokens)
    z = Zero3Inference(model)
    torch.cuda.empty_cache()
    run(2)                                              # warm-up (tensor maps, allocator)
    _, t1 = run(1)                                      # prefill + first token
    toks, tn = run(a.new_tokens)
    if ref is not None:
        assert torch.equal(ref, toks), f"rank {rank}: sharded tokens differ from the unsharded model"
    per_tok = (tn - t1) / max(a.new_tokens - 1, 1)
    t = torch.tensor([t1, per_tok], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        fp = z.bytes_per_gpu()
        print(json.dumps(dict(config="Cambrian-34B-shaped generate, ZeRO-3 sharded decoder layers", n_gpus=world,
                              layers=a.layers, params_total=n_params, params_sharded=n_layer_params,
                              batch_per_gpu=1, prompt_tokens=a.prompt + 599, new_tokens=a.new_tokens,
                              prefill_plus_first_token_ms=float(t[0]), decode_ms_per_token=float(t[1]),
                              tokens_per_s_all_gpus=world * 1000.0 / float(t[1]),
                              gathers=z.gathers, shard_gb_per_gpu=fp["shards"] / 2 ** 30,
                              staging_gb_per_gpu=fp["staging"] / 2 ** 30,
                              peak_mem_gb=torch.cuda.max_memory_allocated() / 2 ** 30,
                              token_identical_to_unsharded=bool(a.check))))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
