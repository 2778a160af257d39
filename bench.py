#!/usr/bin/env python
"""bench.py — Cambrian-1-8B training-step throughput on N x B200 (BASELINE.json metric) + roofline + CPU baseline.

    python bench.py --gpus N --steps K --warmup W            # this framework (hand-written sm_100a kernels)
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm's CPU port on the host cores

Workload: Llama-3-8B decoder (32 L, H 4096, 32/8 heads x 128, FFN 14336, vocab 128256),
four towers (SigLIP-SO400M/14@384, CLIP ViT-L/14@336, DINOv2 ViT-L/14@336, ConvNeXt-XXL@1024 — each interpolated
to 576 tokens), SVA connector depth 3 + 10 in-LLM SVA layers (start 0, stride 3), image_position 91, 576 visual tokens +
24 newlines spliced into a 2048-token sequence, bf16 compute with fp32 master weights + AdamW, one bucketed NCCL
all-reduce of the trainable gradients per step.  Random-init weights, synthetic images / ids (no network).
One "step" = towers fwd + (connector + decoder + loss) fwd/bwd + gradient all-reduce + AdamW on one micro-batch per GPU.
Weak scaling: the per-GPU micro-batch is fixed; `value` is whole-job samples/s.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "train_samples_per_sec"
UNIT = "samples/s"

# algorithmic forward TFLOP per sample (ConvNeXt @1024)
TOWERS_TF = 0.381 + 0.665 + 0.381 + 6.335
TRAIN_FWD_TF = 0.012 + 0.308 + 0.024 + 31.84
STEP_TF = TOWERS_TF + 3 * TRAIN_FWD_TF  # 104.3


def cambrian_8b_config(args):
    from cambrian_b200.model.language_model.cambrian_llama import CambrianConfig
    small = args.small
    cfg = CambrianConfig(
        hidden_size=4096 if not small else 1024, intermediate_size=14336 if not small else 2048,
        num_hidden_layers=32 if not small else 4, num_attention_heads=32 if not small else 8,
        num_key_value_heads=8 if not small else 2, vocab_size=128256 if not small else 8192,
        max_position_embeddings=8192, rope_theta=500000.0, rms_norm_eps=1e-5)
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
    cfg.num_of_vision_sampler_layers = 10 if not small else 2
    cfg.start_of_vision_sampler_layers = 0
    cfg.stride_of_vision_sampler_layers = 3 if not small else 2
    cfg.image_position = 91
    cfg.fused_lm_loss = True
    cfg.lm_loss_chunk = 4096
    if small:
        cfg.convnext_config_overrides = dict(depths=(1, 1, 2, 1), dims=(96, 192, 384, 768), image_size=256)
    return cfg


TOWER_RES = [384, 336, 336, 1024]


def make_host_batch(cfg, B, S, seed, res):
    """Synthetic batch on the host (pinned), shaped like DataCollatorForSupervisedDataset's output
    (train_fsdp.py:1168-1236): expanded ids, labels, attention mask, position ids, one image per tower."""
    g = torch.Generator().manual_seed(seed)
    q = int(cfg.image_token_len ** 0.5)
    span = q * (q + 1)
    p0 = cfg.image_position
    ids = torch.randint(3, cfg.vocab_size, (B, S), generator=g)
    ids[:, p0] = -200
    ids[:, p0 + 1:p0 + span] = 0
    labels = ids.clone()
    labels[:, :p0 + span] = -100
    attn = torch.ones(B, S, dtype=torch.bool)
    pos = torch.arange(S)[None].expand(B, S).contiguous()
    images = [torch.randn(B, 3, r, r, generator=g).bfloat16() for r in res]
    n_valid = int((labels[:, 1:] != -100).sum())
    from cambrian_b200.train.collator import valid_label_ranges
    ranges, nv2 = valid_label_ranges(labels)        # host-side collator hint: rows that can carry a loss term
    assert nv2 == n_valid
    batch = dict(input_ids=ids, labels=labels, attention_mask=attn, position_ids=pos, images=images)
    for k, v in batch.items():
        batch[k] = [t.pin_memory() for t in v] if isinstance(v, list) else v.pin_memory()
    return batch, (n_valid, ranges)


def to_device(batch, dev):
    out = {}
    nbytes = 0
    for k, v in batch.items():
        if isinstance(v, list):
            out[k] = [t.to(dev, non_blocking=True) for t in v]
            nbytes += sum(t.numel() * t.element_size() for t in v)
        else:
            out[k] = v.to(dev, non_blocking=True)
            nbytes += v.numel() * v.element_size()
    return out, nbytes


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([x.strip() for x in o.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max((float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()), default=None),
                "samples": len(self.rows), "reasons": reasons}


class KernelTimer:
    """CUDA-event timing of individual C-ABI launches on the launching stream (torch's current stream)."""

    def __init__(self):
        self.events = []  # (kind, work, start, end)
        self.shapes = []  # per GEMM launch: (out shape, K, a_mn, b_mn, act, bias?, residual?, event index)
        self.enabled = False

    def wrap(self, ops_mod):
        timer = self
        g0, f0, b0 = ops_mod.gemm, ops_mod.sva_window_attn_fwd, ops_mod.sva_window_attn_bwd

        def gemm(a, b, **kw):
            if not timer.enabled:
                return g0(a, b, **kw)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            out = g0(a, b, **kw)
            e.record()
            K = a.shape[-2] if kw.get("a_mn") else a.shape[-1]
            timer.events.append(("gemm", 2.0 * out.numel() * K, s, e))
            timer.shapes.append((tuple(out.shape), K, bool(kw.get("a_mn")), bool(kw.get("b_mn")), kw.get("act"),
                                 kw.get("bias") is not None, kw.get("residual") is not None, len(timer.events) - 1))
            return out

        def sva_f(q, ks, vs, masks, rs, batch, q_side, **kw):
            if not timer.enabled:
                return f0(q, ks, vs, masks, rs, batch, q_side, **kw)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            out = f0(q, ks, vs, masks, rs, batch, q_side, **kw)
            e.record()
            byts = 2 * (2 * sum(k.numel() for k in ks) + 2 * q.numel()) + sum(0 if m is None else m.numel() for m in (masks or []))
            timer.events.append(("sva_fwd", float(byts), s, e))
            return out

        def sva_b(q, out_, dout, lse, ks, vs, masks, rs, batch, q_side, **kw):
            if not timer.enabled:
                return b0(q, out_, dout, lse, ks, vs, masks, rs, batch, q_side, **kw)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            r = b0(q, out_, dout, lse, ks, vs, masks, rs, batch, q_side, **kw)
            e.record()
            byts = 2 * (4 * sum(k.numel() for k in ks) + 4 * q.numel())  # read K,V,Q,O,dO; write dK,dV,dQ
            timer.events.append(("sva_bwd", float(byts), s, e))
            return r

        gs0 = ops_mod.gemm_swiglu

        def gemm_swiglu(x2d, w_gu, gu_out=None, act_out=None):
            # the decoder's fused gate/up projection + SwiGLU (the dominant launch): 2 * M * 2F * K FLOP
            if not timer.enabled:
                return gs0(x2d, w_gu, gu_out, act_out)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            out = gs0(x2d, w_gu, gu_out, act_out)
            e.record()
            M, K = x2d.shape
            timer.events.append(("gemm", 2.0 * M * w_gu.shape[0] * K, s, e))
            timer.shapes.append(((M, w_gu.shape[0]), K, False, False, "swiglu_pair", False, False, len(timer.events) - 1))
            return out

        ops_mod.gemm, ops_mod.sva_window_attn_fwd, ops_mod.sva_window_attn_bwd = gemm, sva_f, sva_b
        ops_mod.gemm_swiglu = gemm_swiglu

    def shape_table(self, top=24):
        """Per-shape GEMM time / rate inside the timed step (CB_BENCH_SHAPES=1 prints it to stderr)."""
        agg = {}
        for rec in self.shapes:
            _, work, s, e = self.events[rec[-1]]
            a = agg.setdefault(rec[:-1], [0.0, 0.0, 0])
            a[0] += work
            a[1] += s.elapsed_time(e)
            a[2] += 1
        rows = sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]
        return [f"{ms:9.2f} ms {n:5d}x {fl / ms / 1e9:7.0f} TF/s  out={k[0]} K={k[1]} a_mn={int(k[2])} b_mn={int(k[3])} "
                f"act={k[4]} bias={int(k[5])} res={int(k[6])}" for k, (fl, ms, n) in rows]

    def dominant(self):
        """The fused gate/up + SwiGLU GEMM launches alone (the kernel `roofline.traffic` was captured on): per-launch
        algorithmic FLOP / average CUDA-event duration."""
        fl = ms = 0.0
        n = 0
        shape = None
        for rec in self.shapes:
            if rec[4] != "swiglu_pair":
                continue
            _, work, s, e = self.events[rec[-1]]
            fl += work
            ms += s.elapsed_time(e)
            n += 1
            shape = (rec[0], rec[1])
        if n == 0 or ms <= 0:
            return None
        return dict(kernel="gemm_bf16_tcgen05_2cta<256,0,0,swiglu_pair> (cb_gemm_swiglu_bf16)",
                    shape=f"M={shape[0][0]} 2F={shape[0][1]} K={shape[1]}", launches_timed=n,
                    flop_per_launch=fl / n, avg_launch_ms=ms / n, achieved=fl / (ms / 1000.0) / 1e12)

    def totals(self):
        agg = {}
        for kind, work, s, e in self.events:
            ms = s.elapsed_time(e)
            a = agg.setdefault(kind, [0.0, 0.0, 0])
            a[0] += work
            a[1] += ms
            a[2] += 1
        return agg


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", 1400.0), d.get("bf16_tflops", 1590.0), "measured"
    return 6650.0, 1400.0, 1590.0, "fallback"


# ------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference algorithm on a bounded sample of the same workload
# ------------------------------------------------------------------------------------------------------------------
def cpu_reference_sample(threads=None):
    """Times, on the host cores, a bounded slice of ONE training sample of the same workload with the fp32 oracle
    (oracle/cambrian_oracle.py — the port of the reference's PyTorch path): 3-layer SVA connector over 4 x 576 x 1024
    grids (fwd+bwd), ONE Llama-3-8B-shaped decoder layer at S=2048 (fwd+bwd), ONE in-LLM SVA layer (fwd+bwd), and the
    loss head on 128 positions (fwd+bwd).  The per-sample step time is extrapolated by algorithmic FLOPs."""
    from oracle import cambrian_oracle as O
    # threads = cores this process may actually run on (cgroup / affinity), not os.cpu_count(): oversubscribing a
    # CPU-limited container makes the fp32 GEMMs an order of magnitude slower
    try:
        usable = len(os.sched_getaffinity(0))
    except AttributeError:
        usable = os.cpu_count() or 1
    try:  # cgroup v2 CPU quota
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            usable = max(1, min(usable, int(int(q) / int(per) + 0.5)))
    except Exception:
        pass
    torch.set_num_threads(threads or usable)
    _w = torch.randn(1024, 1024)
    for _ in range(3):  # spin up the intra-op thread pool before timing
        _w = _w @ _w.t() * 1e-3
    g = torch.Generator().manual_seed(0)
    rn = lambda *s, sc=0.02: (torch.randn(*s, generator=g) * sc)
    H, I, nh, nkv, S, V = 4096, 14336, 32, 8, 2048, 128256
    sd = {}
    p = "model.layers.0."
    sd[p + "input_layernorm.weight"] = torch.ones(H)
    sd[p + "post_attention_layernorm.weight"] = torch.ones(H)
    sd[p + "self_attn.q_proj.weight"] = rn(H, H)
    sd[p + "self_attn.k_proj.weight"] = rn(nkv * 128, H)
    sd[p + "self_attn.v_proj.weight"] = rn(nkv * 128, H)
    sd[p + "self_attn.o_proj.weight"] = rn(H, H)
    sd[p + "mlp.gate_proj.weight"] = rn(I, H)
    sd[p + "mlp.up_proj.weight"] = rn(I, H)
    sd[p + "mlp.down_proj.weight"] = rn(H, I)

    def sva_params(prefix, D, depth):
        for l in range(depth):
            q = f"{prefix}layers.{l}."
            sd[q + "proj_context.weight"] = rn(1024, 1024)
            sd[q + "proj_in.weight"] = rn(1024, D + 1024)
            sd[q + "proj_out.linear_1.weight"] = rn(1024, 1024)
            sd[q + "proj_out.linear_2.weight"] = rn(D, 1024)
            sd[q + "norm.weight"], sd[q + "norm.bias"] = torch.ones(1024), torch.zeros(1024)
            for nm in ["q_proj"] + [f"{k}_proj_{i}" for i in range(4) for k in "kv"]:
                sd[q + f"cross_attn.{nm}.0.weight"], sd[q + f"cross_attn.{nm}.0.bias"] = torch.ones(1024), torch.zeros(1024)
                sd[q + f"cross_attn.{nm}.1.weight"] = rn(1024, 1024)
            sd[q + "cross_attn.o_proj.weight"] = rn(1024, 1024)

    sva_params("conn.", 1024, 3)
    sva_params("inllm.", H, 1)
    sd["lm_head.weight"] = rn(V, H)
    for v in sd.values():
        v.requires_grad_()
    cfg = dict(hidden_size=H, num_attention_heads=nh, num_key_value_heads=nkv, rms_norm_eps=1e-5)
    x = rn(1, S, H, sc=1.0).requires_grad_()
    feats = [rn(576, 1, 1024, sc=1.0) for _ in range(4)]
    ctx = rn(576, 1, 1024, sc=1.0)
    q0 = rn(576, 1, 1024, sc=1.0).requires_grad_()
    qh = rn(576, 1, H, sc=1.0).requires_grad_()
    cos, sin = O.rope_cos_sin(torch.arange(S)[None], 128, 5e5)
    labels = torch.randint(0, V, (1, 129), generator=g)
    t0 = time.perf_counter()
    a = O.sva_sampler(sd, "conn.", q0, ctx, feats, None, 3)
    b = O.llama_layer(sd, "model.layers.0.", x, cos, sin, None, cfg)
    c = O.sva_sampler(sd, "inllm.", qh, ctx, feats, None, 1)
    _, loss = O.lm_loss(sd, b[:, :129], labels)
    (a.float().pow(2).mean() + c.float().pow(2).mean() + loss).backward()
    dt = time.perf_counter() - t0
    # algorithmic TFLOP of the sample (fwd x3): connector 0.054, decoder layer (31.84-2.15 lm_head)/32, in-LLM SVA 0.0254,
    # lm_head on 129 of 2048 positions
    sample_tf = 3 * (0.054 + (31.84 - 2.15) / 32 + 0.0254 + 2.15 * 129 / 2048)
    est_step_s = dt * STEP_TF / sample_tf
    return dict(sample_seconds=dt, sample_tflop=sample_tf, est_seconds_per_sample=est_step_s, value=1.0 / est_step_s,
                cores=torch.get_num_threads(),
                sample=("oracle fp32 on host: SVA connector (3 layers, 4x576x1024 grids) + 1 Llama-3-8B decoder layer @S=2048 "
                        "+ 1 in-LLM SVA layer + loss head on 129 positions, fwd+bwd, B=1; extrapolated to the full "
                        f"{STEP_TF:.1f} TFLOP step by algorithmic FLOPs"))


def run_reference(args, rank, world):
    if rank != 0:
        return
    # one host CPU regardless of N: the CPU arm processes a single sample stream; each "step" is one bounded sample
    vals = [cpu_reference_sample() for _ in range(max(1, min(args.steps, 3)))]
    v = statistics.median(x["value"] for x in vals)
    r = vals[0]
    line = dict(metric=METRIC, value=v, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1000.0 / v, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="fp32",
                data="synthetic", impl="reference",
                config={"workload": "Cambrian-1-8B train step (4 towers, SVA, 576 vis-tok, seq 2048) — CPU port, bounded sample"},
                cpu_baseline=dict(value=v, unit=UNIT, cores=r["cores"], kind="port", sample=r["sample"]),
                e2e=dict(value=v, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--micro-batch", type=int, default=int(os.environ.get("CB_MICRO_BATCH", "4")))
    ap.add_argument("--seq", type=int, default=2048)
    ap.add_argument("--recompute", type=int, default=int(os.environ.get("CB_RECOMPUTE", "0")))
    ap.add_argument("--small", action="store_true", help="tiny shapes for a functional check (not a valid bench)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from cambrian_b200 import _lib, ops
    from cambrian_b200.engine import TrainEngine
    from cambrian_b200.model.language_model.cambrian_llama import CambrianLlamaForCausalLM
    lib = _lib.load()

    cfg = cambrian_8b_config(args)
    res = TOWER_RES if not args.small else [384, 336, 336, 256]
    torch.manual_seed(1234 + rank)
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.bfloat16)
    with torch.device(dev):
        model = CambrianLlamaForCausalLM(cfg)
        for t in model.get_model().vision_tower_aux_list:
            t.load_model()
    torch.set_default_dtype(prev)
    model.train()
    model.get_model().gradient_checkpointing = bool(args.recompute)
    engine = TrainEngine(model, lr=4e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0)
    engine.defer_param_sync = os.environ.get("CB_DEFER_PARAM_SYNC", "1") != "0"   # towers overlap the optimizer tail
    n_train = sum(p.numel() for p in engine.params)
    n_tower = sum(p.numel() for t in model.get_model().vision_tower_aux_list for p in t.parameters())

    B, S = args.micro_batch, args.seq
    host_batches = [make_host_batch(cfg, B, S, 1000 * rank + i, res) for i in range(2)]
    dev_batch, h2d_bytes = to_device(host_batches[0][0], dev)
    n_valid, label_ranges = host_batches[0][1]
    img_pos = [cfg.image_position] * B  # known to the collator (train_fsdp.py:1089-1165); avoids a D2H scan per step
    torch.cuda.synchronize()

    def step_resident():
        engine.zero_grad()
        out = model(**dev_batch, num_valid_labels=n_valid, image_positions=img_pos, label_ranges=label_ranges)
        out.loss.backward()
        engine.step()
        return out.loss

    def step_e2e(i):
        hb, (nv, lr) = host_batches[i % 2]
        db, _ = to_device(hb, dev)
        engine.zero_grad()
        out = model(**db, num_valid_labels=nv, image_positions=img_pos, label_ranges=lr)
        out.loss.backward()
        engine.step()
        return float(out.loss.item())  # device -> host read of the step's result

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        loss = step_resident()
    torch.cuda.synchronize()
    if not torch.isfinite(loss):
        raise RuntimeError(f"non-finite loss after warm-up: {loss}")

    timer = KernelTimer()
    timer.wrap(ops)
    sampler = ClockSampler(local_rank)
    sampler.start()
    # ---- timed region 1: inputs resident in HBM
    barrier()
    timer.enabled = True
    l0 = lib.cb_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t_host0 = time.perf_counter()
    for _ in range(args.steps):
        loss = step_resident()
    host_enqueue_ms = (time.perf_counter() - t_host0) * 1000.0 / args.steps  # CPU time to enqueue one step (no sync)
    e1.record()
    barrier()
    timer.enabled = False
    launches = (lib.cb_launch_count() - l0) // args.steps
    ms = e0.elapsed_time(e1)
    # ---- timed region 2: end to end through the public API with host buffers
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for i in range(args.steps):
        lv = step_e2e(i)
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    sampler.stop_flag = True
    sampler.join(timeout=3)
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_step = ms / args.steps
    value = world * B / (ms_step / 1000.0)
    e2e_value = world * B / (ms_e2e / args.steps / 1000.0)
    hbm_peak, tf_sustained, tf_burst, peak_src = peaks()
    agg = timer.totals()
    if os.environ.get("CB_BENCH_SHAPES"):
        print("\n".join(timer.shape_table()), file=sys.stderr)
    roof = None
    if "gemm" in agg:
        fl, tms, n = agg["gemm"]
        ach = fl / (tms / 1000.0) / 1e12
        roof = dict(bound="tensor", kernel="gemm_bf16_tcgen05", achieved=ach, peak=tf_sustained, unit="TFLOP/s",
                    frac=ach / tf_sustained, traffic=None, launches_timed=n, share_of_step=tms / ms,
                    peak_source=f"{peak_src} bf16_tflops_sustained (kernel timed inside a long step)")
    # ncu --set full DRAM traffic of the dominant GEMM launch — the decoder's fused gate/up projection + SwiGLU
    # (cb_gemm_swiglu_bf16, M=8192 F=14336 K=4096): profiles/r01_kernels_v2_ncu_summary.txt; algorithmic bytes of that
    # launch = (M*K + 2F*K + M*2F + M*F) * 2 = 1.007 GB
    if roof is not None:
        try:
            dom = timer.dominant()
            if dom is not None:
                dom["frac"] = dom["achieved"] / tf_sustained
                roof["dominant_launch"] = dom
        except Exception as exc:  # never let optional reporting break the bench line
            roof["dominant_launch"] = {"error": str(exc)}
        roof["traffic"] = 1.546132e9 + 0.683987e9
        roof["traffic_note"] = ("dram read+write bytes of ONE launch of the dominant kernel: fused gate/up + SwiGLU GEMM "
                                "M=8192 F=14336 K=4096 (algorithmic 1.007e9 B; tensor pipe 97.7 % active in the same "
                                "capture); ncu summary under profiles/")
    roof_sva = None
    if not args.small:
        # second headline metric (BASELINE.json "SVA HBM GB/s"): the fused window-attention kernel alone, on inputs larger
        # than L2 (batch 32: 377 MB for the BASELINE grids, 1.51 GB for the release grids), CUDA events, 20 launches
        roof_sva = {}
        for tag, rs in (("baseline_grids_576x4", [1, 1, 1, 1]), ("release_grids_576x3_9216", [1, 1, 1, 4])):
            Bq, qs = 32, 24
            qq = torch.randn(Bq * qs * qs, 1024, device=dev).bfloat16()
            ks = [torch.randn(Bq, (r * qs) ** 2, 1024, device=dev).bfloat16() for r in rs]
            vs = [torch.randn(Bq, (r * qs) ** 2, 1024, device=dev).bfloat16() for r in rs]
            for _ in range(3):
                ops.sva_window_attn_fwd(qq, ks, vs, None, rs, Bq, qs)
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for _ in range(20):
                ops.sva_window_attn_fwd(qq, ks, vs, None, rs, Bq, qs)
            s1.record()
            torch.cuda.synchronize()
            byts = 2.0 * (2 * sum(k.numel() for k in ks) + 2 * qq.numel())
            ach = byts / (s0.elapsed_time(s1) / 20 / 1000.0) / 1e9
            roof_sva[tag] = dict(achieved=ach, frac=ach / hbm_peak, algorithmic_bytes_per_launch=byts,
                                 mb_per_sample_per_layer=byts / Bq / 1e6)
            del qq, ks, vs
        roof_sva.update(bound="hbm", kernel="sva_window_attn_fwd", peak=hbm_peak, unit="GB/s",
                        achieved=roof_sva["baseline_grids_576x4"]["achieved"],
                        frac=roof_sva["baseline_grids_576x4"]["frac"],
                        traffic=0.339770e9 + 0.008153e9, peak_source=f"{peak_src} hbm_gbs",
                        traffic_note="ncu dram bytes of one batch-32 launch on the BASELINE grids (algorithmic 0.377e9 B; the "
                                     "output write-back had not left L2 when the counter was read)")
        if "sva_fwd" in agg:
            by, tms, n = agg["sva_fwd"]
            roof_sva["in_step_achieved"] = by / (tms / 1000.0) / 1e9  # B=4 inputs (38 MB) are L2-resident in the step
    # executed FLOPs per sample: the fused loss skips the vocabulary GEMMs (3 x 2*H*V per row) of rows whose shifted
    # label is ignore_index — identical loss / gradients, fewer FLOPs than BASELINE.md's 104.3 TFLOP accounting
    rows_done = sum(b - a for a, b in label_ranges)
    skipped_tf = 3 * 2.0 * cfg.hidden_size * cfg.vocab_size * (B * S - rows_done) / B / 1e12
    model_tf = value * (STEP_TF - skipped_tf) if not args.small else None
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(3, args.warmup),
                ms_per_step=ms_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16",
                data="synthetic", per_gpu=value / world,
                config={"workload": "Cambrian-1-8B train step: 4 towers (SigLIP-SO400M@384, CLIP-L@336, DINOv2-L@336, "
                                    "ConvNeXt-XXL@1024) -> SVA (3 + 10 layers) -> Llama-3-8B, 576 vis-tok, seq 2048"
                        if not args.small else "SMALL functional check (not the benchmark config)",
                        "micro_batch_per_gpu": B, "global_batch": B * world, "seq_len": S,
                        "parallelism": f"dp{world}", "activation_recompute": bool(args.recompute),
                        "optimizer": "AdamW fp32 master + bf16 grads, fused", "trainable_params": n_train,
                        "frozen_tower_params": n_tower,
                        "lm_head_rows": f"{rows_done} of {B * S} (rows with an ignored label skip the vocabulary GEMMs; "
                                        f"loss and gradients identical; MFU counts executed FLOPs only)", "l2_policy": "inputs larger than L2 (16 GB weights streamed per pass)"},
                gpu_launches=int(launches), host_enqueue_ms_per_step=round(host_enqueue_ms, 1), loss=float(loss),
                peak_mem_gb=round(torch.cuda.max_memory_allocated() / 2 ** 30, 1),
                peak_reserved_gb=round(torch.cuda.max_memory_reserved() / 2 ** 30, 1),
                e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=int(h2d_bytes), d2h_bytes_per_step=4),
                clocks=sampler.summary(), roofline=roof, roofline_sva=roof_sva,
                model_tflops_per_gpu=(model_tf / world) if model_tf else None,
                mfu_vs_sustained=(model_tf / world / tf_sustained) if model_tf else None)
    if world == 1 and not args.no_cpu_baseline:
        try:
            r = cpu_reference_sample()
            line["cpu_baseline"] = dict(value=r["value"], unit=UNIT, cores=r["cores"], kind="port", sample=r["sample"])
        except Exception as ex:  # the baseline leg must never take the GPU number down with it
            line["cpu_baseline"] = dict(value=None, unit=UNIT, cores=os.cpu_count(), kind="port", sample=f"failed: {ex}")
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
