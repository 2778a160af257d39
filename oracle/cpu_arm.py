"""TEST / BASELINE INFRASTRUCTURE ONLY — the reference algorithm timed on the host cores (bench.py `--impl reference` and
the `cpu_baseline` leg).  Never imported by the product.

What is timed.  A full Cambrian training step on a CPU takes minutes (104 TFLOP per sample for the 8B config), so one
BOUNDED step times every DISTINCT block of the step exactly once at its real shape (fp32, B = 1, forward + backward for the
trainable part, forward only for the frozen towers) and the step time is COMPOSED as  sum_i count_i x t_i  with the counts
the model structure dictates (32 decoder layers, 10 in-LLM SVA layers, 23 CLIP blocks, ...).  This replaces round 1's
extrapolation by FLOPs: every block type contributes its own measured seconds (ConvNeXt's depthwise convolutions or the
q_len=1 SVA attention are far from GEMM speed on a CPU), nothing is scaled by arithmetic intensity.  The row count of the
loss head is the one linear scaling (R of S rows are timed).

Whose code runs.  SVA blocks run the REFERENCE'S OWN `VisionTokenSampler` (oracle/_ref/vision_sampler.py, copied from
/root/reference by oracle/make_ref.py in the build container; pure torch) when that file is present — `kind: "reference"`
for those blocks — and the pinned oracle port (oracle/cambrian_oracle.py) otherwise.  The decoder layer, the loss head and
the towers are the oracle port of what the reference delegates to transformers / timm (pinned in tests/test_oracle_pin.py).
"""
from __future__ import annotations

import importlib.util
import os
import statistics
import time

import torch

from . import cambrian_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))


def usable_cores() -> int:
    """Cores this process may actually run on (affinity / cgroup quota), not os.cpu_count(): oversubscribing a CPU-limited
    container makes fp32 GEMMs an order of magnitude slower."""
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
    return usable


def reference_sampler_module():
    """The reference's vision_sampler.py from oracle/_ref (None when absent, e.g. the recipe never ran)."""
    path = os.path.join(HERE, "_ref", "vision_sampler.py")
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location("cambrian_ref_vision_sampler", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _t(fn):
    t0 = time.perf_counter()
    fn()
    return time.perf_counter() - t0


class _Blocks:
    """Builds the weights / inputs of every distinct block once; `sample()` times each block once."""

    def __init__(self, name, llm, C):
        g = torch.Generator().manual_seed(0)
        self.rn = lambda *s, sc=0.02: (torch.randn(*s, generator=g) * sc)
        self.name, self.llm, self.C = name, llm, C
        self.S = C["seq"]
        self.ref_vs = reference_sampler_module() if C["sva"] else None
        self.blocks = []        # (label, count, callable, kind)
        self._decoder()
        self._loss_head()
        if C["sva"]:
            self._sva()
        else:
            self._mlp_projector()
        self._towers()

    # ---- trainable part (forward + backward) ------------------------------------------------------------------------
    def _decoder(self):
        rn, L = self.rn, self.llm
        H, I, nh, nkv = L["hidden_size"], L["intermediate_size"], L["num_attention_heads"], L["num_key_value_heads"]
        hd = H // nh
        p = "model.layers.0."
        sd = {p + "input_layernorm.weight": torch.ones(H), p + "post_attention_layernorm.weight": torch.ones(H),
              p + "self_attn.q_proj.weight": rn(nh * hd, H), p + "self_attn.k_proj.weight": rn(nkv * hd, H),
              p + "self_attn.v_proj.weight": rn(nkv * hd, H), p + "self_attn.o_proj.weight": rn(H, nh * hd),
              p + "mlp.gate_proj.weight": rn(I, H), p + "mlp.up_proj.weight": rn(I, H), p + "mlp.down_proj.weight": rn(H, I)}
        for v in sd.values():
            v.requires_grad_()
        x = rn(1, self.S, H, sc=1.0).requires_grad_()
        cos, sin = O.rope_cos_sin(torch.arange(self.S)[None], hd, L["rope_theta"])
        cfg = dict(hidden_size=H, num_attention_heads=nh, num_key_value_heads=nkv, rms_norm_eps=L["rms_norm_eps"])

        def run():
            out = O.llama_layer(sd, p, x, cos, sin, None, cfg)
            out.float().pow(2).mean().backward()
        self.blocks.append((f"decoder layer fwd+bwd (S={self.S})", L["num_hidden_layers"], run, "port"))

    def _loss_head(self):
        rn, L = self.rn, self.llm
        H, V, R = L["hidden_size"], L["vocab_size"], 256
        sd = {"lm_head.weight": rn(V, H).requires_grad_()}
        h = rn(1, R + 1, H, sc=1.0).requires_grad_()
        labels = torch.randint(0, V, (1, R + 1))

        def run():
            _, loss = O.lm_loss(sd, h, labels)
            loss.backward()
        self.blocks.append((f"lm_head + shifted CE fwd+bwd on {R} of {self.S} rows", self.S / R, run, "port"))

    def _sva_state(self, prefix, D, depth, T):
        rn, sd = self.rn, {}
        for l in range(depth):
            q = f"{prefix}layers.{l}."
            sd[q + "proj_context.weight"] = rn(1024, 1024)
            sd[q + "proj_in.weight"] = rn(1024, D + 1024)
            sd[q + "proj_out.linear_1.weight"] = rn(1024, 1024)
            sd[q + "proj_out.linear_2.weight"] = rn(D, 1024)
            sd[q + "norm.weight"], sd[q + "norm.bias"] = torch.ones(1024), torch.zeros(1024)
            for nm in ["q_proj"] + [f"{k}_proj_{i}" for i in range(T) for k in "kv"]:
                sd[q + f"cross_attn.{nm}.0.weight"], sd[q + f"cross_attn.{nm}.0.bias"] = torch.ones(1024), torch.zeros(1024)
                sd[q + f"cross_attn.{nm}.1.weight"] = rn(1024, 1024)
            sd[q + "cross_attn.o_proj.weight"] = rn(1024, 1024)
        return sd

    def _sampler_runner(self, D, depth, T, queries, ctx, feats):
        """(callable, kind): the reference's VisionTokenSampler when oracle/_ref holds it, else the oracle port."""
        sd = self._sva_state("", D, depth, T)
        masks = [torch.ones(576, 1, dtype=torch.bool) for _ in range(T)]
        if self.ref_vs is not None:
            m = self.ref_vs.VisionTokenSampler(D, 1024, [1024] * T, [1] * T, 1024, depth)
            m.load_state_dict(sd)

            def run_ref():
                m(queries, ctx, *feats, *masks).float().pow(2).mean().backward()
            return run_ref, "reference"
        for v in sd.values():
            v.requires_grad_()

        def run_port():
            O.sva_sampler(sd, "", queries, ctx, feats, masks, depth).float().pow(2).mean().backward()
        return run_port, "port"

    def _sva(self):
        rn, L, C = self.rn, self.llm, self.C
        H, T = L["hidden_size"], len(C["towers"])
        feats = [rn(576, 1, 1024, sc=1.0).requires_grad_() for _ in range(T)]
        ctx = rn(576, 1, 1024, sc=1.0)
        q0 = rn(576, 1, 1024, sc=1.0).requires_grad_()
        qh = rn(576, 1, H, sc=1.0).requires_grad_()
        run_c, kind_c = self._sampler_runner(1024, 3, T, q0, ctx, feats)
        # aux projectors (Linear-GELU-Linear-LN per tower) + mm_projector (cambrian_arch.py:372-379, :410-411)
        dims = {"siglip": 1152, "clip": 1024, "dinov2": 1024, "convnext": 3072}
        psd = {}
        tower_dims = [next(v for k, v in dims.items() if k in t.lower()) for t in C["towers"]]
        for i, c in enumerate(tower_dims):
            p = f"aux{i}."
            psd.update({p + "0.weight": rn(1024, c), p + "0.bias": torch.zeros(1024), p + "2.weight": rn(1024, 1024),
                        p + "2.bias": torch.zeros(1024), p + "3.weight": torch.ones(1024), p + "3.bias": torch.zeros(1024)})
        psd.update({"mm.0.weight": rn(H, 1024), "mm.0.bias": torch.zeros(H), "mm.2.weight": rn(H, H), "mm.2.bias": torch.zeros(H)})
        for v in psd.values():
            v.requires_grad_()
        tf = [rn(1, 576, c, sc=1.0) for c in tower_dims]

        def run_connector():
            aux = [O.mm_projector_aux(psd, f"aux{i}.", f) for i, f in enumerate(tf)]
            run_c()
            out = O.mlp2x_gelu(psd, "mm.", q0.detach().view(1, 576, 1024))
            (out.float().pow(2).mean() + sum(a.float().pow(2).mean() for a in aux)).backward()
        self.blocks.append(("aux projectors + SVA connector (3 layers, 4x576x1024 grids) + mm_projector fwd+bwd", 1,
                            run_connector, kind_c))
        run_l, kind_l = self._sampler_runner(H, 1, T, qh, ctx, feats)
        self.blocks.append((f"in-LLM SVA layer fwd+bwd (q_dim {H})", C["n_sva"], run_l, kind_l))

    def _mlp_projector(self):
        rn, H = self.rn, self.llm["hidden_size"]
        sd = {"mm.0.weight": rn(H, 1024), "mm.0.bias": torch.zeros(H), "mm.2.weight": rn(H, H), "mm.2.bias": torch.zeros(H)}
        for v in sd.values():
            v.requires_grad_()
        f = rn(1, 576, 1024, sc=1.0)
        self.blocks.append(("mlp2x_gelu projector fwd+bwd (576 tokens)", 1,
                            lambda: O.mlp2x_gelu(sd, "mm.", f).float().pow(2).mean().backward(), "port"))

    # ---- frozen towers (forward only): t(embed + n blocks) = t0 + n * (t1 - t0) -------------------------------------
    def _vit_sd(self, kind, D, I, n_pos, layers):
        rn, sd = self.rn, {}
        if kind == "clip":
            p = "vision_model."
            sd[p + "embeddings.patch_embedding.weight"] = rn(D, 3, 14, 14)
            sd[p + "embeddings.class_embedding"] = rn(D)
            sd[p + "embeddings.position_embedding.weight"] = rn(n_pos + 1, D)
            sd[p + "pre_layrnorm.weight"], sd[p + "pre_layrnorm.bias"] = torch.ones(D), torch.zeros(D)
            for i in range(layers):
                q = f"{p}encoder.layers.{i}."
                for nm in ("q_proj", "k_proj", "v_proj", "out_proj"):
                    sd[q + f"self_attn.{nm}.weight"], sd[q + f"self_attn.{nm}.bias"] = rn(D, D), torch.zeros(D)
                for nm in ("layer_norm1", "layer_norm2"):
                    sd[q + nm + ".weight"], sd[q + nm + ".bias"] = torch.ones(D), torch.zeros(D)
                sd[q + "mlp.fc1.weight"], sd[q + "mlp.fc1.bias"] = rn(I, D), torch.zeros(I)
                sd[q + "mlp.fc2.weight"], sd[q + "mlp.fc2.bias"] = rn(D, I), torch.zeros(D)
        elif kind == "siglip":
            sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"] = rn(D, 3, 14, 14), torch.zeros(D)
            sd["pos_embed"] = rn(1, n_pos, D)
            sd["norm.weight"], sd["norm.bias"] = torch.ones(D), torch.zeros(D)
            for i in range(layers):
                q = f"blocks.{i}."
                sd[q + "attn.qkv.weight"], sd[q + "attn.qkv.bias"] = rn(3 * D, D), torch.zeros(3 * D)
                sd[q + "attn.proj.weight"], sd[q + "attn.proj.bias"] = rn(D, D), torch.zeros(D)
                for nm in ("norm1", "norm2"):
                    sd[q + nm + ".weight"], sd[q + nm + ".bias"] = torch.ones(D), torch.zeros(D)
                sd[q + "mlp.fc1.weight"], sd[q + "mlp.fc1.bias"] = rn(I, D), torch.zeros(I)
                sd[q + "mlp.fc2.weight"], sd[q + "mlp.fc2.bias"] = rn(D, I), torch.zeros(D)
        else:  # dinov2
            sd["embeddings.patch_embeddings.projection.weight"] = rn(D, 3, 14, 14)
            sd["embeddings.patch_embeddings.projection.bias"] = torch.zeros(D)
            sd["embeddings.cls_token"] = rn(1, 1, D)
            sd["embeddings.position_embeddings"] = rn(1, n_pos + 1, D)
            sd["layernorm.weight"], sd["layernorm.bias"] = torch.ones(D), torch.zeros(D)
            for i in range(layers):
                q = f"encoder.layer.{i}."
                for nm in ("query", "key", "value"):
                    sd[q + f"attention.attention.{nm}.weight"], sd[q + f"attention.attention.{nm}.bias"] = rn(D, D), torch.zeros(D)
                sd[q + "attention.output.dense.weight"], sd[q + "attention.output.dense.bias"] = rn(D, D), torch.zeros(D)
                for nm in ("norm1", "norm2"):
                    sd[q + nm + ".weight"], sd[q + nm + ".bias"] = torch.ones(D), torch.zeros(D)
                sd[q + "layer_scale1.lambda1"], sd[q + "layer_scale2.lambda1"] = torch.ones(D), torch.ones(D)
                sd[q + "mlp.fc1.weight"], sd[q + "mlp.fc1.bias"] = rn(I, D), torch.zeros(I)
                sd[q + "mlp.fc2.weight"], sd[q + "mlp.fc2.bias"] = rn(D, I), torch.zeros(D)
        return sd

    def _towers(self):
        rn = self.rn
        for name, R in zip(self.C["towers"], self.C["res"]):
            low = name.lower()
            img = rn(1, 3, R, R, sc=1.0)
            n_pos = (R // 14) ** 2
            if "openai/clip" in low:
                sd = self._vit_sd("clip", 1024, 4096, n_pos, 1)
                f = lambda n, sd=sd, img=img: O.clip_vit(sd, dict(num_hidden_layers=n + 1, patch_size=14, num_attention_heads=16,
                                                                   select_layer=-2), img)
                self._tower_blocks(f"CLIP ViT-L/14@{R}", f, 23)      # hidden_states[-2]: 23 of 24 layers
            elif "siglip" in low:
                sd = self._vit_sd("siglip", 1152, 4304, n_pos, 1)
                f = lambda n, sd=sd, img=img: O.siglip_vit(sd, dict(num_hidden_layers=n, patch_size=14, num_attention_heads=16,
                                                                     interp=576), img)
                self._tower_blocks(f"SigLIP SO400M/14@{R}", f, 27)
            elif "dinov2" in low:
                sd = self._vit_sd("dino", 1024, 4096, n_pos, 1)
                f = lambda n, sd=sd, img=img: O.dinov2_vit(sd, dict(num_hidden_layers=n, patch_size=14, num_attention_heads=16,
                                                                     interp=576), img)
                self._tower_blocks(f"DINOv2 ViT-L/14@{R}", f, 24)
            else:
                self._convnext(R, img)

    def _tower_blocks(self, label, f, n_blocks):
        def run0():
            with torch.no_grad():
                f(0)

        def run1():
            with torch.no_grad():
                f(1)
        self.blocks.append((f"{label}: embed + resize (0 blocks) fwd", 1 - n_blocks, run0, "port"))   # t0 + n (t1 - t0)
        self.blocks.append((f"{label}: embed + 1 block fwd", n_blocks, run1, "port"))

    def _convnext(self, R, img):
        rn = self.rn
        depths, dims = (3, 4, 30, 3), (384, 768, 1536, 3072)
        sd = {"stem.0.weight": rn(dims[0], 3, 4, 4), "stem.0.bias": torch.zeros(dims[0]), "stem.1.weight": torch.ones(dims[0]),
              "stem.1.bias": torch.zeros(dims[0])}
        for s, c in enumerate(dims):
            p = f"stages.{s}."
            if s > 0:
                sd[p + "downsample.0.weight"], sd[p + "downsample.0.bias"] = torch.ones(dims[s - 1]), torch.zeros(dims[s - 1])
                sd[p + "downsample.1.weight"], sd[p + "downsample.1.bias"] = rn(c, dims[s - 1], 2, 2), torch.zeros(c)
            q = f"{p}blocks.0."
            sd[q + "conv_dw.weight"], sd[q + "conv_dw.bias"] = rn(c, 1, 7, 7), torch.zeros(c)
            sd[q + "norm.weight"], sd[q + "norm.bias"] = torch.ones(c), torch.zeros(c)
            sd[q + "mlp.fc1.weight"], sd[q + "mlp.fc1.bias"] = rn(4 * c, c), torch.zeros(4 * c)
            sd[q + "mlp.fc2.weight"], sd[q + "mlp.fc2.bias"] = rn(c, 4 * c), torch.zeros(c)
            sd[q + "gamma"] = torch.ones(c)

        def run(depth):
            with torch.no_grad():
                O.convnext_trunk(sd, dict(depths=depth, interp=576, multi_stage=False), img)
        self.blocks.append((f"ConvNeXt-XXL@{R}: stem + downsamples + resize (0 blocks) fwd", 1 - sum(depths),
                            lambda: run((0, 0, 0, 0)), "port"))
        for s, d in enumerate(depths):
            one = tuple(1 if k == s else 0 for k in range(4))
            self.blocks.append((f"ConvNeXt-XXL@{R}: stem + downsamples + 1 block of stage {s} (C={dims[s]}) fwd", d,
                                lambda one=one: run(one), "port"))

    def sample(self):
        return [_t(fn) for (_, _, fn, _) in self.blocks]


def run(name, llm, C, steps=3, warmup=1, budget_s=None, threads=None):
    """Returns dict(value [samples/s], ms_per_bounded_step, est_seconds_per_sample, fraction, components, cores, kind, sample,
    basis).  `budget_s`: if the warm-up sample alone exceeds it, that sample is the measurement (no further steps)."""
    torch.set_num_threads(threads or usable_cores())
    w = torch.randn(1024, 1024)
    for _ in range(3):  # spin up the intra-op thread pool before timing
        w = w @ w.t() * 1e-3
    blk = _Blocks(name, llm, C)
    samples, walls = [], []
    warm = []
    for i in range(max(1, warmup)):
        t0 = time.perf_counter()
        warm.append(blk.sample())
        wall = time.perf_counter() - t0
    used_warmup = budget_s is not None and wall > budget_s
    if used_warmup:
        samples, walls = [warm[-1]], [wall]
    else:
        for _ in range(steps):
            t0 = time.perf_counter()
            samples.append(blk.sample())
            walls.append(time.perf_counter() - t0)
    med = [statistics.median(s[i] for s in samples) for i in range(len(blk.blocks))]
    est = sum(c * t for (_, c, _, _), t in zip(blk.blocks, med))
    kinds = {k for (_, _, _, k) in blk.blocks}
    comps = [dict(block=lbl, count=round(c, 3), seconds=round(t, 4), kind=k) for (lbl, c, _, k), t in zip(blk.blocks, med)]
    kind = "reference" if kinds == {"reference"} else "port"
    ref_note = ("SVA blocks run the reference's own vision_sampler.py (oracle/_ref); "
                if "reference" in kinds else "oracle/_ref absent: every block is the oracle port; ")
    return dict(value=1.0 / est, ms_per_bounded_step=1000.0 * statistics.median(walls), est_seconds_per_sample=est,
                fraction=sum(med) / est, components=comps, cores=torch.get_num_threads(), kind=kind,
                basis="composed: 1 / sum(count x median measured block seconds); no FLOP extrapolation",
                sample=(f"fp32 on {torch.get_num_threads()} host threads, B=1; {len(samples)} bounded step(s)"
                        f"{' (the warm-up sample: over the time budget)' if used_warmup else f' after {max(1, warmup)} warm-up'}"
                        f": every distinct block of the step timed once at its real shape ({len(blk.blocks)} blocks, "
                        f"{sum(med):.1f} s per bounded step = {100 * sum(med) / est:.1f}% of a composed full step of {est:.0f} s); "
                        + ref_note + "decoder / loss head / towers are the pinned oracle port"))
