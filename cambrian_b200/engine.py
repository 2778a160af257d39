"""Data-parallel training engine for the Cambrian hot path.

One process per GPU.  All trainable parameters live in ONE flat bf16 buffer (the compute copy) with matching flat
buffers for bf16 gradients, fp32 master weights and fp32 Adam moments:

  * weight-gradient GEMMs accumulate straight into `param.main_grad` (a view of the flat gradient buffer), so
    autograd never materialises or sums parameter gradients;
  * gradient reduction = NCCL all-reduce over contiguous buckets of the flat gradient buffer.  Each parameter's number
    of gradient contributions per step is learned on the first step; from then on a bucket's all-reduce is launched
    (asynchronously, on NCCL's stream) the moment its last contribution has been enqueued, so NVLink traffic overlaps
    the remaining backward GEMMs.  Frozen towers are never reduced;
  * the optimizer is one fused AdamW kernel over the flat buffers (fp32 master update -> bf16 compute copy), with the
    1/world gradient average folded in.

The reference does its gradient reduction inside torch_xla FSDP (`xm.all_reduce` helper at
cambrian_trainer.py:181-190) and steps HF Trainer's AdamW (cambrian_trainer.py:242-381).
"""
from __future__ import annotations

import functools

import torch
import torch.distributed as dist

from . import ops


def _round8(n: int) -> int:
    return (n + 7) // 8 * 8


class TrainEngine:
    def __init__(self, model: torch.nn.Module, lr: float = 4e-5, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, bucket_mb: float = 256.0, process_group=None, overlap: bool = True,
                 zero_stage: int = 0):
        """zero_stage = 0: replicated optimizer state (DDP, BASELINE config 3).
        zero_stage = 2: optimizer state (fp32 master + Adam moments, 12 of the 16 bytes/param) sharded across ranks
        (BASELINE config 4, scripts/zero2.json): gradients are reduce-scattered, each rank updates its 1/world slice of
        the flat buffer, the updated bf16 parameters are all-gathered."""
        self.model = model
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.overlap = overlap and zero_stage == 0
        self.zero_stage = zero_stage
        if zero_stage not in (0, 2):
            raise ValueError("zero_stage must be 0 or 2")
        self.rank = dist.get_rank(process_group) if self.world > 1 else 0
        self.step_count = 0
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
        if not named:
            raise ValueError("TrainEngine: model has no trainable parameters")
        dev = named[0][1].device
        self.params = [p for _, p in named]
        self.names = [n for n, _ in named]
        offs, total = [], 0
        for p in self.params:
            if p.dtype != torch.bfloat16:
                raise ValueError("TrainEngine expects bf16 parameters (fp32 masters are kept by the engine)")
            offs.append(total)
            total += _round8(p.numel())
        self.offsets = offs
        self.shard = 0
        if zero_stage == 2:
            self.shard = _round8((total + self.world - 1) // self.world)
            total = self.shard * self.world          # pad so every rank owns an equal, 16-byte aligned slice
        self.total = total
        self.flat_p = torch.zeros(total, dtype=torch.bfloat16, device=dev)
        self.flat_g = torch.zeros(total, dtype=torch.bfloat16, device=dev)
        for i, (p, o) in enumerate(zip(self.params, offs)):
            n = p.numel()
            self.flat_p[o:o + n].copy_(p.data.reshape(-1))
            p.data = self.flat_p[o:o + n].view(p.shape)          # re-point: compute copy now lives in the flat buffer
            p.main_grad = self.flat_g[o:o + n].view(p.shape)
            p._cb_fresh = set()
            p._cb_notify = functools.partial(self._on_write, i)
            p.grad = None
        if zero_stage == 2:
            lo = self.rank * self.shard
            self.master = self.flat_p[lo:lo + self.shard].float()
            self.exp_avg = torch.zeros(self.shard, dtype=torch.float32, device=dev)
            self.exp_avg_sq = torch.zeros(self.shard, dtype=torch.float32, device=dev)
            self.shard_g = torch.zeros(self.shard, dtype=torch.bfloat16, device=dev)
        else:
            self.master = self.flat_p.float()
            self.exp_avg = torch.zeros(total, dtype=torch.float32, device=dev)
            self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev)
        # buckets: contiguous parameter ranges of ~bucket_mb
        self.buckets = []  # (start_elem, end_elem, [param indices])
        cur, cur_start = [], 0
        limit = int(bucket_mb * 1024 * 1024 / 2)
        for i, (p, o) in enumerate(zip(self.params, offs)):
            cur.append(i)
            end = o + _round8(p.numel())
            if end - cur_start >= limit:
                self.buckets.append((cur_start, end, cur))
                cur, cur_start = [], end
        if cur:
            self.buckets.append((cur_start, total, cur))
        self._bucket_of = {}
        for b, (_, _, idx) in enumerate(self.buckets):
            for i in idx:
                self._bucket_of[i] = b
        # contribution accounting for the backward/all-reduce overlap
        self._expected = None                      # writes per parameter per step, learned on the first step
        self._writes = [0] * len(self.params)
        self._remaining = None                     # per bucket: parameters not yet final
        self._launched = [False] * len(self.buckets)
        self._updated = [False] * len(self.buckets)
        self._handles = {}
        self._opt_stream = None
        self._params_pending = False
        self._towers_trainable = any("vision_tower" in n for n in self.names)
        # opt-in (bench / training loops): skip the end-of-step wait on the optimizer stream and let the model wait just
        # before its first trainable module; anyone reading parameters right after step() must call wait_for_params()
        self.defer_param_sync = False
        if hasattr(model, "prepare_inputs_labels_for_multimodal"):
            model._cb_param_sync = self.wait_for_params

    # ---- per-step protocol ---------------------------------------------------------------------------------------
    def zero_grad(self):
        for p in self.params:
            p._cb_fresh.clear()
            p.grad = None
        self._writes = [0] * len(self.params)
        self._launched = [False] * len(self.buckets)
        self._updated = [False] * len(self.buckets)
        self._handles = {}
        if self._expected is not None:
            self._remaining = [sum(1 for i in idx if self._expected[i] > 0) for (_, _, idx) in self.buckets]

    def _launch_bucket(self, b):
        s, e, _ = self.buckets[b]
        self._launched[b] = True
        self._handles[b] = dist.all_reduce(self.flat_g[s:e], op=dist.ReduceOp.SUM, group=self.pg, async_op=True)

    def _update_bucket(self, b, side_stream: bool):
        """Fused AdamW on bucket b's slice of the flat buffers.  With side_stream=True it runs on the optimizer stream,
        ordered after the gradients (main-stream event, or the bucket's all-reduce), so the HBM-bound update overlaps the
        tensor-core-bound remainder of the backward pass: its 28 B/param of traffic would otherwise be a serial tail."""
        s, e, _ = self.buckets[b]
        self._updated[b] = True
        if side_stream and self.master.is_cuda:
            if self._opt_stream is None:
                self._opt_stream = torch.cuda.Stream(device=self.master.device)
            if b in self._handles:
                with torch.cuda.stream(self._opt_stream):
                    self._handles.pop(b).wait()          # optimizer stream waits for NCCL
            else:
                self._opt_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._opt_stream):
                self._adamw(self.master[s:e], self.exp_avg[s:e], self.exp_avg_sq[s:e], self.flat_g[s:e], self.flat_p[s:e],
                            self.step_count + 1)
        else:
            if b in self._handles:
                self._handles.pop(b).wait()
            self._adamw(self.master[s:e], self.exp_avg[s:e], self.exp_avg_sq[s:e], self.flat_g[s:e], self.flat_p[s:e],
                        self.step_count + 1)

    def _on_write(self, i):
        """Called (host side, in stream order) right after a gradient contribution of parameter i was enqueued."""
        self._writes[i] += 1
        if not self.overlap or self._expected is None:
            return
        if self._writes[i] == self._expected[i]:
            b = self._bucket_of[i]
            self._remaining[b] -= 1
            if self._remaining[b] == 0 and not self._launched[b]:
                # a bucket whose parameters also receive plain-autograd gradients is only final at step()
                if all(self._expected[j] > 0 for j in self.buckets[b][2]):
                    if self.world > 1:
                        self._launch_bucket(b)
                    else:
                        self._launched[b] = True
                    self._update_bucket(b, side_stream=True)

    def _finalize_unwritten(self):
        """Fold in gradients that reached a parameter through plain autograd (a parameter used by an ordinary torch
        view/op, e.g. the `vision_query[g:g+1]` slice), then zero parameters that received no gradient at all this step
        (unused modules) so they do not feed stale values to Adam."""
        for i, p in enumerate(self.params):
            if self._updated[self._bucket_of[i]]:
                continue
            if p.grad is not None:
                if p._cb_fresh:
                    p.main_grad.add_(p.grad.to(p.main_grad.dtype))
                else:
                    p.main_grad.copy_(p.grad)
                    p._cb_fresh.add("all")
                p.grad = None
            if not p._cb_fresh:
                p.main_grad.zero_()

    def reduce_gradients(self):
        """All-reduce (sum) every bucket that was not already launched during backward and wait for them (buckets already
        consumed by an overlapped optimizer update are skipped).  The 1/world average is folded into AdamW."""
        if self.world == 1:
            return
        for b in reversed(range(len(self.buckets))):
            if not self._launched[b]:
                self._launch_bucket(b)
        for b in list(self._handles):
            if not self._updated[b]:
                self._handles.pop(b).wait()

    def _zero2_step(self):
        lo = self.rank * self.shard
        if self.world > 1:
            if dist.get_backend(self.pg) == "nccl":
                dist.reduce_scatter_tensor(self.shard_g, self.flat_g, op=dist.ReduceOp.SUM, group=self.pg)
                g = self.shard_g
            else:  # gloo (CPU tests) has no reduce_scatter_tensor: all-reduce, then keep the local slice
                dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM, group=self.pg)
                g = self.flat_g[lo:lo + self.shard]
        else:
            g = self.flat_g
        p16 = self.flat_p[lo:lo + self.shard]
        self._adamw(self.master, self.exp_avg, self.exp_avg_sq, g, p16)
        if self.world > 1:
            dist.all_gather_into_tensor(self.flat_p, p16.clone(), group=self.pg)

    def _adamw(self, master, m, v, g, p16, step=None):
        ops.adamw(master, m, v, g, p16, self.lr, self.betas[0], self.betas[1], self.eps, self.wd,
                  self.step_count if step is None else step, grad_scale=1.0 / self.world)

    def step(self):
        self._finalize_unwritten()
        if self._expected is None:
            self._expected = list(self._writes)
        self.step_count += 1
        if self.zero_stage == 2:
            self._zero2_step()
            return
        self.step_count -= 1                       # _update_bucket stamps step_count + 1
        self.reduce_gradients()
        for b in range(len(self.buckets)):
            if not self._updated[b]:
                self._update_bucket(b, side_stream=False)
        self.step_count += 1
        if self._opt_stream is not None:
            if self.defer_param_sync and hasattr(self.model, "_cb_param_sync") and not self._towers_trainable:
                # the frozen towers of the next forward read no trainable parameter: let them overlap the tail of the
                # optimizer (last buckets: embeddings, connector); the model calls wait_for_params() before the first
                # trainable module runs (cambrian_arch.prepare_inputs_labels_for_multimodal / forward without images)
                self._params_pending = True
            else:
                torch.cuda.current_stream().wait_stream(self._opt_stream)   # next forward sees every updated parameter

    def wait_for_params(self):
        """Make the current stream wait for optimizer updates still running on the side stream (no-op otherwise)."""
        if self._params_pending:
            self._params_pending = False
            torch.cuda.current_stream().wait_stream(self._opt_stream)

    # ---- convenience ---------------------------------------------------------------------------------------------
    def train_step(self, **batch):
        self.zero_grad()
        out = self.model(**batch)
        out.loss.backward()
        self.step()
        return out.loss

    def state_bytes(self):
        opt = (self.shard if self.zero_stage == 2 else self.total) * 12
        return self.total * 4 + opt
