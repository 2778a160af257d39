"""ConvNeXt-XXL stage-2 fc1 GEMM (M=16384 N=6144 K=1536, bias + GELU epilogue) and the same GEMM without epilogue work,
for an `ncu --set full --import-source on` capture of the epilogue-bound case."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cambrian_b200 import ops  # noqa: E402

dev = "cuda"
a = torch.randn(16384, 1536, device=dev).bfloat16()
w = torch.randn(6144, 1536, device=dev).bfloat16() * 0.02
b = torch.randn(6144, device=dev).bfloat16()
out = torch.empty(16384, 6144, device=dev, dtype=torch.bfloat16)


def t(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for name, kw in [("plain", {}), ("bias", dict(bias=b)), ("gelu", dict(act="gelu")), ("bias+gelu", dict(bias=b, act="gelu")),
                 ("bias+quick_gelu", dict(bias=b, act="quick_gelu"))]:
    for bn in (512, 256):
        us = t(lambda: ops.gemm(a, w, out=out, force_bn=bn, **kw))
        print(f"{name:16s} bn={bn}: {us:8.1f} us  {2 * 16384 * 6144 * 1536 / us / 1e6:.0f} TFLOP/s", flush=True)
