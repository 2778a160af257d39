"""ConvNeXt-XXL stage-3 depthwise 7x7 conv (B=4, 64x64, C=1536) for an `ncu --set full` capture."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cambrian_b200 import ops  # noqa: E402

x = torch.randn(4, 64, 64, 1536, device="cuda").bfloat16()
w = torch.randn(7, 7, 1536, device="cuda").bfloat16()
b = torch.randn(1536, device="cuda").bfloat16()
for _ in range(3):
    ops.dwconv7(x, w, b)
torch.cuda.synchronize()
