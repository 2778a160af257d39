"""GPU image preprocessing — mirror of `process_images` (cambrian/mm_utils.py:186-201) for the multi-tower model.

The reference runs, per image and per tower, `expand2square` (pad to a square with the tower's mean colour), a PIL
bicubic resize to the tower's resolution and the tower's `preprocess` (scale to [0, 1], normalise) on the host — four PIL
resizes per image, the input-pipeline bottleneck SURVEY.md §8f rank 4 names.  Here the raw uint8 image is uploaded once
and each tower's tensor is produced by two CUDA kernels that reproduce Pillow's uint8 arithmetic bit for bit
(cambrian_b200/csrc/preprocess.cu)."""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def _to_u8_hwc(image, device):
    if torch.is_tensor(image):
        t = image
        if t.dtype != torch.uint8:
            raise ValueError("process_images: tensor images must be uint8 [H, W, 3]")
    else:
        if hasattr(image, "convert"):                       # PIL.Image
            image = image.convert("RGB")
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(image, dtype=np.uint8)))
    if t.dim() != 3 or t.shape[2] != 3:
        raise ValueError("process_images: expected RGB images [H, W, 3]")
    return t.to(device, non_blocking=True)


def process_images(images, image_processor, model_cfg=None, device="cuda"):
    """images: list of PIL images / uint8 HWC arrays; image_processor: the per-tower processor list
    (`[t.image_processor for t in towers]`, builder.py:165).  Returns one bf16 CUDA tensor [B, 3, R_i, R_i] per tower —
    the structure `process_images` returns (mm_utils.py:199-200; the reference's `.half()` is bf16 here)."""
    per_tower = [[] for _ in image_processor]
    for image in images:
        raw = _to_u8_hwc(image, device)
        for i, proc in enumerate(image_processor):
            if not hasattr(proc, "image_mean"):
                raise NotImplementedError("processors without image_mean (mm_utils.py:192) are not used by Cambrian-1")
            size = int(proc.crop_size["height"])
            pad = tuple(int(x * 255) for x in proc.image_mean)                      # mm_utils.py:194
            std = getattr(proc, "image_std", None)
            if std is None:
                raise ValueError("process_images: processor lacks image_std")
            per_tower[i].append(ops.preprocess_image(raw, size, pad, proc.image_mean, std))
    return [torch.stack(t, 0) for t in per_tower]
