"""`build_vision_tower_aux_list` / `build_vision_tower` — mirror of cambrian/model/multimodal_encoder/builder.py:23-147,
restricted to the four Cambrian-1 towers (the reference's other 15 encoders are study-only)."""
from __future__ import annotations

import copy

from .towers import CLIPConvNextTower, ClipVisionTower, DinoVisionTower, SiglipVisionTower


def _build(name: str, cfg, **kwargs):
    low = name.lower()
    if "openai/clip" in low:
        return ClipVisionTower(name, args=cfg, **kwargs)
    if "siglip" in low:
        return SiglipVisionTower(name, args=cfg, **kwargs)
    if "clip-convnext" in low:
        return CLIPConvNextTower(name, args=cfg, **kwargs)
    if "dinov2" in low:
        return DinoVisionTower(name, args=cfg, **kwargs)
    raise ValueError(f"Unknown vision tower: {name}")  # builder.py:147


def build_vision_tower(vision_tower_cfg, **kwargs):
    name = getattr(vision_tower_cfg, "mm_vision_tower", getattr(vision_tower_cfg, "vision_tower", None))
    return _build(name, vision_tower_cfg, **kwargs)


def build_vision_tower_aux_list(vision_tower_cfg, **kwargs):
    names = getattr(vision_tower_cfg, "mm_vision_tower_aux_list", getattr(vision_tower_cfg, "vision_tower_aux_list", None))
    lens = getattr(vision_tower_cfg, "mm_vision_tower_aux_token_len_list",
                   getattr(vision_tower_cfg, "vision_tower_aux_token_len_list", None))
    towers = []
    for name, tok_len in zip(names, lens):
        cfg = copy.deepcopy(vision_tower_cfg)
        towers.append(_build(name + "-interp{}".format(tok_len), cfg, **kwargs))  # builder.py:92
    return towers
