from .towers import CLIPConvNextTower  # noqa: F401  (same import path as the reference's clip_convnext_encoder.py)
