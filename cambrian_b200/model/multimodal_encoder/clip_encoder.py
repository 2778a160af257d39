from .towers import ClipVisionTower  # noqa: F401  (same import path as the reference's clip_encoder.py)
