from .towers import BaseVisionTower  # noqa: F401  (same import path as the reference's base_encoder.py)
