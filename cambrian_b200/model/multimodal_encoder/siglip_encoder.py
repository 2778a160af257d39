from .towers import SiglipVisionTower  # noqa: F401  (same import path as the reference's siglip_encoder.py)
