from .towers import DinoVisionTower  # noqa: F401  (same import path as the reference's dino_encoder.py)
