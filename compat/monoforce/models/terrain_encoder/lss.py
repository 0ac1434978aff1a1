"""Drop-in for monoforce/src/monoforce/models/terrain_encoder/lss.py (names callers import)."""
from monoforce_b200.terrain_encoder import LiftSplatShoot, CamEncode, BevEncode, Up, ScaledTanh  # noqa: F401
