"""Drop-in for monoforce/src/monoforce/models/traj_predictor/dphysics.py (names callers import)."""
from monoforce_b200.dphysics import (DPhysics, DPhysConfig, generate_controls, vw_to_track_vels,  # noqa: F401
                                     inertia_tensor, normalized, skew_symmetric)
