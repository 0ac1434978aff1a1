"""monoforce_b200: B200-native (sm_100a) implementation of MonoForce's differentiable-physics
trajectory rollout behind the reference's own Python interface.

    from monoforce_b200 import DPhysics, DPhysConfig, generate_controls, LiftSplatShoot

The compute lives in libmonoforce_b200.so (C ABI: include/monoforce_b200.h), built in-tree by
`python -m monoforce_b200.build`.  There is no CPU fallback.
"""
from .dphys_config import DPhysConfig  # noqa: F401
from .dphysics import (DPhysics, generate_controls, vw_to_track_vels, inertia_tensor,  # noqa: F401
                       normalized, skew_symmetric, path_costs_from_forces)

from .terrain_encoder import LiftSplatShoot  # noqa: F401

__version__ = "0.1.0"
