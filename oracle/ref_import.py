"""Import the UNMODIFIED reference DPhysics from /root/reference through the stand-ins.

TEST INFRASTRUCTURE ONLY.  Works only in the build container (the GPU box has no
/root/reference); used by tests/golden/make_golden.py to mint golden vectors and by
tests/test_oracle_vs_reference.py (skipped when the reference tree is absent).
"""
import os
import sys

REFERENCE_SRC = "/root/reference/monoforce/src"
SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def reference_available() -> bool:
    return os.path.isdir(REFERENCE_SRC)


def import_reference():
    """Returns (dphysics_module, dphys_config_module) of the reference."""
    if not reference_available():
        raise RuntimeError("reference tree not present (expected in the build container only)")
    for p in (REFERENCE_SRC, SHIMS):
        if p not in sys.path:
            sys.path.insert(0, p)
    import importlib
    cfg_mod = importlib.import_module("monoforce.models.traj_predictor.dphys_config")
    dp_mod = importlib.import_module("monoforce.models.traj_predictor.dphysics")
    return dp_mod, cfg_mod
