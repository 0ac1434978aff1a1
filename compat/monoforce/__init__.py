"""Import shim: makes `monoforce.*` resolve to monoforce_b200 for the hot-path modules.

Put `<repo>/compat` on PYTHONPATH *before* the reference's `monoforce/src`: the reference's scripts
(`scripts/run.py:11-12`, `scripts/train.py`, `scripts/fit_terrain.py:5-7`) then import the B200
DPhysics / DPhysConfig / physics_loss unchanged, while every module that is NOT on the hot path
(datasets, vis, ros, transformations ...) still comes from the reference tree."""
import os
import pkgutil

# allow the rest of the reference's `monoforce` package (outside the hot path) to be found as well
__path__ = pkgutil.extend_path(__path__, __name__)
