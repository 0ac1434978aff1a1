"""The reference's terrain_encoder/utils.py mixes LSS helpers with PIL / dataset helpers that are not on the
hot path; everything is re-exported from the reference tree, `gen_dx_bx` additionally from monoforce_b200."""
import importlib.util
import os
import sys

_here = os.path.abspath(os.path.dirname(__file__))
for _p in sys.path:
    _cand = os.path.join(_p, "monoforce", "models", "terrain_encoder", "utils.py")
    if os.path.isfile(_cand) and os.path.abspath(os.path.dirname(_cand)) != _here:
        _spec = importlib.util.spec_from_file_location("monoforce.models.terrain_encoder._ref_utils", _cand)
        _mod = importlib.util.module_from_spec(_spec)
        _spec.loader.exec_module(_mod)
        globals().update({k: v for k, v in vars(_mod).items() if not k.startswith("__")})
        break
from monoforce_b200.terrain_encoder import gen_dx_bx  # noqa: E402,F401
