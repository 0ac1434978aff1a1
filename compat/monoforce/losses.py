"""Drop-in for monoforce/src/monoforce/losses.py.  `physics_loss` (losses.py:102-138) is the B200 fused loss + gradient
kernel (csrc/physics_loss.cu), so scripts/train.py:405-406, scripts/fit_terrain.py:57 and scripts/eval.py:151 reach it
unchanged; hm_loss / total_variation / rotation_difference / translation_difference come from monoforce_b200.losses; anything
else (slerp ...) is re-exported from the reference tree when one is on sys.path."""
import importlib.util
import os
import sys

_here = os.path.abspath(os.path.dirname(__file__))
for _p in sys.path:
    _cand = os.path.join(_p, "monoforce", "losses.py")
    if os.path.isfile(_cand) and os.path.abspath(os.path.dirname(_cand)) != _here:
        _spec = importlib.util.spec_from_file_location("monoforce._ref_losses", _cand)
        _mod = importlib.util.module_from_spec(_spec)
        _spec.loader.exec_module(_mod)
        globals().update({k: v for k, v in vars(_mod).items() if not k.startswith("__")})
        break
from monoforce_b200.losses import (physics_loss, rotation_difference, translation_difference, total_variation,  # noqa: E402,F401
                                   hm_loss)
