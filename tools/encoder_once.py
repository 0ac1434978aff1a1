"""One warm + N profiled encoder forwards (fast path) for ncu launch lists:  python tools/encoder_once.py [--cfg4] [--b1] [n]"""
import os
import sys
import warnings

_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [_R, os.path.join(_R, "tests")]
import torch  # noqa: E402
from helpers_lss import default_cfg, make_inputs  # noqa: E402
from monoforce_b200 import LiftSplatShoot  # noqa: E402

cfg4 = "--cfg4" in sys.argv
n = int([a for a in sys.argv[1:] if a.isdigit()][0]) if any(a.isdigit() for a in sys.argv[1:]) else 1
gc, ac = default_cfg()
if cfg4:
    gc["xbound"] = [-6.4, 6.4, 0.05]; gc["ybound"] = [-6.4, 6.4, 0.05]; ac["final_dim"] = [512, 512]
torch.manual_seed(0)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    net = LiftSplatShoot(gc, ac).cuda().eval()
net.fast_inference = True
inputs = [t.cuda() for t in make_inputs(gc, ac, 1 if "--b1" in sys.argv else 16, 0)]
with torch.no_grad():
    for _ in range(1 + n):
        net(*inputs)
torch.cuda.synchronize()
print("done")
