import sys, os
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _R); sys.path.insert(0, os.path.join(_R, 'tests'))
import torch
from helpers_mfb import make_spec, hill_map, rel_err, load_golden
from oracle import dphysics_oracle as O
from monoforce_b200 import DPhysics, DPhysConfig
g = load_golden("marv_hill128_odeint_T60_B2")
for variant in ("odeint", "step"):
    cfg = DPhysConfig(robot="marv", grid_res=0.1); cfg.traj_sim_time = 0.6; cfg.use_odeint = variant == "odeint"
    sim = DPhysics(cfg, device="cuda")
    B = 2
    z = torch.from_numpy(g["z"]); controls = torch.from_numpy(g["controls"])
    T = controls.shape[1]
    r32 = O.rollout(make_spec(cfg), z.repeat(B, 1, 1), controls, variant=variant)
    r64 = O.rollout(make_spec(cfg), z.double().repeat(B, 1, 1), controls.double(), variant=variant, dtype=torch.float64)
    ks, kf = sim(z.cuda().unsqueeze(0), controls.cuda())
    print(variant, "T", T, "controls", controls[:, 0].tolist())
    for n, a, b, c in zip(("Xs", "Xds", "Rs", "Om", "Fs"), ks + kf, r32[0] + r32[1], r64[0] + r64[1]):
        a = a.cpu().double()
        print(" kernel-vs-fp64", n, ["%.1e" % ((a[:, t] - c[:, t]).abs().max()) for t in range(0, T, 6)])
        print(" ref32 -vs-fp64", n, ["%.1e" % ((b[:, t].double() - c[:, t]).abs().max()) for t in range(0, T, 6)])
