"""Launch sequence for ncu: the planner-size forward (64 trajectories x 500 steps, one shared 128x128 map), step-loop variant
unless `odeint` is given.   python tools/profile_small.py [odeint]"""
import os
import sys
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _R)
import torch
from monoforce_b200 import DPhysics, DPhysConfig

cfg = DPhysConfig(robot="marv")
cfg.use_odeint = "odeint" in sys.argv[1:]
sim = DPhysics(cfg, device="cuda")
T = int(cfg.traj_sim_time / cfg.dt)
g = torch.Generator().manual_seed(0)
xg, yg = cfg.x_grid, cfg.y_grid
z = (torch.exp(-(xg - 2) ** 2 / 4) * torch.exp(-yg ** 2 / 2)).cuda()[None]
ctrl = torch.stack([torch.rand(64, T, generator=g) * 2 - 1, torch.rand(64, T, generator=g) * 4 - 2], -1).cuda()
for _ in range(3):
    with torch.no_grad():
        sim(z_grid=z, controls=ctrl)
torch.cuda.synchronize()
print("done")
