"""Host-side cost of one planner-size DPhysics call (64 x 500, shared 128x128 map, no_grad): cProfile over 300 calls + the
CUDA launch list of one call.   python tools/profile_call_overhead.py [odeint]"""
import cProfile
import os
import pstats
import sys
import time
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
zr = z.repeat(64, 1, 1)
ctrl = torch.stack([torch.rand(64, T, generator=g) * 2 - 1, torch.rand(64, T, generator=g) * 4 - 2], -1).cuda()


def call(zz=z):
    with torch.no_grad():
        return sim(z_grid=zz, controls=ctrl)


for zz, name in ((z, "shared (1,H,W) map"), (zr, "64 repeated maps")):
    for _ in range(5):
        call(zz)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200):
        call(zz)
        torch.cuda.synchronize()
    t1 = time.perf_counter()
    for _ in range(200):
        call(zz)
    t2 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"{name}: {1e3 * (t1 - t0) / 200:.3f} ms per call incl. synchronize; host-only enqueue time {1e3 * (t2 - t1) / 200:.3f} ms")

pr = cProfile.Profile()
pr.enable()
for _ in range(300):
    call()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)

from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    call()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
