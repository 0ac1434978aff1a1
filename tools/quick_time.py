"""Scratch timing of the rollout kernels at BASELINE config 2/3 size (not the bench)."""
import sys, os, time
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _R); sys.path.insert(0, os.path.join(_R, 'tests'))
import torch
from monoforce_b200 import DPhysics, DPhysConfig
from helpers_mfb import hill_map

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = 400
cfg = DPhysConfig(robot="marv", grid_res=0.05)
cfg.traj_sim_time, cfg.use_odeint = T * cfg.dt, False
sim = DPhysics(cfg, device="cuda")
z = hill_map(cfg).cuda()
gen = torch.Generator().manual_seed(0)
controls = (torch.rand(B, 1, 2, generator=gen) * torch.tensor([2.0, 4.0]) - torch.tensor([1.0, 2.0])).repeat(1, T, 1).cuda()

def timed(fn, n=5):
    """device time of the library calls only (events inside DPhysics), summed per invocation of fn"""
    for _ in range(2): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        sim.timings = []
        fn(); torch.cuda.synchronize()
        ts.append(sum(a.elapsed_time(b) for _, a, b in sim.timings))
    sim.timings = None
    return min(ts), sum(ts) / len(ts)

with torch.no_grad():
    best, mean = timed(lambda: sim(z.unsqueeze(0), controls))
bytes_ = B * T * 5432
print(f"fwd forces: best {best:.3f} ms mean {mean:.3f} ms -> {B*T/best*1e3:.3e} steps/s, {bytes_/best/1e6:.0f} GB/s ({bytes_/best/1e6/6549.4:.1%} of 6549 GB/s)")
sim.return_forces = False
with torch.no_grad():
    best, mean = timed(lambda: sim(z.unsqueeze(0), controls))
print(f"fwd no-forces: best {best:.3f} ms mean {mean:.3f} ms -> {B*T/best*1e3:.3e} steps/s")
sim.return_forces = True
zk = z.clone().requires_grad_(True)
def fb():
    zk.grad = None
    (Xs, _, _, _), _ = sim(zk.unsqueeze(0), controls)
    Xs.pow(2).mean().backward()
best, mean = timed(fb)
print(f"fwd+bwd (forces materialised): best {best:.3f} ms mean {mean:.3f} ms -> {B*T/best*1e3:.3e} steps/s")
sim.return_forces = False
best, mean = timed(fb)
print(f"fwd+bwd (no forces): best {best:.3f} ms mean {mean:.3f} ms -> {B*T/best*1e3:.3e} steps/s")
ck = controls.clone().requires_grad_(True)
def fb_ctrl():
    ck.grad = None
    (Xs, _, _, _), _ = sim(z.unsqueeze(0), ck)
    Xs.pow(2).mean().backward()
best, mean = timed(fb_ctrl)
print(f"fwd+bwd (no forces, grads to controls only = no atomics): best {best:.3f} ms")
fk = torch.full_like(z, 0.5).requires_grad_(True)
def fb_both():
    zk.grad = None; fk.grad = None
    (Xs, _, _, _), _ = sim(zk.unsqueeze(0), controls, friction=fk.unsqueeze(0))
    Xs.pow(2).mean().backward()
best, mean = timed(fb_both)
print(f"fwd+bwd (no forces, grads to z and friction): best {best:.3f} ms")
