"""Shared test helpers (imported as `helpers_mfb`; tests/ is put on sys.path by conftest.py)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_golden(name):
    d = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return {k: d[k] for k in d.files}


def make_spec(cfg):
    """oracle.RolloutSpec from a monoforce_b200 (or reference) DPhysConfig."""
    from oracle.dphysics_oracle import RolloutSpec
    return RolloutSpec(points=cfg.robot_points, masks=list(cfg.driving_parts), robot_size=cfg.robot_size,
                       mass=cfg.robot_mass, gravity=cfg.gravity, stiffness=cfg.stiffness, damping=float(cfg.damping),
                       grid_res=cfg.grid_res, d_max=cfg.d_max, dt=cfg.dt, traj_sim_time=cfg.traj_sim_time,
                       omega_max=cfg.omega_max, robot=cfg.robot, joint_positions=list(cfg.joint_positions.values()))


def hill_map(cfg, noise=0.0, seed=0, dtype=torch.float32):
    xg, yg = cfg.x_grid.to(dtype), cfg.y_grid.to(dtype)
    z = torch.exp(-(xg - 2) ** 2 / 4) * torch.exp(-(yg - 0) ** 2 / 2)
    if noise:
        g = torch.Generator().manual_seed(seed)
        z = z + noise * torch.randn(z.shape, generator=g, dtype=dtype)
    return z


def rel_err(a, b, floor=1e-12):
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return ((a - b).abs().max() / max(b.abs().max().item(), floor)).item()
