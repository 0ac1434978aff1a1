"""DPhysConfig: host-side mirror of the reference's robot / terrain / simulation constants.

Mirrors the attribute bag of the reference (`dphys_config.py:77-153`) so that callers that
build a config, tweak `traj_sim_time` / `dt` / `use_odeint` and hand it to `DPhysics` keep
working.  Differences, all deliberate:

* the contact points, driving-part masks and robot size are loaded from pre-computed
  tables (`monoforce_b200/data/<robot>.npz`, minted by `tools/make_robot_data.py` with
  the reference's own geometry recipe, `dphys_config.py:8-74`) - no open3d at run time;
* `part_id` (N,) int32 is the per-point driving part the kernel consumes (the last mask
  containing the point wins, as in the reference's assignment loop `dphysics.py:243-246`).
"""
from __future__ import annotations

import os

import numpy as np
import torch
import yaml

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def _canonical_robot(robot: str) -> str:
    # dphys_config.py:20-23 folds e.g. 'tradr2' onto 'tradr'
    if "tradr" in robot:
        return "tradr"
    if "marv" in robot:
        return "marv"
    return robot


def robot_geometry(robot: str):
    """(points (N,3) f32, [masks (N,) bool], (Lx, Ly) 0-dim tensors) - `dphys_config.py:38-74`."""
    path = os.path.join(_DATA, f"{_canonical_robot(robot)}.npz")
    assert os.path.exists(path), f'Mesh file {path} does not exist.'
    d = np.load(path)
    pts = torch.from_numpy(d["points"].astype(np.float32))
    masks = [torch.from_numpy(m.copy()) for m in d["masks"]]
    size = (torch.tensor(d["robot_size"][0]), torch.tensor(d["robot_size"][1]))
    return pts, masks, size


def part_ids(masks, n_points: int) -> torch.Tensor:
    pid = torch.full((n_points,), -1, dtype=torch.int32)
    for i, m in enumerate(masks):
        pid[m] = i
    return pid


_JOINTS = {
    "tradr": dict(fl=[0.250, 0.272, 0.019], fr=[0.250, -0.272, 0.019], rl=[-0.250, 0.272, 0.019], rr=[-0.250, -0.272, 0.019]),
    "marv": dict(fl=[0.250, 0.272, 0.019], fr=[0.250, -0.272, 0.019], rl=[-0.250, 0.272, 0.019], rr=[-0.250, -0.272, 0.019]),
    "husky": dict(fl=[0.256, 0.285, 0.033], fr=[0.256, -0.285, 0.033], rl=[-0.256, 0.285, 0.033], rr=[-0.256, -0.285, 0.033]),
}
_MASS = {"tradr": 40., "marv": 60., "husky": 50.}


class DPhysConfig:
    def __init__(self, robot='marv', grid_res=0.1):
        self.robot = robot
        self.vel_max = 1.0      # m/s
        self.omega_max = 2.0    # rad/s
        kind = next((k for k in ("tradr", "marv", "husky") if k in robot), None)
        if kind is None:
            raise ValueError(f'Robot {robot} not supported. Available robots: tradr, marv, husky')
        self.robot_mass = _MASS[kind]
        self.joint_positions = {k: list(v) for k, v in _JOINTS[kind].items()}
        self.joint_angles = dict(fl=0.0, fr=0.0, rl=0.0, rr=0.0)
        self.robot_points, self.driving_parts, self.robot_size = robot_geometry(robot)
        self.part_id = part_ids(self.driving_parts, self.robot_points.shape[0])

        self.gravity = 9.81
        self.gravity_direction = torch.tensor([0., 0., -1.])

        # height map
        self.grid_res = grid_res
        self.r_min = 0.6
        self.d_max = 6.4
        self.h_max = 2.0
        x_grid = torch.arange(-self.d_max, self.d_max, self.grid_res)
        y_grid = torch.arange(-self.d_max, self.d_max, self.grid_res)
        self.x_grid, self.y_grid = torch.meshgrid(x_grid, y_grid, indexing='ij')
        self.z_grid = torch.zeros_like(self.x_grid)
        self.friction = 1.0 * torch.ones_like(self.z_grid)
        self.stiffness = 50_000.
        self.damping = np.sqrt(4 * self.robot_mass * self.stiffness)   # critical damping
        self.hm_interp_method = None

        # trajectory shooting
        self.traj_sim_time = 5.0
        self.dt = 0.01
        self.n_sim_trajs = 64
        self.integration_mode = 'euler'
        self.use_odeint = True      # reference default (dphys_config.py:153)

    def __str__(self):
        return str(self.__dict__)

    def to_yaml(self, path):
        out = {}
        for k, v in self.__dict__.items():
            if isinstance(v, (np.ndarray, torch.Tensor)):
                v = v.tolist()
            elif isinstance(v, (list, tuple)) and len(v) and isinstance(v[0], torch.Tensor):
                v = [t.tolist() for t in v]
            elif isinstance(v, np.generic):
                v = v.item()
            out[k] = v
        with open(path, 'w') as f:
            yaml.safe_dump(out, f)

    def from_yaml(self, path):
        with open(path, 'r') as f:
            params = yaml.load(f, Loader=yaml.FullLoader)
        for k, v in params.items():
            setattr(self, k, v)
