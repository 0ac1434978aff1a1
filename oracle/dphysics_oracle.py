"""CPU oracle for the DPhysics trajectory rollout  --  TEST INFRASTRUCTURE, NOT PRODUCT.

A batched torch-CPU restatement (fp32 or fp64) of the reference algorithm in
/root/reference/monoforce/src/monoforce/models/traj_predictor/dphysics.py.  Only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this file; the product (monoforce_b200/) never does and has no CPU fallback.

Parity status: PINNED.  The reference ships no tests or golden vectors,
so the goldens under tests/golden/ were minted by executing the unmodified reference
through oracle/shims (tests/golden/make_golden.py) and this oracle is checked against
them bit-for-bit in fp32 (tests/test_oracle_golden.py) and, in the build container, against
the live reference (tests/test_oracle_vs_reference.py).  The two third-party pieces the
reference leans on (torchdiffeq fixed-grid Euler, open3d voxel down-sample) are absent
from /root/reference and are restated from their published algorithms in oracle/shims;
for those two the parity is "unpinned" beyond N_pts = 223 (diff_physics.ipynb:217).

Each function names the reference lines it follows.  The arithmetic order mirrors the
reference op-for-op so that fp32 results agree to the last bit on the same torch build.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import math
import numpy as np
import torch


@dataclass
class RolloutSpec:
    """Constants the reference reads from DPhysConfig (dphys_config.py:77-153)."""
    points: torch.Tensor                 # (N,3) body-frame contact points
    masks: List[torch.Tensor]            # driving-part masks, each (N,) bool
    robot_size: Tuple[float, float]      # (Lx, Ly)  (0-dim tensors in the reference)
    mass: float = 60.0
    gravity: float = 9.81
    stiffness: float = 50_000.0
    damping: float = field(default=None)  # sqrt(4 m k) when None (dphys_config.py:143)
    grid_res: float = 0.1
    d_max: float = 6.4
    dt: float = 0.01
    traj_sim_time: float = 5.0
    omega_max: float = 2.0
    robot: str = "marv"
    joint_positions: Sequence[Sequence[float]] = (
        (0.250, 0.272, 0.019), (0.250, -0.272, 0.019), (-0.250, 0.272, 0.019), (-0.250, -0.272, 0.019))

    def __post_init__(self):
        if self.damping is None:
            self.damping = float(np.sqrt(4 * self.mass * self.stiffness))


# ---------------------------------------------------------------------------------------
# small helpers
# ---------------------------------------------------------------------------------------
def unit(v: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """dphysics.py:7-19 (`normalized`): divide by the norm clamped from below."""
    return v / torch.clamp(torch.norm(v, dim=-1, keepdim=True), min=eps)


def hat(v: torch.Tensor) -> torch.Tensor:
    """dphysics.py:22-40 (`skew_symmetric`)."""
    S = torch.zeros(v.shape[0], 3, 3, dtype=v.dtype)
    S[:, 0, 1] = -v[:, 2]
    S[:, 0, 2] = v[:, 1]
    S[:, 1, 2] = -v[:, 0]
    S[:, 1, 0] = v[:, 2]
    S[:, 2, 0] = -v[:, 1]
    S[:, 2, 1] = v[:, 0]
    return S


def point_mass_inertia(mass: float, pts: torch.Tensor) -> torch.Tensor:
    """dphysics.py:107-141 (`inertia_tensor`): uniform point masses about the body origin."""
    mp = mass / pts.shape[1]
    px, py, pz = pts[:, :, 0], pts[:, :, 1], pts[:, :, 2]
    ixx = torch.sum(mp * (py ** 2 + pz ** 2), dim=1)
    iyy = torch.sum(mp * (px ** 2 + pz ** 2), dim=1)
    izz = torch.sum(mp * (px ** 2 + py ** 2), dim=1)
    ixy = -torch.sum(mp * px * py, dim=1)
    ixz = -torch.sum(mp * px * pz, dim=1)
    iyz = -torch.sum(mp * py * pz, dim=1)
    rows = [torch.stack([ixx, ixy, ixz], dim=1),
            torch.stack([ixy, iyy, iyz], dim=1),
            torch.stack([ixz, iyz, izz], dim=1)]
    return torch.stack(rows, dim=1)


def track_speeds(v, w, Ly, n_tracks: int):
    """dphysics.py:75-104 (`vw_to_track_vels`); note the two spellings of the same product."""
    if n_tracks == 2:
        return torch.stack([v - w * (Ly / 2.0), v + w * (Ly / 2.0)], dim=-1)
    if n_tracks == 4:
        lo = v - w * Ly / 2.0
        hi = v + w * Ly / 2.0
        return torch.stack([lo, hi, v - w * Ly / 2.0, v + w * Ly / 2.0], dim=-1)
    raise ValueError("n_tracks must be 2 or 4")


def sample_grid(grid: torch.Tensor, xq: torch.Tensor, yq: torch.Tensor, d_max: float, res: float,
                want_normals: bool = False):
    """dphysics.py:385-455 (`interpolate_grid`), quirks kept:
    truncation toward zero (`.long()`), flat-index clamp (no per-axis clamp), the
    x+1 / y+1 neighbours carrying each other's bilinear weight, forward-difference
    normals that are constant per cell."""
    B, H, W = grid.shape
    flat = grid.reshape(B, -1)
    gx = (xq.reshape(B, -1) + d_max) / res
    gy = (yq.reshape(B, -1) + d_max) / res
    ix = gx.long()
    iy = gy.long()
    fx = gx - ix.to(grid.dtype)
    fy = gy - iy.to(grid.dtype)
    last = H * W - 1
    k00 = torch.clamp(iy + H * ix, 0, last)
    k10 = torch.clamp(iy + H * (ix + 1), 0, last)        # "front": x+1
    k01 = torch.clamp((iy + 1) + H * ix, 0, last)        # "left":  y+1
    k11 = torch.clamp((iy + 1) + H * (ix + 1), 0, last)
    z00 = flat.gather(1, k00)
    z10 = flat.gather(1, k10)
    z01 = flat.gather(1, k01)
    z11 = flat.gather(1, k11)
    val = (1 - fx) * (1 - fy) * z00 + (1 - fx) * fy * z10 + fx * (1 - fy) * z01 + fx * fy * z11
    if not want_normals:
        return val
    sx = (z10 - z00) / res
    sy = (z01 - z00) / res
    nrm = unit(torch.stack([-sx, -sy, torch.ones_like(sx)], dim=-1))
    return val, nrm


def articulate(spec: RolloutSpec, pts0: torch.Tensor, angles: torch.Tensor) -> torch.Tensor:
    """dphysics.py:326-358 (`update_joints`): static points unless marv with non-zero angles."""
    B = angles.shape[0]
    pts = pts0.repeat(B, 1, 1)
    if spec.robot != "marv" or torch.allclose(angles, torch.zeros_like(angles)):
        return pts
    for i, mask in enumerate(spec.masks):
        piv = torch.as_tensor(spec.joint_positions[i], dtype=pts.dtype).unsqueeze(0)
        a = angles[:, i]
        zero, one = torch.zeros_like(a), torch.ones_like(a)
        Ry = torch.stack([torch.cos(a), zero, torch.sin(a),
                          zero, one, zero,
                          -torch.sin(a), zero, torch.cos(a)], dim=1).view(B, 3, 3)
        sel = pts[:, mask]
        sel = sel - piv.unsqueeze(1)
        sel = sel @ Ry.transpose(1, 2)
        sel = sel + piv.unsqueeze(1)
        pts[:, mask] = sel
    return pts


# ---------------------------------------------------------------------------------------
# one evaluation of the right-hand side
# ---------------------------------------------------------------------------------------
def derivatives(spec: RolloutSpec, pts0, z_grid, friction, ctrl_t, angles_t, state):
    """dphysics.py:172-272 (`forward_kinematics`) for one time index.

    Returns (xdd, dR, omega_d), (F_spring, F_friction)."""
    x, xd, R, om = state
    m, g = spec.mass, spec.gravity
    body = articulate(spec, pts0, angles_t)                                   # :192
    Iinv = torch.linalg.inv(point_mass_inertia(m, body))                      # :196-197
    P = body @ R.transpose(1, 2) + x.unsqueeze(1)                             # :200
    Pd = xd.unsqueeze(1) + torch.linalg.cross(om.unsqueeze(1), P - x.unsqueeze(1))   # :204
    zP, n = sample_grid(z_grid, P[..., 0], P[..., 1], spec.d_max, spec.grid_res, want_normals=True)  # :211
    zP = zP.unsqueeze(-1)
    mu = sample_grid(friction, P[..., 0], P[..., 1], spec.d_max, spec.grid_res).unsqueeze(-1)       # :216
    dh = P[..., 2:3] - zP                                                     # :220
    contact = torch.sigmoid(-10. * dh)                                        # :223
    vn = (Pd * n).sum(dim=2).unsqueeze(2)                                     # :228
    Fs = -torch.mul((spec.stiffness * dh + spec.damping * vn), n)             # :230
    Fs = torch.mul(Fs, contact) / torch.sum(contact, dim=1, keepdim=True)     # :231-232
    Fs = torch.clamp(Fs, min=-m * g, max=m * g)                               # :233
    heading = unit(R[..., 0])                                                 # :237
    Nf = torch.norm(Fs, dim=2)                                                # :238
    Ly = spec.robot_size[1]
    tv = track_speeds(ctrl_t[:, 0], ctrl_t[:, 1], Ly, len(spec.masks))        # :239-240
    cmd = torch.zeros_like(Pd)
    for i, mask in enumerate(spec.masks):                                     # :243-246
        cmd[:, mask] = (tv[:, i].unsqueeze(1) * heading).unsqueeze(1)
    slip = mu * (cmd - Pd)                                                    # :247
    slip_n = (slip * n).sum(dim=2).unsqueeze(2)
    slip_t = slip - slip_n * n                                                # :249
    Ff = torch.clamp(Nf.unsqueeze(2) * slip_t, min=-m * g, max=m * g)         # :250-251
    torque = torch.sum(torch.linalg.cross(P - x.unsqueeze(1), Fs + Ff), dim=1)  # :255
    om_d = (Iinv @ torque.unsqueeze(2)).squeeze(2)                            # :256
    om_d = torch.clamp(om_d, min=-spec.omega_max, max=spec.omega_max)         # :257
    dR = hat(om) @ R                                                          # :258-259
    grav = m * g * torch.tensor([0., 0., -1.], dtype=x.dtype).unsqueeze(0)    # :264
    xdd = (grav + Fs.sum(dim=1) + Ff.sum(dim=1)) / m                          # :265-266
    return (xdd, dR, om_d), (Fs, Ff)


def rotate_rodrigues(R, om, dt, eps=1e-6):
    """dphysics.py:290-324 (`integrate_rotation`): R <- R (I + K sin(th dt) + K K (1 - cos(th dt)))."""
    S = hat(om)
    th = torch.norm(om, dim=-1, keepdim=True).unsqueeze(-1)
    K = S / torch.clamp(th, min=eps)
    eye = torch.eye(3, dtype=R.dtype)
    return R @ (eye + K * torch.sin(th * dt) + K @ K * (1 - torch.cos(th * dt)))


def advance(state, rates, dt):
    """dphysics.py:274-288 + :360-383 (`update_state`, euler mode): semi-implicit Euler."""
    x, xd, R, om = state
    xdd, _, om_d = rates
    xd = xd + xdd * dt
    x = x + xd * dt
    om = om + om_d * dt
    R = rotate_rodrigues(R, om, dt)
    return x, xd, R, om


# ---------------------------------------------------------------------------------------
# whole rollout
# ---------------------------------------------------------------------------------------
def time_grid(spec: RolloutSpec, dtype) -> torch.Tensor:
    """dphysics.py:166-167: ts = linspace(0, T, int(T / dt))."""
    T, dt = spec.traj_sim_time, spec.dt
    return torch.linspace(0, T, int(T / dt), dtype=dtype)


def rollout(spec: RolloutSpec, z_grid, controls, joint_angles=None, state=None, friction=None,
            variant: str = "step", dtype=torch.float32, mutate_state: bool = False):
    """dphysics.py:530-594 (`dphysics`) with `dynamics` (:467-497, variant="step") or
    `dynamics_odeint` (:499-528, variant="odeint").

    All tensors are converted to `dtype` on the CPU.  Returns
    ((Xs, Xds, Rs, Omegas), (F_springs, F_frictions)) like the reference."""
    cast = lambda t: torch.as_tensor(t).detach().to(dtype=dtype, device="cpu") if not (
        isinstance(t, torch.Tensor) and t.dtype == dtype and t.device.type == "cpu") else t
    z_grid = cast(z_grid)
    controls = cast(controls)
    B = z_grid.shape[0]
    pts0 = spec.points.to(dtype).unsqueeze(0)                                  # :149
    if state is None:                                                          # :554-559
        x = torch.zeros(B, 3, dtype=dtype)
        xd = torch.zeros_like(x)
        xd[:, 0] = controls[:, 0, 0]
        R = torch.eye(3, dtype=dtype).repeat(B, 1, 1)
        om = torch.zeros_like(x)
        om[:, 2] = controls[:, 0, 1]
        state = (x, xd, R, om)
    else:
        state = tuple(cast(s) if mutate_state else cast(s).clone() for s in state)
    if friction is None:                                                       # :562
        friction = torch.ones_like(z_grid)
    friction = cast(friction)

    # start-height snap, :567-571
    x = state[0]
    P0 = pts0.repeat(B, 1, 1) @ state[2].transpose(1, 2) + x.unsqueeze(1)
    x[..., 2:3] = sample_grid(z_grid, P0[..., 0], P0[..., 1], spec.d_max, spec.grid_res).mean(dim=1, keepdim=True)

    n_ts = min(int(spec.traj_sim_time / spec.dt), controls.shape[1])           # :573
    assert controls.shape == (B, n_ts, 2), f'Controls shape {controls.shape} != {(B, n_ts, 2)}'
    if joint_angles is None:
        joint_angles = torch.zeros((B, n_ts, 4), dtype=dtype)
    joint_angles = cast(joint_angles)
    assert joint_angles.shape == (B, n_ts, 4), f'Joint angles shape {joint_angles.shape} != {(B, n_ts, 4)}'
    ts = time_grid(spec, dtype)[:n_ts]                                         # :581

    if variant == "step":                                                      # :467-497
        rec = [[] for _ in range(6)]
        for i in range(n_ts):
            rates, forces = derivatives(spec, pts0, z_grid, friction, controls[:, i], joint_angles[:, i], state)
            state = advance(state, rates, spec.dt)
            for lst, val in zip(rec, (*state, *forces)):
                lst.append(val)
        Xs, Xds, Rs, Oms, Fs, Ff = (torch.stack(r, dim=1) for r in rec)
    elif variant == "odeint":                                                  # :499-528 + fixed-grid Euler
        N = pts0.shape[1]
        y = (*state, torch.zeros(B, N, 3, dtype=dtype), torch.zeros(B, N, 3, dtype=dtype))
        rec = [[v] for v in y]
        for i in range(n_ts - 1):
            # t_id = argmin |t - ts| is i itself on the solver grid (:183)
            (xdd, dR, om_d), (fs, ff) = derivatives(spec, pts0, z_grid, friction, controls[:, i],
                                                    joint_angles[:, i], y[:4])
            h = ts[i + 1] - ts[i]
            dy = (y[1], xdd, dR, om_d, fs, ff)                                 # :269, :464
            y = tuple(a + h * b for a, b in zip(y, dy))
            for lst, val in zip(rec, y):
                lst.append(val)
        Xs, Xds, Rs, Oms, Fs, Ff = (torch.stack(r, dim=1) for r in rec)
    else:
        raise ValueError(variant)

    delta_h = spec.mass * spec.gravity / (spec.stiffness + 1e-6)               # :587
    Xs = Xs + Rs[:, :, :3, 2] * delta_h                                        # :589
    return (Xs, Xds, Rs, Oms), (Fs, Ff)


def path_cost(F_springs: torch.Tensor) -> torch.Tensor:
    """monoforce_ros/nodes/monoforce_node.py:91: per-trajectory traversal cost."""
    return torch.norm(F_springs, dim=-1).std(dim=-1).std(dim=-1)


def physics_loss(states_pred, states_gt, pred_ts, gt_ts, gamma=0.9):
    """monoforce/src/monoforce/losses.py:102-138 without the optional rotation term."""
    X, Xp = states_gt[0], states_pred[0]
    ids = torch.argmin(torch.abs(pred_ts.unsqueeze(1) - gt_ts.unsqueeze(2)), dim=2)
    Xp_at = Xp[torch.arange(X.shape[0]).unsqueeze(1), ids]
    w = 1. / (1. + gamma * gt_ts.unsqueeze(2))
    return ((Xp_at * w - X * w) ** 2).mean()
