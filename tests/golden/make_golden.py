#!/usr/bin/env python
"""Mint golden vectors by executing the UNMODIFIED reference DPhysics on CPU.

Build-container only (needs /root/reference); the reference is imported through the
open3d / torchdiffeq stand-ins in oracle/shims.  The reference has
no tests or fixtures of its own, so these files are the parity anchor:

    python tests/golden/make_golden.py        # rewrites tests/golden/*.npz

Every case stores its inputs and the reference outputs (fp32 forward cases) or the
reference's autograd gradients (fp64 gradient cases).  Large force tensors are stored
at a few time indices plus per-step checksums to keep the fixtures small.
"""
import os
import sys
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.ref_import import import_reference  # noqa: E402


def hill(cfg, noise=0.0, gen=None):
    xg, yg = cfg.x_grid, cfg.y_grid
    z = torch.exp(-(xg - 2) ** 2 / 4) * torch.exp(-(yg - 0) ** 2 / 2)      # robot_control.py:52
    if noise:
        z = z + noise * torch.randn(z.shape, generator=gen, dtype=z.dtype)
    return z


def yaw_state(B, gen, dtype):
    x = torch.randn(B, 3, generator=gen, dtype=dtype) * 0.5
    xd = torch.randn(B, 3, generator=gen, dtype=dtype) * 0.2
    a = torch.rand(B, generator=gen, dtype=dtype) * 6.28
    R = torch.zeros(B, 3, 3, dtype=dtype)
    R[:, 0, 0] = a.cos(); R[:, 0, 1] = -a.sin(); R[:, 1, 0] = a.sin(); R[:, 1, 1] = a.cos(); R[:, 2, 2] = 1
    om = torch.randn(B, 3, generator=gen, dtype=dtype) * 0.1
    return x, xd, R, om


def const_controls(B, n, gen, dtype, v=(-1, 1), w=(-2, 2)):
    vv = torch.rand(B, generator=gen, dtype=dtype) * (v[1] - v[0]) + v[0]
    ww = torch.rand(B, generator=gen, dtype=dtype) * (w[1] - w[0]) + w[0]
    return torch.stack([vv[:, None].repeat(1, n), ww[:, None].repeat(1, n)], -1)


def forward_case(dp, cfgm, name, robot, grid_res, T, B, terrain, variant="step", joints=False,
                 given_state=False, fric=False, fixed_ctrl=None, seed=0, keep_steps=(0, 1)):
    gen = torch.Generator().manual_seed(seed)
    cfg = cfgm.DPhysConfig(robot=robot, grid_res=grid_res)
    cfg.traj_sim_time = T * cfg.dt
    cfg.use_odeint = (variant == "odeint")
    sim = dp.DPhysics(cfg)
    if terrain == "flat":
        z = torch.zeros_like(cfg.x_grid)
    elif terrain == "ramp":
        # Diagonal ramp z = 0.25 (x + y): the only sloped terrain on which the reference's sampling
        # (x+1 / y+1 neighbours carry each other's weight, dphysics.py:442-445) is CONTINUOUS across
        # cell borders.  Elsewhere a contact point crossing a border gets a height jump of
        # z[i,j+1] - z[i+1,j], so two fp32 implementations that differ by one ulp in the point
        # position can disagree about the step at which the jump happens (parity).
        z = 0.25 * (cfg.x_grid + cfg.y_grid)
    elif terrain == "hill":
        z = hill(cfg)
    else:
        z = hill(cfg, 0.02, gen)
    n = int(cfg.traj_sim_time / cfg.dt)
    if fixed_ctrl is not None:
        controls = torch.tensor([[list(fixed_ctrl)] * n] * B)
    else:
        controls = const_controls(B, n, gen, torch.float32)
    friction = (0.3 + 0.7 * torch.rand(z.shape, generator=gen)) if fric else None
    ja = None
    if joints:
        ramp = torch.linspace(-1, 1, n).repeat(B, 1)
        ja = torch.stack([ramp, ramp, -ramp, -ramp], -1)
    st = yaw_state(B, gen, torch.float32) if given_state else None
    st_in = tuple(s.clone() for s in st) if st else None
    with torch.no_grad():
        states, forces = sim(z.repeat(B, 1, 1), controls, joint_angles=ja,
                             state=tuple(s.clone() for s in st) if st else None,
                             friction=None if friction is None else friction.repeat(B, 1, 1))
    Xs, Xds, Rs, Oms = states
    Fs, Ff = forces
    keep = sorted(set([k if k >= 0 else Fs.shape[1] + k for k in keep_steps] + [Fs.shape[1] - 1]))
    out = dict(
        robot=robot, grid_res=grid_res, T=T, variant=variant,
        z=z.numpy(), controls=controls.numpy(),
        Xs=Xs.numpy(), Xds=Xds.numpy(), Rs=Rs.numpy(), Omegas=Oms.numpy(),
        F_keep_steps=np.asarray(keep), Fs_keep=Fs[:, keep].numpy(), Ff_keep=Ff[:, keep].numpy(),
        Fs_sum=Fs.double().sum(dim=2).numpy(), Ff_sum=Ff.double().sum(dim=2).numpy(),
        Fs_abs=Fs.double().abs().sum(dim=(2, 3)).numpy(), Ff_abs=Ff.double().abs().sum(dim=(2, 3)).numpy(),
        cost=torch.norm(Fs, dim=-1).std(dim=-1).std(dim=-1).numpy(),        # monoforce_node.py:91
    )
    if friction is not None:
        out["friction"] = friction.numpy()
    if ja is not None:
        out["joint_angles"] = ja.numpy()
    if st_in is not None:
        for k, v in zip(("x0", "xd0", "R0", "om0"), st_in):
            out[k] = v.numpy()
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: Xs[0,-1]={Xs[0, -1].tolist()} cost[0]={out['cost'][0]:.6g} -> {os.path.getsize(path)} B")


def grad_case(dp, cfgm, name, robot, grid_res, T, B, variant="step", seed=0):
    """fp64 reference gradients (procedure: SURVEY.md appendix B)."""
    torch.set_default_dtype(torch.float64)
    try:
        gen = torch.Generator().manual_seed(seed)
        cfg = cfgm.DPhysConfig(robot=robot, grid_res=grid_res)
        cfg.traj_sim_time = T * cfg.dt
        cfg.use_odeint = (variant == "odeint")
        sim = dp.DPhysics(cfg)
        sim.x_points = sim.x_points.double()
        z = hill(cfg, 0.02, gen).double()
        n = int(cfg.traj_sim_time / cfg.dt)
        controls = const_controls(B, n, gen, torch.float64, v=(0.3, 1.0))
        friction = 0.3 + 0.7 * torch.rand(z.shape, generator=gen, dtype=torch.float64)
        st = yaw_state(B, gen, torch.float64)
        zz = z.clone().requires_grad_(True)
        ff = friction.clone().requires_grad_(True)
        cc = controls.clone().requires_grad_(True)
        s_in = [s.clone().requires_grad_(True) for s in st]
        # weights for a scalar objective touching every output
        states, forces = sim(zz.unsqueeze(0).expand(B, -1, -1), cc, state=tuple(s * 1.0 for s in s_in),
                             friction=ff.unsqueeze(0).expand(B, -1, -1))
        outs = list(states) + list(forces)
        # weights are rounded to fp32 so the stored (compact) copy reproduces the loss exactly
        ws = [torch.randn(o.shape, generator=gen, dtype=torch.float64).float().double() for o in outs]
        scales = [1.0, 1.0, 1.0, 1.0, 1e-3, 1e-3]
        loss = sum(s * (o * w).sum() for s, o, w in zip(scales, outs, ws))
        loss.backward()
        out = dict(robot=robot, grid_res=grid_res, T=T, variant=variant, z=z.numpy(), friction=friction.numpy(),
                   controls=controls.numpy(), x0=st[0].numpy(), xd0=st[1].numpy(), R0=st[2].numpy(), om0=st[3].numpy(),
                   loss=loss.item(), scales=np.asarray(scales),
                   g_z=zz.grad.numpy(), g_friction=ff.grad.numpy(), g_controls=cc.grad.numpy(),
                   g_x0=s_in[0].grad.numpy(), g_xd0=s_in[1].grad.numpy(), g_R0=s_in[2].grad.numpy(),
                   g_om0=s_in[3].grad.numpy())
        for i, w in enumerate(ws):
            out[f"w{i}"] = w.numpy().astype(np.float32) if i >= 4 else w.numpy()
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: loss={loss.item():.9g} |g_z|={zz.grad.abs().sum().item():.6g} -> {os.path.getsize(path)} B")
    finally:
        torch.set_default_dtype(torch.float32)


def main():
    dp, cfgm = import_reference()
    # BASELINE config 1: 1 trajectory, flat 64x64, T=100, controls == (1.0, 0.3)
    forward_case(dp, cfgm, "cfg1_marv_flat64_T100", "marv", 0.2, 100, 1, "flat", fixed_ctrl=(1.0, 0.3))
    forward_case(dp, cfgm, "cfg1_tradr_flat64_T100", "tradr", 0.2, 100, 1, "flat", fixed_ctrl=(1.0, 0.3))
    forward_case(dp, cfgm, "marv_hill128_T100_B4", "marv", 0.1, 100, 4, "hill", seed=1)
    forward_case(dp, cfgm, "marv_noise128_state_fric_T100_B4", "marv", 0.1, 100, 4, "noise",
                 given_state=True, fric=True, seed=2)
    forward_case(dp, cfgm, "tradr_noise128_state_fric_T100_B4", "tradr", 0.1, 100, 4, "noise",
                 given_state=True, fric=True, seed=3)
    forward_case(dp, cfgm, "marv_flat256_T400_B2", "marv", 0.05, 400, 2, "flat", seed=4)
    forward_case(dp, cfgm, "marv_hill128_odeint_T60_B2", "marv", 0.1, 60, 2, "hill", variant="odeint", seed=5)
    forward_case(dp, cfgm, "marv_ramp128_odeint_T200_B3", "marv", 0.1, 200, 3, "ramp", variant="odeint", seed=9)
    forward_case(dp, cfgm, "marv_ramp256_T400_B3", "marv", 0.05, 400, 3, "ramp", seed=10)
    forward_case(dp, cfgm, "tradr_ramp128_T300_B3", "tradr", 0.1, 300, 3, "ramp", seed=11)
    forward_case(dp, cfgm, "marv_hill128_joints_T60_B2", "marv", 0.1, 60, 2, "hill", joints=True, seed=6)
    grad_case(dp, cfgm, "grad64_marv_noise128_T40_B2", "marv", 0.1, 40, 2, seed=7)
    grad_case(dp, cfgm, "grad64_tradr_noise64_T40_B2", "tradr", 0.2, 40, 2, seed=8)


if __name__ == "__main__":
    main()
