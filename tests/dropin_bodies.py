"""Bodies of two reference call sites, statement for statement, importing `monoforce.*` the way the scripts do:

  fit_terrain      monoforce/scripts/fit_terrain.py:12-62  (optimize_terrain; n_iters shortened, vis=False)
  predict_states   monoforce/scripts/train.py:96-99,231-246 (TrainerCore.terrain_preproc + predicts_states) followed by
                   train.py:402-406 (physics_loss on the predicted states) and a backward to the terrain maps

Run in a FRESH interpreter with PYTHONPATH deciding who provides `monoforce`:
  reference:   PYTHONPATH=/root/reference/monoforce/src:oracle/shims     (tests/golden/make_golden_dropin.py)
  this repo:   PYTHONPATH=compat[:/root/reference/monoforce/src]          (tests/test_dropin_gpu.py)

    python tests/dropin_bodies.py fit_terrain|predict_states OUT.npz [device]
"""
import sys

import numpy as np
import torch


def fit_terrain(n_iters=3):
    from monoforce.models.traj_predictor.dphysics import DPhysics
    from monoforce.models.traj_predictor.dphys_config import DPhysConfig
    from monoforce.losses import physics_loss, total_variation

    dphys_cfg = DPhysConfig(grid_res=0.4)
    T, dt = 6.0, 0.01
    dphys_cfg.dt = dt
    dphys_cfg.traj_sim_time = T
    x_grid, y_grid = dphys_cfg.x_grid, dphys_cfg.y_grid
    z_grid_gt = torch.exp(-(x_grid - 2.5) ** 2 / 1) * torch.exp(-(y_grid - 0) ** 2 / 4)
    z_grid_gt = z_grid_gt.repeat(1, 1, 1)
    controls = torch.tensor([[[1.0, 0.0]] * int(dphys_cfg.traj_sim_time / dphys_cfg.dt)])
    dphysics = DPhysics(dphys_cfg)
    states_gt, forces_gt = dphysics(z_grid=z_grid_gt, controls=controls)
    z_grid = torch.zeros_like(z_grid_gt, requires_grad=True)
    friction = 0.5 * torch.ones_like(z_grid)
    friction.requires_grad = True
    optimizer = torch.optim.Adam([{'params': z_grid, 'lr': 0.02}, {'params': friction, 'lr': 0.01}])
    losses, tvs = [], []
    ts = torch.arange(0, T, dt)[None]
    g_first = None
    for i in range(n_iters):
        optimizer.zero_grad()
        states, _ = dphysics(z_grid=z_grid, controls=controls, friction=friction)
        loss_traj = physics_loss(states_pred=states, states_gt=states_gt, pred_ts=ts, gt_ts=ts, gamma=0.9)
        loss_terrain = total_variation(z_grid)
        loss = loss_traj
        loss.backward()
        if g_first is None:
            g_first = (z_grid.grad.clone(), friction.grad.clone())
        optimizer.step()
        losses.append(loss_traj.item())
        tvs.append(loss_terrain.item())
    return dict(losses=np.asarray(losses), tv=np.asarray(tvs), Xs_gt=states_gt[0].detach().cpu().numpy(),
                Xs_last=states[0].detach().cpu().numpy(), z_grid=z_grid.detach().cpu().numpy(),
                friction=friction.detach().cpu().numpy(), g_z_first=g_first[0].cpu().numpy(),
                g_friction_first=g_first[1].cpu().numpy())


def predict_states(device='cpu', bsz=4):
    from monoforce.models.traj_predictor.dphysics import DPhysics
    from monoforce.models.traj_predictor.dphys_config import DPhysConfig
    from monoforce.losses import physics_loss

    dphys_cfg = DPhysConfig(robot='marv', grid_res=0.4)            # train.py:439 (--dphys_grid_res 0.4)
    dphys_cfg.traj_sim_time = 5.0                                  # train.py:440
    dphysics = DPhysics(dphys_cfg, device=device)                  # train.py:93
    kernel_size = int(dphys_cfg.grid_res / 0.1)                    # train.py:96-98 (lss_cfg.yaml xbound step 0.1)
    terrain_preproc = torch.nn.AvgPool2d(kernel_size=kernel_size, stride=kernel_size)

    g = torch.Generator().manual_seed(0)
    xs = torch.arange(-6.4, 6.4, 0.1)
    X, Y = torch.meshgrid(xs, xs, indexing='ij')
    terrain = {'terrain': torch.stack([0.3 * torch.exp(-((X - 1.5 - 0.3 * i) ** 2 + (Y + 0.2 * i) ** 2) / 3) for i in range(bsz)])[:, None],
               'friction': 0.4 + 0.5 * torch.rand(bsz, 1, 128, 128, generator=g)}
    terrain = {k: v.to(device).requires_grad_(True) for k, v in terrain.items()}
    pose0 = torch.eye(4).repeat(bsz, 1, 1)
    yaw = torch.linspace(-0.5, 0.5, bsz)
    pose0[:, 0, 0], pose0[:, 0, 1], pose0[:, 1, 0], pose0[:, 1, 1] = yaw.cos(), -yaw.sin(), yaw.sin(), yaw.cos()
    pose0[:, :2, 3] = 0.2 * torch.randn(bsz, 2, generator=g)
    T = int(dphys_cfg.traj_sim_time / dphys_cfg.dt)
    controls = torch.stack([0.4 + 0.4 * torch.rand(bsz, generator=g), 0.6 * (torch.rand(bsz, generator=g) - 0.5)], -1)
    controls = controls[:, None].repeat(1, T, 1)
    control_ts = torch.arange(T)[None].repeat(bsz, 1) * dphys_cfg.dt
    traj_ts = torch.sort(torch.rand(bsz, 40, generator=g) * 4.5)[0]
    Xs = torch.cumsum(0.01 * torch.randn(bsz, 40, 3, generator=g), dim=1)

    # --- train.py:231-246 ---
    terrain_ = {}
    for k, v in terrain.items():
        terrain_[k] = terrain_preproc(v)
    x0 = pose0[:, :3, 3].to(device)
    xd0 = torch.zeros_like(x0)
    R0 = pose0[:, :3, :3].to(device)
    omega0 = torch.zeros_like(xd0)
    state0 = (x0, xd0, R0, omega0)
    states_pred, _ = dphysics(z_grid=terrain_['terrain'].squeeze(1), state=state0,
                              controls=controls.to(device),
                              friction=terrain_['friction'].squeeze(1))
    # --- train.py:402-406 ---
    states_gt = [Xs.to(device), None, None, None]
    loss_phys = physics_loss(states_pred=states_pred, states_gt=states_gt,
                             pred_ts=control_ts.to(device), gt_ts=traj_ts.to(device))
    loss_phys.backward()
    return dict(loss=np.asarray(loss_phys.item()), Xs=states_pred[0].detach().cpu().numpy(),
                Rs=states_pred[2].detach().cpu().numpy(), x0z=x0[:, 2].detach().cpu().numpy(),
                g_terrain=terrain['terrain'].grad.cpu().numpy(), g_friction=terrain['friction'].grad.cpu().numpy())


if __name__ == '__main__':
    which, out = sys.argv[1], sys.argv[2]
    dev = sys.argv[3] if len(sys.argv) > 3 else 'cpu'
    res = fit_terrain() if which == 'fit_terrain' else predict_states(dev)
    np.savez_compressed(out, **res)
    import monoforce.models.traj_predictor.dphysics as m
    print('monoforce provided by', m.__file__)
