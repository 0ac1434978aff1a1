"""GPU parity tests for the terrain encoder (fused lift-splat kernel through the C ABI)."""
import numpy as np
import pytest
import torch

from helpers_mfb import load_golden, rel_err
from helpers_lss import small_cfg, make_inputs, perturb_for_test

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(autouse=True)
def _exact_fp32_convs():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _net():
    from monoforce_b200.terrain_encoder import LiftSplatShoot
    grid_conf, aug_conf = small_cfg()
    torch.manual_seed(0)
    return perturb_for_test(LiftSplatShoot(grid_conf, aug_conf)).eval(), grid_conf, aug_conf


def test_fused_lift_splat_matches_oracle_forward_and_backward():
    """K5 vs the CPU restatement of soft-max (x) features + voxel pooling, values and gradients."""
    from monoforce_b200.terrain_encoder import _LiftSplat
    from oracle.lss_oracle import lift_splat
    net, grid_conf, aug_conf = _net()
    g = torch.Generator().manual_seed(7)
    B, N, D, C, fH, fW = 3, 4, net.D, net.camC, 8, 12
    _, *calib = make_inputs(grid_conf, aug_conf, B, 9)
    geom = net.get_geometry(*calib)
    logits = torch.randn(B * N, D + C, fH, fW, generator=g)
    w = torch.randn(B, C, 64, 64, generator=g)
    lr = logits.clone().requires_grad_(True)
    ref = lift_splat(lr, geom, net.dx, net.bx, net.nx, D, C)
    (ref * w).sum().backward()
    lk = logits.to(DEV).requires_grad_(True)
    vox = net.voxel_index(geom).to(DEV)
    bev = _LiftSplat.apply(lk.permute(0, 2, 3, 1), vox.view(-1), B, N, D, C, 64, 64).permute(0, 3, 1, 2)
    (bev * w.to(DEV)).sum().backward()
    assert rel_err(bev, ref) < 1e-5
    assert rel_err(lk.grad, lr.grad) < 1e-5
    # points outside the grid contribute nothing; total mass is conserved for the kept ones
    depth = logits[:, :D].softmax(1)
    kept = (vox.cpu() >= 0).view(B * N, D, fH, fW)
    want_mass = (depth * kept).unsqueeze(1) * logits[:, D:].unsqueeze(2)
    assert rel_err(bev.sum(dim=(2, 3)).cpu(), want_mass.view(B, N, C, D, fH, fW).sum(dim=(1, 3, 4, 5))) < 1e-4


def test_encoder_forward_matches_reference_golden():
    """Whole network on the GPU (cuDNN fp32 convs + fused lift-splat) vs the unmodified reference on CPU."""
    g = load_golden("lss_small_eval_B2")
    net, grid_conf, aug_conf = _net()
    net = net.to(DEV)
    inputs = [t.to(DEV) for t in make_inputs(grid_conf, aug_conf, 2, 1)]
    with torch.no_grad():
        out = net(*inputs)
        bev = net.get_voxels(*inputs)
    assert rel_err(bev[:, 0], g["bev_ch0"]) < 1e-4
    # a frustum point within an ulp of a voxel border may fall into the neighbouring BEV cell when the 3x3
    # inverses of get_geometry are evaluated on the GPU instead of the CPU (truncating index, lss.py:246):
    # allow a handful of cells to move, hold all others tight
    d = (bev.sum(dim=1).cpu().double() - torch.from_numpy(g["bev_sum"]).double()).abs()
    # (the 64-channel sum cancels heavily, so its error is measured against 64 x the per-channel magnitude)
    scale = 64 * float(np.abs(g["bev_ch0"]).max())
    assert int((d > 1e-4 * scale).sum()) <= 8 and float(d.max()) < 2e-2 * scale
    for k in ("geom", "terrain", "diff", "friction"):
        assert out[k].shape == (2, 1, 64, 64)
        assert rel_err(out[k], g[k]) < 1e-3, k


def test_encoder_backward_reaches_the_images():
    g = load_golden("lss_small_eval_B2")
    net, grid_conf, aug_conf = _net()
    net = net.to(DEV)
    x, *calib = [t.to(DEV) for t in make_inputs(grid_conf, aug_conf, 2, 1)]
    x.requires_grad_(True)
    o = net(x, *calib)
    gen = torch.Generator().manual_seed(2)
    w = {k: torch.randn(v.shape, generator=gen).to(DEV) for k, v in o.items()}
    sum((o[k] * w[k]).sum() for k in ("geom", "diff", "friction")).backward()
    assert rel_err(x.grad[:, 3, :, ::4, ::4], g["g_x_cam3"]) < 5e-3


def test_encoder_feeds_rollout_end_to_end():
    """BASELINE config 4 in miniature: encoder -> terrain / friction maps -> DPhysics rollout -> backward into the encoder."""
    from monoforce_b200 import DPhysics, DPhysConfig
    net, grid_conf, aug_conf = _net()
    net = net.to(DEV)
    inputs = [t.to(DEV) for t in make_inputs(grid_conf, aug_conf, 2, 4)]
    out = net(*inputs)
    cfg = DPhysConfig(robot="tradr", grid_res=0.2)
    cfg.traj_sim_time, cfg.use_odeint = 0.5, True          # the callers' default integrator (train.py:439-440)
    sim = DPhysics(cfg, device=DEV)
    T = 50
    controls = torch.tensor([[[0.8, 0.3]] * T, [[0.5, -0.4]] * T], device=DEV)
    states, forces = sim(out["terrain"].squeeze(1), controls, friction=out["friction"].squeeze(1))
    assert states[0].shape == (2, T, 3) and torch.isfinite(states[0]).all()
    states[0].pow(2).mean().backward()
    gsum = sum(float(p.grad.abs().sum()) for p in net.bevencode.up_geom.parameters() if p.grad is not None)
    assert gsum > 0 and np.isfinite(gsum)


def test_fast_inference_path_close_to_fp32_path():
    """bf16 tcgen05 path for the dense layers vs the fp32 cuDNN path of the same network (eval mode)."""
    net, grid_conf, aug_conf = _net()
    net = net.to(DEV)
    inputs = [t.to(DEV) for t in make_inputs(grid_conf, aug_conf, 2, 6)]
    with torch.no_grad():
        ref = net(*inputs)
        net.fast_inference = True
        fast = net(*inputs)
        net.fast_inference = False
    for k in ("geom", "terrain", "diff", "friction"):
        assert fast[k].shape == ref[k].shape
        assert rel_err(fast[k], ref[k]) < 5e-2, k
        assert (fast[k] - ref[k]).abs().mean().item() < 1e-2 * ref[k].abs().mean().item() + 1e-3, k
    # with grad enabled (training / fine-tuning) the fp32 autograd path is used regardless of the flag
    net.fast_inference = True
    out = net(*inputs)
    assert out["geom"].requires_grad
