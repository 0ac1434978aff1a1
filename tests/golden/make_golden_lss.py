#!/usr/bin/env python
"""Mint encoder goldens by running the UNMODIFIED reference LiftSplatShoot on CPU (build container only).

The reference's `efficientnet_pytorch` import resolves to oracle/shims/efficientnet_pytorch (random-init trunk
restatement, see there).  Inputs and weights are reproducible from seeds (torch CPU generators), so only the
outputs are stored:   python tests/golden/make_golden_lss.py
"""
import os
import sys
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle", "shims"), "/root/reference/monoforce/src"]
from helpers_lss import small_cfg, make_inputs, perturb_for_test  # noqa: E402


def main():
    from monoforce.models.terrain_encoder.lss import LiftSplatShoot
    grid_conf, aug_conf = small_cfg()
    torch.manual_seed(0)
    net = perturb_for_test(LiftSplatShoot(grid_conf, aug_conf)).eval()
    inputs = make_inputs(grid_conf, aug_conf, B=2, seed=1)
    with torch.no_grad():
        geom = net.get_geometry(*inputs[1:])
        cam = net.get_cam_feats(inputs[0])
        bev = net.voxel_pooling(geom, cam)
        out = net.bevencode(bev)
    # gradient of a fixed linear objective w.r.t. the images' last camera, through everything (train-free: eval mode)
    x = inputs[0].clone().requires_grad_(True)
    o = net(x, *inputs[1:])
    g = torch.Generator().manual_seed(2)
    w = {k: torch.randn(v.shape, generator=g) for k, v in o.items()}
    sum((o[k] * w[k]).sum() for k in ("geom", "diff", "friction")).backward()
    path = os.path.join(HERE, "lss_small_eval_B2.npz")
    np.savez_compressed(path, bev_sum=bev.sum(dim=1).numpy(), bev_abs=float(bev.abs().sum()), bev_ch0=bev[:, 0].numpy(),
                        g_x_cam3=x.grad[:, 3, :, ::4, ::4].numpy(), **{k: v.numpy() for k, v in out.items()})
    print({k: (tuple(v.shape), float(v.abs().mean())) for k, v in out.items()}, "bev", tuple(bev.shape), "->",
          os.path.getsize(path), "B")


def default_size():
    """lss_cfg.yaml sizes (4 cameras 256x416 -> 128x128 BEV), one scene: the reference's outputs only (inputs / weights are
    reproducible from seeds)."""
    from monoforce.models.terrain_encoder.lss import LiftSplatShoot
    from helpers_lss import default_cfg
    grid_conf, aug_conf = default_cfg()
    torch.manual_seed(0)
    net = perturb_for_test(LiftSplatShoot(grid_conf, aug_conf)).eval()
    inputs = make_inputs(grid_conf, aug_conf, B=1, seed=3)
    with torch.no_grad():
        out = net(*inputs)
    path = os.path.join(HERE, "lss_default_eval_B1.npz")
    np.savez_compressed(path, **{k: v.numpy().astype(np.float32) for k, v in out.items()})
    print({k: (tuple(v.shape), float(v.abs().mean())) for k, v in out.items()}, "->", os.path.getsize(path), "B")


if __name__ == "__main__":
    main()
    default_size()
