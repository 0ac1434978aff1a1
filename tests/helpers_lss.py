"""Shared encoder test fixtures: a small LSS configuration and reproducible synthetic inputs."""
import math
import torch


def small_cfg():
    """Same structure as monoforce/config/lss_cfg.yaml, smaller image (128x192 -> 8x12 feature map) and a
    64x64 BEV grid (0.2 m) so the CPU reference runs in seconds."""
    grid_conf = {"xbound": [-6.4, 6.4, 0.2], "ybound": [-6.4, 6.4, 0.2], "zbound": [-3.2, 3.2, 6.4], "dbound": [0.6, 6.4, 0.1]}
    aug_conf = {"final_dim": [128, 192], "H": 1200, "W": 1920, "rand_flip": False, "bot_pct_lim": [0.0, 0.0],
                "resize_lim": [0.193, 0.225], "rot_lim": [-5.4, 5.4]}
    return grid_conf, aug_conf


def default_cfg():
    """monoforce/config/lss_cfg.yaml:1-36."""
    grid_conf = {"xbound": [-6.4, 6.4, 0.1], "ybound": [-6.4, 6.4, 0.1], "zbound": [-3.2, 3.2, 6.4], "dbound": [0.6, 6.4, 0.1]}
    aug_conf = {"final_dim": [256, 416], "H": 1200, "W": 1920, "rand_flip": False, "bot_pct_lim": [0.0, 0.0],
                "resize_lim": [0.193, 0.225], "rot_lim": [-5.4, 5.4]}
    return grid_conf, aug_conf


def make_inputs(grid_conf, aug_conf, B, seed, n_cams=4):
    """imgs ~ N(0,1); four cameras yawed 0/90/180/270 deg, pitched 20 deg down, 0.5 m above the origin
    (SURVEY.md 8d recipe); small per-sample augmentation in post_rots / post_trans."""
    g = torch.Generator().manual_seed(seed)
    H, W = aug_conf["final_dim"]
    imgs = torch.randn(B, n_cams, 3, H, W, generator=g)
    f = 0.6 * W
    K = torch.tensor([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1.]]).repeat(B, n_cams, 1, 1)
    rots = torch.zeros(B, n_cams, 3, 3)
    # camera frame: z forward, x right, y down -> ego frame: x forward, y left, z up
    base = torch.tensor([[0., 0., 1.], [-1., 0., 0.], [0., -1., 0.]])
    pitch = math.radians(20.0)
    Rp = torch.tensor([[1., 0, 0], [0, math.cos(pitch), -math.sin(pitch)], [0, math.sin(pitch), math.cos(pitch)]])
    for n in range(n_cams):
        a = n * math.pi / 2
        Rz = torch.tensor([[math.cos(a), -math.sin(a), 0], [math.sin(a), math.cos(a), 0], [0, 0, 1.]])
        rots[:, n] = Rz @ base @ Rp
    trans = torch.zeros(B, n_cams, 3)
    trans[..., 2] = 0.5
    trans[..., :2] = 0.1 * torch.randn(B, n_cams, 2, generator=g)
    ang = 0.05 * torch.randn(B, n_cams, generator=g)
    post_rots = torch.eye(3).repeat(B, n_cams, 1, 1)
    post_rots[..., 0, 0] = ang.cos(); post_rots[..., 0, 1] = -ang.sin()
    post_rots[..., 1, 0] = ang.sin(); post_rots[..., 1, 1] = ang.cos()
    post_trans = torch.zeros(B, n_cams, 3)
    post_trans[..., :2] = 3.0 * torch.randn(B, n_cams, 2, generator=g)
    return imgs, rots, trans, K, post_rots, post_trans


def perturb_for_test(net, seed=3):
    """Random-init weights leave the ReLU heads dead and every BatchNorm at identity statistics; give the
    running statistics, the residual-branch BN gains (zero-initialised by resnet18) and the head biases
    non-trivial values so that eval-mode parity is informative.  Deterministic; applied identically to the
    reference network and to ours."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.copy_(0.1 * torch.randn(m.running_mean.shape, generator=g))
                m.running_var.copy_(0.5 + torch.rand(m.running_var.shape, generator=g))
                m.weight.copy_(0.5 + torch.rand(m.weight.shape, generator=g))
                m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
        for head in (net.bevencode.up_diff, net.bevencode.up_friction):
            head[4].bias.fill_(0.2)
    return net
