"""GPU tests of the encoder's inference kernels (K4 generalised tensor-core convolution, K7 memory-bound layers) against
their fp32 torch statements (tests/helpers_fast_emul.py) on bf16-rounded inputs, through the C ABI.

Tolerances: operands are bf16 (8 significant bits), accumulation fp32, outputs rounded to bf16: per element
|err| <= 2^-8 |y| + accumulated operand rounding; checked as max |err| <= 2e-2 * max|y| and mean |err| <= 4e-3 * mean|y|.
"""
import numpy as np
import pytest
import torch

import helpers_fast_emul as emul
from helpers_mfb import load_golden, rel_err
from helpers_lss import small_cfg, default_cfg, make_inputs, perturb_for_test

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _bf(t):
    return t.to(torch.bfloat16)


def _close(got, want, max_tol=2e-2, mean_tol=4e-3):
    got, want = got.float().cpu(), want.float().cpu()
    assert got.shape == want.shape, (got.shape, want.shape)
    err = (got - want).abs()
    assert err.max().item() <= max_tol * want.abs().max().item() + 1e-6, (err.max().item(), want.abs().max().item())
    assert err.mean().item() <= mean_tol * want.abs().mean().item() + 1e-6, (err.mean().item(), want.abs().mean().item())


CONV_CASES = [
    # name,            N, H,  W,  Cin, Cout, K, stride, pad, act,            extras
    ("3x3_s1_gelu",    2, 16, 26, 432, 512, 3, 1, 1, emul.ACT_GELU, {}),                 # camera Up.conv[0]: ragged tiles, Cin % 64 != 0
    ("3x3_s2_relu",    2, 32, 32, 64, 128, 3, 2, 1, emul.ACT_RELU, {}),                   # ResNet layer2.0.conv1
    ("1x1_s2_none",    2, 32, 32, 64, 128, 1, 2, 0, emul.ACT_NONE, {}),                   # ResNet downsample
    ("7x7_s2_relu",    1, 64, 64, 64, 64, 7, 2, 3, emul.ACT_RELU, {}),                    # BevEncode.conv1
    ("3x3_residual",   2, 24, 40, 128, 128, 3, 1, 1, emul.ACT_RELU, {"residual": True}),  # BasicBlock.conv2 + identity + ReLU
    ("1x1_expand",     3, 20, 28, 16, 96, 1, 1, 0, emul.ACT_SILU, {}),                    # MBConv expand: Cin 16 (one quarter of a K chunk)
    ("1x1_project",    3, 9, 13, 1152, 320, 1, 1, 0, emul.ACT_NONE, {"per_image": True}),            # MBConv project, SE folded per image
    ("1x1_project_skip", 2, 16, 26, 672, 112, 1, 1, 0, emul.ACT_NONE, {"per_image": True, "residual": True}),
    ("1x1_narrow_out", 2, 33, 17, 96, 24, 1, 1, 0, emul.ACT_NONE, {}),                    # Cout 24: one 64-column tile, 40 columns masked
    ("3x3_s2_odd",     1, 17, 31, 40, 72, 3, 2, 1, emul.ACT_RELU, {}),                    # odd sizes, Cout % 64 != 0
    # deep K, Cout % 256 == 0 and at least one tile per SM: the 128 x 256 tile instantiation (one CTA per SM)
    ("3x3_wide_gelu",  16, 32, 48, 192, 256, 3, 1, 1, emul.ACT_GELU, {}),
    ("3x3_wide_res",   12, 40, 40, 136, 512, 3, 1, 1, emul.ACT_RELU, {"residual": True}),  # ragged K chunks, ragged tiles, residual
]


@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv2d_tcgen05_matches_fp32_statement(case):
    from monoforce_b200 import ops
    name, N, H, W, Cin, Cout, K, stride, pad, act, ex = case
    g = torch.Generator().manual_seed(hash(name) % 1000)
    x = _bf(torch.randn(N, H, W, Cin, generator=g))
    wshape = (N, Cout, K, K, Cin) if ex.get("per_image") else (Cout, K, K, Cin)
    w = _bf(torch.randn(*wshape, generator=g) * (1.0 / np.sqrt(K * K * Cin)))
    scale = 0.5 + torch.rand(Cout, generator=g)
    shift = 0.2 * torch.randn(Cout, generator=g)
    Ho, Wo = emul.conv_out_size(H, K, stride, pad, pad), emul.conv_out_size(W, K, stride, pad, pad)
    res = _bf(torch.randn(N, Ho, Wo, Cout, generator=g)) if ex.get("residual") else None
    want = emul.conv2d_nhwc(x, w, scale, shift, act, stride=stride, pad=(pad, pad), out_hw=(Ho, Wo), residual=res)
    got = ops.conv2d_nhwc(x.to(DEV), w.to(DEV), scale.to(DEV), shift.to(DEV), act, stride=stride, pad=(pad, pad), out_hw=(Ho, Wo),
                          residual=None if res is None else res.to(DEV))
    assert got.dtype == torch.bfloat16
    _close(got, want)


def test_conv2d_asymmetric_static_same_padding():
    """TF 'SAME' at stride 2 pads one pixel more AFTER than before ((0,1) for k=3): only the low side is a parameter, the
    high side is the TMA unit's zero fill."""
    from monoforce_b200 import ops
    g = torch.Generator().manual_seed(5)
    x = _bf(torch.randn(2, 16, 26, 64, generator=g))
    w = _bf(torch.randn(64, 3, 3, 64, generator=g) / 24)
    one, zero = torch.ones(64), torch.zeros(64)
    want = emul.conv2d_nhwc(x, w, one, zero, emul.ACT_NONE, stride=2, pad=(0, 0), out_hw=(8, 13))
    got = ops.conv2d_nhwc(x.to(DEV), w.to(DEV), one.to(DEV), zero.to(DEV), ops.ACT_NONE, stride=2, pad=(0, 0), out_hw=(8, 13))
    _close(got, want)


def test_conv2d_fused_heads_epilogue():
    """Three 3x3 256 -> 128 convs + BN + GELU as one 256 -> 384 launch, each head's 1x1 conv + ScaledTanh / ReLU in the epilogue
    (lss.py:117-139): fp32 (N,3,H,W) out, the 384-channel tensor never exists."""
    from monoforce_b200 import ops
    g = torch.Generator().manual_seed(11)
    N, H, W = 2, 24, 40
    x = _bf(torch.randn(N, H, W, 256, generator=g))
    w = _bf(torch.randn(384, 3, 3, 256, generator=g) / 48)
    scale, shift = 0.5 + torch.rand(384, generator=g), 0.1 * torch.randn(384, generator=g)
    head_w = torch.randn(384, generator=g) / 11
    heads = (head_w, [0.05, 0.2, -0.1], [emul.HEAD_SCALED_TANH, emul.HEAD_RELU, emul.HEAD_RELU], [-1.0, 0.0, 0.0], [1.0, 0.0, 0.0])
    want = emul.conv2d_nhwc(x, w, scale, shift, emul.ACT_GELU, pad=(1, 1), heads=heads)
    got = ops.conv2d_nhwc(x.to(DEV), w.to(DEV), scale.to(DEV), shift.to(DEV), ops.ACT_GELU, pad=(1, 1),
                          heads=(head_w.to(DEV),) + heads[1:])
    assert got.dtype == torch.float32 and got.shape == (N, 3, H, W)
    _close(got, want, max_tol=1e-2, mean_tol=3e-3)
    assert (got[:, 1:] >= 0).all() and got[:, 0].abs().max() <= 1.0


def test_conv2d_rejects_bad_arguments():
    from monoforce_b200 import ops
    x = torch.zeros(1, 8, 8, 64, dtype=torch.bfloat16, device=DEV)
    w = torch.zeros(64, 3, 3, 64, dtype=torch.bfloat16, device=DEV)
    v = torch.zeros(64, device=DEV)
    with pytest.raises(RuntimeError, match="stride"):
        ops.conv2d_nhwc(x, w, v, v, 0, stride=3, pad=(1, 1), out_hw=(3, 3))
    with pytest.raises(RuntimeError, match="does not fit"):
        ops.conv2d_nhwc(x, w, v, v, 0, stride=1, pad=(1, 1), out_hw=(12, 8))
    with pytest.raises(RuntimeError, match="multiple of 8"):
        ops.conv2d_nhwc(x[..., :60].contiguous(), w[..., :60].contiguous(), v, v, 0, pad=(1, 1))


@pytest.mark.parametrize("C,K,stride,pad,H,W", [(32, 3, 1, (1, 1), 40, 52), (96, 3, 2, (0, 1), 40, 52), (144, 5, 2, (1, 2), 20, 26),
                                                (672, 5, 1, (2, 2), 16, 26), (1152, 3, 1, (1, 1), 8, 13), (240, 3, 2, (0, 1), 17, 9)])
def test_depthwise_conv_bn_swish_and_pool(C, K, stride, pad, H, W):
    from monoforce_b200 import ops
    g = torch.Generator().manual_seed(C + K)
    N = 3
    x = _bf(torch.randn(N, H, W, C, generator=g))
    w = torch.randn(K * K, C, generator=g) / K
    shift = 0.3 * torch.randn(C, generator=g)
    pool_want = torch.zeros(N, C)
    want = emul.dwconv_bn_silu(x, w, shift, K, stride, pad, pool_want)
    pool = torch.zeros(N, C, device=DEV)
    got = ops.dwconv_bn_silu(x.to(DEV), w.to(DEV), shift.to(DEV), K, stride, pad, pool)
    _close(got, want)
    # the pool sums the fp32 values before the bf16 rounding of the stored output
    assert rel_err(pool, got.float().sum((1, 2))) < 5e-3
    assert rel_err(pool, pool_want) < 5e-3
    again = ops.dwconv_bn_silu(x.to(DEV), w.to(DEV), shift.to(DEV), K, stride, pad, None)
    assert torch.equal(again, got)


def test_squeeze_excite_fold_into_projection_weights():
    from monoforce_b200 import ops
    g = torch.Generator().manual_seed(3)
    for N, C, Sq, Cout in ((4, 32, 8, 16), (3, 1152, 48, 320), (2, 240, 10, 80)):
        pool = torch.rand(N, C, generator=g) * 50
        wr, br = torch.randn(Sq, C, generator=g) / C ** 0.5, 0.1 * torch.randn(Sq, generator=g)
        we, be = torch.randn(Sq, C, generator=g) / Sq ** 0.5, 0.1 * torch.randn(C, generator=g)      # expand weights, transposed
        proj = _bf(torch.randn(Cout, C, generator=g))
        want = emul.se_fold(pool, 1 / 49.0, wr, br, we, be, proj)
        got = ops.se_fold(pool.to(DEV), 1 / 49.0, wr.to(DEV), br.to(DEV), we.to(DEV), be.to(DEV), proj.to(DEV))
        assert got.shape == (N, Cout, 1, 1, C) and got.dtype == torch.bfloat16
        _close(got, want, max_tol=1e-2, mean_tol=3e-3)


def test_stem_conv_and_upsample_concat_and_cast():
    from monoforce_b200 import ops
    g = torch.Generator().manual_seed(8)
    img = torch.randn(3, 3, 38, 50, generator=g)
    w = torch.randn(3, 3, 3, 32, generator=g) / 5
    shift = 0.2 * torch.randn(32, generator=g)
    _close(ops.stem_conv(img.to(DEV), w, shift, (0, 1)), emul.stem_conv(img, w, shift, (0, 1)))
    for scale, Cs, Cl, Cout in ((2, 112, 320, 432), (4, 64, 256, 320), (2, 0, 256, 256), (2, 16, 24, 64)):
        low = _bf(torch.randn(2, 8, 13, Cl, generator=g))
        skip = _bf(torch.randn(2, 8 * scale, 13 * scale, Cs, generator=g)) if Cs else None
        want = emul.upsample_concat_nhwc(skip, low, (8 * scale, 13 * scale), Cout)
        got = ops.upsample_concat_nhwc(None if skip is None else skip.to(DEV), low.to(DEV), (8 * scale, 13 * scale), Cout)
        _close(got, want, max_tol=1e-2, mean_tol=3e-3)
        if Cs:
            assert torch.equal(got[..., :Cs].cpu(), skip)                 # the skip half is a pure copy
    # odd ratios, a 1-pixel source, and an output SMALLER than the source (cells that own no output pixel): every output pixel is
    # still written exactly once by the per-source-cell kernel
    for (Hl, Wl), (H, W) in (((5, 7), (13, 9)), ((1, 1), (6, 4)), ((9, 12), (4, 5)), ((3, 2), (3, 2)), ((7, 5), (1, 11))):
        low = _bf(torch.randn(2, Hl, Wl, 24, generator=g))
        skip = _bf(torch.randn(2, H, W, 8, generator=g))
        got = ops.upsample_concat_nhwc(skip.to(DEV), low.to(DEV), (H, W), 40)
        _close(got, emul.upsample_concat_nhwc(skip, low, (H, W), 40), max_tol=1e-2, mean_tol=3e-3)
        assert torch.equal(got[..., :8].cpu(), skip) and float(got[..., 32:].abs().max()) == 0.0
    x = torch.randn(5, 7, 64, generator=g)
    assert torch.equal(ops.cast_bf16(x.to(DEV)).cpu(), x.to(torch.bfloat16))


def test_lift_splat_on_bf16_logits_matches_fp32_kernel():
    from monoforce_b200 import ops
    from monoforce_b200.terrain_encoder import LiftSplatShoot, _LiftSplat
    gc, ac = small_cfg()
    net = LiftSplatShoot(gc, ac)
    B, N, D, C, fH, fW = 2, 4, net.D, net.camC, 8, 12
    g = torch.Generator().manual_seed(4)
    _, *calib = make_inputs(gc, ac, B, 9)
    vox = net.voxel_index(net.get_geometry(*calib)).to(DEV)
    logits = _bf(torch.randn(B * N, fH, fW, 128, generator=g)).to(DEV)
    want = _LiftSplat.apply(logits[..., :D + C].float().contiguous(), vox.view(-1), B, N, D, C, 64, 64)
    got = ops.lift_splat_bf16(logits, vox.view(-1), B, N, D, C, 64, 64)
    assert rel_err(got, want) < 1e-5


def _net(cfg_fn, seed=0):
    from monoforce_b200.terrain_encoder import LiftSplatShoot
    gc, ac = cfg_fn()
    torch.manual_seed(seed)
    return perturb_for_test(LiftSplatShoot(gc, ac)).eval().to(DEV), gc, ac


def _fast(net, inputs):
    net.fast_inference = True
    with torch.no_grad():
        out = net(*inputs)
    net.fast_inference = False
    return out


def test_fast_path_stages_match_their_fp32_statement():
    """Stage by stage on the GPU (bf16 kernels) vs the SAME host logic driven by the fp32 emulation: trunk endpoints,
    camera Up, BEV backbone."""
    import importlib
    from monoforce_b200 import encoder_fast
    net, gc, ac = _net(small_cfg)
    x, *calib = make_inputs(gc, ac, 2, 6)
    B, N = x.shape[:2]
    imgs = x.view(B * N, *x.shape[2:])
    with torch.no_grad():
        P = encoder_fast.prepare(net)
        f16, f32 = encoder_fast.trunk_endpoints(P, imgs.to(DEV))
        up = encoder_fast.up_block(P["cam_up"], f16, f32, 2)
        xb = _bf(torch.randn(2, 64, 64, 64))
        g1, g3 = encoder_fast.bev_backbone(P, xb.to(DEV))
        # the same functions with every kernel replaced by its fp32 statement, on the CPU copy of the prepared weights
        cpu_net = net.__class__(gc, ac)
        cpu_net.load_state_dict({k: v.cpu() for k, v in net.state_dict().items()})
        cpu_net.eval()
        old = encoder_fast.ops
        encoder_fast.ops = emul
        try:
            Pc = encoder_fast.prepare(cpu_net)
            e16, e32 = encoder_fast.trunk_endpoints(Pc, imgs)
            eup = encoder_fast.up_block(Pc["cam_up"], e16, e32, 2)
            h1, h3 = encoder_fast.bev_backbone(Pc, xb.float())
        finally:
            encoder_fast.ops = old
    for name, got, want, tol in (("reduction_4", f16, e16, 0.05), ("reduction_5", f32, e32, 0.05), ("cam_up", up, eup, 0.05),
                                 ("layer1", g1, h1, 0.03), ("layer3", g3, h3, 0.05)):
        err = (got.float().cpu() - want).abs()
        print(name, "max rel", (err.max() / want.abs().max()).item(), "mean rel", (err.mean() / want.abs().mean()).item())
        assert err.max() <= tol * want.abs().max() and err.mean() <= 0.3 * tol * want.abs().mean(), name


@pytest.mark.parametrize("which", ["small", "lss_cfg.yaml"])
def test_fast_inference_whole_network_vs_reference_golden(which):
    """The whole network on repo kernels (bf16 tensor-core path) vs the UNMODIFIED reference on the CPU (goldens minted by
    tests/golden/make_golden_lss.py): the small fixture (2 scenes) and the reference's own lss_cfg.yaml sizes
    (4 cameras 256x416 -> 128x128 BEV, 1 scene)."""
    from monoforce_b200 import _lib
    if which == "small":
        net, gc, ac = _net(small_cfg)
        inputs = [t.to(DEV) for t in make_inputs(gc, ac, 2, 1)]
        g = load_golden("lss_small_eval_B2")
    else:
        net, gc, ac = _net(default_cfg)
        inputs = [t.to(DEV) for t in make_inputs(gc, ac, 1, 3)]
        g = load_golden("lss_default_eval_B1")
    n0 = _lib.kernel_launches()
    out = _fast(net, inputs)
    launched = _lib.kernel_launches() - n0
    assert launched >= 80, launched             # 1 stem + 16 x (<=4) MBConv + 4 + 1 + 1 + 1 + 15 + 3 + 2 kernels: the repo's, not cuDNN's
    for k in ("geom", "terrain", "diff", "friction"):
        got, want = out[k].float().cpu().numpy(), g[k]
        assert got.shape == want.shape
        err = np.abs(got - want)
        print(which, k, "max abs", err.max(), "mean abs", err.mean(), "ref mean abs", np.abs(want).mean())
        # bf16 operands through ~60 layers: a few percent of the output scale at worst, well under a percent on average
        assert err.max() < 0.15 * max(np.abs(want).max(), 0.1) and err.mean() < 0.02 * np.abs(want).mean() + 2e-3, k
    # with grad enabled (training / fine-tuning) the fp32 autograd path is used regardless of the flag
    net.fast_inference = True
    assert net(*inputs)["geom"].requires_grad


def test_fast_path_follows_weight_updates_and_calibration_cache():
    from monoforce_b200 import encoder_fast
    net, gc, ac = _net(small_cfg)
    inputs = [t.to(DEV) for t in make_inputs(gc, ac, 2, 2)]
    a = _fast(net, inputs)
    with torch.no_grad():
        net.bevencode.up_friction[4].bias.add_(0.5)                     # e.g. an optimizer step / load_state_dict
    b = _fast(net, inputs)
    # (fp32 atomics in the lift-splat and in the squeeze-excite pool make two runs agree to rounding, not bit for bit)
    assert (b["friction"] - a["friction"]).mean().item() > 0.3 and torch.allclose(a["geom"], b["geom"], atol=2e-3)
    # the voxel index is cached per calibration: same tensors -> the same object; rebuilt tensors with equal values -> a hit too
    v1 = net.cached_voxel_index(*inputs[1:])
    assert net.cached_voxel_index(*inputs[1:]) is v1
    assert net.cached_voxel_index(*[t.clone() for t in inputs[1:]]) is v1
    moved = [t.clone() for t in inputs[1:]]
    moved[1][..., 0] += 0.7
    v2 = net.cached_voxel_index(*moved)
    assert v2 is not v1 and not torch.equal(v1, v2)
    assert torch.equal(v2, net.voxel_index(net.get_geometry(*moved)))


def test_terrain_and_path_postprocessing_kernels():
    """F4: terrain = geom - diff + AvgPool2d(k) of the physics inputs in one pass (lss.py:158, train.py:96-99,234-235); poses and
    the inclination cost of the planner (monoforce_node.py:80-85, diff_physics.py:262-266 via scipy's Euler angles)."""
    from scipy.spatial.transform import Rotation
    from monoforce_b200 import ops
    g = torch.Generator().manual_seed(21)
    heads = torch.randn(3, 3, 50, 36, generator=g).to(DEV)
    geom, diff, fric = heads[:, 0:1], heads[:, 1:2], heads[:, 2:3]
    for k in (1, 2, 4, 7):
        terrain, zp, mp = ops.terrain_postproc(geom, diff, fric, k)
        assert torch.equal(terrain, geom - diff)
        assert torch.allclose(zp, torch.nn.functional.avg_pool2d(geom - diff, k), atol=1e-6)
        assert torch.allclose(mp, torch.nn.functional.avg_pool2d(fric, k), atol=1e-6)
    B, T = 5, 77
    rot = Rotation.random(B * T, random_state=3)
    Rs = torch.as_tensor(rot.as_matrix(), dtype=torch.float32).view(B, T, 3, 3)
    Xs = torch.randn(B, T, 3, generator=g)
    poses, cost = ops.path_postproc(Xs.to(DEV), Rs.to(DEV))
    rpy = torch.as_tensor(Rotation.from_matrix(Rs.view(-1, 3, 3).numpy()).as_euler('xyz'))
    want = rpy[:, 0].reshape(B, -1).abs().mean(-1) + rpy[:, 1].reshape(B, -1).abs().mean(-1)
    assert torch.allclose(cost.cpu().double(), want, rtol=1e-4, atol=1e-5)
    ref = torch.zeros(B, T, 4, 4)
    ref[:, :, :3, 3], ref[:, :, :3, :3], ref[:, :, 3, 3] = Xs, Rs, 1.0
    assert torch.equal(poses.cpu(), ref)


def test_fast_path_cuda_graph_replay_matches_eager_launches():
    """fast_graph: the inference launches captured once and replayed; same numbers (up to the atomics' summation order), new
    images are picked up, a weight update or a new calibration re-captures."""
    net, gc, ac = _net(small_cfg)
    a_in = [t.to(DEV) for t in make_inputs(gc, ac, 2, 2)]
    b_in = [t.to(DEV) for t in make_inputs(gc, ac, 2, 5)]
    eager_a, eager_b = _fast(net, a_in), _fast(net, b_in[:1] + a_in[1:])
    net.fast_graph = True
    g_a = _fast(net, a_in)
    graph = net._mfb_graph["graph"]
    g_b = _fast(net, b_in[:1] + a_in[1:])
    assert net._mfb_graph["graph"] is graph                                   # replayed, not re-captured
    for k in ("geom", "terrain", "diff", "friction"):
        assert torch.allclose(g_a[k], eager_a[k], atol=2e-3) and torch.allclose(g_b[k], eager_b[k], atol=2e-3), k
        assert not torch.allclose(g_a[k], g_b[k], atol=1e-4)
    with torch.no_grad():
        net.bevencode.up_diff[4].bias.add_(0.25)
    g_c = _fast(net, a_in)
    assert net._mfb_graph["graph"] is not graph and (g_c["diff"] - g_a["diff"]).mean().item() > 0.1
    g_d = _fast(net, b_in)                                                     # other calibration -> other voxel index -> new graph
    net.fast_graph = False
    eager_d = _fast(net, b_in)
    assert torch.allclose(g_d["terrain"], eager_d["terrain"], atol=2e-3)


def test_conv2d_random_shapes_fuzz():
    """40 random small problems (any Cin / Cout multiple of 8, K in 1..7, stride 1 / 2, every low-side padding that fits,
    ragged spatial sizes down to 1 x 1, all activations, residual on / off, per-image weights on / off) against the fp32 statement:
    the tile walk, TMA zero fill (spatial and channel), the transposed / direct store paths and the 64 / 128-column variants all
    get hit by construction."""
    from monoforce_b200 import ops
    rng = np.random.RandomState(1234)
    for case in range(40):
        K = int(rng.choice([1, 1, 2, 3, 3, 5, 7]))
        stride = int(rng.choice([1, 1, 2]))
        pad = int(rng.randint(0, K))
        Cin, Cout = 8 * int(rng.randint(1, 24)), 8 * int(rng.randint(1, 24))
        N = int(rng.randint(1, 4))
        H, W = int(rng.randint(max(1, K - 2 * pad), 40)), int(rng.randint(max(1, K - 2 * pad), 40))
        Ho, Wo = emul.conv_out_size(H, K, stride, pad, pad), emul.conv_out_size(W, K, stride, pad, pad)
        if Ho < 1 or Wo < 1:
            continue
        act = int(rng.randint(0, 4))
        per_image, with_res = bool(rng.randint(0, 2)) and K == 1, bool(rng.randint(0, 2))
        g = torch.Generator().manual_seed(case)
        x = _bf(torch.randn(N, H, W, Cin, generator=g))
        w = _bf(torch.randn(*((N,) if per_image else ()), Cout, K, K, Cin, generator=g) / np.sqrt(K * K * Cin))
        scale, shift = 0.5 + torch.rand(Cout, generator=g), 0.2 * torch.randn(Cout, generator=g)
        res = _bf(torch.randn(N, Ho, Wo, Cout, generator=g)) if with_res else None
        want = emul.conv2d_nhwc(x, w, scale, shift, act, stride=stride, pad=(pad, pad), out_hw=(Ho, Wo), residual=res)
        got = ops.conv2d_nhwc(x.to(DEV), w.to(DEV), scale.to(DEV), shift.to(DEV), act, stride=stride, pad=(pad, pad), out_hw=(Ho, Wo),
                              residual=None if res is None else res.to(DEV))
        err = (got.float().cpu() - want).abs()
        ok = err.max() <= 3e-2 * want.abs().max() + 1e-3
        assert ok, (case, dict(K=K, stride=stride, pad=pad, Cin=Cin, Cout=Cout, N=N, H=H, W=W, act=act, per_image=per_image, res=with_res),
                    err.max().item(), want.abs().max().item())


def test_depthwise_random_shapes_fuzz():
    from monoforce_b200 import ops
    rng = np.random.RandomState(99)
    for case in range(24):
        K = int(rng.choice([3, 5]))
        stride = int(rng.choice([1, 2]))
        pad = ((K - 1) // 2, (K - 1) // 2) if stride == 1 else ((0, 1) if K == 3 else (1, 2))
        C = 8 * int(rng.randint(1, 40))
        N, H, W = int(rng.randint(1, 4)), int(rng.randint(K, 45)), int(rng.randint(K, 45))
        g = torch.Generator().manual_seed(1000 + case)
        x = _bf(torch.randn(N, H, W, C, generator=g))
        w, shift = torch.randn(K * K, C, generator=g) / K, 0.3 * torch.randn(C, generator=g)
        pool_want = torch.zeros(N, C)
        want = emul.dwconv_bn_silu(x, w, shift, K, stride, pad, pool_want)
        pool = torch.zeros(N, C, device=DEV)
        got = ops.dwconv_bn_silu(x.to(DEV), w.to(DEV), shift.to(DEV), K, stride, pad, pool)
        assert got.shape == want.shape, (case, got.shape, want.shape)
        err = (got.float().cpu() - want).abs()
        assert err.max() <= 2e-2 * want.abs().max() + 1e-3, (case, K, stride, C, N, H, W, err.max().item())
        assert rel_err(pool, pool_want) < 1e-2, (case, K, stride, C, N, H, W)
