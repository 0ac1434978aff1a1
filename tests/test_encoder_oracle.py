"""CPU tests for the terrain encoder: module surface, state_dict compatibility, oracle vs reference."""
import numpy as np
import pytest
import torch

from helpers_mfb import load_golden
from helpers_lss import small_cfg, make_inputs, perturb_for_test


def test_state_dict_layout_matches_reference_checkpoints():
    """Key names / shapes a released `val.pth` would carry (lss.py:177-186, efficientnet_pytorch 0.7.1 layout)."""
    from monoforce_b200.terrain_encoder import LiftSplatShoot
    grid_conf, aug_conf = small_cfg()
    net = LiftSplatShoot(grid_conf, aug_conf)
    sd = net.state_dict()
    assert len(sd) == 504
    for key, shape in {"dx": (3,), "bx": (3,), "nx": (3,), "frustum": (59, 8, 12, 3),
                       "camencode.trunk._conv_stem.weight": (32, 3, 3, 3),
                       "camencode.trunk._blocks.0._depthwise_conv.weight": (32, 1, 3, 3),
                       "camencode.trunk._blocks.1._expand_conv.weight": (96, 16, 1, 1),
                       "camencode.trunk._blocks.15._project_conv.weight": (320, 1152, 1, 1),
                       "camencode.trunk._blocks.5._se_reduce.bias": (10,),
                       "camencode.trunk._fc.weight": (1000, 1280),
                       "camencode.up1.conv.0.weight": (512, 432, 3, 3), "camencode.up1.conv.4.running_var": (512,),
                       "camencode.depthnet.weight": (59 + 64, 512, 1, 1),
                       "bevencode.conv1.weight": (64, 64, 7, 7), "bevencode.layer3.1.bn2.weight": (256,),
                       "bevencode.up1.conv.3.weight": (256, 256, 3, 3), "bevencode.up_geom.1.weight": (128, 256, 3, 3),
                       "bevencode.up_friction.4.bias": (1,)}.items():
        assert tuple(sd[key].shape) == shape, key
    from monoforce_b200.efficientnet import EfficientNet
    assert sum(p.numel() for p in EfficientNet.from_name().parameters()) == 5_288_548       # EfficientNet-B0
    assert float(net.dx[0]) == pytest.approx(0.2) and int(net.nx[0]) == 64 and net.D == 59
    with pytest.raises(RuntimeError, match="CUDA only"):
        net.eval()(*make_inputs(grid_conf, aug_conf, 1, 0))


def test_lift_splat_oracle_matches_reference_when_present():
    """Build container only: oracle/lss_oracle.lift_splat == reference get_depth_feat + voxel_pooling."""
    from oracle.ref_import import reference_available
    if not reference_available():
        pytest.skip("reference tree not present on this machine")
    import sys, os
    from helpers_mfb import ROOT
    for p in (os.path.join(ROOT, "oracle", "shims"), "/root/reference/monoforce/src"):
        if p not in sys.path:
            sys.path.insert(0, p)
    from monoforce.models.terrain_encoder.lss import LiftSplatShoot as Ref
    from oracle.lss_oracle import lift_splat
    grid_conf, aug_conf = small_cfg()
    torch.manual_seed(0)
    ref = Ref(grid_conf, aug_conf).eval()
    x, *calib = make_inputs(grid_conf, aug_conf, 2, 5)
    with torch.no_grad():
        geom = ref.get_geometry(*calib)
        want = ref.voxel_pooling(geom, ref.get_cam_feats(x))
        B, N = x.shape[:2]
        logits = ref.camencode.depthnet(ref.camencode.get_eff_depth(x.view(B * N, *x.shape[2:])))
        got = lift_splat(logits, geom, ref.dx, ref.bx, ref.nx, ref.D, ref.camC)
    # the reference sums by differencing one global cumsum (utils.py:144-152): ~1e-6 absolute cancellation noise
    assert torch.allclose(got, want, rtol=1e-4, atol=1e-5)


def test_network_restatement_matches_reference_golden_up_to_the_fused_stage():
    """Everything except the CUDA-only lift-splat runs on CPU: same seed -> same weights; trunk + Up + depthnet logits,
    frustum geometry and voxel indices feed the ORACLE lift-splat, then our BevEncode must reproduce the reference's
    outputs (golden minted from the unmodified reference)."""
    from monoforce_b200.terrain_encoder import LiftSplatShoot
    from oracle.lss_oracle import lift_splat
    g = load_golden("lss_small_eval_B2")
    grid_conf, aug_conf = small_cfg()
    torch.manual_seed(0)
    net = perturb_for_test(LiftSplatShoot(grid_conf, aug_conf)).eval()
    x, *calib = make_inputs(grid_conf, aug_conf, 2, 1)
    with torch.no_grad():
        geom = net.get_geometry(*calib)
        B, N = x.shape[:2]
        logits = net.camencode.depth_logits_and_feats(x.view(B * N, *x.shape[2:]))
        bev = lift_splat(logits, geom, net.dx, net.bx, net.nx, net.D, net.camC)
        out = net.bevencode(bev)
        # voxel index helper agrees with the oracle's in-grid filter
        vox = net.voxel_index(geom)
    assert np.allclose(bev[:, 0].numpy(), g["bev_ch0"], rtol=1e-4, atol=1e-5)
    assert np.allclose(bev.sum(dim=1).numpy(), g["bev_sum"], rtol=1e-4, atol=1e-4)
    for k in ("geom", "terrain", "diff", "friction"):
        assert np.allclose(out[k].numpy(), g[k], rtol=1e-4, atol=1e-5), k
    assert vox.dtype == torch.int32 and int((vox >= 0).sum()) > 0 and int(vox.max()) < 64 * 64


def test_efficientnet_restatement_is_structurally_torchvision_b0():
    """The trunk internals cannot be pinned against efficientnet_pytorch 0.7.1 (absent from /root/reference and from this
    image), so they are anchored against an INDEPENDENT implementation of the same published architecture:
    torchvision.models.efficientnet_b0.  Weights are copied across key by key (which proves every block's kernel size,
    stride, expansion, squeeze-excite width and channel count), then the two networks must produce the same five
    endpoint feature maps.  The only modelled difference between the two packages is the padding convention of the four
    stride-2 layers (TF 'SAME' = (0,1)/(1,2) here, symmetric in torchvision); for this comparison our stride-2 pads are
    switched to symmetric, and the 'SAME' rule itself is checked against its closed form below."""
    from functools import partial
    import torchvision
    from monoforce_b200.efficientnet import EfficientNet, _same_pad
    torch.manual_seed(0)
    tv = torchvision.models.efficientnet_b0(weights=None, norm_layer=partial(torch.nn.BatchNorm2d, eps=1e-3, momentum=0.01)).eval()
    ours = EfficientNet.from_name("efficientnet-b0").eval()
    assert sum(p.numel() for p in tv.parameters()) == sum(p.numel() for p in ours.parameters()) == 5_288_548
    with torch.no_grad():
        for m in tv.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.3); m.running_var.uniform_(0.5, 1.5); m.weight.uniform_(0.5, 1.5); m.bias.normal_(0, 0.2)

    def cp(dst, src):
        assert {k: tuple(v.shape) for k, v in dst.state_dict().items()} == {k: tuple(v.shape) for k, v in src.state_dict().items()}
        dst.load_state_dict(src.state_dict())
    cp(ours._conv_stem, tv.features[0][0]); cp(ours._bn0, tv.features[0][1])
    tv_blocks = [b for stage in tv.features[1:8] for b in stage]
    assert len(tv_blocks) == len(ours._blocks) == 16
    for mine, theirs in zip(ours._blocks, tv_blocks):
        parts = list(theirs.block)
        a = mine._block_args
        if a.expand_ratio != 1:
            e = parts.pop(0)
            cp(mine._expand_conv, e[0]); cp(mine._bn0, e[1])
            assert isinstance(e[2], torch.nn.SiLU)
        dw, se, pr = parts
        assert dw[0].kernel_size == (a.kernel_size,) * 2 and dw[0].stride == (a.stride,) * 2 and dw[0].groups == dw[0].in_channels
        cp(mine._depthwise_conv, dw[0]); cp(mine._bn1, dw[1])
        cp(mine._se_reduce, se.fc1); cp(mine._se_expand, se.fc2)
        assert isinstance(se.activation, torch.nn.SiLU) and isinstance(se.scale_activation, torch.nn.Sigmoid)
        cp(mine._project_conv, pr[0]); cp(mine._bn2, pr[1])
        assert len(pr) == 2                                          # no activation after the projection
        assert theirs.use_res_connect == (a.id_skip and a.stride == 1 and a.input_filters == a.output_filters)
    cp(ours._conv_head, tv.features[8][0]); cp(ours._bn1, tv.features[8][1])
    # TF 'SAME' at the nominal 224 input (what Conv2dStaticSamePadding bakes in): stride-1 symmetric, stride-2 one more
    # pixel after than before
    assert _same_pad(224, 3, 2) == (0, 1) and _same_pad(112, 3, 1) == (1, 1) and _same_pad(56, 5, 2) == (1, 2)
    assert _same_pad(28, 3, 2) == (0, 1) and _same_pad(14, 5, 1) == (2, 2) and _same_pad(14, 5, 2) == (1, 2)
    n_sym = 0
    for m in ours.modules():
        if hasattr(m, "static_padding") and isinstance(m.static_padding, torch.nn.ZeroPad2d):
            l, r, t, b = m.static_padding.padding
            if l != r:
                assert m.stride == (2, 2) and (l, r) == (t, b) == ((0, 1) if m.kernel_size[0] == 3 else (1, 2))
                p = (m.kernel_size[0] - 1) // 2
                m.static_padding = torch.nn.ZeroPad2d((p, p, p, p))
                n_sym += 1
    assert n_sym == 5                                                # stem + the four stride-2 depthwise convs
    x = torch.randn(2, 3, 96, 128)
    with torch.no_grad():
        y = ours._swish(ours._bn0(ours._conv_stem(x)))
        z = tv.features[0](x)
        assert torch.allclose(y, z, rtol=1e-4, atol=1e-5)
        for mine, theirs in zip(ours._blocks, tv_blocks):
            y, z = mine(y), theirs(z)
            assert y.shape == z.shape and torch.allclose(y, z, rtol=1e-4, atol=1e-4), mine._block_args
        assert torch.allclose(ours._swish(ours._bn1(ours._conv_head(y))), tv.features[8](z), rtol=1e-4, atol=1e-4)


def test_fast_path_host_logic_matches_module_path_in_fp32(monkeypatch):
    """encoder_fast.prepare() + forward() with every kernel replaced by its fp32 torch statement (helpers_fast_emul) must
    reproduce the module path: BatchNorm folding, squeeze-excite folded into per-image projection weights, static-same
    padding, strides and output sizes of the ResNet layers, residual placement, concat order of the Up blocks and the
    fused 1x1 heads.  (The kernels themselves are compared with the same emulation on the GPU, tests/test_encoder_gpu.py.)"""
    import helpers_fast_emul as emul
    from monoforce_b200 import LiftSplatShoot, encoder_fast
    monkeypatch.setattr(encoder_fast, "ops", emul)
    monkeypatch.setattr(encoder_fast, "WDTYPE", torch.float32)
    gc, ac = small_cfg()
    torch.manual_seed(0)
    net = perturb_for_test(LiftSplatShoot(gc, ac)).eval()
    x, *calib = make_inputs(gc, ac, 2, 1)
    B, N = x.shape[:2]
    rel = lambda a, b: ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()
    with torch.no_grad():
        P = encoder_fast.prepare(net)
        assert encoder_fast.prepare(net) is P                                   # cached
        # trunk endpoints
        t = net.camencode.trunk
        y = t._swish(t._bn0(t._conv_stem(x.view(B * N, *x.shape[2:]))))
        feats, prev = [], y
        for blk in t._blocks:
            y = blk(y)
            if prev.size(2) > y.size(2):
                feats.append(prev)
            prev = y
        feats.append(y)
        f16, f32 = encoder_fast.trunk_endpoints(P, x.view(B * N, *x.shape[2:]))
        assert rel(f16.permute(0, 3, 1, 2), feats[3]) < 1e-4 and rel(f32.permute(0, 3, 1, 2), feats[4]) < 1e-4
        # camera Up + depthnet
        up = encoder_fast.up_block(P["cam_up"], f16, f32, 2)
        assert rel(up.permute(0, 3, 1, 2), net.camencode.up1(feats[4], feats[3])) < 1e-4
        # BEV backbone
        be = net.bevencode
        xb = torch.randn(2, be.conv1.in_channels, 64, 64)
        x1 = be.layer1(be.relu(be.bn1(be.conv1(xb)))); x3 = be.layer3(be.layer2(x1))
        g1, g3 = encoder_fast.bev_backbone(P, xb.permute(0, 2, 3, 1).contiguous())
        assert rel(g1.permute(0, 3, 1, 2), x1) < 1e-4 and rel(g3.permute(0, 3, 1, 2), x3) < 1e-4
        # whole network vs the reference golden (the lift-splat emulation stands in for K5)
        vox = net.voxel_index(net.get_geometry(*calib))
        out = encoder_fast.forward(net, x, vox)
        g = load_golden("lss_small_eval_B2")
        for k in ("geom", "terrain", "diff", "friction"):
            assert out[k].shape == g[k].shape and np.allclose(out[k].numpy(), g[k], rtol=1e-3, atol=1e-4), k
        # an in-place weight update must rebuild the folded weights
        be.conv1.weight.mul_(1.1)
        assert encoder_fast.prepare(net) is not P


def test_upsample_cell_ownership_partitions_the_output():
    """The up-sample kernel works per SOURCE cell and finds the output pixels a cell owns with `first_dst`
    (monoforce_b200/csrc/encoder_ops.cu); restated here in float32: for every (n_in, n_out) each output index belongs to exactly
    one cell and that cell is the forward mapping src = min(int(r * dst), n_in - 1) of torch's align_corners=True bilinear."""
    import numpy as np
    f32 = np.float32

    def src_index(dst, r, n_in):
        return min(int(f32(r) * f32(dst)), n_in - 1)

    def first_dst(s, r, n_out, n_in):
        if s <= 0:
            return 0
        if r <= 0:
            return n_out
        d = min(max(int(np.ceil(f32(s) / f32(r))), 0), n_out)
        while d > 0 and src_index(d - 1, r, n_in) >= s:
            d -= 1
        while d < n_out and src_index(d, r, n_in) < s:
            d += 1
        return d
    for n_in in list(range(1, 40)) + [64, 128, 129]:
        for n_out in list(range(1, 80)) + [128, 256, 257]:
            r = f32(n_in - 1) / f32(n_out - 1) if n_out > 1 else f32(0)
            owner = [-1] * n_out
            for s in range(n_in):
                lo = first_dst(s, r, n_out, n_in)
                hi = first_dst(s + 1, r, n_out, n_in) if s + 1 < n_in else n_out
                for d in range(lo, hi):
                    assert owner[d] == -1, (n_in, n_out, d)
                    owner[d] = s
            assert owner == [src_index(d, r, n_in) for d in range(n_out)], (n_in, n_out)
