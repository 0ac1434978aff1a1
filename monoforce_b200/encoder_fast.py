"""Inference path of the terrain encoder on repo kernels only: NHWC bf16 from the camera images to the BEV heads.

What `LiftSplatShoot.forward` (terrain_encoder/lss.py:282-291) does in eval mode, layer for layer, with every eval-mode
BatchNorm folded into its convolution (host, once per set of weights) and nothing in between going through the framework:

    images (B*N,3,H,W) fp32
      -> stem 3x3/2 + BN + swish                                   K7 stem_conv          (efficientnet_pytorch `_conv_stem`, lss.py:78)
      -> 16 MBConv blocks:  1x1 expand + BN + swish                 K4 (tcgen05)          (lss.py:83-90)
                            depthwise kxk + BN + swish + SE pool    K7 dwconv
                            SE MLP folded into per-image weights    K7 se_fold
                            1x1 project + BN (+ skip)               K4, per-image weights
      -> Up(320 + 112 -> 512): upsample x2 + concat                 K7 upsample_concat    (lss.py:27-46,92-94)
                               2 x (3x3 + BN + GELU)                K4
      -> depthnet 1x1 (D + C logits)                                K4                    (lss.py:58)
      -> depth soft-max (x) features, voxel pooling                 K5 lift_splat (bf16 logits)   (lss.py:63-71,238-280)
      -> cast to bf16                                               K7 cast
      -> conv1 7x7/2 + BN + ReLU, ResNet-18 layer1..3               K4 (stride 2, residual + ReLU epilogue)   (lss.py:104-116,140-151)
      -> Up(64 + 256 -> 256, x4)                                    K7 + K4
      -> x2 upsample; 3 heads: 3x3 + BN + GELU, 1x1, ScaledTanh | ReLU   K7 + ONE K4 launch with the fused head epilogue (lss.py:117-139)
      -> terrain = geom - diff                                      (lss.py:158)

The only framework ops left are allocations, one memset of the squeeze-excite pool buffer and the final subtraction on
(B,1,X,Y).  Weights are prepared by `prepare()` and cached on the module keyed on `params_stamp` (storage + version of every
parameter / buffer), so `load_state_dict` / an optimizer step rebuild them.
"""
from __future__ import annotations

import torch
from torch import nn

from . import ops
from .efficientnet import params_stamp


def _bn_fold(bn, conv_bias, cout, device):
    """(scale, shift) fp32 of an eval-mode BatchNorm (or of a plain bias when bn is None)."""
    bias = conv_bias.detach().float() if conv_bias is not None else torch.zeros(cout, device=device)
    if bn is None:
        return torch.ones(cout, device=device), bias
    scale = bn.weight.detach().float() * torch.rsqrt(bn.running_var.detach().float() + bn.eps)
    return scale, bn.bias.detach().float() + (bias - bn.running_mean.detach().float()) * scale


WDTYPE = torch.bfloat16      # operand dtype of the tensor-core convolutions (the CPU unit tests of prepare() use float32)


def _conv_w(conv, pad_cout_to=None):
    """(Cout,KH,KW,Cin) bf16 K-major weights; BatchNorm stays an fp32 (scale, shift) epilogue."""
    w = conv.weight.detach().permute(0, 2, 3, 1).to(WDTYPE)
    if pad_cout_to is not None and w.shape[0] < pad_cout_to:
        w = torch.cat([w, torch.zeros(pad_cout_to - w.shape[0], *w.shape[1:], dtype=w.dtype, device=w.device)])
    return w.contiguous()


def _pad_vec(v, n, fill):
    return v if v.shape[0] == n else torch.cat([v, torch.full((n - v.shape[0],), fill, device=v.device)])


def _conv_spec(conv, bn, act, pad_cout_to=None):
    cout = conv.out_channels if pad_cout_to is None else pad_cout_to
    scale, shift = _bn_fold(bn, conv.bias, conv.out_channels, conv.weight.device)
    return dict(w=_conv_w(conv, pad_cout_to), scale=_pad_vec(scale, cout, 1.0).contiguous(), shift=_pad_vec(shift, cout, 0.0).contiguous(),
                act=act, stride=conv.stride[0], k=conv.kernel_size[0], pad=conv.padding[0] if isinstance(conv.padding, tuple) else 0)


def _fold_pixels(spec, cin, cout):
    """1x1 convolution with a NARROW input (Cin < 64): measured on B200, the TMA unit moves 32-byte activation rows (Cin = 16)
    at less than half the rate of 128-byte rows, so the layer ran at 1.6 TB/s instead of 4.2.  A 1x1 convolution over [pixels, Cin]
    is the same matrix product as one over [pixels / f, f * Cin] with the weights repeated block-diagonally (f x f blocks): the
    views are free (NHWC rows of f consecutive pixels are contiguous), the rows become >= 128 bytes, and the f-fold redundant
    MMA work is irrelevant for a memory-bound layer."""
    f = next((k for k in (2, 4) if k * cin >= 64), 4)
    if cin >= 64 or spec["k"] != 1 or spec["stride"] != 1 or f * cout > 1152:
        return
    w = spec["w"]                                               # (Cout, 1, 1, Cin)
    wf = torch.zeros(f * cout, 1, 1, f * cin, dtype=w.dtype, device=w.device)
    for j in range(f):
        wf[j * cout:(j + 1) * cout, :, :, j * cin:(j + 1) * cin] = w
    spec["fold"] = (f, wf.contiguous(), spec["scale"].repeat(f).contiguous(), spec["shift"].repeat(f).contiguous())


def _static_pad(conv):
    """(low, high) zero padding of an efficientnet_pytorch Conv2dStaticSamePadding."""
    p = conv.static_padding
    if isinstance(p, nn.ZeroPad2d):
        l, r, t, b = p.padding
        assert (l, r) == (t, b), "square static padding expected"
        return l, r
    return 0, 0


def prepare(net):
    """Folded inference weights of a LiftSplatShoot module (cached; rebuilt when any parameter / buffer changed)."""
    stamp = (params_stamp(net), WDTYPE)
    cache = net.__dict__.get("_mfb_fast")
    if cache is not None and cache["stamp"] == stamp:
        return cache
    with torch.no_grad():
        cam, bev = net.camencode, net.bevencode
        t = cam.trunk
        dev = t._conv_stem.weight.device
        P = {"stamp": stamp, "XY": (int(net.nx[0]), int(net.nx[1]))}       # host copies: no device read on the launch path
        # --- stem: (3,3,3,32) fp32 [dy][dx][ci][co] with the BN scale folded into the weights
        s_scale, s_shift = _bn_fold(t._bn0, None, 32, dev)
        assert t._conv_stem.in_channels == 3 and t._conv_stem.out_channels == 32
        # (host copies: the stem kernel takes its 3.5 KB of weights by value, as constant-bank operands)
        P["stem_w"] = (t._conv_stem.weight.detach().float() * s_scale.view(-1, 1, 1, 1)).permute(2, 3, 1, 0).contiguous().cpu()
        P["stem_shift"] = s_shift.contiguous().cpu()
        P["stem_pad"] = _static_pad(t._conv_stem)
        # --- MBConv blocks
        blocks = []
        for blk in t._blocks:
            a = blk._block_args
            oup = a.input_filters * a.expand_ratio
            b = {"k": a.kernel_size, "stride": blk._depthwise_conv.stride[0], "dw_pad": _static_pad(blk._depthwise_conv),
                 "skip": bool(a.id_skip and a.stride == 1 and a.input_filters == a.output_filters), "oup": oup}
            if a.expand_ratio != 1:
                b["expand"] = _conv_spec(blk._expand_conv, blk._bn0, ops.ACT_SILU)
                _fold_pixels(b["expand"], a.input_filters, oup)
            d_scale, d_shift = _bn_fold(blk._bn1, None, oup, dev)
            b["dw_w"] = (blk._depthwise_conv.weight.detach().float()[:, 0] * d_scale.view(-1, 1, 1)).permute(1, 2, 0).reshape(-1, oup).contiguous()
            b["dw_shift"] = d_shift.contiguous()
            sq = blk._se_reduce.out_channels
            b["se"] = (blk._se_reduce.weight.detach().float().view(sq, oup).contiguous(), blk._se_reduce.bias.detach().float().contiguous(),
                       blk._se_expand.weight.detach().float().view(oup, sq).t().contiguous(), blk._se_expand.bias.detach().float().contiguous())
            b["proj"] = _conv_spec(blk._project_conv, blk._bn2, ops.ACT_NONE)
            b["proj_w2d"] = b["proj"]["w"].view(blk._project_conv.out_channels, oup).contiguous()
            blocks.append(b)
        P["blocks"] = blocks
        P["pool_elems"] = sum(b["oup"] for b in blocks)
        # --- camera Up + depthnet (Cout = D + C = 123 is not a multiple of 8: weights / epilogue vectors padded to 128)
        P["cam_up"] = [_conv_spec(cam.up1.conv[0], cam.up1.conv[1], ops.ACT_GELU), _conv_spec(cam.up1.conv[3], cam.up1.conv[4], ops.ACT_GELU)]
        dn = cam.depthnet
        P["depthnet"] = _conv_spec(dn, None, ops.ACT_NONE, pad_cout_to=-(-dn.out_channels // 64) * 64)
        # --- BEV backbone
        P["bev_conv1"] = _conv_spec(bev.conv1, bev.bn1, ops.ACT_RELU)
        layers = []
        for layer in (bev.layer1, bev.layer2, bev.layer3):
            for blk in layer:
                layers.append({"c1": _conv_spec(blk.conv1, blk.bn1, ops.ACT_RELU), "c2": _conv_spec(blk.conv2, blk.bn2, ops.ACT_RELU),
                               "ds": _conv_spec(blk.downsample[0], blk.downsample[1], ops.ACT_NONE) if blk.downsample is not None else None})
        P["bev_layers"], P["n_layer1"] = layers, len(bev.layer1)
        P["bev_up"] = [_conv_spec(bev.up1.conv[0], bev.up1.conv[1], ops.ACT_GELU), _conv_spec(bev.up1.conv[3], bev.up1.conv[4], ops.ACT_GELU)]
        P["bev_up_scale"] = int(bev.up1.up.scale_factor)
        # --- heads: three 3x3 256 -> 128 convs as ONE 256 -> 384 launch; each head's 1x1 conv + output activation in its epilogue
        heads = (bev.up_geom, bev.up_diff, bev.up_friction)
        P["heads_fused"] = all(h[4].out_channels == 1 and h[1].out_channels == 128 for h in heads)
        specs = [_conv_spec(h[1], h[2], ops.ACT_GELU) for h in heads]
        P["heads"] = dict(w=torch.cat([s["w"] for s in specs]).contiguous(), scale=torch.cat([s["scale"] for s in specs]).contiguous(),
                          shift=torch.cat([s["shift"] for s in specs]).contiguous(), act=ops.ACT_GELU, stride=1, k=3, pad=1)
        P["heads_up_scale"] = int(heads[0][0].scale_factor)
        if P["heads_fused"]:
            head_act, lo, hi = [], [], []
            for h in heads:
                m = h[5]
                if isinstance(m, nn.ReLU):
                    head_act.append(ops.HEAD_RELU); lo.append(0.0); hi.append(0.0)
                elif hasattr(m, "min_val") and hasattr(m, "max_val"):
                    head_act.append(ops.HEAD_SCALED_TANH); lo.append(float(m.min_val)); hi.append(float(m.max_val))
                else:
                    P["heads_fused"] = False
            P["head_epilogue"] = (torch.cat([h[4].weight.detach().float().view(-1) for h in heads]).contiguous(),
                                  [float(h[4].bias.detach()) if h[4].bias is not None else 0.0 for h in heads], head_act, lo, hi)
    net.__dict__["_mfb_fast"] = P
    return P


def _conv(x, spec, residual=None, out_hw=None, pad=None, w=None):
    k = spec["k"]
    p = spec["pad"] if pad is None else pad
    fold = spec.get("fold")
    if fold is not None and w is None and residual is None and x.shape[2] % fold[0] == 0:
        f, wf, scale, shift = fold
        N, H, W, Cin = x.shape
        y = ops.conv2d_nhwc(x.reshape(N, H, W // f, f * Cin), wf, scale, shift, spec["act"])
        return y.reshape(N, H, W, y.shape[3] // f)
    return ops.conv2d_nhwc(x, spec["w"] if w is None else w, spec["scale"], spec["shift"], spec["act"], stride=spec["stride"],
                           pad=(p, p), out_hw=out_hw, residual=residual)


def _same_out(n, k, stride, pad):
    return (n + 2 * pad - k) // stride + 1


def trunk_endpoints(P, imgs):
    """EfficientNet-B0 trunk on (BN,3,H,W) fp32 images: the /16 (112 ch) and /32 (320 ch) feature maps, NHWC bf16."""
    x = ops.stem_conv(imgs.contiguous(), P["stem_w"], P["stem_shift"], P["stem_pad"])
    BN = x.shape[0]
    pools = torch.zeros(BN * P["pool_elems"], dtype=torch.float32, device=x.device)      # one memset for all 16 blocks
    off = 0
    feats, prev = [], x
    for b in P["blocks"]:
        inp = x
        if "expand" in b:
            x = _conv(x, b["expand"], pad=0)
        oup = b["oup"]
        pool = pools[off:off + BN * oup].view(BN, oup)
        off += BN * oup
        x = ops.dwconv_bn_silu(x, b["dw_w"], b["dw_shift"], b["k"], b["stride"], b["dw_pad"], pool)
        wn = ops.se_fold(pool, 1.0 / (x.shape[1] * x.shape[2]), *b["se"], b["proj_w2d"])
        if oup < 64 and not b["skip"] and x.shape[2] % 2 == 0:
            # narrow input rows again (block 0: 32 channels): two pixels per GEMM row, per-image weights block-diagonal (see _fold_pixels)
            N_, H_, W_, _ = x.shape
            cout = wn.shape[1]
            wf = torch.zeros(N_, 2 * cout, 1, 1, 2 * oup, dtype=wn.dtype, device=wn.device)
            wf[:, :cout, :, :, :oup] = wn
            wf[:, cout:, :, :, oup:] = wn
            pr = b["proj"]
            y = ops.conv2d_nhwc(x.reshape(N_, H_, W_ // 2, 2 * oup), wf, pr["scale"].repeat(2), pr["shift"].repeat(2), pr["act"])
            x = y.reshape(N_, H_, W_, cout)
        else:
            x = _conv(x, b["proj"], residual=inp if b["skip"] else None, pad=0, w=wn)
        if prev.shape[1] > x.shape[1]:
            feats.append(prev)
        prev = x
    feats.append(x)
    return feats[3], feats[4]


def up_block(specs, skip, low, scale):
    """Up.forward (lss.py:44-46) on NHWC bf16: skip at the output resolution, low at 1/scale of it."""
    H, W = low.shape[1] * scale, low.shape[2] * scale
    assert skip.shape[1:3] == (H, W), (tuple(skip.shape), tuple(low.shape), scale)
    x = ops.upsample_concat_nhwc(skip, low, (H, W), skip.shape[3] + low.shape[3])
    return _conv(_conv(x, specs[0]), specs[1])


def bev_backbone(P, x):
    """conv1 7x7/2 + BN + ReLU, layer1..3 of ResNet-18 (torchvision BasicBlock) on NHWC bf16 -> (layer1 out, layer3 out)."""
    c1 = P["bev_conv1"]
    x = _conv(x, c1, out_hw=(_same_out(x.shape[1], 7, 2, 3), _same_out(x.shape[2], 7, 2, 3)))
    x1 = None
    for i, blk in enumerate(P["bev_layers"]):
        s = blk["c1"]["stride"]
        ohw = (_same_out(x.shape[1], 3, s, 1), _same_out(x.shape[2], 3, s, 1))
        idt = x if blk["ds"] is None else _conv(x, blk["ds"], out_hw=ohw, pad=0)
        y = _conv(x, blk["c1"], out_hw=ohw)
        x = _conv(y, blk["c2"], residual=idt)
        if i == P["n_layer1"] - 1:
            x1 = x
    return x1, x


def forward(net, imgs, vox):
    """Eval-mode LiftSplatShoot.forward on repo kernels.  imgs (B,N,3,H,W) fp32 CUDA; vox from net.cached_voxel_index."""
    P = prepare(net)
    B, N, Cin, H, W = imgs.shape
    f16, f32 = trunk_endpoints(P, imgs.reshape(B * N, Cin, H, W).float())
    y = up_block(P["cam_up"], f16, f32, 2)
    logits = _conv(y, P["depthnet"], pad=0)                                   # (BN, fH, fW, 128) bf16, D + C used
    X, Y = P["XY"]
    bev = ops.lift_splat_bf16(logits, vox.view(-1), B, N, net.D, net.camC, X, Y)
    x1, x3 = bev_backbone(P, ops.cast_bf16(bev))
    y = up_block(P["bev_up"], x1, x3, P["bev_up_scale"])
    s = P["heads_up_scale"]
    up = ops.upsample_concat_nhwc(None, y, (y.shape[1] * s, y.shape[2] * s), y.shape[3])
    hs = P["heads"]
    if P["heads_fused"]:
        out = ops.conv2d_nhwc(up, hs["w"], hs["scale"], hs["shift"], hs["act"], pad=(1, 1), heads=P["head_epilogue"])
        geom, diff, friction = out[:, 0:1], out[:, 1:2], out[:, 2:3]
        return {'geom': geom, 'terrain': ops.terrain_postproc(geom, diff, friction, 1)[0], 'diff': diff, 'friction': friction}
    else:      # outC != 1: the 1x1 output convolutions go through torch on the (B,X,Y,384) tensor
        z = _conv(up, hs).permute(0, 3, 1, 2).float()
        heads = (net.bevencode.up_geom, net.bevencode.up_diff, net.bevencode.up_friction)
        geom, diff, friction = (h[5](h[4](z[:, 128 * i:128 * (i + 1)])) for i, h in enumerate(heads))
    return {'geom': geom, 'terrain': geom - diff, 'diff': diff, 'friction': friction}
