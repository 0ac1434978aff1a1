"""Thin tensor-level wrappers over the C ABI for the encoder's dense layers (inference path)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2


def conv_bn_act_nhwc(x: torch.Tensor, wgt: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor, act: int) -> torch.Tensor:
    """y = act(scale * conv_same(x, wgt) + shift) on the tcgen05 tensor cores (csrc/conv_tcgen05.cuh).

    x (N,H,W,Cin) bf16 contiguous, wgt (Cout,KS,KS,Cin) bf16 contiguous, scale/shift (Cout,) fp32 -> (N,H,W,Cout) bf16."""
    if not x.is_cuda:
        raise RuntimeError("conv_bn_act_nhwc runs on CUDA only (tcgen05 kernel, no CPU fallback)")
    assert x.dtype == torch.bfloat16 and wgt.dtype == torch.bfloat16 and x.is_contiguous() and wgt.is_contiguous()
    N, H, W, Cin = x.shape
    Cout, KS, KS2, Cin2 = wgt.shape
    assert KS == KS2 and Cin == Cin2, (x.shape, wgt.shape)
    y = torch.empty(N, H, W, Cout, dtype=torch.bfloat16, device=x.device)
    scale, shift = scale.float().contiguous(), shift.float().contiguous()
    lib = _lib.load()
    p = lambda t: C.c_void_p(t.data_ptr())
    with torch.cuda.device(x.device):
        st = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(lib.mfb_conv_bn_act_bf16(p(x), p(wgt), p(scale), p(shift), p(y), N, H, W, Cin, Cout, KS, act, C.c_void_p(st)),
                   "mfb_conv_bn_act_bf16")
    return y


def fold_conv_bn(conv: torch.nn.Conv2d, bn: torch.nn.BatchNorm2d | None, pad_cin_to: int = 64):
    """(wgt (Cout,KS,KS,Cin_padded) bf16, scale, shift) of Conv2d [+ eval-mode BatchNorm2d].

    BatchNorm stays an fp32 epilogue (scale / shift) instead of being multiplied into the bf16 weights."""
    w = conv.weight.detach()
    Cout, Cin, KS, _ = w.shape
    cin_p = (Cin + pad_cin_to - 1) // pad_cin_to * pad_cin_to
    wk = torch.zeros(Cout, KS, KS, cin_p, dtype=torch.bfloat16, device=w.device)
    wk[..., :Cin] = w.permute(0, 2, 3, 1).to(torch.bfloat16)
    bias = conv.bias.detach().float() if conv.bias is not None else torch.zeros(Cout, device=w.device)
    if bn is None:
        return wk.contiguous(), torch.ones(Cout, device=w.device), bias
    inv = torch.rsqrt(bn.running_var.detach().float() + bn.eps)
    scale = bn.weight.detach().float() * inv
    shift = bn.bias.detach().float() + (bias - bn.running_mean.detach().float()) * scale
    return wk.contiguous(), scale, shift


# ---------------------------------------------------------------------------------------------------------------
# general tensor-core convolution + the memory-bound NHWC bf16 layers (csrc/conv_tcgen05.cuh, csrc/encoder_ops.cu)
# ---------------------------------------------------------------------------------------------------------------
ACT_SILU = 3
HEAD_NONE, HEAD_RELU, HEAD_SCALED_TANH = 0, 1, 2


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _need_cuda(x, what):
    if not x.is_cuda:
        raise RuntimeError(f"{what} runs on CUDA only (sm_100a kernel, no CPU fallback)")


def conv_out_size(n, k, stride, pad_lo, pad_hi):
    return (n + pad_lo + pad_hi - k) // stride + 1


def conv2d_nhwc(x, wgt, scale, shift, act, *, stride=1, pad=(0, 0), out_hw=None, residual=None, heads=None):
    """K4 (mfb_conv2d_bf16).  x (N,H,W,Cin) bf16; wgt (Cout,KH,KW,Cin) or per-image (N,Cout,KH,KW,Cin) bf16; pad = low-side
    (pad_h, pad_w); out_hw defaults to the "same"-style size ceil(H / stride).  heads = (head_w (Cout,) fp32, bias list,
    act list, lo list, hi list) switches to the fused 1x1-head epilogue and returns (N, n_heads, Ho, Wo) fp32."""
    _need_cuda(x, "conv2d_nhwc")
    assert x.dtype == torch.bfloat16 and wgt.dtype == torch.bfloat16 and x.is_contiguous() and wgt.is_contiguous()
    N, H, W, Cin = x.shape
    per_image = wgt.dim() == 5
    Cout, KH, KW, Cin2 = wgt.shape[-4:]
    assert Cin == Cin2 and (not per_image or wgt.shape[0] == N), (tuple(x.shape), tuple(wgt.shape))
    Ho, Wo = out_hw if out_hw is not None else (-(-H // stride), -(-W // stride))
    d = _lib.ConvDesc(N=N, H=H, W=W, Cin=Cin, Ho=Ho, Wo=Wo, Cout=Cout, KH=KH, KW=KW, stride=stride, pad_h=pad[0], pad_w=pad[1],
                      act=act, per_image_weights=int(per_image), n_heads=0)
    y = head_w = head_out = None
    if heads is not None:
        head_w, bias, acts, lo, hi = heads
        d.n_heads = len(bias)
        for i in range(len(bias)):
            d.head_bias[i], d.head_act[i], d.head_lo[i], d.head_hi[i] = float(bias[i]), int(acts[i]), float(lo[i]), float(hi[i])
        head_out = torch.empty(N, len(bias), Ho, Wo, dtype=torch.float32, device=x.device)
    else:
        y = torch.empty(N, Ho, Wo, Cout, dtype=torch.bfloat16, device=x.device)
        if residual is not None:
            assert residual.shape == y.shape and residual.dtype == torch.bfloat16 and residual.is_contiguous()
    lib = _lib.load()
    with torch.cuda.device(x.device):
        _lib.check(lib.mfb_conv2d_bf16(C.byref(d), _p(x), _p(wgt), _p(scale), _p(shift), _p(residual), _p(y), _p(head_w),
                                       _p(head_out), _stream(x.device)), "mfb_conv2d_bf16")
    return head_out if heads is not None else y


def upsample_concat_nhwc(skip, low, out_hw, c_out):
    """[skip, bilinear_up(low) (align_corners=True), zero padding] along channels, NHWC bf16 (skip may be None)."""
    _need_cuda(low, "upsample_concat_nhwc")
    N, Hl, Wl, Cl = low.shape
    H, W = out_hw
    Cs = 0 if skip is None else skip.shape[3]
    assert low.dtype == torch.bfloat16 and low.is_contiguous() and (skip is None or (skip.is_contiguous() and skip.shape[:3] == (N, H, W)))
    out = torch.empty(N, H, W, c_out, dtype=torch.bfloat16, device=low.device)
    lib = _lib.load()
    with torch.cuda.device(low.device):
        _lib.check(lib.mfb_upsample_concat_nhwc_bf16(_p(skip), _p(low), _p(out), N, H, W, Cs, Hl, Wl, Cl, c_out, _stream(low.device)),
                   "mfb_upsample_concat_nhwc_bf16")
    return out


def stem_conv(img, w, shift, pad):
    """EfficientNet stem: img (N,3,H,W) fp32 -> (N,Ho,Wo,32) bf16, 3x3 stride 2, low-side pad, BN folded, swish.
    w (3,3,3,32) and shift (32,) are HOST fp32 tensors (passed to the kernel by value)."""
    _need_cuda(img, "stem_conv")
    assert img.dtype == torch.float32 and img.is_contiguous() and img.shape[1] == 3
    assert not w.is_cuda and not shift.is_cuda and w.is_contiguous() and w.numel() == 864 and shift.numel() == 32
    N, _, H, W = img.shape
    lo, hi = pad
    Ho, Wo = conv_out_size(H, 3, 2, lo, hi), conv_out_size(W, 3, 2, lo, hi)
    y = torch.empty(N, Ho, Wo, 32, dtype=torch.bfloat16, device=img.device)
    lib = _lib.load()
    with torch.cuda.device(img.device):
        _lib.check(lib.mfb_stem_conv_bf16(_p(img), _p(w), _p(shift), _p(y), N, H, W, Ho, Wo, lo, lo, _stream(img.device)), "mfb_stem_conv_bf16")
    return y


def dwconv_bn_silu(x, w, shift, K, stride, pad, pool=None):
    """Depthwise KxK + folded BN + swish, NHWC bf16; `pool` (N,C) fp32 accumulates the spatial sum of the output."""
    _need_cuda(x, "dwconv_bn_silu")
    assert x.dtype == torch.bfloat16 and x.is_contiguous()
    N, H, W, Cc = x.shape
    lo, hi = pad
    Ho, Wo = conv_out_size(H, K, stride, lo, hi), conv_out_size(W, K, stride, lo, hi)
    y = torch.empty(N, Ho, Wo, Cc, dtype=torch.bfloat16, device=x.device)
    lib = _lib.load()
    with torch.cuda.device(x.device):
        _lib.check(lib.mfb_dwconv_bn_silu_bf16(_p(x), _p(w), _p(shift), _p(y), _p(pool), N, H, W, Cc, Ho, Wo, K, stride, lo, lo,
                                               _stream(x.device)), "mfb_dwconv_bn_silu_bf16")
    return y


def se_fold(pool, inv_hw, w_reduce, b_reduce, w_expand_t, b_expand, proj_w):
    """Per-image projection matrices with the squeeze-excite scale folded in: (N, Cout, 1, 1, C) bf16.  w_reduce (Sq, C) and
    w_expand_t (Sq, C) = the TRANSPOSE of the expand weights (coalesced reads); `pool` is overwritten with the excite scale."""
    _need_cuda(pool, "se_fold")
    N, Cc = pool.shape
    Cout = proj_w.shape[0]
    Sq, Cse = w_reduce.shape
    out = torch.empty(N, Cout, 1, 1, Cc, dtype=torch.bfloat16, device=pool.device)
    lib = _lib.load()
    with torch.cuda.device(pool.device):
        assert w_expand_t.shape == w_reduce.shape
        _lib.check(lib.mfb_se_fold_bf16(_p(pool), float(inv_hw), _p(w_reduce), _p(b_reduce), _p(w_expand_t), _p(b_expand), _p(proj_w),
                                        _p(out), N, Cc, Cse, Sq, Cout, _stream(pool.device)), "mfb_se_fold_bf16")
    return out


def cast_bf16(x):
    _need_cuda(x, "cast_bf16")
    assert x.dtype == torch.float32 and x.is_contiguous() and x.numel() % 8 == 0
    y = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    lib = _lib.load()
    with torch.cuda.device(x.device):
        _lib.check(lib.mfb_cast_f32_to_bf16(_p(x), _p(y), x.numel(), _stream(x.device)), "mfb_cast_f32_to_bf16")
    return y


def lift_splat_bf16(logits, vox, B, N, D, Cc, X, Y):
    """K5 forward on bf16 logits rows (BN, fH, fW, row_stride >= D + C) -> bev (B, X, Y, C) fp32 (inference only)."""
    _need_cuda(logits, "lift_splat_bf16")
    assert logits.dtype == torch.bfloat16 and logits.is_contiguous()
    BN, fH, fW, rs = logits.shape
    bev = torch.zeros(B, X, Y, Cc, dtype=torch.float32, device=logits.device)
    lib = _lib.load()
    with torch.cuda.device(logits.device):
        _lib.check(lib.mfb_lift_splat_forward_bf16(_p(logits), rs, _p(vox), _p(bev), B, N, D, Cc, fH, fW, X, Y, _stream(logits.device)),
                   "mfb_lift_splat_forward_bf16")
    return bev


def terrain_postproc(geom, diff, friction, pool=1, want_terrain=True):
    """terrain = geom - diff (lss.py:158) and AvgPool2d(pool) of (terrain, friction) (train.py:96-99,234-235) in one kernel.
    geom / diff / friction: (B,1,X,Y) fp32 CUDA (channel slices of one tensor are fine).  Returns (terrain | None,
    terrain_pooled (B,1,X/pool,Y/pool), friction_pooled)."""
    _need_cuda(geom, "terrain_postproc")
    B, one, X, Y = geom.shape
    assert one == 1 and diff.shape == geom.shape == friction.shape and geom.dtype == torch.float32
    bs = {t.stride(0) for t in (geom, diff, friction)}
    ok = len(bs) == 1 and all(t.stride()[2:] == (Y, 1) for t in (geom, diff, friction))
    if not ok:
        geom, diff, friction = geom.contiguous(), diff.contiguous(), friction.contiguous()
    terrain = torch.empty(B, 1, X, Y, dtype=torch.float32, device=geom.device) if want_terrain else None
    zp = torch.empty(B, 1, X // pool, Y // pool, dtype=torch.float32, device=geom.device)
    mp = torch.empty_like(zp)
    lib = _lib.load()
    with torch.cuda.device(geom.device):
        _lib.check(lib.mfb_terrain_postproc(_p(geom), _p(diff), _p(friction), geom.stride(0), _p(terrain), _p(zp), _p(mp), B, X, Y, pool,
                                            _stream(geom.device)), "mfb_terrain_postproc")
    return terrain, zp, mp


def path_postproc(Xs, Rs, want_poses=True):
    """(poses (B,T,4,4) | None, inclination cost (B,)) of a rollout: monoforce_node.py:80-85, diff_physics.py:262-266."""
    _need_cuda(Xs, "path_postproc")
    B, T, _ = Xs.shape
    Xs, Rs = Xs.detach().float().contiguous(), Rs.detach().float().contiguous()
    poses = torch.empty(B, T, 4, 4, dtype=torch.float32, device=Xs.device) if want_poses else None
    cost = torch.empty(B, dtype=torch.float32, device=Xs.device)
    lib = _lib.load()
    with torch.cuda.device(Xs.device):
        _lib.check(lib.mfb_path_postproc(_p(Xs), _p(Rs), _p(poses), _p(cost), B, T, _stream(Xs.device)), "mfb_path_postproc")
    return poses, cost
