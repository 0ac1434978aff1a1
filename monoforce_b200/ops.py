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
