"""GPU parity of the tcgen05 implicit-GEMM convolution (K4) against torch's fp32 conv on bf16-rounded operands."""
import pytest
import torch
import torch.nn.functional as F

from helpers_mfb import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _reference(x, w, scale, shift, act):
    # x (N,H,W,Cin) bf16, w (Cout,KS,KS,Cin) bf16 -> fp32 conv of the SAME rounded operands
    y = F.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), padding=w.shape[1] // 2)
    y = y * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    y = {0: lambda t: t, 1: torch.relu, 2: lambda t: F.gelu(t)}[act](y)
    return y.permute(0, 2, 3, 1)


@pytest.mark.parametrize("N,H,W,Cin,Cout,KS,act", [
    (1, 8, 16, 64, 128, 1, 0),        # single tile, single K chunk
    (1, 8, 16, 64, 128, 3, 0),        # 3x3 taps with zero padding on all sides
    (2, 16, 32, 128, 128, 3, 2),      # several tiles / chunks, GELU epilogue
    (1, 16, 26, 448, 512, 3, 2),      # CamEncode.up1 conv at the default 16x26 feature map (432 ch padded to 448), ragged width
    (2, 64, 64, 320, 256, 3, 2),      # BevEncode.up1 first conv
    (1, 128, 128, 256, 128, 3, 2),    # BevEncode head conv at full BEV resolution
    (1, 16, 26, 512, 128, 1, 0),      # depthnet-like 1x1 (Cout padded)
    (1, 12, 20, 64, 64, 3, 1),        # BLOCK_N = 64 path, ragged in both directions, ReLU
])
def test_conv_bn_act_matches_torch(N, H, W, Cin, Cout, KS, act):
    from monoforce_b200.ops import conv_bn_act_nhwc
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator().manual_seed(N * 1000 + H + Cin + KS)
    x = torch.randn(N, H, W, Cin, generator=g).to(DEV).to(torch.bfloat16)
    w = (torch.randn(Cout, KS, KS, Cin, generator=g) / (KS * KS * Cin) ** 0.5).to(DEV).to(torch.bfloat16)
    scale = (0.5 + torch.rand(Cout, generator=g)).to(DEV)
    shift = (0.2 * torch.randn(Cout, generator=g)).to(DEV)
    y = conv_bn_act_nhwc(x, w, scale, shift, act)
    ref = _reference(x, w, scale, shift, act)
    torch.cuda.synchronize()
    assert y.shape == ref.shape
    # bf16 output rounding is 2^-9 relative; accumulation is fp32 in both
    assert rel_err(y.float(), ref) < 8e-3
    assert (y.float() - ref).abs().mean().item() < 2e-3 * ref.abs().mean().item() + 1e-4
