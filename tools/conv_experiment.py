"""Where does a memory-bound K4 layer spend its time?  Times MBConv block-1 expand (64 x 256 x 256 x 16 -> 96, 1x1, SiLU) and a few
other shapes (round 2 used it with a development switch in the kernel that skipped the stores / the TMEM loads; the switch is gone)."""
import os, sys
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [_R, os.path.join(_R, "tests")]
import torch
from monoforce_b200 import ops

def timed(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

dev = "cuda"
only = [a for a in sys.argv[1:] if not a.startswith("-")]
for name, (N, H, W, Cin, Cout, act) in {"b1_expand folded x4: 64->384 @256x64": (64, 256, 64, 64, 384, ops.ACT_SILU),
                                      "b2_expand folded x4: 96->576 @128x32": (64, 128, 32, 96, 576, ops.ACT_SILU),
                                      "b1_expand 16->96 @256": (64, 256, 256, 16, 96, ops.ACT_SILU), "b2_expand 24->144 @128": (64, 128, 128, 24, 144, ops.ACT_SILU),
                                      "64->96 @256 (full K chunk)": (64, 256, 256, 64, 96, ops.ACT_SILU), "16->128 @256": (64, 256, 256, 16, 128, ops.ACT_SILU),
                                      "b0_project 32->16 @256": (64, 256, 256, 32, 16, ops.ACT_NONE)}.items():
    if only and not any(o in name for o in only):
        continue
    x = torch.randn(N, H, W, Cin, device=dev).to(torch.bfloat16)
    w = (torch.randn(Cout, 1, 1, Cin, device=dev) * 0.1).to(torch.bfloat16)
    sc, sh = torch.ones(Cout, device=dev), torch.zeros(Cout, device=dev)
    ms = timed(lambda: ops.conv2d_nhwc(x, w, sc, sh, act))
    mb = (x.numel() + N * H * W * Cout) * 2 / 1e6
    print(f"{name:40s} {ms * 1e3:8.1f} us   {mb:7.0f} MB  -> {mb / ms / 1e3:6.2f} TB/s", flush=True)
