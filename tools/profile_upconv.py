"""ncu target: the camera Up.conv[0] layer at BASELINE config 4 (64 x 32 x 32 x 432 -> 512, 3x3, GELU), the K4 launch furthest below the tensor peak."""
import os, sys
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [_R, os.path.join(_R, "tests")]
import torch
from monoforce_b200 import ops
dev = "cuda"
x = torch.randn(64, 32, 32, 432, device=dev).to(torch.bfloat16)
w = (torch.randn(512, 3, 3, 432, device=dev) * 0.02).to(torch.bfloat16)
sc, sh = torch.ones(512, device=dev), torch.zeros(512, device=dev)
for _ in range(4):
    ops.conv2d_nhwc(x, w, sc, sh, ops.ACT_GELU, pad=(1, 1))
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    ops.conv2d_nhwc(x, w, sc, sh, ops.ACT_GELU, pad=(1, 1))
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
print(f"{ms * 1e3:.1f} us  {2 * 64 * 1024 * 512 * 432 * 9 / ms / 1e9:.0f} TFLOP/s")
