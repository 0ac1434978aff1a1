"""ncu target: the K4 tcgen05 convolution at the BevEncode head shape (16 x 128 x 128 x 256 -> 384) and K5 lift-splat."""
import os, sys
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [_R, os.path.join(_R, "tests")]
import torch
from monoforce_b200 import ops, LiftSplatShoot
from helpers_lss import default_cfg, make_inputs
dev = "cuda"
x = torch.randn(16, 128, 128, 256, device=dev).to(torch.bfloat16)
w = (torch.randn(384, 3, 3, 256, device=dev) * 0.02).to(torch.bfloat16)
sc, sh = torch.ones(384, device=dev), torch.zeros(384, device=dev)
for _ in range(3):
    ops.conv_bn_act_nhwc(x, w, sc, sh, ops.ACT_GELU)
grid_conf, aug_conf = default_cfg()
torch.manual_seed(0)
net = LiftSplatShoot(grid_conf, aug_conf).to(dev).eval()
inputs = [t.to(dev) for t in make_inputs(grid_conf, aug_conf, 16, 0)]
with torch.no_grad():
    for _ in range(2):
        net.get_voxels(*inputs)
torch.cuda.synchronize()
print("done")
