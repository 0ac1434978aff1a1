"""Small launch sequence for ncu: 2 warm-up + 1 profiled forward (forces+cost) and backward at BASELINE config 3 size."""
import sys, os
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _R); sys.path.insert(0, os.path.join(_R, 'tests'))
import torch
from bench import synth_inputs
from monoforce_b200 import DPhysics
from monoforce_b200.losses import physics_loss

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
d = synth_inputs(B, 0)
sim = DPhysics(d["cfg"], device="cuda"); sim.fused_cost = True
controls = d["controls"].cuda(); ts = d["ts"].cuda()
with torch.no_grad():
    gt, _ = sim(d["z_gt"].cuda().unsqueeze(0), controls)
z = d["z0"].cuda().unsqueeze(0).requires_grad_(True)
fr = d["fr0"].cuda().unsqueeze(0).requires_grad_(True)
for _ in range(n):
    z.grad = None; fr.grad = None
    st, _ = sim(z, controls, friction=fr)
    physics_loss(st, gt, ts, ts, 0.9).backward()
torch.cuda.synchronize()
print("done")
