"""Scratch timing of the forward / backward library calls at BASELINE config 3 size (not the bench)."""
import sys, os
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _R); sys.path.insert(0, os.path.join(_R, 'tests'))
import torch
from bench import synth_inputs
from monoforce_b200 import DPhysics
from monoforce_b200.losses import physics_loss

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
d = synth_inputs(B, 0)
sim = DPhysics(d["cfg"], device="cuda"); sim.fused_cost = True
controls = d["controls"].cuda(); ts = d["ts"].cuda()
with torch.no_grad():
    gt, _ = sim(d["z_gt"].cuda().unsqueeze(0), controls)
z = d["z0"].cuda().unsqueeze(0).requires_grad_(True)
fr = d["fr0"].cuda().unsqueeze(0).requires_grad_(True)
for tape in (True, False):
    sim.adjoint_tape = tape
    res = {}
    for it in range(7):
        z.grad = None; fr.grad = None
        sim.timings = []
        st, _ = sim(z, controls, friction=fr)
        physics_loss(st, gt, ts, ts, 0.9).backward()
        torch.cuda.synchronize()
        if it >= 2:
            for n, a, b in sim.timings:
                res.setdefault(n, []).append(a.elapsed_time(b))
    print(f"tape={tape}: " + "  ".join(f"{n} min {min(v):.3f} ms mean {sum(v)/len(v):.3f} ms" for n, v in res.items()),
          f" |g_z| {z.grad.abs().sum().item():.6e} |g_fr| {fr.grad.abs().sum().item():.6e}")
