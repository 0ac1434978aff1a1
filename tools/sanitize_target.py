"""Tiny workload for compute-sanitizer (memcheck / racecheck / synccheck): every kernel once, small sizes."""
import os, sys
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [_R, os.path.join(_R, "tests")]
import torch
from monoforce_b200 import DPhysics, DPhysConfig, LiftSplatShoot, ops
from helpers_lss import small_cfg, make_inputs
dev = "cuda"
for thr, robot, variant in (("0", "marv", False), ("0", "tradr", True), (None, "marv", False), (None, "tradr", True)):
    # both forward kernels: MFB_FWD_WIDE_MAX_B=0 forces one warp per trajectory (K1), unset picks K1w for these small batches
    # (the same switch for the adjoint: MFB_BWD_WIDE_MAX_B, one warp vs one CTA per trajectory in the single-sweep kernel)
    for var in ("MFB_FWD_WIDE_MAX_B", "MFB_BWD_WIDE_MAX_B"):
        if thr is None:
            os.environ.pop(var, None)
        else:
            os.environ[var] = thr
    cfg = DPhysConfig(robot=robot, grid_res=0.4); cfg.traj_sim_time, cfg.use_odeint = 0.07, variant
    sim = DPhysics(cfg, device=dev); sim.fused_cost = not variant
    B, T = 5, 7
    g = torch.Generator().manual_seed(0)
    z = (0.1 * torch.randn(32, 32, generator=g)).to(dev).requires_grad_(True)
    fr = (0.5 + 0.5 * torch.rand(32, 32, generator=g)).to(dev).requires_grad_(True)
    c = (torch.rand(B, T, 2, generator=g) * 2 - 1).to(dev).requires_grad_(True)
    x0 = torch.zeros(B, 3, device=dev); x0[0, 0] = 6.5; x0[1, 1] = -6.6     # two robots partly off the map
    st = (x0, torch.zeros(B, 3, device=dev), torch.eye(3, device=dev).repeat(B, 1, 1), torch.zeros(B, 3, device=dev))
    for tape in (True, False):          # single-sweep adjoint (contact_sum tape) and three-pass adjoint
        sim.adjoint_tape = tape
        (Xs, Xd, Rs, Om), (Fs, Ff) = sim(z.unsqueeze(0), c, state=st, friction=fr.unsqueeze(0))
        (Xs.sum() + Rs.sum() + 1e-3 * Fs.sum() + 1e-3 * Ff.sum()).backward()
    if robot == "marv":
        with torch.no_grad():
            sim(z.detach().unsqueeze(0), c.detach(), joint_angles=torch.full((B, T, 4), 0.3, device=dev))
grid_conf, aug_conf = small_cfg()
torch.manual_seed(0)
net = LiftSplatShoot(grid_conf, aug_conf).to(dev).eval()
inp = [t.to(dev) for t in make_inputs(grid_conf, aug_conf, 1, 0)]
x = inp[0].clone().requires_grad_(True)
net(x, *inp[1:])["terrain"].sum().backward()
with torch.no_grad():
    net.fast_inference = True
    out = net(*inp)        # whole inference path on repo kernels: stem, K4 (all variants incl. per-image weights, stride 2, residual,
                           # transposed stores, fused heads), depthwise + SE, upsample/concat, bf16 lift-splat, cast, terrain post-processing
    # map groups (one map per scene) + planner post-processing
    cfg = DPhysConfig(robot="marv", grid_res=0.2); cfg.traj_sim_time, cfg.use_odeint = 0.05, False
    sim = DPhysics(cfg, device=dev); sim.fused_cost = True
    zz = torch.zeros(2, 64, 64, device=dev); cc = torch.rand(6, 5, 2, device=dev)
    (Xs, _, Rs, _), _ = sim(zz, cc)
    ops.path_postproc(Xs, Rs)
    ops.terrain_postproc(out["geom"], out["diff"], out["friction"], 2)
zz = torch.zeros(2, 64, 64, device=dev, requires_grad=True)
(Xs, _, _, _), _ = sim(zz, cc)
Xs.sum().backward()
torch.cuda.synchronize()
print("sanitize target done")
