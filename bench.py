#!/usr/bin/env python
"""Benchmark of the B200-native DPhysics rollout (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of synthetic input: BASELINE config 3
(4096 trajectories x T=400 per GPU on one shared synthetic 256x256 height map): forward rollout
(all outputs of the reference API materialised, incl. the two (B,T,N,3) force tensors, plus the
fused per-trajectory cost), physics_loss against ground-truth poses, adjoint backward to
d/dz_grid and d/dfriction (the fit_terrain.py training-loss path).  For N > 1 the trajectory batch
is sharded (weak scaling, 4096 per GPU); the per-shard costs are all-gathered and the two map
gradients all-reduced over NCCL inside the timed step.

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU restatement of the reference
(oracle/, proven bit-identical to the reference's PyTorch-CPU DPhysics) on a bounded sample.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

METRIC = "trajectory-steps/sec (batch×T) at 256² map, T=400; fwd+bwd"
UNIT = "trajectory-steps/s"
T_STEPS = 400
GRID_RES = 0.05           # 256 x 256 map
N_POINTS_BYTES = {"marv": 5432, "tradr": 4280}   # algorithmic bytes per trajectory-step


def synth_inputs(B, seed, device=None, pin=False):
    """the reference's demo hill, shooting controls of monoforce_node.py:42-52."""
    from monoforce_b200 import DPhysConfig
    cfg = DPhysConfig(robot="marv", grid_res=GRID_RES)
    cfg.traj_sim_time = T_STEPS * cfg.dt
    cfg.use_odeint = False
    g = torch.Generator().manual_seed(seed)
    xg, yg = cfg.x_grid, cfg.y_grid
    z_gt = torch.exp(-(xg - 2) ** 2 / 4) * torch.exp(-(yg - 0) ** 2 / 2)       # robot_control.py:52,101
    half = B // 2
    v = torch.cat([torch.rand(half, generator=g) * 0.5 + 0.5, -(torch.rand(B - half, generator=g) * 0.5 + 0.5)])
    w = torch.rand(B, generator=g) * 4.0 - 2.0
    controls = torch.stack([v, w], -1).unsqueeze(1).repeat(1, T_STEPS, 1).contiguous()
    z0 = torch.zeros_like(z_gt)                                                # fit_terrain.py:41-43
    fr0 = 0.5 * torch.ones_like(z_gt)
    ts = torch.arange(0, T_STEPS * cfg.dt, cfg.dt)[None][:, :T_STEPS]
    out = dict(cfg=cfg, z_gt=z_gt, z0=z0, fr0=fr0, controls=controls, ts=ts)
    if pin:
        for k in ("z_gt", "z0", "fr0", "controls", "ts"):
            out[k] = out[k].pin_memory()
    return out


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else 0x8: "hw_slowdown",
                 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_reference_run(B, steps, warmup, fwd_only=False):
    """The reference's own algorithm on the host cores (oracle port, all threads), fwd + autograd bwd."""
    from oracle import dphysics_oracle as O
    from helpers_mfb import make_spec
    from monoforce_b200.losses import physics_loss
    torch.set_num_threads(os.cpu_count() or 1)
    d = synth_inputs(B, seed=0)
    spec = make_spec(d["cfg"])
    with torch.no_grad():
        gt, _ = O.rollout(spec, d["z_gt"].unsqueeze(0).expand(B, -1, -1), d["controls"])
    times = []
    for i in range(warmup + steps):
        z = d["z0"].clone().requires_grad_(not fwd_only)
        fr = d["fr0"].clone().requires_grad_(not fwd_only)
        t0 = time.perf_counter()
        if fwd_only:
            with torch.no_grad():
                O.rollout(spec, z.unsqueeze(0).expand(B, -1, -1), d["controls"], friction=fr.unsqueeze(0).expand(B, -1, -1))
        else:
            st, _ = O.rollout(spec, z.unsqueeze(0).expand(B, -1, -1), d["controls"], friction=fr.unsqueeze(0).expand(B, -1, -1))
            physics_loss(st, gt, d["ts"], d["ts"], 0.9).backward()
        dt_ = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt_)
    return sum(times) / len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = args.cpu_sample
    sec = cpu_reference_run(B, args.steps, args.warmup)
    val = B * T_STEPS / sec
    cores = os.cpu_count() or 1
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"BASELINE config 3 (fwd+bwd, 256x256 shared map, T=400) on a bounded sample of {B} "
                               f"trajectories per step (the reference's autograd graph needs ~110 MB per trajectory)",
                   "robot": "marv", "n_points": 223, "map": "256x256", "T": T_STEPS, "sample_trajectories": B},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{B} trajectories x {T_STEPS} steps fwd+bwd per step, torch-CPU oracle "
                                   f"(bit-identical to the reference's PyTorch-CPU DPhysics), {cores} threads"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--traj-per-gpu", type=int, default=4096)
    ap.add_argument("--cpu-sample", type=int, default=64, help="trajectories per CPU-baseline step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-encoder", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from monoforce_b200 import DPhysics, _lib
    from monoforce_b200.losses import physics_loss
    from monoforce_b200.dist import allreduce_map_grads

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    B = args.traj_per_gpu
    d = synth_inputs(B, seed=rank, pin=True)        # each rank owns a different shard of control sequences
    cfg = d["cfg"]
    N = cfg.robot_points.shape[0]
    sim = DPhysics(cfg, device=dev)
    sim.fused_cost = True
    controls = d["controls"].to(dev)
    ts = d["ts"].to(dev)
    z_gt = d["z_gt"].to(dev).unsqueeze(0)
    with torch.no_grad():
        states_gt, _ = sim(z_gt, controls)
        states_gt = tuple(s.clone() for s in states_gt)
    z = d["z0"].to(dev).unsqueeze(0).requires_grad_(True)
    fr = d["fr0"].to(dev).unsqueeze(0).requires_grad_(True)
    costs_all = torch.empty(world * B, device=dev) if world > 1 else None

    ev = lambda: torch.cuda.Event(enable_timing=True)

    def step(record=False):
        z.grad = None
        fr.grad = None
        states, forces = sim(z, controls, friction=fr)
        loss = physics_loss(states, states_gt, ts, ts, 0.9)
        loss.backward()
        if world > 1:
            dist.all_gather_into_tensor(costs_all, sim.last_cost)
            allreduce_map_grads(z.grad, fr.grad)           # both 256 KiB maps in one flat all-reduce
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    sim.timings = []          # CUDA events bracketing exactly the C-ABI calls on the launching stream
    n0 = _lib.kernel_launches()
    t_a, t_b = ev(), ev()
    t_a.record()
    for _ in range(args.steps):
        loss = step(record=True)
    t_b.record()
    barrier()
    launches = _lib.kernel_launches() - n0
    clocks = sampler.stop()
    ms_total = t_a.elapsed_time(t_b)
    ms_step = ms_total / args.steps
    fwd_t = [a.elapsed_time(b_) for n, a, b_ in sim.timings if n == "forward"]
    bwd_t = [a.elapsed_time(b_) for n, a, b_ in sim.timings if n == "backward"]
    sim.timings = None
    fwd_ms, bwd_ms = sum(fwd_t) / len(fwd_t), sum(bwd_t) / len(bwd_t)
    if world > 1:
        t = torch.tensor([ms_step, fwd_ms, bwd_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step, fwd_ms, bwd_ms = t.tolist()
    value = world * B * T_STEPS / (ms_step * 1e-3)

    # ---- forward-only (BASELINE config 2) ----
    with torch.no_grad():
        for _ in range(3):
            sim(z, controls, friction=fr)
        torch.cuda.synchronize()
        f_a, f_b = ev(), ev()
        f_a.record()
        for _ in range(args.steps):
            sim(z, controls, friction=fr)
        f_b.record()
        torch.cuda.synchronize()
    fwd_only_ms = f_a.elapsed_time(f_b) / args.steps
    if world > 1:
        t = torch.tensor([fwd_only_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        fwd_only_ms = t.item()

    # ---- training step that never materialises the two force tensors (reported separately, never as an HBM fraction) ----
    sim.return_forces = False
    for _ in range(2):
        step()
    barrier()
    n_a, n_b = ev(), ev()
    n_a.record()
    for _ in range(args.steps):
        step()
    n_b.record()
    barrier()
    sim.return_forces = True
    noforce_ms = n_a.elapsed_time(n_b) / args.steps
    if world > 1:
        t = torch.tensor([noforce_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        noforce_ms = t.item()

    # ---- end to end through the public API with HOST inputs / outputs ----
    h_controls, h_z, h_fr = d["controls"], d["z0"], d["fr0"]
    h_gz = torch.empty_like(h_z).pin_memory()
    h_gfr = torch.empty_like(h_fr).pin_memory()
    h_cost = torch.empty(B).pin_memory()
    h_loss = torch.empty(()).pin_memory()

    def e2e_step():
        c_dev = h_controls.to(dev, non_blocking=True)
        z_dev = h_z.to(dev, non_blocking=True).unsqueeze(0).requires_grad_(True)
        f_dev = h_fr.to(dev, non_blocking=True).unsqueeze(0).requires_grad_(True)
        states, _ = sim(z_dev, c_dev, friction=f_dev)
        l = physics_loss(states, states_gt, ts, ts, 0.9)
        l.backward()
        if world > 1:
            dist.all_gather_into_tensor(costs_all, sim.last_cost)
            allreduce_map_grads(z_dev.grad, f_dev.grad)
        h_gz.copy_(z_dev.grad[0], non_blocking=True)
        h_gfr.copy_(f_dev.grad[0], non_blocking=True)
        h_cost.copy_(sim.last_cost, non_blocking=True)
        h_loss.copy_(l.detach(), non_blocking=True)
        torch.cuda.current_stream().synchronize()      # the caller reads the loss every step
        return float(h_loss)

    for _ in range(3):
        e2e_step()
    barrier()
    e_a, e_b = ev(), ev()
    e_a.record()
    for _ in range(args.steps):
        e2e_step()
    e_b.record()
    barrier()
    e2e_ms = e_a.elapsed_time(e_b) / args.steps
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = t.item()
    h2d = h_controls.numel() * 4 + h_z.numel() * 4 + h_fr.numel() * 4
    d2h = h_gz.numel() * 4 + h_gfr.numel() * 4 + h_cost.numel() * 4 + 4

    peak, peak_src = measured_hbm_peak()
    bytes_per_launch = B * T_STEPS * N_POINTS_BYTES["marv"]
    achieved = bytes_per_launch / (fwd_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "fwd_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None

    # ---- BASELINE config 4 in brief: terrain encoder (16 scenes x 4 cameras, lss_cfg.yaml sizes, eval fast path) ----
    encoder = None
    if rank == 0 and world == 1 and not args.no_encoder:
        try:
            from helpers_lss import default_cfg, make_inputs
            from monoforce_b200 import LiftSplatShoot
            gc, ac = default_cfg()
            torch.manual_seed(0)
            net = LiftSplatShoot(gc, ac).to(dev).eval()
            net.fast_inference = True
            enc_in = [t.to(dev) for t in make_inputs(gc, ac, 16, 0)]
            with torch.no_grad():
                for _ in range(3):
                    net(*enc_in)
                torch.cuda.synchronize()
                g_a, g_b = ev(), ev()
                g_a.record()
                for _ in range(5):
                    net(*enc_in)
                g_b.record()
                torch.cuda.synchronize()
            enc_ms = g_a.elapsed_time(g_b) / 5
            encoder = {"workload": "LiftSplatShoot forward, 16 scenes x 4 cams 256x416 -> 128x128 BEV, eval, "
                                   "tcgen05 dense layers + fused lift-splat", "ms": enc_ms, "scenes_per_s": 16 / (enc_ms * 1e-3)}
            del net, enc_in
        except Exception as e:     # the encoder is an extra; never let it break the headline line
            encoder = {"error": repr(e)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        Bc = args.cpu_sample
        sec = cpu_reference_run(Bc, steps=1, warmup=1)
        cores = os.cpu_count() or 1
        cpu = {"value": Bc * T_STEPS / sec, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{Bc} trajectories x {T_STEPS} steps fwd+bwd (same map/controls recipe), 1 warm-up + 1 timed pass, "
                         f"torch-CPU oracle with {cores} threads"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "BASELINE config 3: 4096 trajectories x T=400 per GPU, forward (states + both force "
                                   "tensors + fused cost materialised) + physics_loss + adjoint backward to z_grid and friction",
                       "robot": "marv", "n_points": N, "map": "256x256 shared (grid_res 0.05)", "T": T_STEPS,
                       "trajectories_per_gpu": B, "global_trajectories": world * B,
                       "l2": "each step writes 8.9 GB of outputs >> 126 MB L2 (inputs 13.6 MB), so no explicit flush",
                       "collectives": "none" if world == 1 else "all_gather(costs) + one flat all_reduce(grad z | grad friction) per step"},
            "roofline": {"bound": "hbm", "kernel": "rollout_fwd_kernel<float,7,step,forces,cost>", "achieved": achieved,
                         "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                         "algorithmic_bytes_per_launch": bytes_per_launch, "kernel_ms": fwd_ms, "traffic": traffic},
            "kernels_ms": {"forward_call (cell table + rollout_fwd)": fwd_ms,
                           "backward_call (cell table + rollout_bwd + grad scatter)": bwd_ms},
            "forward_only": {"value": world * B * T_STEPS / (fwd_only_ms * 1e-3), "unit": UNIT, "ms_per_step": fwd_only_ms,
                             "workload": "BASELINE config 2 (forward only, all outputs materialised)"},
            "states_only_training": {"value": world * B * T_STEPS / (noforce_ms * 1e-3), "unit": UNIT, "ms_per_step": noforce_ms,
                                     "workload": "same step with DPhysics.return_forces=False (states + fused cost only; issue-bound, "
                                                 "not an HBM number)"},
            "e2e": {"value": world * B * T_STEPS / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "path": "DPhysics.forward with pinned host controls/maps -> loss, map gradients and costs read back"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "cpu_baseline": cpu,
            "encoder": encoder,
            "loss": float(loss.detach()),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
