#!/usr/bin/env python
"""Benchmark of the B200-native DPhysics rollout (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of synthetic input: BASELINE config 3
(4096 trajectories x T=400 per GPU on one shared synthetic 256x256 height map): forward rollout
(all outputs of the reference API materialised, incl. the two (B,T,N,3) force tensors, plus the
fused per-trajectory cost), physics_loss against ground-truth poses, adjoint backward to
d/dz_grid and d/dfriction (the fit_terrain.py training-loss path).  For N > 1 the trajectory batch
is sharded (weak scaling, 4096 per GPU); the per-shard costs are all-gathered (issued right after the
forward, overlapping the adjoint) and the two map gradients all-reduced over NCCL inside the timed step.

Prints ONE JSON line (rank 0).  Besides the contract keys the line carries the other BASELINE configs as
extra blocks: `forward_only` (config 2), `encoder_cfg4` + `cfg4_e2e` (config 4), `cfg5` (config 5, at --gpus 8),
`strong_scaling` (N > 1), `reference_published` / `encoder_published` (the two timings the reference publishes), `odeint`
(the callers' default integrator at config 2/3 size), `small_batch` latencies and `shooting_e2e`.
`--impl reference` times the CPU restatement of the reference (oracle/, proven bit-identical to the
reference's PyTorch-CPU DPhysics) on a bounded sample.
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
N_SM, SUBPARTS = 148, 4


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


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return {"hbm": float(d["hbm_gbs"]), "bf16": float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1427.0))),
                    "src": "measured (MEASURED_PEAKS.json)"}
        except Exception:
            pass
    return {"hbm": 6650.0, "bf16": 1400.0, "src": "fallback (B200_PROFILING.md)"}


def _profile_json(name):
    p = os.path.join(ROOT, "profiles", name)
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return {}
    return {}


def cpu_reference_run(B, steps, warmup, fwd_only=False):
    """The reference's own algorithm on the host cores (oracle port, all threads), fwd + autograd bwd."""
    from oracle import dphysics_oracle as O
    from helpers_mfb import make_spec
    from monoforce_b200.losses import physics_loss
    torch.set_num_threads(os.cpu_count() or 1)
    d = synth_inputs(B, seed=0)
    spec = make_spec(d["cfg"])
    gt = None
    if not fwd_only:
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


# ---------------------------------------------------------------------------------------------------------------
# timing helpers
# ---------------------------------------------------------------------------------------------------------------
def _ev():
    return torch.cuda.Event(enable_timing=True)


def timed_ms(fn, steps, warmup=3, sync=None):
    """CUDA events on the current (launching) stream around `steps` calls, after `warmup` untimed ones."""
    sync = sync or torch.cuda.synchronize
    for _ in range(warmup):
        fn()
    sync()
    a, b = _ev(), _ev()
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    sync()
    return a.elapsed_time(b) / steps


def wall_ms(fn, steps, warmup=3):
    """Host wall clock with a device synchronise inside every call: a latency, what a blocking caller sees."""
    for _ in range(warmup):
        fn()
        torch.cuda.synchronize()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    ts.sort()
    return {"median_ms": ts[len(ts) // 2], "min_ms": ts[0], "max_ms": ts[-1]}


def guarded(fn):
    """Extra blocks never break the headline line."""
    try:
        return fn()
    except Exception as e:      # noqa: BLE001
        import traceback
        return {"error": repr(e), "trace": traceback.format_exc(limit=3)}


# ---------------------------------------------------------------------------------------------------------------
# extra blocks (rank 0, one GPU)
# ---------------------------------------------------------------------------------------------------------------
def block_reference_published(dev, steps):
    """The ONE timing the reference publishes: examples/diff_physics.ipynb cells 2-7 (64 trajectories x T=600, marv,
    128x128 map `repeat`ed per trajectory, default config => use_odeint=True, autograd enabled): 1.154-1.170 s per call."""
    from monoforce_b200 import DPhysics, DPhysConfig, generate_controls
    cfg = DPhysConfig(robot="marv")
    cfg.grid_res, cfg.d_max, cfg.traj_sim_time, cfg.dt = 0.1, 6.4, 6.0, 0.01
    n, T, dt = cfg.n_sim_trajs, cfg.traj_sim_time, cfg.dt
    H = int(2 * cfg.d_max / cfg.grid_res)
    xg, yg = torch.meshgrid(torch.linspace(-cfg.d_max, cfg.d_max, H), torch.linspace(-cfg.d_max, cfg.d_max, H), indexing="ij")
    z = (1.0 * torch.exp(-1.0 * ((xg - 0) ** 2 + (yg - 4) ** 2)) + 4.0 * torch.exp(-5.0 * ((xg - 1) ** 2 + (yg + 2) ** 2)) +
         2.0 * torch.exp(-3.0 * ((xg + 2) ** 2 + (yg + 4) ** 2))) / 3.0
    z = z.repeat(n, 1, 1).to(dev)
    torch.manual_seed(0)
    cf, _ = generate_controls(n_trajs=n // 2, v_range=(cfg.vel_max / 2, cfg.vel_max), w_range=(-cfg.omega_max, cfg.omega_max),
                              time_horizon=T, dt=dt)
    cb, _ = generate_controls(n_trajs=n // 2, v_range=(-cfg.vel_max, -cfg.vel_max / 2), w_range=(-cfg.omega_max, cfg.omega_max),
                              time_horizon=T, dt=dt)
    controls = torch.cat([cf, cb], 0).to(dev)
    sim = DPhysics(cfg, device=dev)

    def call():
        states, forces = sim(z_grid=z, controls=controls)
        return torch.norm(forces[0], dim=-1).std(dim=-1).std(dim=-1)          # notebook cell 8
    lat = wall_ms(call, max(steps, 5))
    steps_total = n * int(T / dt)
    return {"workload": "diff_physics.ipynb cells 2-8: 64 trajectories x T=600, marv, 128x128 map repeated per trajectory, "
                        "use_odeint=True (default), autograd enabled, forces materialised + path cost",
            "published_s_per_call": [1.154, 1.170], "published_hardware": "unspecified CUDA GPU (notebook output)",
            "ms_per_call": lat["median_ms"], "min_ms": lat["min_ms"], "timing": "host wall clock incl. synchronize, median",
            "trajectory_steps_per_s": steps_total / (lat["median_ms"] * 1e-3),
            "speedup_vs_published": 1154.0 / lat["median_ms"]}


def block_odeint(dev, steps, d, states_gt, ts):
    """The callers' default integrator (dphys_config.py:153 use_odeint=True) at config 2/3 size."""
    from monoforce_b200 import DPhysics
    from monoforce_b200.losses import physics_loss
    import copy
    cfg = copy.copy(d["cfg"])
    cfg.use_odeint = True
    sim = DPhysics(cfg, device=dev)
    controls = d["controls"].to(dev)
    z = d["z0"].to(dev).unsqueeze(0).requires_grad_(True)
    fr = d["fr0"].to(dev).unsqueeze(0).requires_grad_(True)
    B = controls.shape[0]

    def fwd():
        with torch.no_grad():
            sim(z, controls, friction=fr)

    def train():
        z.grad = None
        fr.grad = None
        st, _ = sim(z, controls, friction=fr)
        physics_loss(st, states_gt, ts, ts, 0.9).backward()
    f_ms = timed_ms(fwd, steps)
    t_ms = timed_ms(train, steps)
    return {"workload": "config 2/3 with use_odeint=True (torchdiffeq fixed-grid Euler semantics: forces are time integrals)",
            "forward_ms": f_ms, "forward_value": B * T_STEPS / (f_ms * 1e-3),
            "fwd_bwd_ms": t_ms, "fwd_bwd_value": B * T_STEPS / (t_ms * 1e-3), "unit": UNIT}


def block_small_batch(dev, steps):
    """The reference's real call sizes: ROS shooting (monoforce_node.py:75: 64 trajectories, T=500, 128x128 map repeated,
    no_grad, default odeint) and a training batch (train.py:231-246: bsz distinct 32x32 maps, given initial state)."""
    from monoforce_b200 import DPhysics, DPhysConfig, generate_controls
    from monoforce_b200.losses import physics_loss
    out = {}
    cfg = DPhysConfig(robot="marv")                     # grid_res 0.1 -> 128x128, traj_sim_time 5.0, use_odeint True
    sim = DPhysics(cfg, device=dev)
    torch.manual_seed(1)
    controls, _ = generate_controls(n_trajs=cfg.n_sim_trajs, time_horizon=cfg.traj_sim_time, dt=cfg.dt,
                                    v_range=(-1, 1), w_range=(-2, 2))
    controls = controls.to(dev)
    xg, yg = cfg.x_grid, cfg.y_grid
    z = (torch.exp(-(xg - 2) ** 2 / 4) * torch.exp(-yg ** 2 / 2)).repeat(cfg.n_sim_trajs, 1, 1).to(dev)

    def shoot():
        with torch.no_grad():
            states, forces = sim(z_grid=z, controls=controls)
            return torch.norm(forces[0], dim=-1).std(dim=-1).std(dim=-1).argmin()
    lat = wall_ms(shoot, max(steps, 10))
    out["ros_shooting_B64_T500_128x128_odeint"] = {**lat, "trajectory_steps_per_s": 64 * 500 / (lat["median_ms"] * 1e-3)}
    # the same call replayed from a CUDA graph (DPhysics.graphed): the host-side work of a call is paid once
    run = sim.graphed(z, controls)

    def shoot_graph():
        states, forces = run(z_grid=z)
        return torch.norm(forces[0], dim=-1).std(dim=-1).std(dim=-1).argmin()
    lat = wall_ms(shoot_graph, max(steps, 10))
    out["ros_shooting_graph_replay"] = {**lat, "trajectory_steps_per_s": 64 * 500 / (lat["median_ms"] * 1e-3)}

    cfg2 = DPhysConfig(robot="marv", grid_res=0.4)      # 32x32 maps
    sim2 = DPhysics(cfg2, device=dev)
    bsz, T = 16, int(cfg2.traj_sim_time / cfg2.dt)
    g = torch.Generator().manual_seed(2)
    terrain = (0.1 * torch.randn(bsz, 32, 32, generator=g)).to(dev).requires_grad_(True)
    fric = (0.5 + 0.3 * torch.rand(bsz, 32, 32, generator=g)).to(dev).requires_grad_(True)
    ctrl = torch.stack([torch.rand(bsz, generator=g) * 0.5 + 0.3, torch.rand(bsz, generator=g) - 0.5], -1)
    ctrl = ctrl.unsqueeze(1).repeat(1, T, 1).to(dev)
    ts = (torch.arange(T) * cfg2.dt)[None].to(dev)
    x0 = torch.zeros(bsz, 3, device=dev)
    st0 = lambda: (x0.clone(), torch.zeros_like(x0), torch.eye(3, device=dev).repeat(bsz, 1, 1), torch.zeros_like(x0))
    with torch.no_grad():
        gt, _ = sim2(z_grid=torch.zeros_like(terrain), controls=ctrl, state=st0())

    def train():
        terrain.grad = None
        fric.grad = None
        st, _ = sim2(z_grid=terrain, controls=ctrl, state=st0(), friction=fric)
        physics_loss(st, gt, ts, ts).backward()
    lat2 = wall_ms(train, max(steps, 10))
    out["training_B16_T500_distinct_32x32_odeint_fwd_bwd"] = {**lat2, "trajectory_steps_per_s": bsz * T / (lat2["median_ms"] * 1e-3)}
    out["note"] = ("forward = the small-batch kernel K1w (one CTA per trajectory, two contact points per thread; the library picks "
                   "it below 512 / 1024 trajectories, rollout_fwd.cu); these are latencies, not throughputs")
    return out


def block_shooting_e2e(dev, steps, d):
    """Forward-only end to end: pinned host controls in -> fused per-trajectory costs back on the host, config 2 size."""
    from monoforce_b200 import DPhysics
    sim = DPhysics(d["cfg"], device=dev)
    sim.fused_cost, sim.return_forces = True, False
    B = d["controls"].shape[0]
    z = d["z_gt"].to(dev).unsqueeze(0)
    h_cost = torch.empty(B).pin_memory()

    def call():
        with torch.no_grad():
            c = d["controls"].to(dev, non_blocking=True)
            sim(z, c)
            h_cost.copy_(sim.last_cost, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        return int(h_cost.argmin())
    ms = timed_ms(call, steps)
    return {"workload": "controls (host, pinned) -> rollout (states + fused cost, forces not materialised) -> costs on the host; "
                        "4096 x 400, shared 256x256 map",
            "ms_per_step": ms, "value": B * T_STEPS / (ms * 1e-3), "unit": UNIT,
            "h2d_bytes_per_step": d["controls"].numel() * 4, "d2h_bytes_per_step": B * 4}


def _cfg4_confs():
    from helpers_lss import default_cfg
    gc, ac = default_cfg()
    gc["xbound"] = [-6.4, 6.4, 0.05]
    gc["ybound"] = [-6.4, 6.4, 0.05]
    ac["final_dim"] = [512, 512]
    return gc, ac


def block_encoder(dev, steps, cfg4, peaks):
    """TerrainEncoder forward (eval, bf16 tensor-core path), 16 scenes x 4 cameras; images start in pinned HOST memory
    and their copy is inside the timed region."""
    from helpers_lss import default_cfg, make_inputs
    from monoforce_b200 import LiftSplatShoot, _lib
    import warnings
    gc, ac = _cfg4_confs() if cfg4 else default_cfg()
    torch.manual_seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        net = LiftSplatShoot(gc, ac).to(dev).eval()
    net.fast_inference = True
    host = [t.pin_memory() for t in make_inputs(gc, ac, 16, 0)]
    calib = [t.to(dev) for t in host[1:]]

    def call():
        with torch.no_grad():
            return net(host[0].to(dev, non_blocking=True), *calib)
    n0 = _lib.kernel_launches()
    call()
    launches = _lib.kernel_launches() - n0      # kernels of ONE forward (counted on an eager call)
    eager_ms = timed_ms(call, 5)
    net.fast_graph = True                 # the same launches replayed from a CUDA graph (static shapes and calibration)
    serial_ms = timed_ms(call, 5)
    # steady-state loader: batch i+1 is uploaded on a copy stream while batch i is encoded (two device buffers); every
    # step's upload is inside the timed region, it just overlaps the previous step's kernels
    copy_stream = torch.cuda.Stream(device=dev)
    bufs = [host[0].to(dev), host[0].to(dev)]
    ev_up = [torch.cuda.Event(), torch.cuda.Event()]
    ev_done = [torch.cuda.Event(), torch.cuda.Event()]
    for e in ev_up + ev_done:
        e.record()
    state = {"i": 0}

    def call():
        cur = state["i"] & 1
        main = torch.cuda.current_stream()
        main.wait_event(ev_up[cur])
        with torch.no_grad():
            out = net(bufs[cur], *calib)
        ev_done[cur].record(main)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_done[1 - cur])
            bufs[1 - cur].copy_(host[0], non_blocking=True)
            ev_up[1 - cur].record(copy_stream)
        state["i"] += 1
        return out
    ms = timed_ms(call, max(steps, 10), warmup=4)
    gflop_scene = 230.9 if cfg4 else 65.7          # SURVEY.md 8(a) A14 [FlopCounterMode]
    tf = 16 * gflop_scene / ms
    res = {"workload": ("BASELINE config 4 encoder: 16 scenes x 4 cams 512x512 -> 256x256 BEV" if cfg4 else
                        "lss_cfg.yaml: 16 scenes x 4 cams 256x416 -> 128x128 BEV") + ", eval, all layers on repo kernels (NHWC bf16, "
                       "CUDA-graph replay), every step's images copied from pinned host memory inside the timed region (upload of the next "
                       "batch overlaps the encoding of the current one)",
           "ms": ms, "ms_upload_then_encode_one_stream": serial_ms, "ms_eager_launches_one_stream": eager_ms, "scenes_per_s": 16 / (ms * 1e-3), "h2d_bytes": host[0].numel() * 4,
           "repo_kernel_launches": int(launches),
           "tflops": tf, "frac_of_sustained_bf16_peak": tf / peaks["bf16"], "peak_tflops": peaks["bf16"]}
    return res, net, host, calib


def block_encoder_published(dev, steps):
    """The reference's other published number: LiftSplatShoot forward on ONE scene (4 cameras 256x416 -> 128x128 BEV, lss_cfg.yaml),
    0.391 s for the first call in monoforce/examples/monoforce_inference_with_rough_data.ipynb:392.
    Here: latency of one forward with the images coming from pinned host memory and the terrain map read back, eager launches
    and CUDA-graph replay."""
    from helpers_lss import default_cfg, make_inputs
    from monoforce_b200 import LiftSplatShoot
    import warnings
    gc, ac = default_cfg()
    torch.manual_seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        net = LiftSplatShoot(gc, ac).to(dev).eval()
    net.fast_inference = True
    host = [t.pin_memory() for t in make_inputs(gc, ac, 1, 0)]
    calib = [t.to(dev) for t in host[1:]]

    def call():
        with torch.no_grad():
            out = net(host[0].to(dev, non_blocking=True), *calib)
            return out["terrain"].cpu()                      # 64 KB back to the host: a blocking caller's view
    eager = wall_ms(call, max(steps, 10))
    net.fast_graph = True
    graph = wall_ms(call, max(steps, 10))
    return {"workload": "monoforce_inference_with_rough_data.ipynb: LiftSplatShoot forward, 1 scene x 4 cams 256x416 -> 128x128 BEV, eval; "
                        "images from pinned host memory, terrain map read back; host wall clock incl. synchronize, median",
            "published_s_first_call": 0.391, "published_hardware": "unspecified CUDA GPU (notebook output, first call incl. warm-up)",
            "ms_eager_launches": eager["median_ms"], "ms_graph_replay": graph["median_ms"], "min_ms": graph["min_ms"]}


def block_cfg4_e2e(dev, steps, net, host, calib):
    """BASELINE config 4 end to end: images (host) -> encoder -> terrain / friction maps -> 16 scenes x 256 control
    sequences (one map per scene, map groups) -> fused costs -> host."""
    from monoforce_b200 import DPhysics, DPhysConfig
    cfg = DPhysConfig(robot="marv", grid_res=0.05)
    cfg.traj_sim_time, cfg.use_odeint = T_STEPS * cfg.dt, False
    sim = DPhysics(cfg, device=dev)
    sim.fused_cost = True
    per_scene, scenes = 256, 16
    n = per_scene * scenes
    g = torch.Generator().manual_seed(0)
    ctrl = torch.stack([torch.rand(n, generator=g) * 0.5 + 0.5, torch.rand(n, generator=g) * 4 - 2], -1)
    ctrl = ctrl.unsqueeze(1).repeat(1, T_STEPS, 1).contiguous().pin_memory()
    h_cost = torch.empty(n).pin_memory()
    parts = {}

    def call(forces=True):
        sim.return_forces = forces
        with torch.no_grad():
            out = net(host[0].to(dev, non_blocking=True), *calib)
            c = ctrl.to(dev, non_blocking=True)
            sim(out["terrain"].squeeze(1), c, friction=out["friction"].squeeze(1))      # (16,256,256) maps, 4096 trajectories
            h_cost.copy_(sim.last_cost, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        return h_cost.view(scenes, per_scene).argmin(dim=1)
    ms_serial = timed_ms(call, max(steps // 2, 5))
    ms_nf = timed_ms(lambda: call(False), max(steps // 2, 5))
    sim.return_forces = True

    # steady state of a planner loop: the NEXT frame's images and controls are uploaded on a copy stream while the current
    # frame is encoded and rolled out; every step still uploads one frame and reads one set of costs back
    copy_stream = torch.cuda.Stream(device=dev)
    img = [host[0].to(dev), host[0].to(dev)]
    cdev = [ctrl.to(dev), ctrl.to(dev)]
    ev_up = [torch.cuda.Event(), torch.cuda.Event()]
    ev_done = [torch.cuda.Event(), torch.cuda.Event()]
    for e in ev_up + ev_done:
        e.record()
    st = {"i": 0}

    def piped():
        cur = st["i"] & 1
        main = torch.cuda.current_stream()
        main.wait_event(ev_up[cur])
        with torch.no_grad():
            out = net(img[cur], *calib)
            sim(out["terrain"].squeeze(1), cdev[cur], friction=out["friction"].squeeze(1))
            h_cost.copy_(sim.last_cost, non_blocking=True)
        ev_done[cur].record(main)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_done[1 - cur])
            img[1 - cur].copy_(host[0], non_blocking=True)
            cdev[1 - cur].copy_(ctrl, non_blocking=True)
            ev_up[1 - cur].record(copy_stream)
        main.synchronize()                                   # the planner reads this frame's costs before it goes on
        st["i"] += 1
        return h_cost.view(scenes, per_scene).argmin(dim=1)
    ms = timed_ms(piped, max(steps // 2, 5))
    return {"workload": "images (host) -> LiftSplatShoot (eval) -> terrain/friction (16 x 256x256) -> DPhysics rollout, 16 scenes x "
                        "256 trajectories x T=400 (one map per scene), all reference outputs materialised + fused cost -> costs on host",
            "ms": ms, "ms_upload_then_compute_one_stream": ms_serial, "scenes_per_s": scenes / (ms * 1e-3),
            "trajectory_steps_per_s": n * T_STEPS / (ms * 1e-3), "ms_without_force_tensors_one_stream": ms_nf,
            "pipelining": "the next frame's upload (images + controls) overlaps the current frame's kernels; costs are read back every step", "h2d_bytes": host[0].numel() * 4 + ctrl.numel() * 4, "d2h_bytes": n * 4, **parts}


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
    ap.add_argument("--no-extras", action="store_true", help="headline + roofline + e2e only")
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
    peaks = measured_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(*vals):
        if world == 1:
            return list(vals)
        t = torch.tensor(vals, device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    def make_job(B, seed):
        """Config 3 on B trajectories of this rank: returns (step, sim, tensors)."""
        d = synth_inputs(B, seed=seed, pin=True)
        sim = DPhysics(d["cfg"], device=dev)
        sim.fused_cost = True
        controls = d["controls"].to(dev)
        ts = d["ts"].to(dev)
        z_gt = d["z_gt"].to(dev).unsqueeze(0)
        with torch.no_grad():
            states_gt, _ = sim(z_gt, controls)
            states_gt = tuple(s.clone() for s in states_gt)
        z = d["z0"].to(dev).unsqueeze(0).requires_grad_(True)
        fr = d["fr0"].to(dev).unsqueeze(0).requires_grad_(True)
        coll_events = [] if world > 1 else None      # CUDA events around the exposed collective tail of every step
        costs_all = torch.empty(world * B, device=dev)
        # the kernel writes this rank's costs straight into its slice of the gather buffer (in-place all-gather)
        sim.cost_buffer = costs_all[rank * B:(rank + 1) * B]

        def step():
            z.grad = None
            fr.grad = None
            states, forces = sim(z, controls, friction=fr)
            work = None
            if world > 1:     # depends on the forward only: travels while the adjoint runs
                work = dist.all_gather_into_tensor(costs_all, sim.cost_buffer, async_op=True)
            loss = physics_loss(states, states_gt, ts, ts, 0.9)
            loss.backward()
            if world > 1:
                if coll_events is not None:
                    e0, e1 = _ev(), _ev()
                    e0.record()
                allreduce_map_grads(z.grad, fr.grad)           # both 256 KiB maps: one flat in-place all-reduce
                work.wait()
                if coll_events is not None:
                    e1.record()
                    coll_events.append((e0, e1))
            return loss
        return step, sim, dict(d=d, controls=controls, ts=ts, states_gt=states_gt, z=z, fr=fr, costs_all=costs_all, coll=coll_events)

    B = args.traj_per_gpu
    step, sim, J = make_job(B, seed=rank)        # each rank owns a different shard of control sequences
    d, controls, ts, states_gt, z, fr, costs_all = (J[k] for k in ("d", "controls", "ts", "states_gt", "z", "fr", "costs_all"))
    cfg = d["cfg"]
    N = cfg.robot_points.shape[0]

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    sim.timings = []          # CUDA events bracketing exactly the C-ABI calls on the launching stream
    n0 = _lib.kernel_launches()
    t_a, t_b = _ev(), _ev()
    t_a.record()
    for _ in range(args.steps):
        loss = step()
    t_b.record()
    barrier()
    launches = _lib.kernel_launches() - n0
    clocks = sampler.stop()
    coll_ms = None
    if J["coll"]:
        tail = [a.elapsed_time(b) for a, b in J["coll"][-args.steps:]]
        mine = sum(tail) / len(tail)
        (hi,), (neg_lo,) = max_over_ranks(mine), max_over_ranks(-mine)
        # the rank that arrives last waits least: min over ranks ~ the collectives themselves, max - min ~ skew between ranks
        coll_ms = {"max_over_ranks": hi, "min_over_ranks": -neg_lo}
    ms_step = t_a.elapsed_time(t_b) / args.steps
    fwd_t = [a.elapsed_time(b_) for n, a, b_ in sim.timings if n == "forward"]
    bwd_t = [a.elapsed_time(b_) for n, a, b_ in sim.timings if n == "backward"]
    sim.timings = None
    fwd_ms, bwd_ms = sum(fwd_t) / len(fwd_t), sum(bwd_t) / len(bwd_t)
    ms_step, fwd_ms, bwd_ms = max_over_ranks(ms_step, fwd_ms, bwd_ms)
    value = world * B * T_STEPS / (ms_step * 1e-3)

    # ---- forward-only (BASELINE config 2) ----
    def fwd_only():
        with torch.no_grad():
            sim(z, controls, friction=fr)
    (fwd_only_ms,) = max_over_ranks(timed_ms(fwd_only, args.steps))

    # ---- training step that never materialises the two force tensors (reported separately, never as an HBM fraction) ----
    sim.return_forces = False
    (noforce_ms,) = max_over_ranks(timed_ms(step, args.steps, warmup=2, sync=barrier))
    sim.return_forces = True

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
        work = dist.all_gather_into_tensor(costs_all, sim.cost_buffer, async_op=True) if world > 1 else None
        l = physics_loss(states, states_gt, ts, ts, 0.9)
        l.backward()
        if world > 1:
            allreduce_map_grads(z_dev.grad, f_dev.grad)
            work.wait()
        h_gz.copy_(z_dev.grad[0], non_blocking=True)
        h_gfr.copy_(f_dev.grad[0], non_blocking=True)
        h_cost.copy_(sim.last_cost, non_blocking=True)
        h_loss.copy_(l.detach(), non_blocking=True)
        torch.cuda.current_stream().synchronize()      # the caller reads the loss every step
        return float(h_loss)
    (e2e_ms,) = max_over_ranks(timed_ms(e2e_step, args.steps, sync=barrier))
    h2d = h_controls.numel() * 4 + h_z.numel() * 4 + h_fr.numel() * 4
    d2h = h_gz.numel() * 4 + h_gfr.numel() * 4 + h_cost.numel() * 4 + 4

    bytes_per_launch = B * T_STEPS * N_POINTS_BYTES["marv"]
    achieved = bytes_per_launch / (fwd_ms * 1e-3) / 1e9
    traffic = _profile_json("fwd_traffic.json").get("dram_bytes_per_launch")
    # adjoint: issue-bound (165 MB of DRAM traffic per launch).  Roofline = warp-instructions issued per second against one
    # instruction per cycle per SM sub-partition; instructions per trajectory-step from the committed ncu capture.
    bwd_prof = _profile_json("bwd_inst.json")
    roofline_bwd = None
    if bwd_prof.get("warp_inst_per_traj_step"):
        clk = (clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965) * 1e6
        inst = float(bwd_prof["warp_inst_per_traj_step"]) * B * T_STEPS
        ach = inst / (bwd_ms * 1e-3)
        pk = N_SM * SUBPARTS * clk
        roofline_bwd = {"bound": "issue", "kernel": bwd_prof.get("kernel", "rollout_bwd_sweep_kernel<float,step>"),
                        "achieved": ach / 1e9, "peak": pk / 1e9, "unit": "Gwarp-inst/s", "frac": ach / pk,
                        "warp_inst_per_traj_step": bwd_prof["warp_inst_per_traj_step"], "kernel_ms": bwd_ms,
                        "sm_mhz": clk / 1e6, "source": bwd_prof.get("source"),
                        "dram_bytes_per_launch": bwd_prof.get("dram_bytes_per_launch")}

    # ---- multi-GPU extras: strong scaling of config 3 (4096 trajectories in total) and BASELINE config 5 ----
    strong = cfg5 = None
    if world > 1:
        Bs = 4096 // world
        s_step, s_sim, _ = make_job(Bs, seed=100 + rank)
        (s_ms,) = max_over_ranks(timed_ms(s_step, args.steps, sync=barrier))
        strong = {"workload": f"config 3 with 4096 trajectories IN TOTAL ({Bs} per GPU), same collectives", "ms_per_step": s_ms,
                  "value": 4096 * T_STEPS / (s_ms * 1e-3), "unit": UNIT, "scaling": "strong"}
        del s_step, s_sim
        if world == 8:
            B5 = 8192
            c_step, c_sim, C5 = make_job(B5, seed=200 + rank)

            def shoot5():
                with torch.no_grad():
                    c_sim(C5["z"], C5["controls"], friction=C5["fr"])
                    dist.all_gather_into_tensor(C5["costs_all"], c_sim.cost_buffer)
                    return C5["costs_all"].argmin()
            (f5,) = max_over_ranks(timed_ms(shoot5, args.steps, sync=barrier))
            (t5,) = max_over_ranks(timed_ms(c_step, max(args.steps // 2, 3), sync=barrier))
            cfg5 = {"workload": "BASELINE config 5: 65536 trajectories x T=400 over 8 GPUs (8192 per GPU), all reference outputs "
                                "materialised, NCCL all-gather of the per-shard costs inside the timed step",
                    "forward_allgather_ms": f5, "forward_value": 8 * B5 * T_STEPS / (f5 * 1e-3),
                    "fwd_bwd_ms": t5, "fwd_bwd_value": 8 * B5 * T_STEPS / (t5 * 1e-3), "unit": UNIT}
            del c_step, c_sim, C5
        torch.cuda.empty_cache()

    extras = {}
    solo = rank == 0 and world == 1
    if solo and not args.no_extras:
        extras["reference_published"] = guarded(lambda: block_reference_published(dev, args.steps))
        extras["odeint"] = guarded(lambda: block_odeint(dev, args.steps, d, states_gt, ts))
        extras["small_batch"] = guarded(lambda: block_small_batch(dev, args.steps))
        extras["shooting_e2e"] = guarded(lambda: block_shooting_e2e(dev, args.steps, d))
    # ---- BASELINE config 4: terrain encoder alone (default + config-4 sizes) and encoder -> rollout end to end ----
    if solo and not args.no_encoder:
        del states_gt
        torch.cuda.empty_cache()

        def enc_blocks():
            res, net, host, calib = block_encoder(dev, args.steps, False, peaks)
            extras["encoder"] = res
            del net, host, calib
            res4, net4, host4, calib4 = block_encoder(dev, args.steps, True, peaks)
            extras["encoder_cfg4"] = res4
            extras["cfg4_e2e"] = guarded(lambda: block_cfg4_e2e(dev, args.steps, net4, host4, calib4))
            del net4, host4, calib4
            extras["encoder_published"] = guarded(lambda: block_encoder_published(dev, args.steps))
            return True
        r = guarded(enc_blocks)
        if r is not True:
            extras.setdefault("encoder", r)

    cpu = None
    if solo and not args.no_cpu_baseline:
        Bc = args.cpu_sample
        sec = cpu_reference_run(Bc, steps=1, warmup=1)
        cores = os.cpu_count() or 1
        cpu = {"value": Bc * T_STEPS / sec, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{Bc} trajectories x {T_STEPS} steps fwd+bwd (same map/controls recipe), 1 warm-up + 1 timed pass, "
                         f"torch-CPU oracle with {cores} threads"}
        if not args.no_extras:
            sec_f = cpu_reference_run(256, steps=1, warmup=0, fwd_only=True)
            cpu["forward_no_grad"] = {"value": 256 * T_STEPS / sec_f, "unit": UNIT,
                                      "sample": "256 trajectories x 400 steps forward under no_grad, 1 timed pass (BASELINE.md section 4)"}

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
                       "collectives": "none" if world == 1 else "all_gather(costs), issued after the forward and overlapped with the "
                                      "adjoint + one flat in-place all_reduce(grad z | grad friction) per step"},
            "roofline": {"bound": "hbm", "kernel": "rollout_fwd_kernel<float,7,step,forces,cost>", "achieved": achieved,
                         "peak": peaks["hbm"], "peak_source": peaks["src"], "unit": "GB/s", "frac": achieved / peaks["hbm"],
                         "algorithmic_bytes_per_launch": bytes_per_launch, "kernel_ms": fwd_ms, "traffic": traffic},
            "roofline_bwd": roofline_bwd,
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
            "collective_tail_ms": coll_ms,
            "strong_scaling": strong,
            "cfg5": cfg5,
            **extras,
            "loss": float(loss.detach()),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
