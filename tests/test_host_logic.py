"""CPU tests: host-side logic, the C ABI surface (no compute calls without a GPU), multi-process plumbing."""
import ctypes as C
import os
import re
import socket
import sys

import numpy as np
import pytest
import torch

from helpers_mfb import ROOT


def test_library_exports_every_declared_symbol():
    from monoforce_b200 import _lib
    header = open(os.path.join(ROOT, "include", "monoforce_b200.h")).read()
    declared = set(re.findall(r"\b(mfb_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    lib = _lib.load()
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert lib.mfb_abi_version() == int(re.search(r"#define MFB_ABI_VERSION (\d+)", header).group(1)) == 3


def test_ctypes_structs_mirror_the_header_field_for_field():
    """The ctypes Structures in _lib.py must list the header's struct members in the same order (the boundary is plain C)."""
    from monoforce_b200 import _lib
    header = open(os.path.join(ROOT, "include", "monoforce_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)

    def members(name):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), header, re.S).group(1)
        out = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):
                out.append(re.search(r"([A-Za-z_][A-Za-z0-9_]*)\s*(\[\d+\])?\s*$", part.strip()).group(1))
        return out

    assert members("mfb_rollout_desc") == [f[0] for f in _lib.RolloutDesc._fields_]
    assert members("mfb_rollout_buffers") == [f[0] for f in _lib.RolloutBuffers._fields_]
    assert members("mfb_rollout_grads") == [f[0] for f in _lib.RolloutGrads._fields_]
    assert _lib.kernel_launches() >= 0


def test_struct_layouts_match_header_sizes():
    from monoforce_b200 import _lib
    # 8 int32 + int64 + 9 doubles + 9 doubles
    assert C.sizeof(_lib.RolloutDesc) == 8 * 4 + 8 + 9 * 8 + 9 * 8 + 12 * 8
    assert C.sizeof(_lib.RolloutBuffers) == 20 * 8 + 8 + 8      # 11 inputs + 9 outputs (incl. the contact_sum tape) + workspace
    assert C.sizeof(_lib.RolloutGrads) == 15 * 8


def test_c_abi_rejects_bad_arguments_without_touching_the_gpu():
    from monoforce_b200 import _lib
    lib = _lib.load()
    d = _lib.RolloutDesc(B=4, T=10, N=223, H=64, W=64, n_tracks=4, variant=0, map_stride=0, mass=60., gravity=9.81,
                         stiffness=5e4, damping=3464., grid_res=0.1, d_max=6.4, dt=0.01, omega_max=2., robot_Ly=.5)
    assert lib.mfb_rollout_workspace_bytes(C.byref(d), _lib.MFB_F32) == 64 * 64 * (12 + 8) * 4   # sampling record + corner-gradient record per cell
    io = _lib.RolloutBuffers()
    assert lib.mfb_rollout_forward(C.byref(d), C.byref(io), _lib.MFB_F32, None) == -1
    assert b"NULL" in lib.mfb_last_error()
    for field, val, msg in (("N", 300, b"N must be"), ("W", 32, b"H must equal W"), ("n_tracks", 3, b"n_tracks"),
                            ("variant", 7, b"variant"), ("B", 0, b"B and T")):
        bad = _lib.RolloutDesc.from_buffer_copy(d)
        setattr(bad, field, val)
        assert lib.mfb_rollout_forward(C.byref(bad), C.byref(io), _lib.MFB_F32, None) == -1
        assert msg in lib.mfb_last_error(), (field, lib.mfb_last_error())
        assert lib.mfb_rollout_workspace_bytes(C.byref(bad), _lib.MFB_F32) == -1


def test_config_mirrors_reference_attribute_bag():
    from monoforce_b200 import DPhysConfig
    cfg = DPhysConfig(robot="marv", grid_res=0.05)
    assert cfg.robot_points.shape == (223, 3) and cfg.robot_points.dtype == torch.float32     # diff_physics.ipynb:217
    assert [int(m.sum()) for m in cfg.driving_parts] == [34, 34, 35, 34]
    assert cfg.x_grid.shape == (256, 256) and cfg.z_grid.shape == (256, 256)
    assert cfg.robot_mass == 60. and cfg.use_odeint is True and cfg.dt == 0.01
    assert abs(float(cfg.damping) - np.sqrt(4 * 60 * 50_000.)) < 1e-9
    t = DPhysConfig(robot="tradr")
    assert t.robot_points.shape == (175, 3) and len(t.driving_parts) == 2 and t.robot_mass == 40.
    assert int((t.part_id >= 0).sum()) == 90
    with pytest.raises(ValueError):
        DPhysConfig(robot="spot")
    with pytest.raises(AssertionError):
        DPhysConfig(robot="husky")        # no husky mesh is shipped by the reference either (dphys_config.py:25)


def test_module_surface_and_no_cpu_fallback():
    from monoforce_b200 import DPhysics, DPhysConfig, generate_controls, vw_to_track_vels
    cfg = DPhysConfig(robot="tradr", grid_res=0.4)
    cfg.traj_sim_time = 0.2
    sim = DPhysics(cfg, device="cpu")
    assert sim.x_points.shape == (1, 175, 3) and sim.I_inv.shape == (1, 3, 3) and sim.ts.shape == (20,)
    controls, stamps = generate_controls(n_trajs=3, time_horizon=0.2, dt=0.01)
    assert controls.shape == (3, 20, 2) and stamps.shape == (20,)
    assert torch.equal(controls[:, 0], controls[:, -1])
    tv = vw_to_track_vels(torch.tensor([1.0]), torch.tensor([0.5]), cfg.robot_size, 2)
    assert tv.shape == (1, 2) and tv[0, 0] < tv[0, 1]
    with pytest.raises(ValueError):
        vw_to_track_vels(torch.tensor([1.0]), torch.tensor([0.5]), cfg.robot_size, 3)
    if not torch.cuda.is_available():
        # device='cpu' means host tensors in / out, staged through the GPU: without one it must fail loudly
        with pytest.raises(RuntimeError, match="needs a CUDA device"):
            sim(cfg.z_grid.repeat(3, 1, 1), controls)
    state = (torch.zeros(3, 3), torch.zeros(3, 3), torch.eye(3).repeat(3, 1, 1), torch.zeros(3, 3))
    with pytest.raises(AssertionError, match="Controls shape"):       # same message as dphysics.py:575
        sim(cfg.z_grid.repeat(3, 1, 1), controls[:2], state=state)


def test_per_call_attributes_and_default_friction_cache():
    """Host-side shortcuts of the planner's latency path keep the module's semantics: per-call tensors are plain attributes
    (never parameters / buffers, so state_dict is unchanged), everything else still goes through nn.Module.__setattr__, and
    the cached device copy of dphys_cfg.friction follows in-place edits and replacement of the config tensor."""
    from monoforce_b200 import DPhysics, DPhysConfig
    cfg = DPhysConfig(robot="tradr", grid_res=0.4)
    sim = DPhysics(cfg, device="cpu")
    sim.controls = torch.zeros(2, 5, 2)
    sim.z_grid = torch.nn.Parameter(torch.zeros(1, 4, 4))              # even a Parameter stays a plain per-call attribute
    assert "controls" in sim.__dict__ and "z_grid" in sim.__dict__ and len(list(sim.parameters())) == 0
    assert len(sim.state_dict()) == 0
    sim.extra = torch.nn.Parameter(torch.ones(3))                      # anything else is registered as usual
    assert [n for n, _ in sim.named_parameters()] == ["extra"]
    f0 = sim._default_friction()
    assert f0.shape == (1,) + tuple(cfg.friction.shape) and sim._default_friction() is f0          # cached
    cfg.friction.mul_(0.5)                                             # in-place edit bumps the tensor version
    f1 = sim._default_friction()
    assert f1 is not f0 and torch.equal(f1[0], cfg.friction)
    cfg.friction = torch.full_like(cfg.friction, 0.3)                  # replaced tensor
    assert torch.equal(sim._default_friction()[0], cfg.friction)
    z = sim._zero_scalar(torch.float32).expand(2, 5, 4)
    assert z.shape == (2, 5, 4) and z.stride() == (0, 0, 0) and float(z.abs().sum()) == 0.0


def test_unsupported_integration_mode_is_refused_not_run_as_euler():
    """the reference honours integration_mode ('rk4' in update_state :361-383, `method=` of odeint :510-511);
    the kernels are Euler only, so any other mode must raise instead of silently producing Euler trajectories."""
    from monoforce_b200 import DPhysics, DPhysConfig, generate_controls
    cfg = DPhysConfig(robot="tradr", grid_res=0.4)
    cfg.traj_sim_time = 0.2
    controls, _ = generate_controls(n_trajs=2, time_horizon=0.2, dt=0.01)
    for odeint in (False, True):
        cfg.use_odeint, cfg.integration_mode = odeint, "rk4"
        with pytest.raises(NotImplementedError, match="rk4"):
            DPhysics(cfg, device="cpu")(cfg.z_grid.repeat(2, 1, 1), controls)
    cfg.use_odeint, cfg.integration_mode = False, "midpoint"
    with pytest.raises(ValueError, match="Unknown integration mode"):      # same error as dphysics.py:382
        DPhysics(cfg, device="cpu")(cfg.z_grid.repeat(2, 1, 1), controls)


def test_efficientnet_trunk_weights_are_plumbed_and_absence_is_loud(tmp_path, monkeypatch):
    """lss.py:55 starts from ImageNet weights; here a weights file can be passed (argument or environment
    variable) and a missing one is announced instead of silently training from scratch."""
    import warnings
    from monoforce_b200.efficientnet import EfficientNet
    from monoforce_b200.terrain_encoder import CamEncode
    monkeypatch.delenv("MFB_EFFICIENTNET_B0_WEIGHTS", raising=False)
    with pytest.warns(RuntimeWarning, match="RANDOMLY initialised"):
        src = EfficientNet.from_pretrained("efficientnet-b0")
    with torch.no_grad():
        for p_ in src.parameters():
            p_.add_(0.5)
    f = tmp_path / "b0.pth"
    torch.save(src.state_dict(), f)
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        cam = CamEncode(8, 4, trunk_weights=str(f))
        monkeypatch.setenv("MFB_EFFICIENTNET_B0_WEIGHTS", str(f))
        cam2 = CamEncode(8, 4)
    for k, v in src.state_dict().items():
        assert torch.equal(cam.trunk.state_dict()[k], v) and torch.equal(cam2.trunk.state_dict()[k], v)
    torch.save({"not": torch.zeros(1)}, f)
    with pytest.raises(RuntimeError, match="not an efficientnet_pytorch"):
        CamEncode(8, 4, trunk_weights=str(f))


def test_physics_loss_rotation_term_and_duplicate_stamps():
    """losses.py:129-136 (scripts/eval.py:151 unpacks two values) and the identity shortcut only for strictly
    increasing stamps (argmin returns the first of tied stamps)."""
    from monoforce_b200.losses import physics_loss, rotation_difference
    from oracle.ref_import import reference_available
    g = torch.Generator().manual_seed(1)
    B, T = 3, 12
    Xp, Xg = torch.randn(B, T, 3, generator=g), torch.randn(B, T, 3, generator=g)

    def rots(n):
        q, _ = torch.linalg.qr(torch.randn(B, n, 3, 3, generator=g))
        return q * torch.sign(torch.linalg.det(q))[..., None, None]
    Rp, Rg = rots(T), rots(T)
    ts = torch.arange(T, dtype=torch.float32)[None] * 0.1
    loss, loss_rot = physics_loss((Xp, None, Rp), (Xg, None, Rg), ts, ts, 0.9, rotation_loss=True)
    w = 1. / (1. + 0.9 * ts.unsqueeze(2))
    assert torch.allclose(loss_rot, (rotation_difference(Rp, Rg, reduction='none') * w).mean())
    assert torch.allclose(loss, ((Xp * w - Xg * w) ** 2).mean())
    dup = ts.clone()
    dup[0, 5] = dup[0, 4]                                   # tie: the reference gathers index 4 for stamp 5
    ids = torch.argmin(torch.abs(dup.unsqueeze(1) - dup.unsqueeze(2)), dim=2)
    ref = ((Xp[torch.arange(B).unsqueeze(1), ids] * (1. / (1. + 0.9 * dup.unsqueeze(2))) - Xg * (1. / (1. + 0.9 * dup.unsqueeze(2)))) ** 2).mean()
    assert torch.allclose(physics_loss((Xp,), (Xg,), dup, dup, 0.9), ref)
    if reference_available():
        import importlib.util
        spec = importlib.util.spec_from_file_location("_ref_losses_t", "/root/reference/monoforce/src/monoforce/losses.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        rl, rr = mod.physics_loss((Xp, None, Rp), (Xg, None, Rg), ts, ts, 0.9, rotation_loss=True)
        assert torch.allclose(loss, rl) and torch.allclose(loss_rot, rr)
        assert torch.allclose(physics_loss((Xp,), (Xg,), dup, dup, 0.9), mod.physics_loss((Xp,), (Xg,), dup, dup, 0.9))


def test_physics_loss_matches_reference_definition():
    from monoforce_b200.losses import physics_loss
    from oracle.dphysics_oracle import physics_loss as ref_loss
    g = torch.Generator().manual_seed(0)
    Xp, Xg = torch.randn(5, 40, 3, generator=g), torch.randn(5, 40, 3, generator=g)
    ts = torch.arange(0, 0.4, 0.01)[None][:, :40]
    assert torch.allclose(physics_loss((Xp,), (Xg,), ts, ts, 0.9), ref_loss((Xp,), (Xg,), ts, ts, 0.9), rtol=1e-6)
    gt_ts = torch.tensor([[0.0, 0.1, 0.25, 0.39]])
    assert torch.allclose(physics_loss((Xp,), (Xg[:, :4],), ts, gt_ts, 0.5), ref_loss((Xp,), (Xg[:, :4],), ts, gt_ts, 0.5))


def test_shard_bounds_cover_batch_exactly():
    from monoforce_b200.dist import shard_bounds
    for n, w in ((4096, 8), (65536, 8), (10, 3), (7, 8), (1, 1)):
        spans = [shard_bounds(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from monoforce_b200.dist import shard, gather_costs, best_trajectory, allreduce_map_grads
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    all_costs = torch.arange(n, dtype=torch.float32).flip(0) + 0.5          # the same "global" problem on every rank
    mine = shard(all_costs, rank, world)
    gathered = gather_costs(mine, n)
    idx, val = best_trajectory(mine, n)
    g = torch.full((4, 4), float(rank + 1))
    g2 = torch.arange(3, dtype=torch.float32) * (rank + 1)                   # second map: both travel as one flat all-reduce
    allreduce_map_grads(g, None, g2)
    ok_g2 = torch.equal(g2, torch.arange(3, dtype=torch.float32) * 3) and g.shape == (4, 4) and bool((g == g[0, 0]).all())
    # the two halves of one buffer (what DPhysics' backward returns) are reduced in place through one flat view
    both = torch.stack([torch.full((2, 5), float(rank + 1)), torch.full((2, 5), 10.0 * (rank + 1))])
    ptr = both.data_ptr()
    allreduce_map_grads(both[0], both[1])
    ok_g2 = ok_g2 and both.data_ptr() == ptr and bool((both[0] == 3).all()) and bool((both[1] == 30).all())
    from monoforce_b200.dist import _adjacent_view
    ok_g2 = ok_g2 and _adjacent_view([both[0], both[1]]) is not None and _adjacent_view([both[1], both[0]]) is None
    q.put((rank, torch.equal(gathered, all_costs) and ok_g2, idx, val, float(g[0, 0])))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [64, 37])
def test_two_process_cost_gather_and_grad_allreduce_gloo(n):
    """world_size-2 gloo run of the N>1 host logic: shard -> gather costs -> argmin; all-reduce of map grads."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, idx, val, gsum in res:
        assert ok and idx == n - 1 and val == 0.5 and gsum == 3.0


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the reference's algorithm on the host cores, no GPU) must print ONE JSON line with the
    keys the driver reads; a non-zero rank under torchrun prints nothing and exits 0."""
    import json
    import subprocess
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-sample", "2"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env={**os.environ, "RANK": "0"})
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "config",
                "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    other = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env={**os.environ, "RANK": "1"})
    assert other.returncode == 0 and other.stdout.strip() == ""
