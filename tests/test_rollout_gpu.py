"""GPU parity tests: CUDA rollout (through the C ABI) vs the CPU oracle and the goldens.

Protocol (needed because the rollout is chaotic on rough terrain):
  P1  fp64 kernel vs fp64 oracle, full horizon, every terrain:            <= 1e-9
  P2  fp32 kernel vs the reference's fp32 goldens: cfg1 and flat maps:     <= 1e-4 (north_star tolerance)
      teacher-forced single step (T=1 from random states), all terrains:  <= 1e-5 states, 2e-4 forces
  P3  rough terrain, full horizon: error envelope against the fp64 oracle
  P4  gradients: fp64 adjoint vs the reference's fp64 autograd goldens
"""
import numpy as np
import pytest
import torch

from helpers_mfb import load_golden, make_spec, hill_map, rel_err

pytestmark = pytest.mark.gpu

DEV = "cuda"


ADJOINT_TAPE = True      # flipped by the `adjoint_kernel` fixture: single-sweep (tape) vs three-pass adjoint


@pytest.fixture(params=["warp", "wide"], autouse=True)
def forward_kernel(request, monkeypatch):
    """Every test of this module runs once per forward kernel: K1 (one warp per trajectory, the large-batch kernel) and K1w
    (one CTA per trajectory, one thread per contact point, the small-batch kernel).  The library reads the batch-size threshold
    from MFB_FWD_WIDE_MAX_B at each launch (moving flippers always take K1)."""
    monkeypatch.setenv("MFB_FWD_WIDE_MAX_B", "0" if request.param == "warp" else str(1 << 30))
    yield request.param


@pytest.fixture(params=["sweep", "sweep_wide", "three_pass"])
def adjoint_kernel(request, monkeypatch):
    """Runs an adjoint test once per backward kernel: K2s (single sweep, reads the forward's contact_sum tape) with one warp per
    trajectory, K2s with one CTA per trajectory (the small-batch instantiation; MFB_BWD_WIDE_MAX_B is the batch-size switch) and
    K2 (three-pass, recomputes the tape)."""
    global ADJOINT_TAPE
    ADJOINT_TAPE = request.param != "three_pass"
    monkeypatch.setenv("MFB_BWD_WIDE_MAX_B", str(1 << 30) if request.param == "sweep_wide" else "0")
    yield request.param
    ADJOINT_TAPE = True


def _sim(cfg):
    from monoforce_b200 import DPhysics
    sim = DPhysics(cfg, device=DEV)
    sim.adjoint_tape = ADJOINT_TAPE
    if _FORCE_PER_MAP:
        sim.shared_map = False
    return sim


def _module(robot, grid_res, T, variant="step", dtype=torch.float32):
    from monoforce_b200 import DPhysConfig
    cfg = DPhysConfig(robot=robot, grid_res=grid_res)
    cfg.traj_sim_time = T * cfg.dt
    cfg.use_odeint = (variant == "odeint")
    return _sim(cfg), cfg


def _t(a, dtype=torch.float32):
    return torch.as_tensor(a, dtype=dtype, device=DEV)


def _run_golden(g, dtype=torch.float32, expand=True):
    sim, cfg = _module(str(g["robot"]), float(g["grid_res"]), int(g["T"]), str(g["variant"]), dtype)
    B = g["controls"].shape[0]
    z = _t(g["z"], dtype).unsqueeze(0)
    z = z.expand(B, -1, -1) if expand else z.repeat(B, 1, 1)
    fr = None
    if "friction" in g:
        fr = _t(g["friction"], dtype).unsqueeze(0)
        fr = fr.expand(B, -1, -1) if expand else fr.repeat(B, 1, 1)
    st = None
    if "x0" in g:
        st = tuple(_t(g[k], dtype) for k in ("x0", "xd0", "R0", "om0"))
    return sim(z, _t(g["controls"], dtype), state=st, friction=fr), cfg


# Per-golden tolerances = 3 x the error MEASURED on B200 (tools/measure_parity.py -> profiles/r02_parity_measured.json),
# rounded up to one digit, floor 1e-6; every pose is within 3e-6 and everything within 3e-4 of the reference's fp32 output,
# i.e. inside the north_star 1e-4 for poses with two orders of margin (round 1 used blanket 1e-4 / 20x / 50x bounds).
GOLDEN_TOL = {
    # measured: Xs 1.7e-07, Rs 1.8e-07, Xds 1.4e-07, Omegas 1.1e-06, Fs_keep 3.5e-07, Ff_keep 4.9e-07, Fs_sum 4.0e-07
    "cfg1_marv_flat64_T100": dict(Xs=1e-06, Rs=1e-06, Xds=1e-06, Omegas=4e-06, Fs_keep=2e-06, Ff_keep=2e-06, Fs_sum=2e-06),
    # measured: Xs 5.2e-08, Rs 3.0e-07, Xds 1.8e-07, Omegas 2.6e-06, Fs_keep 2.6e-07, Ff_keep 2.6e-07, Fs_sum 3.9e-07
    "cfg1_tradr_flat64_T100": dict(Xs=1e-06, Rs=1e-06, Xds=1e-06, Omegas=8e-06, Fs_keep=1e-06, Ff_keep=1e-06, Fs_sum=2e-06),
    # measured: Xs 8.5e-07, Rs 7.1e-06, Xds 3.0e-05, Omegas 3.4e-05, Fs_keep 1.2e-05, Ff_keep 9.5e-06, Fs_sum 4.8e-07
    "marv_flat256_T400_B2": dict(Xs=3e-06, Rs=3e-05, Xds=9e-05, Omegas=2e-04, Fs_keep=4e-05, Ff_keep=3e-05, Fs_sum=2e-06),
    # measured: Xs 1.2e-06, Rs 3.1e-05, Xds 5.0e-05, Omegas 2.6e-04, Fs_keep 1.4e-05, Ff_keep 1.7e-05, Fs_sum 6.2e-06
    "marv_ramp256_T400_B3": dict(Xs=4e-06, Rs=1e-04, Xds=2e-04, Omegas=8e-04, Fs_keep=5e-05, Ff_keep=6e-05, Fs_sum=2e-05),
    # measured: Xs 2.8e-06, Rs 2.1e-05, Xds 1.4e-05, Omegas 2.5e-05, Fs_keep 2.7e-05, Ff_keep 1.9e-05, Fs_sum 6.4e-06
    "tradr_ramp128_T300_B3": dict(Xs=9e-06, Rs=7e-05, Xds=5e-05, Omegas=8e-05, Fs_keep=9e-05, Ff_keep=6e-05, Fs_sum=2e-05),
    # measured: Xs 2.4e-07, Rs 2.4e-07, Xds 4.6e-06, Omegas 1.9e-06, Fs_keep 6.9e-07, Ff_keep 4.5e-07, Fs_sum 1.9e-07
    "marv_ramp128_odeint_T200_B3": dict(Xs=1e-06, Rs=1e-06, Xds=2e-05, Omegas=6e-06, Fs_keep=3e-06, Ff_keep=2e-06, Fs_sum=1e-06),
    # measured: Xs 2.8e-07, Rs 1.2e-06, Xds 2.7e-06, Omegas 2.1e-06, Fs_keep 1.6e-06, Ff_keep 8.8e-07, Fs_sum 3.6e-06
    "marv_hill128_T100_B4": dict(Xs=1e-06, Rs=4e-06, Xds=9e-06, Omegas=7e-06, Fs_keep=5e-06, Ff_keep=3e-06, Fs_sum=2e-05),
    # measured: Xs 2.8e-07, Rs 1.1e-06, Xds 2.2e-06, Omegas 9.6e-07, Fs_keep 1.9e-06, Ff_keep 8.2e-06, Fs_sum 2.9e-06
    "marv_noise128_state_fric_T100_B4": dict(Xs=1e-06, Rs=4e-06, Xds=7e-06, Omegas=3e-06, Fs_keep=6e-06, Ff_keep=3e-05, Fs_sum=9e-06),
    # measured: Xs 1.8e-07, Rs 1.7e-06, Xds 2.6e-06, Omegas 3.6e-06, Fs_keep 2.1e-06, Ff_keep 2.1e-06, Fs_sum 5.7e-06
    "tradr_noise128_state_fric_T100_B4": dict(Xs=1e-06, Rs=5e-06, Xds=8e-06, Omegas=2e-05, Fs_keep=7e-06, Ff_keep=7e-06, Fs_sum=2e-05),
    # measured: Xs 1.9e-07, Rs 2.5e-07, Xds 1.7e-06, Omegas 7.5e-07, Fs_keep 3.4e-07, Ff_keep 4.8e-07, Fs_sum 4.8e-07
    "marv_hill128_odeint_T60_B2": dict(Xs=1e-06, Rs=1e-06, Xds=6e-06, Omegas=3e-06, Fs_keep=2e-06, Ff_keep=2e-06, Fs_sum=2e-06),
    # measured: Xs 2.3e-07, Rs 4.8e-07, Xds 1.7e-06, Omegas 1.1e-06, Fs_keep 1.4e-06, Ff_keep 1.6e-06, Fs_sum 8.9e-06
    "marv_hill128_joints_T60_B2": dict(Xs=1e-06, Rs=2e-06, Xds=6e-06, Omegas=4e-06, Fs_keep=5e-06, Ff_keep=5e-06, Fs_sum=3e-05),
}
FWD_GOLDENS = [k for k in GOLDEN_TOL if "joints" not in k]


def _check_golden(states, forces, g, tol):
    Xs, Xds, Rs, Oms = states
    Fs, Ff = forces
    keep = g["F_keep_steps"]
    got = {"Xs": rel_err(Xs, g["Xs"]), "Rs": rel_err(Rs, g["Rs"]), "Xds": rel_err(Xds, g["Xds"]),
           "Omegas": rel_err(Oms, g["Omegas"]), "Fs_keep": rel_err(Fs[:, keep], g["Fs_keep"]),
           "Ff_keep": rel_err(Ff[:, keep], g["Ff_keep"]), "Fs_sum": rel_err(Fs.double().sum(dim=2), g["Fs_sum"])}
    for k, v in got.items():
        assert v < tol[k], (k, v, tol[k])
    assert got["Xs"] < 1e-4 and got["Rs"] < 1e-4            # north_star: poses within 1e-4 relative


@pytest.mark.parametrize("name", FWD_GOLDENS)
def test_fp32_kernel_vs_reference_goldens(name):
    """P2: every output within 3x the measured distance from the reference's own fp32 CPU output (poses <= 1e-4)."""
    g = load_golden(name)
    (states, forces), cfg = _run_golden(g)
    _check_golden(states, forces, g, GOLDEN_TOL[name])


@pytest.mark.parametrize("name", ["marv_hill128_T100_B4", "marv_noise128_state_fric_T100_B4", "marv_ramp128_odeint_T200_B3"])
def test_repeated_maps_equal_shared_map(name):
    """B materialised copies of a map (what the reference's callers pass) == one shared map, bit for bit: both when the
    copies are recognised on the device (default) and when every trajectory really reads its own copy."""
    g = load_golden(name)
    (s1, f1), _ = _run_golden(g, expand=True)
    (s2, f2), _ = _run_golden(g, expand=False)                      # repeat()ed maps, recognised on the device
    global _FORCE_PER_MAP
    _FORCE_PER_MAP = True
    try:
        (s3, f3), _ = _run_golden(g, expand=False)                  # shared_map=False: B cell tables
    finally:
        _FORCE_PER_MAP = False
    for a, b, c in zip(s1 + f1, s2 + f2, s3 + f3):
        assert torch.equal(a, b) and torch.equal(a, c)


def test_repeated_map_detection_rules():
    """Copies are merged only when that cannot change what the caller observes: equal VALUES, and no gradient flowing
    into the individual copies (a leaf of B copies must receive B separate gradients, as in the reference)."""
    sim, cfg = _module("tradr", 0.4, 20)
    B = 6
    z = hill_map(cfg).to(DEV)
    controls = torch.rand(B, 20, 2, device=DEV)
    view = lambda grid: sim._shared_view(grid, B).shape[0]
    assert view(z.unsqueeze(0)) == 1 and view(z.unsqueeze(0).expand(B, -1, -1)) == 1
    rep = z.repeat(B, 1, 1)
    assert view(rep) == 1
    rep2 = rep.clone()
    rep2[B - 1, 3, 3] += 1e-3
    assert view(rep2) == B                                            # one differing cell in the LAST copy
    leaf = rep.clone().requires_grad_(True)
    assert view(leaf) == B                                            # gradients per copy must stay per copy
    with torch.no_grad():
        assert view(leaf) == 1
    sim.shared_map = False
    assert view(rep) == B
    sim.shared_map = True
    assert view(rep2) == 1                                            # the caller's promise
    sim.shared_map = None
    st, _ = sim(leaf, controls)
    st[0].sum().backward()
    assert leaf.grad.shape == (B, 32, 32) and (leaf.grad.flatten(1).abs().sum(1) > 0).all()


_FORCE_PER_MAP = False


def _random_case(cfg, B, T, seed, dtype, terrain="noise", with_state=True, with_fric=True):
    gen = torch.Generator().manual_seed(seed)
    if terrain == "flat":
        z = torch.zeros_like(cfg.x_grid).to(dtype)
    elif terrain == "hill":
        z = hill_map(cfg, dtype=dtype)
    else:
        z = hill_map(cfg, 0.02, seed, dtype=dtype)
    v = torch.rand(B, generator=gen, dtype=dtype) * 2 - 1
    w = torch.rand(B, generator=gen, dtype=dtype) * 4 - 2
    controls = torch.stack([v[:, None].repeat(1, T), w[:, None].repeat(1, T)], -1)
    controls = controls + 0.05 * torch.randn(controls.shape, generator=gen, dtype=dtype)
    fr = (0.3 + 0.7 * torch.rand(z.shape, generator=gen, dtype=dtype)) if with_fric else None
    st = None
    if with_state:
        x = torch.randn(B, 3, generator=gen, dtype=dtype) * 1.0
        xd = torch.randn(B, 3, generator=gen, dtype=dtype) * 0.3
        a = torch.rand(B, generator=gen, dtype=dtype) * 6.28
        tilt = torch.randn(B, generator=gen, dtype=dtype) * 0.1
        Rz = torch.zeros(B, 3, 3, dtype=dtype)
        Rz[:, 0, 0] = a.cos(); Rz[:, 0, 1] = -a.sin(); Rz[:, 1, 0] = a.sin(); Rz[:, 1, 1] = a.cos(); Rz[:, 2, 2] = 1
        Ry = torch.zeros(B, 3, 3, dtype=dtype)
        Ry[:, 0, 0] = tilt.cos(); Ry[:, 0, 2] = tilt.sin(); Ry[:, 2, 0] = -tilt.sin(); Ry[:, 2, 2] = tilt.cos(); Ry[:, 1, 1] = 1
        om = torch.randn(B, 3, generator=gen, dtype=dtype) * 0.2
        st = (x, xd, Rz @ Ry, om)
    return z, controls, fr, st


def _both(robot, grid_res, T, B, seed, dtype, variant="step", **kw):
    from oracle import dphysics_oracle as O
    sim, cfg = _module(robot, grid_res, T, variant, dtype)
    z, controls, fr, st = _random_case(cfg, B, T, seed, dtype, **kw)
    ref = O.rollout(make_spec(cfg), z.repeat(B, 1, 1), controls, state=st,
                    friction=None if fr is None else fr.repeat(B, 1, 1), variant=variant, dtype=dtype)
    dev_state = None if st is None else tuple(s.to(DEV) for s in st)
    out = sim(z.to(DEV).unsqueeze(0).expand(B, -1, -1), controls.to(DEV), state=dev_state,
              friction=None if fr is None else fr.to(DEV).unsqueeze(0).expand(B, -1, -1))
    return out, ref, dev_state


@pytest.mark.parametrize("robot,grid_res,terrain,variant", [
    ("marv", 0.1, "noise", "step"), ("marv", 0.05, "hill", "step"), ("tradr", 0.1, "noise", "step"),
    ("marv", 0.2, "flat", "step"), ("marv", 0.1, "noise", "odeint"), ("tradr", 0.2, "hill", "odeint")])
def test_fp64_kernel_matches_fp64_oracle_full_horizon(robot, grid_res, terrain, variant):
    """P1: the algorithm is the reference's - double precision agrees to 1e-9 over the whole horizon."""
    T = 200
    (states, forces), (rs, rf), dev_state = _both(robot, grid_res, T, 6, 11, torch.float64, variant, terrain=terrain)
    for a, b in zip(states, rs):
        assert rel_err(a, b) < 1e-9
    for a, b in zip(forces, rf):
        assert rel_err(a, b) < 1e-8


def _per_traj_rel(a, b):
    a = a.double().cpu().reshape(a.shape[0], -1)
    b = b.double().cpu().reshape(b.shape[0], -1)
    return (a - b).abs().amax(dim=1) / b.abs().amax().clamp_min(1e-12)


@pytest.mark.parametrize("robot", ["marv", "tradr"])
@pytest.mark.parametrize("terrain", ["flat", "hill", "noise"])
def test_fp32_teacher_forced_single_step(robot, terrain):
    """P2: one step from 256 random states against the fp32 oracle: states <= 1e-5, forces <= 2e-4.

    The maps here (random friction, hill, noise) make the reference's sampling jump at cell borders
    (dphysics.py:442-445), so a trajectory whose contact point sits within an ulp of a border may land in
    the neighbouring cell in one of the two fp32 evaluations.  Expected: ~1 such trajectory per 256; we
    allow 2 % outliers and hold every other trajectory to the tolerance."""
    (states, forces), (rs, rf), _ = _both(robot, 0.1, 1, 256, 5, torch.float32, terrain=terrain)
    B = states[0].shape[0]
    bad = torch.zeros(B, dtype=torch.bool)
    for a, b in zip(states, rs):
        bad |= _per_traj_rel(a, b) >= 1e-5
    for a, b in zip(forces, rf):
        bad |= _per_traj_rel(a, b) >= 2e-4
    assert int(bad.sum()) <= max(1, B // 50), f"{int(bad.sum())} of {B} trajectories off"
    # outliers are border events, not garbage: still close in absolute terms
    for a, b in zip(states, rs):
        assert rel_err(a, b) < 5e-3


def test_fp32_teacher_forced_single_step_continuous_terrain():
    """Same check on terrains where the reference's sampling is continuous (flat / diagonal ramp, friction 1):
    every trajectory within tolerance, no outliers allowed."""
    from oracle import dphysics_oracle as O
    for robot in ("marv", "tradr"):
        for ramp in (0.0, 0.25):
            sim, cfg = _module(robot, 0.1, 1)
            z = ramp * (cfg.x_grid + cfg.y_grid)
            _, controls, _, st = _random_case(cfg, 256, 1, 13, torch.float32)
            rs, rf = O.rollout(make_spec(cfg), z.repeat(256, 1, 1), controls, state=st)
            ks, kf = sim(z.to(DEV).unsqueeze(0), controls.to(DEV), state=tuple(s.to(DEV) for s in st))
            for a, b in zip(ks, rs):
                assert rel_err(a, b) < 1e-5
            for a, b in zip(kf, rf):
                assert rel_err(a, b) < 2e-4


def test_fp32_rough_terrain_error_envelope():
    """P3: on rough terrain the fp32 kernel may drift from the fp64 truth no more than a small multiple of
    what the reference's own fp32 arithmetic (the oracle in fp32) drifts."""
    from oracle import dphysics_oracle as O
    T, B = 400, 32
    sim, cfg = _module("marv", 0.05, T)
    z, controls, fr, st = _random_case(cfg, B, T, 3, torch.float64, terrain="noise", with_state=False, with_fric=False)
    spec = make_spec(cfg)
    zz = z.repeat(B, 1, 1)
    truth = O.rollout(spec, zz, controls, dtype=torch.float64)[0][0]
    ref32 = O.rollout(spec, zz.float(), controls.float(), dtype=torch.float32)[0][0]
    (Xs, _, _, _), _ = sim(z.float().to(DEV).unsqueeze(0), controls.float().to(DEV))
    scale = truth.abs().amax(dim=(1, 2)).clamp_min(1e-6)
    e_ref = ((ref32.double() - truth).abs().amax(dim=(1, 2)) / scale)
    e_ker = ((Xs.double().cpu() - truth).abs().amax(dim=(1, 2)) / scale)
    print(f"fp32-vs-fp64 relative position error over {B} trajectories, T={T}: "
          f"reference median {e_ref.median():.2e} max {e_ref.max():.2e}; kernel median {e_ker.median():.2e} max {e_ker.max():.2e}")
    assert e_ker.median() <= 5 * e_ref.median() + 1e-6
    assert e_ker.max() <= 10 * e_ref.max() + 1e-5


@pytest.mark.parametrize("name", ["grad64_marv_noise128_T40_B2", "grad64_tradr_noise64_T40_B2"])
def test_fp64_adjoint_matches_reference_autograd(name, adjoint_kernel):
    """P4: hand-written adjoint == autograd of the unmodified reference (fp64 goldens)."""
    g = load_golden(name)
    dtype = torch.float64
    sim, cfg = _module(str(g["robot"]), float(g["grid_res"]), int(g["T"]), str(g["variant"]), dtype)
    B = g["controls"].shape[0]
    z = _t(g["z"], dtype).requires_grad_(True)
    fr = _t(g["friction"], dtype).requires_grad_(True)
    controls = _t(g["controls"], dtype).requires_grad_(True)
    st = [_t(g[k], dtype).requires_grad_(True) for k in ("x0", "xd0", "R0", "om0")]
    states, forces = sim(z.unsqueeze(0).expand(B, -1, -1), controls, state=tuple(s * 1.0 for s in st),
                         friction=fr.unsqueeze(0).expand(B, -1, -1))
    outs = list(states) + list(forces)
    loss = sum(float(s) * (o * _t(g[f"w{i}"], dtype)).sum() for i, (s, o) in enumerate(zip(g["scales"], outs)))
    assert abs(loss.item() - float(g["loss"])) <= 1e-9 * max(1.0, abs(float(g["loss"])))
    loss.backward()
    assert rel_err(z.grad, g["g_z"]) < 1e-7
    assert rel_err(fr.grad, g["g_friction"]) < 1e-7
    assert rel_err(controls.grad, g["g_controls"]) < 1e-7
    for t, k in zip(st, ("g_x0", "g_xd0", "g_R0", "g_om0")):
        assert rel_err(t.grad, g[k]) < 1e-7, k


@pytest.mark.parametrize("variant", ["step", "odeint"])
def test_fp64_adjoint_matches_oracle_autograd(variant, adjoint_kernel):
    """P4 on a second objective (positions only, default initial state -> gradient reaches controls[:,0]
    through the initial velocity too), both integrator variants."""
    from oracle import dphysics_oracle as O
    dtype = torch.float64
    T, B = 30, 3
    sim, cfg = _module("marv", 0.2, T, variant, dtype)
    z, controls, fr, _ = _random_case(cfg, B, T, 21, dtype, with_state=False)
    spec = make_spec(cfg)
    zr, fr_r, cr = z.clone().requires_grad_(True), fr.clone().requires_grad_(True), controls.clone().requires_grad_(True)
    rs, rf = O.rollout(spec, zr.unsqueeze(0).expand(B, -1, -1), cr, friction=fr_r.unsqueeze(0).expand(B, -1, -1),
                       variant=variant, dtype=dtype)
    wgt = torch.linspace(0.2, 1.0, T, dtype=dtype).view(1, T, 1)
    (rs[0] * wgt).pow(2).sum().add((rf[0] * 1e-3).pow(2).sum()).backward()
    zk, fk, ck = (t.clone().to(DEV).requires_grad_(True) for t in (z, fr, controls))
    ks, kf = sim(zk.unsqueeze(0).expand(B, -1, -1), ck, friction=fk.unsqueeze(0).expand(B, -1, -1))
    (ks[0] * wgt.to(DEV)).pow(2).sum().add((kf[0] * 1e-3).pow(2).sum()).backward()
    assert rel_err(zk.grad, zr.grad) < 1e-7
    assert rel_err(fk.grad, fr_r.grad) < 1e-7
    assert rel_err(ck.grad, cr.grad) < 1e-7


def test_fp32_adjoint_close_to_fp64(adjoint_kernel):
    """fp32 adjoint vs fp64 autograd, T=50 on the noisy hill.  Measured on B200 (profiles/r02_parity_measured.json):
    see ADJOINT_T50_TOL, set to <= 10x the measured error (was a blanket 2e-2 in round 1)."""
    r = fp32_adjoint_errors(ADJOINT_TAPE)
    print(r)
    assert r["g_z"] < ADJOINT_T50_TOL["g_z"] and r["g_controls"] < ADJOINT_T50_TOL["g_controls"], r


ADJOINT_T50_TOL = {"g_z": 1e-5, "g_controls": 1e-5}    # measured 5.7e-7 .. 7.9e-7 (both adjoint kernels)


def fp32_adjoint_errors(tape):
    """T=50, B=4, noisy hill: fp32 adjoint vs fp64 autograd of the oracle -> (d/dz, d/dcontrols) relative errors."""
    from monoforce_b200 import DPhysics
    from oracle import dphysics_oracle as O
    T, B = 50, 4
    sim, cfg = _module("marv", 0.1, T)
    sim.adjoint_tape = tape
    z, controls, fr, _ = _random_case(cfg, B, T, 8, torch.float64, with_state=False)
    spec = make_spec(cfg)
    zr, cr = z.clone().requires_grad_(True), controls.clone().requires_grad_(True)
    rs, _ = O.rollout(spec, zr.unsqueeze(0).expand(B, -1, -1), cr, friction=fr.unsqueeze(0).expand(B, -1, -1),
                      dtype=torch.float64)
    rs[0].pow(2).mean().backward()
    zk, ck = z.float().to(DEV).requires_grad_(True), controls.float().to(DEV).requires_grad_(True)
    ks, _ = sim(zk.unsqueeze(0), ck, friction=fr.float().to(DEV).unsqueeze(0))
    ks[0].pow(2).mean().backward()
    return {"g_z": rel_err(zk.grad, zr.grad), "g_controls": rel_err(ck.grad, cr.grad)}


GRAD_ENVELOPE_CASES = ("bench", "hill", "noise")


def _l2rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


_ENVELOPE_TRUTH = {}


def _envelope_truth(case, B, T):
    """Inputs + fp64 / fp32 oracle autograd of one envelope case (cached: both adjoint kernels are checked against it)."""
    if (case, B, T) in _ENVELOPE_TRUTH:
        return _ENVELOPE_TRUTH[(case, B, T)]
    from bench import synth_inputs
    from oracle import dphysics_oracle as O
    d = synth_inputs(B, seed=5)
    cfg = d["cfg"]
    assert int(cfg.traj_sim_time / cfg.dt) == T
    spec = make_spec(cfg)
    g = torch.Generator().manual_seed(9)
    z_gt = d["z_gt"].double()
    if case == "bench":
        z0, f0 = d["z0"].double(), d["fr0"].double()
    elif case == "hill":
        z0, f0 = 0.8 * z_gt, torch.full_like(z_gt, 0.7)
    else:
        z0 = z_gt + 0.02 * torch.randn(z_gt.shape, generator=g, dtype=torch.float64)
        f0 = 0.3 + 0.7 * torch.rand(z_gt.shape, generator=g, dtype=torch.float64)
    ts = d["ts"].double()
    controls = d["controls"].double()
    with torch.no_grad():
        gt = O.rollout(spec, z_gt.unsqueeze(0).expand(B, -1, -1), controls, dtype=torch.float64)[0]

    def oracle(dtype):
        z, fr, c = (t.to(dtype).clone().requires_grad_(True) for t in (z0, f0, controls))
        st, _ = O.rollout(spec, z.unsqueeze(0).expand(B, -1, -1), c, friction=fr.unsqueeze(0).expand(B, -1, -1), dtype=dtype)
        loss = O.physics_loss(st, tuple(t.to(dtype) for t in gt), ts.to(dtype), ts.to(dtype), 0.9)
        loss.backward()
        return loss.item(), z.grad, fr.grad, c.grad
    out = dict(cfg=cfg, z0=z0, f0=f0, controls=controls, ts=ts, gt=gt, o64=oracle(torch.float64), o32=oracle(torch.float32))
    _ENVELOPE_TRUTH[(case, B, T)] = out
    return out


def gradient_envelope(case, B=12, T=400):
    """fp32 adjoint at the BENCH horizon (T=400, 256x256 map, bench.py controls recipe, physics_loss objective) against
    fp64 autograd of the oracle, next to the error of the reference's OWN fp32 autograd (the fp32 oracle) against the
    same truth.  Cases: 'bench' = exactly bench.py's step (flat initial map, friction 0.5, targets from the hill),
    'hill' = gradient taken on the smooth hill, 'noise' = hill + 2 cm noise (discontinuous sampling, chaotic)."""
    from monoforce_b200.losses import physics_loss
    tr = _envelope_truth(case, B, T)
    sim = _sim(tr["cfg"])
    sim.return_forces = False
    zk, fk, ck = (t.float().to(DEV).requires_grad_(True) for t in (tr["z0"], tr["f0"], tr["controls"]))
    tk = tr["ts"].float().to(DEV)
    st, _ = sim(zk.unsqueeze(0), ck, friction=fk.unsqueeze(0))
    lk = physics_loss(st, tuple(t.float().to(DEV) for t in tr["gt"]), tk, tk, 0.9)
    lk.backward()
    l64, *g64 = tr["o64"]
    l32, *g32 = tr["o32"]
    out = {"loss64": l64, "loss_err_kernel": abs(lk.item() - l64) / abs(l64), "loss_err_ref32": abs(l32 - l64) / abs(l64)}
    for name, a64, a32, ak in zip(("g_z", "g_friction", "g_controls"), g64, g32, (zk.grad, fk.grad, ck.grad)):
        out[name] = {"kernel_max": rel_err(ak, a64), "ref32_max": rel_err(a32, a64),
                     "kernel_l2": _l2rel(ak, a64), "ref32_l2": _l2rel(a32, a64)}
    return out


# measured on B200 (tools/measure_parity.py, profiles/r02_parity_measured.json); bound = 3x the reference's own fp32
# error + a floor for quantities the reference happens to get to the last bit
@pytest.mark.parametrize("case", GRAD_ENVELOPE_CASES)
def test_fp32_gradient_envelope_T400(case, adjoint_kernel):
    """||g_kernel32 - g_oracle64|| <= 3 x ||g_oracle32(autograd) - g_oracle64|| for d/dz_grid,
    d/dfriction, d/dcontrols at T=400 on the bench map and recipe (and on the hill / noisy hill)."""
    r = gradient_envelope(case)
    print(case, r)
    assert r["loss_err_kernel"] <= 3 * r["loss_err_ref32"] + 1e-4
    for k in ("g_z", "g_friction", "g_controls"):
        assert r[k]["kernel_l2"] <= 3 * r[k]["ref32_l2"] + GRAD_FLOOR[case], (k, r[k])
        assert r[k]["kernel_max"] <= ENVELOPE_MAX_FACTOR * r[k]["ref32_max"] + GRAD_FLOOR[case], (k, r[k])


# Measured on B200 (profiles/r02_parity_measured.json), kernel error / reference-fp32 error, both against fp64 autograd:
#   bench  g_z 1.00 (l2) 1.00 (max)   g_friction 0.94 / 0.99   g_controls 0.79 / 0.73     (reference's own g_z error: 3 % max-rel!)
#   hill   g_z 1.7 / 3.5              g_friction 1.7 / 2.9     g_controls 2.4 / 3.3
#   noise  g_z 1.15 / 1.23            g_friction 1.3 / 1.06    g_controls 1.26 / 1.00     (both ~6-30 % off: chaotic)
# The l2 norm is held to 3x as asked; the max norm of 12 trajectories x 65k cells is a single worst element, measured up to 3.5x.
ENVELOPE_MAX_FACTOR = 5.0
GRAD_FLOOR = {"bench": 1e-4, "hill": 1e-4, "noise": 1e-3}


def bench_forward_envelope(case, B=64):
    """Forward parity on EXACTLY bench.py's inputs (synth_inputs: 256x256 hill, shooting controls, T=400), a 64-trajectory
    slice: kernel fp32 and reference fp32 (oracle) against the fp64 oracle, per trajectory."""
    from bench import synth_inputs
    from oracle import dphysics_oracle as O
    d = synth_inputs(4096, seed=0)
    cfg = d["cfg"]
    spec = make_spec(cfg)
    idx = torch.arange(0, 4096, 4096 // B)
    controls = d["controls"][idx]
    z, fr = (d["z_gt"], torch.ones_like(d["z_gt"])) if case == "hill" else (d["z0"], d["fr0"])
    zz = lambda dt: z.to(dt).unsqueeze(0).expand(B, -1, -1)
    ff = lambda dt: fr.to(dt).unsqueeze(0).expand(B, -1, -1)
    truth = O.rollout(spec, zz(torch.float64), controls.double(), friction=ff(torch.float64), dtype=torch.float64)[0]
    ref32 = O.rollout(spec, zz(torch.float32), controls, friction=ff(torch.float32), dtype=torch.float32)[0]
    sim = _sim(cfg)
    with torch.no_grad():
        ker, _ = sim(z.to(DEV).unsqueeze(0), controls.to(DEV), friction=fr.to(DEV).unsqueeze(0))
    out = {}
    for name, t, r, k in zip(("Xs", "Xds", "Rs", "Omegas"), truth, ref32, ker):
        scale = t.abs().amax().clamp_min(1e-9)
        e_r = (r.double() - t).abs().flatten(1).amax(1) / scale
        e_k = (k.double().cpu() - t).abs().flatten(1).amax(1) / scale
        out[name] = {"ref32_median": e_r.median().item(), "ref32_max": e_r.max().item(),
                     "kernel_median": e_k.median().item(), "kernel_max": e_k.max().item(),
                     "kernel_vs_ref32_max": rel_err(k, r)}
    return out


@pytest.mark.parametrize("case", ["hill", "flat"])
def test_forward_envelope_on_bench_inputs(case):
    """full-horizon oracle comparison on the bench workload itself (smooth hill, 256^2, T=400,
    marv, bench controls): the kernel's fp32 drift from the fp64 truth stays within 3x the reference's own."""
    r = bench_forward_envelope(case)
    print(case, r)
    for k, v in r.items():
        assert v["kernel_median"] <= 3 * v["ref32_median"] + 1e-6, (k, v)
        assert v["kernel_max"] <= 3 * v["ref32_max"] + 1e-5, (k, v)


@pytest.mark.parametrize("variant", ["step", "odeint"])
def test_forward_kernel_dispatch_by_batch_size(variant, monkeypatch, forward_kernel):
    """Without MFB_FWD_WIDE_MAX_B the library picks K1w for planner-size batches and K1 for large ones (rollout_fwd.cu:
    wide_max_b): the default's output is bit-identical to the forced kernel it should have picked, and the two kernels
    agree to fp32 round-off (they associate the per-step sums differently)."""
    if forward_kernel != "warp":
        pytest.skip("one run is enough: the test sets the threshold itself")
    T = 60
    sim, cfg = _module("marv", 0.1, T, variant)
    z = hill_map(cfg).to(DEV)[None]

    def run(B, thr):
        if thr is None:
            monkeypatch.delenv("MFB_FWD_WIDE_MAX_B", raising=False)
        else:
            monkeypatch.setenv("MFB_FWD_WIDE_MAX_B", thr)
        ctrl = torch.stack([torch.rand(B, 1, generator=torch.Generator().manual_seed(B)) * 2 - 1,
                            torch.rand(B, 1, generator=torch.Generator().manual_seed(B + 1)) * 2 - 1], -1).repeat(1, T, 1).to(DEV)
        with torch.no_grad():
            st, fo = sim(z, ctrl)
        return [t.clone() for t in st + fo]
    for B, picked in ((64, str(1 << 30)), (2048, "0")):
        default, wide, warp = run(B, None), run(B, str(1 << 30)), run(B, "0")
        forced = wide if picked != "0" else warp
        assert all(torch.equal(a, b) for a, b in zip(default, forced)), (B, picked)
        assert not all(torch.equal(a, b) for a, b in zip(wide, warp))            # they are different kernels
        for a, b, tol in zip(wide[:4], warp[:4], (1e-4, 5e-3, 1e-4, 5e-3)):      # poses 1e-4; velocities jump at contact changes
            assert rel_err(a, b) < tol


@pytest.mark.parametrize("variant", ["step", "odeint"])
def test_graph_replay_of_the_planner_call_matches_eager(variant):
    """DPhysics.graphed: one captured no_grad call replayed with new maps / controls gives bit-identical states, forces and
    fused cost to the eager call (same kernels, same launch parameters); repeated map copies are read as one shared map."""
    T, B = 40, 64
    sim, cfg = _module("marv", 0.1, T, variant)
    sim.fused_cost = variant == "step"
    g = torch.Generator().manual_seed(21)
    maps = [hill_map(cfg, noise=0.02, seed=k).to(DEV)[None] for k in range(3)]
    ctrls = [torch.stack([torch.rand(B, 1, generator=g) * 2 - 1, torch.rand(B, 1, generator=g) * 2 - 1], -1).repeat(1, T, 1).to(DEV)
             for _ in range(2)]
    run = sim.graphed(maps[0].repeat(B, 1, 1), ctrls[0])                 # what monoforce_node.py passes: B copies of the map
    for z, u in ((maps[0], ctrls[0]), (maps[1], ctrls[0]), (maps[2], ctrls[1])):
        st_g, fo_g = run(z_grid=z.repeat(B, 1, 1), controls=u)
        got = [t.clone() for t in st_g + fo_g]
        cost_g = None if run.cost is None else run.cost.clone()
        with torch.no_grad():
            st_e, fo_e = sim(z, u)
        for a, b in zip(got, st_e + fo_e):
            assert torch.equal(a, b)
        if variant == "step":
            assert torch.equal(cost_g, sim.last_cost)
    with pytest.raises(ValueError, match="shape"):
        run(controls=ctrls[0][:, :-1])


def test_adjoint_kernel_dispatch_by_batch_size(monkeypatch):
    """Without MFB_BWD_WIDE_MAX_B small batches take the one-CTA-per-trajectory shape of the single-sweep adjoint and large ones
    the one-warp-per-trajectory shape: d/dcontrols (written without atomics) of the default call is bit-identical to the forced
    shape it should have picked; the map gradients (accumulated with atomics) agree to round-off."""
    T = 50
    sim, cfg = _module("marv", 0.1, T, "step")
    z0 = hill_map(cfg).to(DEV)[None]

    def run(B, thr):
        if thr is None:
            monkeypatch.delenv("MFB_BWD_WIDE_MAX_B", raising=False)
        else:
            monkeypatch.setenv("MFB_BWD_WIDE_MAX_B", thr)
        gen = torch.Generator().manual_seed(B)
        ctrl = torch.stack([torch.rand(B, 1, generator=gen) * 2 - 1, torch.rand(B, 1, generator=gen) * 2 - 1], -1)
        ctrl = ctrl.repeat(1, T, 1).to(DEV).requires_grad_(True)
        z = z0.clone().requires_grad_(True)
        (Xs, _, Rs, _), _ = sim(z, ctrl)
        (Xs.sum() + Rs.sum()).backward()
        return ctrl.grad.clone(), z.grad.clone()
    for B, picked in ((64, str(1 << 30)), (2048, "0")):
        (gc_d, gz_d), (gc_w, gz_w), (gc_n, gz_n) = run(B, None), run(B, str(1 << 30)), run(B, "0")
        gc_f, gz_f = (gc_w, gz_w) if picked != "0" else (gc_n, gz_n)
        assert torch.equal(gc_d, gc_f), (B, picked)
        assert not torch.equal(gc_w, gc_n)                                   # different kernels, different association
        assert rel_err(gc_w, gc_n) < 1e-3 and rel_err(gz_w, gz_n) < 1e-3 and rel_err(gz_d, gz_f) < 1e-4


def test_per_trajectory_maps_and_off_map_clamp():
    """Distinct map per trajectory (training path) + robots that start outside the map exercise the
    reference's flat-index clamp (dphysics.py:432-435)."""
    from oracle import dphysics_oracle as O
    dtype = torch.float64
    T, B = 40, 5
    sim, cfg = _module("tradr", 0.4, T, dtype=dtype)
    gen = torch.Generator().manual_seed(4)
    H = cfg.x_grid.shape[0]
    z = 0.2 * torch.randn(B, H, H, generator=gen, dtype=dtype)
    fr = 0.3 + 0.7 * torch.rand(B, H, H, generator=gen, dtype=dtype)
    _, controls, _, st = _random_case(cfg, B, T, 9, dtype)
    x = st[0].clone()
    x[0, 0] = 6.9; x[1, 1] = -7.3; x[2, 0] = -6.45; x[3, :2] = torch.tensor([6.35, 6.38], dtype=dtype)
    st = (x, st[1], st[2], st[3])
    rs, rf = O.rollout(make_spec(cfg), z, controls, state=st, friction=fr, dtype=dtype)
    ks, kf = sim(z.to(DEV), controls.to(DEV), state=tuple(s.to(DEV) for s in st), friction=fr.to(DEV))
    for a, b in zip(ks + kf, rs + rf):
        assert rel_err(a, b) < 1e-8


@pytest.mark.parametrize("variant", ["step", "odeint"])
@pytest.mark.parametrize("shared", [False, True])
def test_adjoint_off_map_and_per_trajectory_maps(variant, shared, adjoint_kernel):
    """Gradients when contact points leave the map (the reference's clamped flat indices, dphysics.py:432-435) and
    when every trajectory has its own map (the training path): d/dz_grid, d/dfriction, d/dcontrols and d/dstate
    vs autograd of the fp64 oracle.  All output gradients are dense (states and both force tensors)."""
    from oracle import dphysics_oracle as O
    dtype = torch.float64
    T, B = 25, 6
    sim, cfg = _module("tradr", 0.4, T, variant, dtype)
    gen = torch.Generator().manual_seed(14)
    H = cfg.x_grid.shape[0]
    nm = 1 if shared else B
    z = 0.2 * torch.randn(nm, H, H, generator=gen, dtype=dtype)
    fr = 0.3 + 0.7 * torch.rand(nm, H, H, generator=gen, dtype=dtype)
    _, controls, _, st = _random_case(cfg, B, T, 19, dtype)
    x = st[0].clone()
    x[0, 0] = 6.9; x[1, 1] = -7.3; x[2, 0] = -6.45; x[3, :2] = torch.tensor([6.35, 6.38], dtype=dtype)
    x[4, :2] = torch.tensor([-6.2, 6.3], dtype=dtype)
    st = (x, st[1], st[2], st[3])
    gw = [torch.randn(B, T, 3, generator=gen, dtype=dtype), torch.randn(B, T, 3, generator=gen, dtype=dtype),
          torch.randn(B, T, 3, 3, generator=gen, dtype=dtype), torch.randn(B, T, 3, generator=gen, dtype=dtype)]
    fw = 1e-3 * torch.randn(B, T, cfg.robot_points.shape[0], 3, generator=gen, dtype=dtype)

    def objective(states, forces, dev):
        l = sum((o * w.to(dev)).sum() for o, w in zip(states, gw))
        return l + (forces[0] * fw.to(dev)).sum() + 0.5 * (forces[1] * fw.to(dev)).sum()

    leaves_r = [t.clone().requires_grad_(True) for t in (z, fr, controls) + st]
    zr, frr = leaves_r[0], leaves_r[1]
    rs, rf = O.rollout(make_spec(cfg), zr.expand(B, -1, -1) if shared else zr, leaves_r[2],
                       state=tuple(t * 1.0 for t in leaves_r[3:]), friction=frr.expand(B, -1, -1) if shared else frr,
                       variant=variant, dtype=dtype)
    objective(rs, rf, "cpu").backward()
    leaves_k = [t.clone().to(DEV).requires_grad_(True) for t in (z, fr, controls) + st]
    zk, frk = leaves_k[0], leaves_k[1]
    ks, kf = sim(zk.expand(B, -1, -1) if shared else zk, leaves_k[2], state=tuple(t * 1.0 for t in leaves_k[3:]),
                 friction=frk.expand(B, -1, -1) if shared else frk)
    for a, b in zip(ks + kf, rs + rf):
        assert rel_err(a, b) < 1e-8
    objective(ks, kf, DEV).backward()
    for name, a, b in zip(("z", "friction", "controls", "x0", "xd0", "R0", "om0"), leaves_k, leaves_r):
        assert rel_err(a.grad, b.grad, 1e-9) < 1e-6, name


@pytest.mark.parametrize("variant", ["step", "odeint"])
def test_map_groups_forward_and_adjoint_fp64(variant, adjoint_kernel):
    """M maps for B = k*M trajectories (one map per scene, k control sequences per scene - BASELINE config 4): the same
    numbers as materialising every trajectory's map (`repeat_interleave`) in the oracle, and the map gradients come back
    per scene (M,H,W)."""
    from oracle import dphysics_oracle as O
    dtype = torch.float64
    T, M, k = 30, 3, 5
    B = M * k
    sim, cfg = _module("tradr", 0.4, T, variant, dtype)
    gen = torch.Generator().manual_seed(77)
    H = cfg.x_grid.shape[0]
    z = 0.2 * torch.randn(M, H, H, generator=gen, dtype=dtype)
    fr = 0.3 + 0.7 * torch.rand(M, H, H, generator=gen, dtype=dtype)
    _, controls, _, st = _random_case(cfg, B, T, 41, dtype)
    zr, fr_r, cr = (t.clone().requires_grad_(True) for t in (z, fr, controls))
    rs, rf = O.rollout(make_spec(cfg), zr.repeat_interleave(k, 0), cr, state=st, friction=fr_r.repeat_interleave(k, 0),
                       variant=variant, dtype=dtype)
    (rs[0].pow(2).sum() + rs[3].sum() + 1e-6 * rf[0].pow(2).sum()).backward()
    zk, fk, ck = (t.clone().to(DEV).requires_grad_(True) for t in (z, fr, controls))
    ks, kf = sim(zk, ck, state=tuple(t.to(DEV) for t in st), friction=fk)
    for a, b in zip(ks + kf, rs + rf):
        assert rel_err(a, b) < 1e-8
    (ks[0].pow(2).sum() + ks[3].sum() + 1e-6 * kf[0].pow(2).sum()).backward()
    assert zk.grad.shape == (M, H, H)
    assert rel_err(zk.grad, zr.grad) < 1e-6 and rel_err(fk.grad, fr_r.grad) < 1e-6 and rel_err(ck.grad, cr.grad) < 1e-6
    # a shared friction map next to per-scene height maps is materialised at the finer grouping
    with torch.no_grad():
        a, _ = sim(z.to(DEV), controls.to(DEV), state=tuple(t.to(DEV) for t in st), friction=fr[:1].to(DEV))
        b, _ = O.rollout(make_spec(cfg), z.repeat_interleave(k, 0), controls, state=st, friction=fr[:1].repeat(B, 1, 1),
                         variant=variant, dtype=dtype)
    assert rel_err(a[0], b[0]) < 1e-8


def test_host_tensor_mode_runs_on_the_gpu_and_returns_host_tensors():
    """`DPhysics(cfg)` with the reference's default device='cpu' (scripts/fit_terrain.py:34): host tensors in and out,
    gradients included, staged through the GPU - same numbers as the CUDA-tensor call, launched by the same kernels."""
    from monoforce_b200 import DPhysics, _lib
    T, B = 40, 3
    _, cfg = _module("marv", 0.2, T, "odeint")
    z, controls, fr, _ = _random_case(cfg, B, T, 3, torch.float32, with_state=False)
    host = DPhysics(cfg)                                    # device='cpu'
    zh, fh = z.clone().requires_grad_(True), fr.clone().requires_grad_(True)
    n0 = _lib.kernel_launches()
    (Xs, Xds, Rs, Oms), (Fs, Ff) = host(z_grid=zh.repeat(B, 1, 1), controls=controls, friction=fh.repeat(B, 1, 1))
    assert _lib.kernel_launches() > n0 and Xs.device.type == "cpu" and Fs.device.type == "cpu"
    Xs.pow(2).sum().backward()
    dev = DPhysics(cfg, device=DEV)
    zd, fd = z.clone().to(DEV).requires_grad_(True), fr.clone().to(DEV).requires_grad_(True)
    (Xd, _, Rd, _), (Fd, _) = dev(z_grid=zd.repeat(B, 1, 1), controls=controls.to(DEV), friction=fd.repeat(B, 1, 1))
    Xd.pow(2).sum().backward()
    assert torch.equal(Xs, Xd.cpu()) and torch.equal(Rs, Rd.cpu()) and torch.equal(Fs, Fd.cpu())
    assert zh.grad.device.type == "cpu" and rel_err(zh.grad, zd.grad) < 1e-5 and rel_err(fh.grad, fd.grad, 1e-9) < 1e-5
    # the in-place start-height snap reaches the caller's HOST state tensor (dphysics.py:571)
    st = (torch.zeros(B, 3), torch.zeros(B, 3), torch.eye(3).repeat(B, 1, 1), torch.zeros(B, 3))
    st_d = tuple(t.to(DEV) for t in st)
    with torch.no_grad():
        host(z_grid=(z + 0.3).repeat(B, 1, 1), controls=controls, state=st)
        dev(z_grid=(z + 0.3).to(DEV).repeat(B, 1, 1), controls=controls.to(DEV), state=st_d)
    assert st[0].device.type == "cpu" and st[0][:, 2].min() > 0.3 and torch.equal(st[0], st_d[0].cpu())


def test_cost_buffer_receives_the_fused_cost_in_place():
    T, B = 50, 8
    sim, cfg = _module("marv", 0.1, T)
    sim.fused_cost = True
    z, controls, _, _ = _random_case(cfg, B, T, 2, torch.float32, with_state=False, with_fric=False)
    with torch.no_grad():
        sim(z.to(DEV).unsqueeze(0), controls.to(DEV))
        ref = sim.last_cost.clone()
        gather = torch.zeros(3 * B, device=DEV)
        sim.cost_buffer = gather[B:2 * B]
        sim(z.to(DEV).unsqueeze(0), controls.to(DEV))
    assert sim.last_cost.data_ptr() == gather[B:2 * B].data_ptr()
    assert torch.equal(gather[B:2 * B], ref) and gather[:B].abs().sum() == 0 and gather[2 * B:].abs().sum() == 0
    sim.cost_buffer = torch.zeros(B + 1, device=DEV)
    with pytest.raises(ValueError, match="cost_buffer"):
        sim(z.to(DEV).unsqueeze(0), controls.to(DEV))


def test_fused_cost_matches_torch_definition():
    T, B = 100, 16
    sim, cfg = _module("marv", 0.1, T)
    sim.fused_cost = True
    z, controls, _, _ = _random_case(cfg, B, T, 2, torch.float32, with_state=False, with_fric=False)
    (_, forces) = sim(z.to(DEV).unsqueeze(0), controls.to(DEV))
    ref = torch.norm(forces[0], dim=-1).std(dim=-1).std(dim=-1)        # monoforce_node.py:91
    assert rel_err(sim.last_cost, ref) < 1e-3
    sim.return_forces = False
    (_, f2) = sim(z.to(DEV).unsqueeze(0), controls.to(DEV))
    assert f2[0].numel() == 0
    assert rel_err(sim.last_cost, ref) < 1e-3


def test_c_abi_host_entry_point_matches_device_path():
    """mfb_rollout_forward_host: host buffers in, host buffers out, same numbers as the device path."""
    import ctypes as C
    from monoforce_b200 import _lib
    T, B = 60, 7
    sim, cfg = _module("marv", 0.1, T)
    z, controls, fr, st = _random_case(cfg, B, T, 6, torch.float32)
    (states, forces) = sim(z.to(DEV).unsqueeze(0), controls.to(DEV), state=tuple(s.to(DEV) for s in st),
                           friction=fr.to(DEV).unsqueeze(0))
    lib = _lib.load()
    N = cfg.robot_points.shape[0]
    desc = _lib.RolloutDesc()
    desc.B, desc.T, desc.N, desc.H, desc.W = B, T, N, z.shape[0], z.shape[1]
    desc.n_tracks, desc.variant, desc.map_stride = len(cfg.driving_parts), _lib.MFB_STEP_LOOP, 0
    desc.mass, desc.gravity, desc.stiffness, desc.damping = cfg.robot_mass, cfg.gravity, cfg.stiffness, float(cfg.damping)
    desc.grid_res, desc.d_max, desc.dt, desc.omega_max = cfg.grid_res, cfg.d_max, cfg.dt, cfg.omega_max
    desc.robot_Ly = float(cfg.robot_size[1])
    I_inv = sim._constants(torch.device(DEV), torch.float32)[2]
    for i in range(9):
        desc.I_inv[i] = float(I_inv[i])
    h = lambda a: np.ascontiguousarray(a.numpy() if isinstance(a, torch.Tensor) else a)
    ins = dict(z_grid=h(z), friction=h(fr), controls=h(controls), x0=h(st[0]), xd0=h(st[1]), R0=h(st[2]), omega0=h(st[3]),
               points=h(cfg.robot_points), part_id=h(cfg.part_id))
    outs = dict(Xs=np.empty((B, T, 3), np.float32), Xds=np.empty((B, T, 3), np.float32), Rs=np.empty((B, T, 3, 3), np.float32),
                Omegas=np.empty((B, T, 3), np.float32), F_springs=np.empty((B, T, N, 3), np.float32),
                F_frictions=np.empty((B, T, N, 3), np.float32), x0z=np.empty((B,), np.float32), cost=np.empty((B,), np.float32))
    io = _lib.RolloutBuffers(**{k: C.c_void_p(v.ctypes.data) for k, v in {**ins, **outs}.items()}, ts=None)
    _lib.check(lib.mfb_rollout_forward_host(C.byref(desc), C.byref(io), _lib.MFB_F32, 0), "host entry")
    for k, t in zip(("Xs", "Xds", "Rs", "Omegas"), states):
        assert np.array_equal(outs[k], t.cpu().numpy()), k
    assert np.array_equal(outs["F_springs"], forces[0].cpu().numpy())
    assert np.array_equal(outs["F_frictions"], forces[1].cpu().numpy())
    # moving flippers through the host entry point: the (B,T,4) angles are uploaded like every other input
    gen = torch.Generator().manual_seed(12)
    ja = 0.5 * torch.randn(B, T, 4, generator=gen)
    (sj, fj) = sim(z.to(DEV).unsqueeze(0), controls.to(DEV), joint_angles=ja.to(DEV), state=tuple(s.to(DEV) for s in st),
                   friction=fr.to(DEV).unsqueeze(0))
    for i, piv in enumerate(list(cfg.joint_positions.values())[:4]):
        for k in range(3):
            desc.joint_positions[i * 3 + k] = float(piv[k])
    io.joint_angles = C.c_void_p(h(ja).ctypes.data)
    io.cost = None
    _lib.check(lib.mfb_rollout_forward_host(C.byref(desc), C.byref(io), _lib.MFB_F32, 0), "host entry (joints)")
    assert np.array_equal(outs["Xs"], sj[0].cpu().numpy()) and np.array_equal(outs["F_springs"], fj[0].cpu().numpy())
    assert not np.array_equal(outs["Xs"], states[0].cpu().numpy())
    io.joint_angles = None
    # bad arguments are rejected with a message, not a crash
    desc.N = 1000
    assert lib.mfb_rollout_forward_host(C.byref(desc), C.byref(io), _lib.MFB_F32, 0) != 0
    assert b"N must be" in lib.mfb_last_error()


def test_full_size_properties_cfg2():
    """BASELINE config 2 size (4096 x 400, 256^2 shared map): size-independent properties -
    duplicated controls give duplicated trajectories bit for bit, rotations stay orthonormal,
    sum of per-point forces balances the recorded acceleration."""
    T, B = 400, 4096
    sim, cfg = _module("marv", 0.05, T)
    z = hill_map(cfg).to(DEV)
    gen = torch.Generator().manual_seed(0)
    half = torch.stack([torch.rand(B // 2, generator=gen) * 2 - 1, torch.rand(B // 2, generator=gen) * 4 - 2], -1)
    controls = torch.cat([half, half], 0).unsqueeze(1).repeat(1, T, 1).to(DEV)
    (Xs, Xds, Rs, Oms), (Fs, Ff) = sim(z.unsqueeze(0), controls)
    assert torch.isfinite(Xs).all() and torch.isfinite(Fs).all()
    assert torch.equal(Xs[: B // 2], Xs[B // 2:]) and torch.equal(Fs[: B // 2, -1], Fs[B // 2:, -1])
    RtR = Rs[:, -1].transpose(1, 2) @ Rs[:, -1]
    assert (RtR - torch.eye(3, device=DEV)).abs().max() < 1e-3
    # Newton: m (v[t] - v[t-1]) / dt == sum F + gravity
    t = 250
    acc = (Xds[:, t] - Xds[:, t - 1]) / cfg.dt * cfg.robot_mass
    total = Fs[:, t].sum(1) + Ff[:, t].sum(1)
    total[:, 2] -= cfg.robot_mass * cfg.gravity
    assert ((acc - total).abs().max() / (cfg.robot_mass * cfg.gravity)) < 2e-2


def test_moving_flippers_match_reference_golden_fp32():
    """A8: marv with non-zero joint angles (per-step point articulation + inverse inertia, dphysics.py:192-197, :326-358)."""
    g = load_golden("marv_hill128_joints_T60_B2")
    sim, cfg = _module("marv", float(g["grid_res"]), int(g["T"]), "step")
    B = g["controls"].shape[0]
    with torch.no_grad():
        states, forces = sim(_t(g["z"]).unsqueeze(0).expand(B, -1, -1), _t(g["controls"]), joint_angles=_t(g["joint_angles"]))
    _check_golden(states, forces, g, GOLDEN_TOL["marv_hill128_joints_T60_B2"])


@pytest.mark.parametrize("variant", ["step", "odeint"])
def test_moving_flippers_fp64_match_oracle(variant):
    from oracle import dphysics_oracle as O
    dtype = torch.float64
    T, B = 80, 4
    sim, cfg = _module("marv", 0.1, T, variant, dtype)
    z, controls, fr, st = _random_case(cfg, B, T, 17, dtype)
    ramp = torch.linspace(-1.2, 1.2, T, dtype=dtype).repeat(B, 1)
    ja = torch.stack([ramp, 0.5 * ramp, -ramp, -0.3 * ramp], -1)
    rs, rf = O.rollout(make_spec(cfg), z.repeat(B, 1, 1), controls, joint_angles=ja, state=st,
                       friction=fr.repeat(B, 1, 1), variant=variant, dtype=dtype)
    with torch.no_grad():
        ks, kf = sim(z.to(DEV).unsqueeze(0), controls.to(DEV), joint_angles=ja.to(DEV), state=tuple(s.to(DEV) for s in st),
                     friction=fr.to(DEV).unsqueeze(0))
    for a, b in zip(ks, rs):
        assert rel_err(a, b) < 1e-9
    for a, b in zip(kf, rf):
        assert rel_err(a, b) < 1e-8
    # zero angles take the static-geometry kernel
    zk = z.to(DEV).requires_grad_(True)
    out, _ = sim(zk.unsqueeze(0), controls.to(DEV), joint_angles=torch.zeros_like(ja).to(DEV))
    out[0].sum().backward()
    assert torch.isfinite(zk.grad).all()


@pytest.mark.parametrize("variant", ["step", "odeint"])
def test_moving_flippers_adjoint_fp64_matches_oracle_autograd(variant):
    """A8 + A11: gradients w.r.t. joint angles (through the articulated points AND the per-step inverse inertia),
    height map, friction and controls vs autograd of the fp64 oracle."""
    from oracle import dphysics_oracle as O
    dtype = torch.float64
    T, B = 25, 3
    sim, cfg = _module("marv", 0.2, T, variant, dtype)
    z, controls, fr, st = _random_case(cfg, B, T, 31, dtype)
    gen = torch.Generator().manual_seed(5)
    ja = 0.6 * torch.randn(B, T, 4, generator=gen, dtype=dtype)
    spec = make_spec(cfg)
    leaves_r = [t.clone().requires_grad_(True) for t in (z, fr, controls, ja)]
    rs, rf = O.rollout(spec, leaves_r[0].unsqueeze(0).expand(B, -1, -1), leaves_r[2], joint_angles=leaves_r[3], state=st,
                       friction=leaves_r[1].unsqueeze(0).expand(B, -1, -1), variant=variant, dtype=dtype)
    wgt = torch.linspace(0.3, 1.0, T, dtype=dtype).view(1, T, 1)
    ((rs[0] * wgt).pow(2).sum() + (rs[3] * wgt).sum() + 1e-6 * rf[0].pow(2).sum() + 1e-6 * rf[1].pow(2).sum()).backward()
    leaves_k = [t.clone().to(DEV).requires_grad_(True) for t in (z, fr, controls, ja)]
    ks, kf = sim(leaves_k[0].unsqueeze(0), leaves_k[2], joint_angles=leaves_k[3], state=tuple(s.to(DEV) for s in st),
                 friction=leaves_k[1].unsqueeze(0))
    w_d = wgt.to(DEV)
    ((ks[0] * w_d).pow(2).sum() + (ks[3] * w_d).sum() + 1e-6 * kf[0].pow(2).sum() + 1e-6 * kf[1].pow(2).sum()).backward()
    for name, a, b in zip(("z", "friction", "controls", "joint_angles"), leaves_k, leaves_r):
        assert rel_err(a.grad, b.grad) < 1e-6, name


def _custom_robot(n_points, seed=0):
    """A DPhysConfig whose body is the first / a resampled subset of marv's contact points (exercises every
    points-per-lane instantiation, ragged last slots and the TMA row-alignment phases)."""
    from monoforce_b200 import DPhysConfig
    from monoforce_b200.dphys_config import part_ids
    cfg = DPhysConfig(robot="marv", grid_res=0.2)
    g = torch.Generator().manual_seed(seed)
    base = cfg.robot_points
    idx = torch.randint(0, base.shape[0], (n_points,), generator=g) if n_points > base.shape[0] else torch.randperm(base.shape[0], generator=g)[:n_points]
    pts = base[idx] + (0.01 * torch.randn(n_points, 3, generator=g) if n_points > base.shape[0] else 0)
    cfg.robot_points = pts.contiguous()
    cfg.driving_parts = [m[idx].clone() for m in cfg.driving_parts]
    cfg.part_id = part_ids(cfg.driving_parts, n_points)
    return cfg


@pytest.mark.parametrize("n_points,B,T,variant", [
    (1, 1, 1, "step"), (5, 3, 2, "step"), (31, 5, 3, "odeint"), (32, 2, 5, "step"), (33, 7, 7, "odeint"),
    (64, 9, 6, "step"), (65, 2, 4, "odeint"), (100, 4, 9, "step"), (129, 3, 5, "step"), (161, 6, 5, "odeint"), (200, 3, 11, "step"),
    (256, 5, 4, "step"),        # 65 / 129: first sizes that need a second / third warp in the one-CTA-per-trajectory forward kernel
])
def test_ragged_sizes_fp64_forward_and_adjoint(n_points, B, T, variant, adjoint_kernel):
    """Edge sizes: 1..256 contact points (all PPL instantiations), batch not a multiple of the CTA size, horizons
    that hit every 16-byte phase of the force rows; forward and gradients vs the fp64 oracle."""
    from monoforce_b200 import DPhysics
    from oracle import dphysics_oracle as O
    dtype = torch.float64
    cfg = _custom_robot(n_points, seed=n_points)
    cfg.traj_sim_time, cfg.use_odeint = T * cfg.dt + 1e-9, variant == "odeint"
    sim = _sim(cfg)
    z, controls, fr, st = _random_case(cfg, B, T, 100 + n_points, dtype)
    spec = make_spec(cfg)
    zr, cr = z.clone().requires_grad_(True), controls.clone().requires_grad_(True)
    rs, rf = O.rollout(spec, zr.unsqueeze(0).expand(B, -1, -1), cr, state=st, friction=fr.unsqueeze(0).expand(B, -1, -1),
                       variant=variant, dtype=dtype)
    (rs[0].pow(2).sum() + 1e-6 * rf[0].pow(2).sum() + rs[2].sum()).backward()
    zk, ck = z.to(DEV).requires_grad_(True), controls.to(DEV).requires_grad_(True)
    ks, kf = sim(zk.unsqueeze(0), ck, state=tuple(s.to(DEV) for s in st), friction=fr.to(DEV).unsqueeze(0))
    (ks[0].pow(2).sum() + 1e-6 * kf[0].pow(2).sum() + ks[2].sum()).backward()
    for a, b in zip(ks + kf, rs + rf):
        assert a.shape == b.shape
        assert rel_err(a, b) < 1e-8
    assert rel_err(zk.grad, zr.grad) < 1e-6
    assert rel_err(ck.grad, cr.grad, 1e-9) < 1e-6


def test_fp32_ragged_sizes_forces_bit_pattern():
    """fp32 rows leave through TMA bulk stores + scalar head/tail stores: every element of every row must be written
    (no stale bytes) for all row phases; compare against the device's own no-TMA reference: the fp64 kernel."""
    from monoforce_b200 import DPhysics
    for n_points, B, T in ((7, 5, 9), (175, 3, 6), (223, 2, 5), (130, 4, 7)):
        cfg = _custom_robot(n_points, seed=n_points + 1)
        cfg.traj_sim_time, cfg.use_odeint = T * cfg.dt + 1e-9, False
        sim = DPhysics(cfg, device=DEV)
        z, controls, fr, st = _random_case(cfg, B, T, 7 + n_points, torch.float64, terrain="flat", with_fric=False)
        with torch.no_grad():
            _, f64 = sim(z.to(DEV).unsqueeze(0), controls.to(DEV), state=tuple(s.to(DEV) for s in st))
            sentinel = torch.full((B, T, n_points, 3), float("nan"), device=DEV)      # poison recycled allocations
            del sentinel
            _, f32 = sim(z.float().to(DEV).unsqueeze(0), controls.float().to(DEV), state=tuple(s.float().to(DEV) for s in st))
        assert torch.isfinite(f32[0]).all() and torch.isfinite(f32[1]).all()
        assert rel_err(f32[0], f64[0]) < 1e-3 and rel_err(f32[1], f64[1], 1e-6) < 5e-3


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("case", ["same_grid", "shared_rows", "per_trajectory_rows"])
def test_fused_physics_loss_matches_reference_definition(case, dtype):
    """K6 (losses.py:102-138 fused with its gradient) vs the oracle's restatement of the reference and torch autograd:
    identical time grids (identity gather), one (1,T) row of stamps shared by the batch, and per-trajectory stamps with
    T2 != T1 (nearest-stamp search, several ground-truth stamps hitting the same predicted index)."""
    from monoforce_b200.losses import physics_loss
    from oracle.dphysics_oracle import physics_loss as ref_loss
    g = torch.Generator().manual_seed(3)
    B, T1 = 37, 90
    Xp = torch.randn(B, T1, 3, generator=g, dtype=dtype)
    if case == "same_grid":
        T2 = T1
        pred_ts = gt_ts = (torch.arange(T1, dtype=dtype) * 0.01)[None]
    elif case == "shared_rows":
        T2 = 23
        pred_ts = (torch.arange(T1, dtype=dtype) * 0.01)[None]
        gt_ts = torch.sort(torch.rand(1, T2, generator=g, dtype=dtype) * 0.9)[0]
    else:
        T2 = 41
        pred_ts = torch.cumsum(torch.rand(B, T1, generator=g, dtype=dtype) * 0.02 + 1e-3, dim=1)
        gt_ts = torch.sort(torch.rand(B, T2, generator=g, dtype=dtype) * 1.2)[0]
    Xg = torch.randn(B, T2, 3, generator=g, dtype=dtype)
    Xr = Xp.clone().requires_grad_(True)
    lr = ref_loss((Xr,), (Xg,), pred_ts, gt_ts, 0.7)
    (3.0 * lr).backward()
    Xk = Xp.clone().to(DEV).requires_grad_(True)
    pk = pred_ts.to(DEV)
    gk = pk if case == "same_grid" else gt_ts.to(DEV)
    n0 = __import__("monoforce_b200")._lib.kernel_launches()
    lk = physics_loss((Xk,), (Xg.to(DEV),), pk, gk, 0.7)
    assert __import__("monoforce_b200")._lib.kernel_launches() - n0 == 2       # the fused kernel + its one-block finish
    (3.0 * lk).backward()
    tol = 1e-5 if dtype == torch.float32 else 1e-12
    assert abs(lk.item() - lr.item()) <= tol * abs(lr.item())
    assert rel_err(Xk.grad, Xr.grad) < (1e-5 if dtype == torch.float32 else 1e-12)


def test_full_size_adjoint_properties_cfg3():
    """BASELINE config 3 size (4096 x 400, 256^2 shared map, fp32): size-independent properties of the adjoint.
    (i) the single-sweep kernel (contact_sum tape + kappa channel) and the three-pass kernel differentiate the same
    recorded rollout, so their gradients agree to summation-order rounding; (ii) duplicated controls give duplicated
    control gradients; (iii) the adjoint is linear in the seed: doubling the objective doubles every gradient;
    (iv) the map gradients of the shared map are finite and non-trivial, the friction gradient only touches cells the
    robots visited."""
    from monoforce_b200.losses import physics_loss
    T, B = 400, 4096
    sim, cfg = _module("marv", 0.05, T)
    sim.return_forces = False                      # the training path: the objective reads states only
    gen = torch.Generator().manual_seed(1)
    half = torch.stack([torch.rand(B // 2, generator=gen) * 0.5 + 0.5, torch.rand(B // 2, generator=gen) * 4 - 2], -1)
    controls = torch.cat([half, half], 0).unsqueeze(1).repeat(1, T, 1).to(DEV)
    ts = (torch.arange(T, dtype=torch.float32) * cfg.dt)[None].to(DEV)
    with torch.no_grad():
        gt, _ = sim(hill_map(cfg).to(DEV).unsqueeze(0), controls)

    def grads(tape, scale):
        sim.adjoint_tape = tape
        z = torch.zeros(1, 256, 256, device=DEV, requires_grad=True)
        fr = torch.full((1, 256, 256), 0.5, device=DEV, requires_grad=True)
        c = controls.clone().requires_grad_(True)
        st, _ = sim(z, c, friction=fr)
        (scale * physics_loss(st, gt, ts, ts, 0.9)).backward()
        return z.grad, fr.grad, c.grad

    gz, gf, gc = grads(True, 1.0)
    gz3, gf3, gc3 = grads(False, 1.0)
    gz2, gf2, gc2 = grads(True, 2.0)
    for a in (gz, gf, gc):
        assert torch.isfinite(a).all()
    assert gz.abs().max() > 0 and gf.abs().max() > 0 and gc.abs().max() > 0
    assert rel_err(gz, gz3) < 1e-3 and rel_err(gf, gf3) < 1e-3 and rel_err(gc, gc3) < 1e-3          # (i)
    assert torch.equal(gc[: B // 2], gc[B // 2:])                                                   # (ii)
    assert rel_err(gz2, 2 * gz) < 1e-4 and rel_err(gf2, 2 * gf) < 1e-4 and rel_err(gc2, 2 * gc) < 1e-5   # (iii)
    assert (gf != 0).float().mean() < 0.5                                                            # (iv)
