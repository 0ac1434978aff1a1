"""CPU tests: the oracle reproduces the golden vectors minted from the unmodified reference."""
import numpy as np
import pytest
import torch

from helpers_mfb import load_golden, make_spec


def _cfg(g):
    from monoforce_b200 import DPhysConfig
    cfg = DPhysConfig(robot=str(g["robot"]), grid_res=float(g["grid_res"]))
    cfg.traj_sim_time = int(g["T"]) * cfg.dt
    return cfg


FWD = ["cfg1_marv_flat64_T100", "cfg1_tradr_flat64_T100", "marv_hill128_T100_B4", "marv_noise128_state_fric_T100_B4",
       "tradr_noise128_state_fric_T100_B4", "marv_flat256_T400_B2", "marv_hill128_odeint_T60_B2",
       "marv_hill128_joints_T60_B2", "marv_ramp128_odeint_T200_B3", "marv_ramp256_T400_B3", "tradr_ramp128_T300_B3"]


@pytest.mark.parametrize("name", FWD)
def test_oracle_reproduces_reference_bit_for_bit(name):
    from oracle import dphysics_oracle as O
    g = load_golden(name)
    cfg = _cfg(g)
    B = g["controls"].shape[0]
    z = torch.from_numpy(g["z"]).repeat(B, 1, 1)
    fr = torch.from_numpy(g["friction"]).repeat(B, 1, 1) if "friction" in g else None
    st = tuple(torch.from_numpy(g[k]) for k in ("x0", "xd0", "R0", "om0")) if "x0" in g else None
    ja = torch.from_numpy(g["joint_angles"]) if "joint_angles" in g else None
    (Xs, Xds, Rs, Oms), (Fs, Ff) = O.rollout(make_spec(cfg), z, torch.from_numpy(g["controls"]), joint_angles=ja,
                                              state=st, friction=fr, variant=str(g["variant"]))
    assert np.array_equal(Xs.numpy(), g["Xs"])
    assert np.array_equal(Xds.numpy(), g["Xds"])
    assert np.array_equal(Rs.numpy(), g["Rs"])
    assert np.array_equal(Oms.numpy(), g["Omegas"])
    keep = g["F_keep_steps"]
    assert np.array_equal(Fs[:, keep].numpy(), g["Fs_keep"])
    assert np.array_equal(Ff[:, keep].numpy(), g["Ff_keep"])
    assert np.allclose(Fs.double().sum(dim=2).numpy(), g["Fs_sum"], rtol=0, atol=1e-9)
    assert np.array_equal(O.path_cost(Fs).numpy(), g["cost"])


def test_cfg1_pinned_numbers():
    """The survey's independently probed values for BASELINE config 1."""
    g = load_golden("cfg1_marv_flat64_T100")
    assert np.allclose(g["Xs"][0, 99], [0.7115149, 0.0426049, 0.0029562], atol=2e-6)
    assert np.allclose(g["Omegas"][0, 99], [9.354e-4, -3.392e-4, 0.0955751], atol=2e-6)
    assert abs(float(g["cost"][0]) - 0.0773055) < 1e-5
    g = load_golden("cfg1_tradr_flat64_T100")
    assert np.allclose(g["Xs"][0, 99], [0.5685527, 0.0458160, -0.0028749], atol=2e-6)


@pytest.mark.parametrize("name", ["grad64_marv_noise128_T40_B2", "grad64_tradr_noise64_T40_B2"])
def test_oracle_autograd_matches_reference_gradients(name):
    from oracle import dphysics_oracle as O
    g = load_golden(name)
    cfg = _cfg(g)
    dt = torch.float64
    B = g["controls"].shape[0]
    z = torch.from_numpy(g["z"]).requires_grad_(True)
    fr = torch.from_numpy(g["friction"]).requires_grad_(True)
    c = torch.from_numpy(g["controls"]).requires_grad_(True)
    st = [torch.from_numpy(g[k]).requires_grad_(True) for k in ("x0", "xd0", "R0", "om0")]
    states, forces = O.rollout(make_spec(cfg), z.unsqueeze(0).expand(B, -1, -1), c, state=tuple(s * 1.0 for s in st),
                               friction=fr.unsqueeze(0).expand(B, -1, -1), dtype=dt, mutate_state=True)
    outs = list(states) + list(forces)
    loss = sum(float(s) * (o * torch.from_numpy(g[f"w{i}"]).double()).sum() for i, (s, o) in enumerate(zip(g["scales"], outs)))
    assert abs(loss.item() - float(g["loss"])) < 1e-10
    loss.backward()
    assert np.allclose(z.grad.numpy(), g["g_z"], rtol=1e-9, atol=1e-12)
    assert np.allclose(c.grad.numpy(), g["g_controls"], rtol=1e-9, atol=1e-12)
    assert np.allclose(st[2].grad.numpy(), g["g_R0"], rtol=1e-9, atol=1e-12)


def test_oracle_matches_live_reference_when_present():
    """Build container only: run the unmodified reference next to the oracle on a fresh random case."""
    from oracle.ref_import import reference_available, import_reference
    if not reference_available():
        pytest.skip("reference tree not present on this machine")
    from oracle import dphysics_oracle as O
    dp, cfgm = import_reference()
    torch.manual_seed(123)
    cfg = cfgm.DPhysConfig(robot="marv", grid_res=0.1)
    cfg.traj_sim_time, cfg.use_odeint = 0.5, False
    sim = dp.DPhysics(cfg)
    B, n = 3, 50
    z = (torch.exp(-(cfg.x_grid - 2) ** 2 / 4) * torch.exp(-cfg.y_grid ** 2 / 2) + 0.02 * torch.randn_like(cfg.x_grid)).repeat(B, 1, 1)
    controls = torch.rand(B, 1, 2).repeat(1, n, 1) * torch.tensor([2.0, 4.0]) - torch.tensor([1.0, 2.0])
    with torch.no_grad():
        a, b = sim(z, controls)
    c, e = O.rollout(make_spec(cfg), z, controls)
    for p, q in zip(a + b, c + e):
        assert torch.equal(p, q)
