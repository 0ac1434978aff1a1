"""DPhysics: drop-in host-side mirror of the reference's differentiable-physics module,
backed by the sm_100a rollout kernels behind the C ABI (include/monoforce_b200.h).

Reference interface mirrored here (same names, argument meaning, return structure, asserts):
    monoforce/src/monoforce/models/traj_predictor/dphysics.py
        DPhysics.__init__            :145-170
        DPhysics.forward / dphysics  :596-605, :530-594
        generate_controls            :42-72
        vw_to_track_vels             :75-104
        inertia_tensor               :107-141
The T-step loop (`dynamics` :467-497 / `dynamics_odeint` :499-528), the per-step force model
(`forward_kinematics` :172-272), grid sampling (:385-455) and autograd's backward are ONE CUDA
launch each.  There is no CPU path: tensors must live on a CUDA device and the shared
library must have been built (`python -m monoforce_b200.build`).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .dphys_config import DPhysConfig

__all__ = ["DPhysics", "DPhysConfig", "generate_controls", "vw_to_track_vels", "inertia_tensor",
           "normalized", "skew_symmetric", "path_costs_from_forces"]


# ---------------------------------------------------------------------------------------------
# small host-side helpers that callers import from the reference module
# ---------------------------------------------------------------------------------------------
def normalized(x, eps=1e-6, dim=-1):
    """dphysics.py:7-19."""
    return x / torch.clamp(torch.norm(x, dim=dim, keepdim=True), min=eps)


def skew_symmetric(v):
    """dphysics.py:22-40."""
    assert v.dim() == 2 and v.shape[1] == 3
    z = torch.zeros_like(v[:, 0])
    return torch.stack([torch.stack([z, -v[:, 2], v[:, 1]], dim=1),
                        torch.stack([v[:, 2], z, -v[:, 0]], dim=1),
                        torch.stack([-v[:, 1], v[:, 0], z], dim=1)], dim=1)


def generate_controls(n_trajs=10, time_horizon=5.0, dt=0.01, v_range=(-1.0, 1.0), w_range=(-1.0, 1.0)):
    """Constant (v, w) per trajectory, uniformly sampled - dphysics.py:42-72.

    Returns (controls (n_trajs, N, 2), time_stamps (N,)) with N = int(time_horizon / dt)."""
    n = int(time_horizon / dt)
    stamps = torch.linspace(0, time_horizon, n)
    v = torch.rand(n_trajs) * (v_range[1] - v_range[0]) + v_range[0]
    w = torch.rand(n_trajs) * (w_range[1] - w_range[0]) + w_range[0]
    return torch.stack([v.unsqueeze(1).repeat(1, n), w.unsqueeze(1).repeat(1, n)], dim=-1), stamps


def vw_to_track_vels(v, w, robot_size, n_tracks):
    """dphysics.py:75-104: (left, right) or (FL, FR, RL, RR) track speeds."""
    _, Ly = robot_size
    lo, hi = v - w * (Ly / 2.0), v + w * (Ly / 2.0)
    if n_tracks == 2:
        return torch.stack([lo, hi], dim=-1)
    if n_tracks == 4:
        return torch.stack([lo, hi, lo, hi], dim=-1)
    raise ValueError('n_tracks must be 2 or 4')


def inertia_tensor(mass, points):
    """Point-mass inertia about the body origin, (B,N,3) -> (B,3,3) - dphysics.py:107-141."""
    assert points.dim() == 3
    mp = mass / points.shape[1]
    x, y, z = points[:, :, 0], points[:, :, 1], points[:, :, 2]
    Ixx = torch.sum(mp * (y ** 2 + z ** 2), dim=1)
    Iyy = torch.sum(mp * (x ** 2 + z ** 2), dim=1)
    Izz = torch.sum(mp * (x ** 2 + y ** 2), dim=1)
    Ixy = -torch.sum(mp * x * y, dim=1)
    Ixz = -torch.sum(mp * x * z, dim=1)
    Iyz = -torch.sum(mp * y * z, dim=1)
    return torch.stack([torch.stack([Ixx, Ixy, Ixz], dim=1),
                        torch.stack([Ixy, Iyy, Iyz], dim=1),
                        torch.stack([Ixz, Iyz, Izz], dim=1)], dim=1)


def path_costs_from_forces(F_springs):
    """monoforce_ros/nodes/monoforce_node.py:91 (torch ops; the kernel can also emit it fused)."""
    return torch.norm(F_springs, dim=-1).std(dim=-1).std(dim=-1)


# ---------------------------------------------------------------------------------------------
# autograd bridge to the C ABI
# ---------------------------------------------------------------------------------------------
def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class _RolloutMeta:
    """Everything that is not a differentiable tensor."""
    __slots__ = ("desc", "points", "part_id", "ts", "want_forces", "want_cost", "dtype_code", "B", "T", "N", "timings",
                 "use_tape", "cost_buffer")


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "monoforce_b200.DPhysics runs on CUDA only (sm_100a kernels, no CPU fallback); "
                f"got a tensor on {t.device}. Construct DPhysics(cfg, device='cuda').")


def _events(meta, name):
    """CUDA events bracketing exactly the library call on the launching stream (DPhysics.timings opt-in)."""
    if meta.timings is None:
        return None
    ev = (name, torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    ev[1].record()
    meta.timings.append(ev)
    return ev


def _workspace(lib, meta, dev):
    n = int(lib.mfb_rollout_workspace_bytes(C.byref(meta.desc), meta.dtype_code))
    if n < 0:
        raise RuntimeError("mfb_rollout_workspace_bytes rejected the rollout description")
    return torch.empty(n, dtype=torch.uint8, device=dev)


class _Rollout(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, mu, controls, x0, xd0, R0, om0, joint_angles, meta: _RolloutMeta):
        lib = _lib.load()
        dt_ = z.dtype
        dev = z.device
        B, T, N = meta.B, meta.T, meta.N
        z, mu, controls = z.contiguous(), mu.contiguous(), controls.contiguous()
        # x0 is cloned: the caller's tensor is overwritten in place afterwards (start-height snap, :571)
        x0, xd0, R0, om0 = x0.contiguous().clone(), xd0.contiguous(), R0.contiguous(), om0.contiguous()
        new = lambda *s: torch.empty(*s, dtype=dt_, device=dev)
        Xs, Xds, Rs, Oms = new(B, T, 3), new(B, T, 3), new(B, T, 3, 3), new(B, T, 3)
        Fs = new(B, T, N, 3) if meta.want_forces else None
        Ff = new(B, T, N, 3) if meta.want_forces else None
        x0z = new(B)
        cost = None
        if meta.want_cost:
            cost = meta.cost_buffer if meta.cost_buffer is not None else new(B)
        # adjoint tape (4 bytes per trajectory-step): only when a backward can follow
        csum = new(B, T) if meta.use_tape else None
        ws = _workspace(lib, meta, dev)
        io = _lib.RolloutBuffers(
            workspace=_ptr(ws), workspace_bytes=ws.numel(),
            z_grid=_ptr(z), friction=_ptr(mu), controls=_ptr(controls), x0=_ptr(x0), xd0=_ptr(xd0), R0=_ptr(R0),
            omega0=_ptr(om0), points=_ptr(meta.points), part_id=_ptr(meta.part_id), ts=_ptr(meta.ts),
            joint_angles=_ptr(joint_angles),
            Xs=_ptr(Xs), Xds=_ptr(Xds), Rs=_ptr(Rs), Omegas=_ptr(Oms), F_springs=_ptr(Fs), F_frictions=_ptr(Ff),
            x0z=_ptr(x0z), cost=_ptr(cost), contact_sum=_ptr(csum))
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            ev = _events(meta, "forward")
            _lib.check(lib.mfb_rollout_forward(C.byref(meta.desc), C.byref(io), meta.dtype_code, C.c_void_p(stream)),
                       "mfb_rollout_forward")
            if ev is not None:
                ev[2].record()
        ctx.meta = meta
        ctx.set_materialize_grads(False)     # unused outputs (e.g. the 2 x (B,T,N,3) forces) give None, not zeros
        ctx.save_for_backward(z, mu, controls, x0, xd0, R0, om0, Xs, Xds, Rs, Oms, x0z)
        ctx.joint_angles = joint_angles
        ctx.csum = csum
        empty = torch.empty(0, dtype=dt_, device=dev)
        outs = (Xs, Xds, Rs, Oms, Fs if Fs is not None else empty, Ff if Ff is not None else empty, x0z,
                cost if cost is not None else empty)
        ctx.mark_non_differentiable(outs[7])
        if Fs is None:
            ctx.mark_non_differentiable(outs[4], outs[5])
        return outs

    @staticmethod
    def backward(ctx, gXs, gXds, gRs, gOms, gFs, gFf, gx0z, gcost):
        lib = _lib.load()
        meta = ctx.meta
        z, mu, controls, x0, xd0, R0, om0, Xs, Xds, Rs, Oms, x0z = ctx.saved_tensors
        need = ctx.needs_input_grad
        dev, dt_ = z.device, z.dtype
        c = lambda g: None if g is None else g.contiguous()
        gXs, gXds, gRs, gOms, gFs, gFf, gx0z = map(c, (gXs, gXds, gRs, gOms, gFs, gFf, gx0z))
        if not meta.want_forces:
            gFs = gFf = None
        if need[0] and need[1] and z.shape == mu.shape:
            # both map gradients in ONE buffer (two contiguous halves): a data-parallel caller all-reduces them as a
            # single flat tensor without a gather / scatter copy (monoforce_b200/dist.py::allreduce_map_grads)
            g_both = torch.zeros((2,) + tuple(z.shape), dtype=dt_, device=dev)
            g_z, g_mu = g_both[0], g_both[1]
        else:
            g_z = torch.zeros_like(z) if need[0] else None
            g_mu = torch.zeros_like(mu) if need[1] else None
        g_c = torch.empty_like(controls) if need[2] else None
        g_x0 = torch.empty_like(x0) if need[3] else None
        g_xd0 = torch.empty_like(xd0) if need[4] else None
        g_R0 = torch.empty_like(R0) if need[5] else None
        g_om0 = torch.empty_like(om0) if need[6] else None
        ja = ctx.joint_angles
        g_ja = torch.empty_like(ja) if (ja is not None and need[7]) else None
        ws = _workspace(lib, meta, dev)
        io = _lib.RolloutBuffers(
            workspace=_ptr(ws), workspace_bytes=ws.numel(),
            z_grid=_ptr(z), friction=_ptr(mu), controls=_ptr(controls), x0=_ptr(x0), xd0=_ptr(xd0), R0=_ptr(R0),
            omega0=_ptr(om0), points=_ptr(meta.points), part_id=_ptr(meta.part_id), ts=_ptr(meta.ts),
            joint_angles=_ptr(ja),
            Xs=_ptr(Xs), Xds=_ptr(Xds), Rs=_ptr(Rs), Omegas=_ptr(Oms), F_springs=None, F_frictions=None,
            x0z=_ptr(x0z), cost=None, contact_sum=_ptr(ctx.csum))
        grads = _lib.RolloutGrads(
            g_Xs=_ptr(gXs), g_Xds=_ptr(gXds), g_Rs=_ptr(gRs), g_Omegas=_ptr(gOms), g_F_springs=_ptr(gFs),
            g_F_frictions=_ptr(gFf), g_x0z=_ptr(gx0z),
            g_z_grid=_ptr(g_z), g_friction=_ptr(g_mu), g_controls=_ptr(g_c), g_x0=_ptr(g_x0), g_xd0=_ptr(g_xd0),
            g_R0=_ptr(g_R0), g_omega0=_ptr(g_om0), g_joint_angles=_ptr(g_ja))
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            ev = _events(meta, "backward")
            _lib.check(lib.mfb_rollout_backward(C.byref(meta.desc), C.byref(io), C.byref(grads), meta.dtype_code,
                                                C.c_void_p(stream)), "mfb_rollout_backward")
            if ev is not None:
                ev[2].record()
        return g_z, g_mu, g_c, g_x0, g_xd0, g_R0, g_om0, g_ja, None


# ---------------------------------------------------------------------------------------------
# the module
# ---------------------------------------------------------------------------------------------
class DPhysics(torch.nn.Module):
    """Same constructor / call contract as the reference `DPhysics` (dphysics.py:144-170, :596-605).

    Extra, opt-in attributes (not in the reference, defaults keep its behaviour):
      * ``return_forces`` (True): set False to skip materialising the two (B,T,N,3) force
        tensors (they are returned as empty tensors) - the training path only needs states.
      * ``fused_cost`` (False): when True the kernel also emits the per-trajectory traversal
        cost ``norm(F_springs).std(-1).std(-1)`` (monoforce_node.py:91) in ``last_cost``.
      * ``timings`` (None): set to a list to collect CUDA events around each library call.
      * ``shared_map`` (None): B `repeat`ed copies of one map are detected on the device and read as ONE shared map
        (bit-identical results, no B cell tables); True promises it without looking, False disables the test.
      * z_grid / friction with M maps for B = k*M trajectories: k consecutive trajectories share each map.
      * ``adjoint_tape`` (True): record the per-step soft-contact normaliser (4 B per trajectory-step) when
        gradients are required, which lets the backward run the single-sweep adjoint kernel.
    """

    # attributes rewritten on every call with plain tensors / flags (the reference does the same, :561-581): they skip
    # nn.Module.__setattr__'s Parameter / buffer / sub-module bookkeeping, ~10 assignments per call on the planner's latency path
    _PER_CALL = frozenset({'controls', 'joint_angles', 'ts', 'z_grid', 'friction', '_z_arg', '_mu_arg', '_x0z', 'last_cost',
                           '_moving_joints'})

    def __setattr__(self, name, value):
        if name in DPhysics._PER_CALL:
            self.__dict__[name] = value
        else:
            super().__setattr__(name, value)

    def __init__(self, dphys_cfg=None, device='cpu'):
        super().__init__()
        self.dphys_cfg = DPhysConfig() if dphys_cfg is None else dphys_cfg
        self.device = device
        self.x_points = self.dphys_cfg.robot_points.to(self.device).unsqueeze(0)       # (1, N, 3)
        # inertia about the body origin and its inverse, computed once like the reference's __init__ (:152-153);
        # the reference re-derives the same constant every step (:196-197) because geometry is static
        self.I = inertia_tensor(mass=self.dphys_cfg.robot_mass, points=self.x_points).to(self.device)
        self.I_inv = torch.linalg.inv(self.I)
        self.z_grid = None
        self.friction = None
        self.stiffness = self.dphys_cfg.stiffness
        self.damping = self.dphys_cfg.damping
        self.controls = None
        self.joint_angles = None
        T, dt = self.dphys_cfg.traj_sim_time, self.dphys_cfg.dt
        self.ts = torch.linspace(0, T, int(T / dt)).to(self.device)
        self.integrator = self.dynamics_odeint if self.dphys_cfg.use_odeint else self.dynamics
        self.return_forces = True
        self.fused_cost = False
        self.last_cost = None
        self.cost_buffer = None   # optional (B,) tensor the fused cost is written into (e.g. this rank's slice of an all-gather buffer)
        self.shared_map = None    # None: recognise B repeated copies of one map on the device; True / False: caller's word
        self.adjoint_tape = True  # False: no contact_sum tape, the backward runs the three-pass adjoint (development / tests)
        self.timings = None      # set to a list to collect (name, start_event, end_event) around every library call
        self._const_cache = {}

    # The two integrators exist as named methods because callers select them through
    # `use_odeint`; both are one fused CUDA launch.
    def dynamics(self, state):
        return self._launch(state, _lib.MFB_STEP_LOOP)

    def dynamics_odeint(self, state):
        return self._launch(state, _lib.MFB_ODEINT_EULER)

    # -- constants on the device, cached per (device, dtype) -----------------------------------
    def _constants(self, device, dtype):
        key = (str(device), dtype)
        if key not in self._const_cache:
            cfg = self.dphys_cfg
            pts = cfg.robot_points.to(device=device, dtype=dtype).contiguous()
            part = cfg.part_id.to(device=device, dtype=torch.int32).contiguous()
            # I_inv in the working precision (the reference inverts in the tensor dtype, :153)
            I = inertia_tensor(mass=cfg.robot_mass, points=cfg.robot_points.to(dtype).unsqueeze(0))
            I_inv = torch.linalg.inv(I)[0].double().cpu().numpy().reshape(-1)
            self._const_cache[key] = (pts, part, I_inv)
        return self._const_cache[key]

    def _launch(self, state, variant):
        x0, xd0, R0, om0 = state
        z, mu, controls = self._z_arg, self._mu_arg, self.controls
        dev, dtype = controls.device, controls.dtype
        _require_cuda(z, mu, controls, x0, xd0, R0, om0)
        if dtype not in (torch.float32, torch.float64):
            raise TypeError(f"DPhysics supports float32 and float64, got {dtype}")
        cfg = self.dphys_cfg
        pts, part, I_inv = self._constants(dev, dtype)
        B, T = controls.shape[0], controls.shape[1]
        H, W = z.shape[-2], z.shape[-1]
        assert H == W, f'height map must be square (the reference strides rows by H), got {(H, W)}'
        desc = _lib.RolloutDesc()
        desc.B, desc.T, desc.N, desc.H, desc.W = B, T, pts.shape[0], H, W
        desc.n_tracks = len(cfg.driving_parts)
        desc.variant = variant
        desc.map_stride = 0 if z.shape[0] == 1 else H * W
        desc.traj_per_map = 0 if z.shape[0] in (1, B) else B // z.shape[0]
        desc.mass, desc.gravity = float(cfg.robot_mass), float(cfg.gravity)
        desc.stiffness, desc.damping = float(self.stiffness), float(self.damping)
        desc.grid_res, desc.d_max, desc.dt = float(cfg.grid_res), float(cfg.d_max), float(cfg.dt)
        desc.omega_max = float(cfg.omega_max)
        desc.robot_Ly = float(cfg.robot_size[1])
        for i in range(9):
            desc.I_inv[i] = float(I_inv[i])
        for i, piv in enumerate(list(cfg.joint_positions.values())[:4]):
            for k in range(3):
                desc.joint_positions[i * 3 + k] = float(piv[k])
        meta = _RolloutMeta()
        meta.desc, meta.points, meta.part_id = desc, pts, part
        meta.ts = None
        if variant == _lib.MFB_ODEINT_EULER:
            ts = self.ts
            if ts.dtype != dtype:   # the reference builds linspace in the default dtype (:167)
                n_full = int(cfg.traj_sim_time / cfg.dt)
                ts = torch.linspace(0, cfg.traj_sim_time, n_full, dtype=dtype)[:T]
            meta.ts = ts.to(device=dev).contiguous()
        moving = self._moving_joints
        joint_angles = self.joint_angles.to(device=dev, dtype=dtype).contiguous() if moving else None
        meta.want_forces = bool(self.return_forces) or moving
        meta.want_cost = bool(self.fused_cost) and variant == _lib.MFB_STEP_LOOP and not moving
        meta.dtype_code = _lib.MFB_F32 if dtype == torch.float32 else _lib.MFB_F64
        meta.B, meta.T, meta.N = B, T, pts.shape[0]
        meta.timings = self.timings
        cb = self.cost_buffer
        if cb is not None and not (cb.shape == (B,) and cb.dtype == dtype and cb.device == dev and cb.is_contiguous()):
            raise ValueError(f"cost_buffer must be a contiguous ({B},) {dtype} tensor on {dev}")
        meta.cost_buffer = cb
        cast = lambda t: t.to(device=dev, dtype=dtype)
        # contact_sum tape for the single-sweep adjoint: only when a backward can follow this call
        meta.use_tape = bool(self.adjoint_tape) and torch.is_grad_enabled() and any(
            t is not None and t.requires_grad for t in (z, mu, controls, x0, xd0, R0, om0, joint_angles))
        Xs, Xds, Rs, Oms, Fs, Ff, x0z, cost = _Rollout.apply(cast(z), cast(mu), controls, cast(x0), cast(xd0),
                                                              cast(R0), cast(om0), joint_angles, meta)
        self._x0z = x0z
        self.last_cost = cost if meta.want_cost else None
        return Xs, Xds, Rs, Oms, Fs, Ff

    def _zero_scalar(self, dtype):
        key = ('zero', str(self.device), dtype)
        if key not in self._const_cache:
            self._const_cache[key] = torch.zeros((), device=self.device, dtype=dtype)
        return self._const_cache[key]

    def _default_friction(self):
        """`dphys_cfg.friction` as a (1,H,W) tensor on self.device, uploaded once per (tensor, version, device) instead of once
        per call (a pageable host->device copy blocks the host behind the previous call's kernel)."""
        f = self.dphys_cfg.friction
        key = (id(f), f._version, str(self.device))
        if self.__dict__.get('_fric_key') != key:
            self.__dict__['_fric_dev'] = f.to(self.device).unsqueeze(0)
            self.__dict__['_fric_key'] = key
        return self.__dict__['_fric_dev']

    def _shared_view(self, grid, B):
        """(1,H,W) view when all B trajectories read one map: a broadcast / `expand`ed input, or - what the reference's
        callers pass (monoforce_node.py:156, diff_physics.ipynb cell 3) - B materialised `repeat`s of one map.  Repeats
        are recognised by comparing all copies with the first on the device (one pass over the maps + one host read per
        call: far cheaper than B cell tables) when no gradient can flow into the copies; `self.shared_map = True / False`
        overrides the test (True is a promise by the caller, False never looks)."""
        if grid.shape[0] == 1:
            return grid
        if grid.stride(0) == 0:
            return grid[:1]
        hint = self.shared_map
        if hint is True:
            return grid[:1]
        if hint is None and grid.shape[0] == B and B > 1 and not (torch.is_grad_enabled() and grid.requires_grad):
            if bool((grid == grid[:1]).all()):
                return grid[:1]
        return grid

    def _compute_device(self):
        """The CUDA device the kernels run on.  `DPhysics(cfg)` with the reference's default device='cpu'
        (scripts/fit_terrain.py:34) means HOST tensors in and out: they are staged through cuda:<current> (differentiable
        `.to()` copies) - the arithmetic is still the sm_100a kernels, there is no CPU implementation to fall back to."""
        d = torch.device(self.device)
        if d.type == 'cuda':
            return d
        if not torch.cuda.is_available():
            raise RuntimeError("monoforce_b200.DPhysics needs a CUDA device (sm_100a kernels, no CPU fallback); "
                               "torch.cuda.is_available() is False")
        return torch.device('cuda', torch.cuda.current_device())

    def dphysics(self, z_grid, controls, joint_angles=None, state=None, friction=None):
        """dphysics.py:530-594.  `z_grid` / `friction` may also be given as (1,H,W): one map shared
        by all `controls.shape[0]` trajectories (the reference's callers pass B copies)."""
        cfg = self.dphys_cfg
        dt, T = cfg.dt, cfg.traj_sim_time
        # the reference reads integration_mode in update_state (:282-284, 'euler' | 'rk4') and hands it to torchdiffeq as
        # `method` (:510-511); both kernels implement Euler only, so anything else must not silently run as Euler
        mode = getattr(cfg, 'integration_mode', 'euler')
        if mode != 'euler':
            if mode != 'rk4' and not cfg.use_odeint:
                raise ValueError(f'Unknown integration mode: {mode}')          # dphysics.py:382
            raise NotImplementedError(
                f"integration_mode={mode!r} is not implemented by the sm_100a rollout kernels (Euler only, which is what "
                f"every caller of the reference uses); refusing to run it as Euler")
        host_io = torch.device(self.device).type != 'cuda'
        controls = torch.as_tensor(controls).to(self.device)
        # the reference takes B from z_grid.shape[0] (:551).  Extension: M maps for B = k*M control sequences means
        # "k consecutive trajectories per map" (one map per scene, many shots per scene); M == 1 is one shared map
        M, Bc = z_grid.shape[0], controls.shape[0]
        batch_size = Bc if (M != Bc and Bc % M == 0) else M
        dtype = controls.dtype

        if state is None:                                                       # :554-559
            x = torch.zeros(batch_size, 3, device=self.device, dtype=dtype)
            xd = torch.zeros_like(x)
            xd[:, 0] = controls[:, 0, 0]
            R = torch.eye(3, device=self.device, dtype=dtype).repeat(batch_size, 1, 1)
            omega = torch.zeros_like(x)
            omega[:, 2] = controls[:, 0, 1]
            state = (x, xd, R, omega)

        N_ts = min(int(T / dt), controls.shape[1])                              # :573
        B = state[0].shape[0]
        assert controls.shape == (B, N_ts, 2), f'Controls shape {controls.shape} != {(B, N_ts, 2)}'
        self.controls = controls
        self._moving_joints = False
        if joint_angles is not None:
            assert joint_angles.shape == (B, N_ts, 4), f'Joint angles shape {joint_angles.shape} != {(B, N_ts, 4)}'
            # the reference articulates the body only for marv with non-zero angles (:340; one host sync per CALL
            # here instead of one per step there)
            self._moving_joints = cfg.robot == 'marv' and bool(torch.any(joint_angles != 0))
        # (zero angles are never read by the kernels: a stride-0 view keeps the attribute's shape without a fill per call)
        self.joint_angles = joint_angles if joint_angles is not None else \
            self._zero_scalar(dtype).expand(B, N_ts, 4)
        self.ts = self.ts[:N_ts]                                                # :581
        if self.ts.shape[0] != N_ts:
            raise AssertionError(f'time grid has {self.ts.shape[0]} samples, need {N_ts}')

        cdev = self._compute_device()
        if friction is None:                                                    # :562 (no B copies: stride-0 view)
            friction = self._default_friction()
        self.z_grid = z_grid.to(self.device)                                    # :563-564
        self.friction = friction.to(self.device)
        self._z_arg = self._shared_view(self.z_grid, batch_size).to(cdev)
        self._mu_arg = self._shared_view(self.friction, batch_size).to(cdev)
        nz, nm = self._z_arg.shape[0], self._mu_arg.shape[0]
        if nz != nm:
            # different sharing for height and friction: materialise the coarser one at the finer grouping
            if max(nz, nm) % min(nz, nm) != 0:
                raise AssertionError(f'z_grid has {nz} maps but friction has {nm}')
            if nz < nm:
                self._z_arg = torch.repeat_interleave(self._z_arg, nm // nz, dim=0)
            else:
                self._mu_arg = torch.repeat_interleave(self._mu_arg, nz // nm, dim=0)
        if batch_size % self._z_arg.shape[0] != 0:
            raise AssertionError(f'{self._z_arg.shape[0]} maps cannot be shared evenly by {batch_size} trajectories')

        integrator = self.dynamics_odeint if cfg.use_odeint else self.dynamics   # read at call time
        if host_io:
            self.controls = controls.to(cdev)
            try:
                outs = integrator(tuple(s.to(cdev) for s in state))
            finally:
                self.controls = controls
            Xs, Xds, Rs, Omegas, F_springs, F_frictions = (o.to(self.device) for o in outs)
        else:
            Xs, Xds, Rs, Omegas, F_springs, F_frictions = integrator(state)

        # the reference snaps the start height into the caller's tensor in place (:571)
        with torch.no_grad():
            state[0][..., 2] = self._x0z.to(device=state[0].device, dtype=state[0].dtype)
        return (Xs, Xds, Rs, Omegas), (F_springs, F_frictions)

    def forward(self, z_grid, controls, joint_angles=None, state=None, vis=False, friction=None):
        states, forces = self.dphysics(z_grid=z_grid, controls=controls, joint_angles=joint_angles, state=state,
                                       friction=friction)
        if vis:
            with torch.no_grad():
                self.visualize(states=states, z_grid=z_grid)
        return states, forces

    def graphed(self, z_grid, controls, friction=None):
        """CUDA-graph replay of the planner's call (monoforce_node.py:55-96: every cycle the same shapes, a new height map and the
        same or new control sequences, no gradients): `run = sim.graphed(z_grid, controls)` captures ONE no_grad call
        `self(z_grid, controls, friction=friction)` (default initial state) and returns a callable;
        `states, forces = run(z_grid=new_map)` copies the new inputs into the captured buffers and replays the launches - the
        host-side work of a call (tensor bookkeeping, ctypes marshalling: ~0.26 ms, a third of the 64 x 500 shooting call) is
        paid once.  The returned tensors are the captured outputs and are overwritten by the next replay; `run.cost` is the
        fused per-trajectory cost when `fused_cost` is set.  Shapes, dtype, device and the sharing structure of the maps
        ((1,H,W) or stride-0 views are shared; B repeated copies are taken as ONE shared map here) are fixed at capture."""
        if torch.device(self.device).type != 'cuda':
            raise RuntimeError("DPhysics.graphed needs device='cuda' (the captured buffers live on the device)")
        dev = torch.device(self.device)
        controls = torch.as_tensor(controls).to(dev)
        B = controls.shape[0]

        def static_map(m):
            if m is None:
                return None
            m = m.to(dev)
            shared = m.shape[0] == 1 or m.stride(0) == 0 or (m.shape[0] == B and B > 1 and bool((m == m[:1]).all()))
            return m[:1].clone() if shared else m.clone()
        s_z, s_mu, s_u = static_map(z_grid), static_map(friction), controls.clone()
        hint = self.shared_map
        self.shared_map = False                 # the captured call must not look at the maps (a device read-back cannot be captured)
        try:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side), torch.no_grad():      # warm-up outside the capture: workspaces, lazy module loading
                self(s_z, s_u, friction=s_mu)
            torch.cuda.current_stream(dev).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph), torch.no_grad():
                out = self(s_z, s_u, friction=s_mu)
                cost = self.last_cost
        finally:
            self.shared_map = hint

        def fill(dst, src, name):
            if src is None:
                return
            src = src.to(dev)
            if dst.shape[0] == 1 and src.shape[0] != 1:
                src = src[:1]                                   # repeated copies of the one shared map
            if src.shape != dst.shape:
                raise ValueError(f"{name}: shape {tuple(src.shape)} differs from the captured {tuple(dst.shape)}")
            dst.copy_(src, non_blocking=True)

        def run(z_grid=None, controls=None, friction=None):
            fill(s_z, z_grid, "z_grid")
            fill(s_u, controls, "controls")
            if friction is not None:
                if s_mu is None:
                    raise ValueError("friction was not part of the captured call")
                fill(s_mu, friction, "friction")
            graph.replay()
            return out
        run.graph, run.cost, run.outputs = graph, cost, out
        return run

    def update_joints(self, joint_angles):
        """Body points for the given flipper angles (B,4) -> (B,N,3), dphysics.py:326-358 (host-side,
        used by visualisation)."""
        B = joint_angles.shape[0]
        x_points = self.x_points.repeat(B, 1, 1)
        if self.dphys_cfg.robot != 'marv' or not bool(torch.any(joint_angles != 0)):
            return x_points
        pivots = list(self.dphys_cfg.joint_positions.values())
        for i, mask in enumerate(self.dphys_cfg.driving_parts):
            piv = torch.as_tensor(pivots[i], dtype=x_points.dtype, device=x_points.device).view(1, 1, 3)
            a = joint_angles[:, i]
            c, s, o, l = torch.cos(a), torch.sin(a), torch.zeros_like(a), torch.ones_like(a)
            Ry = torch.stack([c, o, s, o, l, o, -s, o, c], dim=1).view(B, 3, 3)
            x_points[:, mask] = (x_points[:, mask] - piv) @ Ry.transpose(1, 2) + piv
        return x_points

    def visualize(self, states, z_grid, forces=None, states_gt=None, friction=None):
        """dphysics.py:607-669 needs mayavi (a GUI dependency outside the hot path)."""
        try:
            from mayavi import mlab  # noqa: F401
        except ImportError as e:
            raise RuntimeError("DPhysics.visualize needs mayavi, which is not part of monoforce_b200") from e
        raise NotImplementedError("visualisation is out of scope of the B200 hot path (SURVEY.md section 8)")
