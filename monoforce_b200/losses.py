"""physics_loss: the training objective that seeds the rollout adjoint.

Mirrors `monoforce/src/monoforce/losses.py:102-138` (time-weighted MSE between predicted and
ground-truth positions at the nearest predicted time stamps; the optional rotation term, used only
by scripts/eval.py:151, is a handful of torch ops on top).

CUDA tensors go through ONE fused kernel (csrc/physics_loss.cu, C entry point `mfb_physics_loss`) that
does the nearest-stamp search, the weighted squared residuals and d loss / d X_pred in the same pass,
so neither the reference's (N, T2, T1) distance matrix nor its gather exist.  Host tensors (the unit
tests of the definition) use the same formula written with torch ops.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


_INCREASING = {}      # (data_ptr, version, shape, stride) -> bool; one device->host read per distinct stamp tensor


def _strictly_increasing(ts):
    key = (ts.data_ptr(), ts._version, tuple(ts.shape), ts.stride(), str(ts.device))
    hit = _INCREASING.get(key)
    if hit is None:
        if len(_INCREASING) > 64:
            _INCREASING.clear()
        hit = bool((ts[..., 1:] > ts[..., :-1]).all()) if ts.shape[-1] > 1 else True
        _INCREASING[key] = hit
    return hit


def _same_grid(pred_ts, gt_ts):
    """True when the nearest-stamp gather (losses.py:116-119) is provably the identity: both arguments are the same
    storage AND the stamps are strictly increasing (with duplicated stamps argmin returns the FIRST of the ties, which
    is not the identity)."""
    same = (pred_ts is gt_ts) or (pred_ts.shape == gt_ts.shape and pred_ts.data_ptr() == gt_ts.data_ptr()
                                  and pred_ts.stride() == gt_ts.stride())
    return same and _strictly_increasing(gt_ts)


def rotation_difference(R1, R2, reduction='mean'):
    """Squared geodesic angle between rotations (losses.py:48-65)."""
    assert R1.shape == R2.shape and R1.shape[-2:] == (3, 3)
    dR = R1 @ R2.transpose(dim0=-2, dim1=-1)
    tr = dR.diagonal(dim1=-2, dim2=-1).sum(dim=-1, keepdim=True)
    theta = torch.arccos(torch.clip((tr - 1) / 2., min=-1, max=1.)) ** 2
    if reduction == 'mean':
        return theta.mean()
    if reduction == 'sum':
        return theta.sum()
    return theta


def translation_difference(x1, x2, reduction='mean'):
    """losses.py:36-45."""
    assert x1.shape == x2.shape and x1.shape[-1] == 3
    d = torch.norm(x1 - x2, dim=-1)
    return d.mean() if reduction == 'mean' else d.sum() if reduction == 'sum' else d


def total_variation(heightmap):
    """losses.py:68-74 (fit_terrain.py:58)."""
    h, w = heightmap.shape[-2:]
    tv = torch.sum(torch.abs(heightmap[..., :, :-1] - heightmap[..., :, 1:])) + \
        torch.sum(torch.abs(heightmap[..., :-1, :] - heightmap[..., 1:, :]))
    return tv / (h * w)


def hm_loss(height_pred, height_gt, weights=None, h_max=None):
    """Weighted MSE between height maps, NaN cells ignored - losses.py:77-99 (train.py:387-395)."""
    assert height_pred.shape == height_gt.shape, 'Height prediction and ground truth must have the same shape'
    if weights is None:
        weights = torch.ones_like(height_gt)
    assert weights.shape == height_gt.shape, 'Weights and height ground truth must have the same shape'
    if h_max is not None:
        height_pred = h_max * torch.tanh(height_pred)
    valid = ~(torch.isnan(height_pred) | torch.isnan(height_gt))
    height_gt, height_pred, weights = height_gt[valid], height_pred[valid], weights[valid]
    return ((height_pred * weights - height_gt * weights) ** 2).mean()


def _rotation_term(states_pred, states_gt, pred_ts, gt_ts, gamma):
    """losses.py:129-134 (torch ops: eval.py:151 is the only caller; not on the training hot path)."""
    R = states_gt[2]
    dev = R.device
    pred_ts, gt_ts = pred_ts.to(dev), gt_ts.to(dev)
    if _same_grid(pred_ts, gt_ts) and states_pred[2].shape[1] == gt_ts.shape[-1]:
        R_pred = states_pred[2]
    else:
        ts_ids = torch.argmin(torch.abs(pred_ts.unsqueeze(1) - gt_ts.unsqueeze(2)), dim=2)
        R_pred = states_pred[2][torch.arange(R.shape[0], device=dev).unsqueeze(1), ts_ids]
    time_weights = 1. / (1. + gamma * gt_ts.unsqueeze(2))
    return (rotation_difference(R_pred, R, reduction='none') * time_weights).mean()


class _FusedPhysicsLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X_pred, X_gt, pred_ts, gt_ts, gamma):
        lib = _lib.load()
        dev, dt = X_pred.device, X_pred.dtype
        B, T1, _ = X_pred.shape
        T2 = X_gt.shape[1]
        same = _same_grid(pred_ts, gt_ts) and T1 == T2
        Xp, Xg = X_pred.contiguous(), X_gt.to(dt).contiguous()

        def rows(ts):            # (B|1, T) -> (contiguous 2-D tensor, elements between two trajectories' rows)
            ts = ts.to(device=dev, dtype=dt)
            ts = ts.reshape(1, -1) if ts.dim() == 1 else ts
            assert ts.shape[0] in (1, B), f"time stamps must be (B, T) or (1, T), got {tuple(ts.shape)}"
            if ts.shape[0] == B and B > 1 and ts.stride(0) == 0:
                ts = ts[:1]
            ts = ts.contiguous()
            return ts, (0 if ts.shape[0] == 1 else ts.shape[1])

        gts, gs = rows(gt_ts)
        pts, ps = (gts, gs) if same else rows(pred_ts)
        assert gts.shape[1] == T2 and pts.shape[1] == T1, (tuple(gts.shape), tuple(pts.shape), T1, T2)
        want_grad = ctx.needs_input_grad[0]
        loss = torch.empty((), dtype=dt, device=dev)
        g = (torch.empty_like(Xp) if same else torch.zeros_like(Xp)) if want_grad else None
        scratch = torch.empty(_lib.MFB_PHYSICS_LOSS_MAX_BLOCKS, dtype=torch.float64, device=dev)
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(lib.mfb_physics_loss(p(Xp), p(Xg), p(pts), p(gts), ps, gs, B, T1, T2, float(gamma), int(same),
                                            p(loss), p(g), p(scratch), _lib.MFB_F32 if dt == torch.float32 else _lib.MFB_F64,
                                            C.c_void_p(st)), "mfb_physics_loss")
        ctx.save_for_backward(g)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        (g,) = ctx.saved_tensors
        return (None if g is None else g * grad_out), None, None, None, None


def _host_physics_loss(X_pred, X, pred_ts, gt_ts, gamma):
    """HOST tensors only: the reference's definition written with torch ops, for the CPU unit tests of the definition
    (tests/test_host_logic.py).  CUDA tensors never take this path (see physics_loss)."""
    if _same_grid(pred_ts, gt_ts) and X_pred.shape[1] == gt_ts.shape[-1]:
        X_pred_gt_ts = X_pred
    else:
        ts_ids = torch.argmin(torch.abs(pred_ts.unsqueeze(1) - gt_ts.unsqueeze(2)), dim=2)
        X_pred_gt_ts = X_pred[torch.arange(X.shape[0], device=X.device).unsqueeze(1), ts_ids]
    time_weights = 1. / (1. + gamma * gt_ts.unsqueeze(2))
    return ((X_pred_gt_ts * time_weights - X * time_weights) ** 2).mean()


def physics_loss(states_pred, states_gt, pred_ts, gt_ts, gamma=0.9, rotation_loss=False):
    """losses.py:102-138.  Returns `loss`, or `(loss, loss_rot)` with rotation_loss=True (scripts/eval.py:151)."""
    X = states_gt[0]
    X_pred = states_pred[0]
    if X_pred.is_cuda:
        if X_pred.dtype not in (torch.float32, torch.float64):
            raise TypeError(f"physics_loss supports float32 / float64 on CUDA, got {X_pred.dtype}")
        if X.requires_grad:
            raise NotImplementedError("physics_loss: gradients w.r.t. the ground-truth states are not provided by the "
                                      "fused kernel (no caller of the reference differentiates them)")
        loss = _FusedPhysicsLoss.apply(X_pred, X, pred_ts, gt_ts, gamma)
    else:
        loss = _host_physics_loss(X_pred, X, pred_ts, gt_ts, gamma)
    if rotation_loss:
        return loss, _rotation_term(states_pred, states_gt, pred_ts, gt_ts, gamma)
    return loss
