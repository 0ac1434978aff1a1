"""physics_loss: the training objective that seeds the rollout adjoint.

Mirrors `monoforce/src/monoforce/losses.py:102-138` (time-weighted MSE between predicted and
ground-truth positions at the nearest predicted time stamps; optional rotation term omitted
as in every shipped caller: train.py:405-406, fit_terrain.py:57).

CUDA tensors go through ONE fused kernel (csrc/physics_loss.cu, C entry point `mfb_physics_loss`) that
does the nearest-stamp search, the weighted squared residuals and d loss / d X_pred in the same pass,
so neither the reference's (N, T2, T1) distance matrix nor its gather exist.  Host tensors (the unit
tests of the definition) use the same formula written with torch ops.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def _same_grid(pred_ts, gt_ts):
    return (pred_ts is gt_ts) or (pred_ts.shape == gt_ts.shape and pred_ts.data_ptr() == gt_ts.data_ptr()
                                  and pred_ts.stride() == gt_ts.stride())


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
    if rotation_loss:
        raise NotImplementedError("rotation_loss=True is not used by any caller on the hot path")
    X = states_gt[0]
    X_pred = states_pred[0]
    if X_pred.is_cuda:
        if X_pred.dtype not in (torch.float32, torch.float64):
            raise TypeError(f"physics_loss supports float32 / float64 on CUDA, got {X_pred.dtype}")
        if X.requires_grad:
            raise NotImplementedError("physics_loss: gradients w.r.t. the ground-truth states are not provided by the "
                                      "fused kernel (no caller of the reference differentiates them)")
        return _FusedPhysicsLoss.apply(X_pred, X, pred_ts, gt_ts, gamma)
    return _host_physics_loss(X_pred, X, pred_ts, gt_ts, gamma)
