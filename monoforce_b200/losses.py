"""physics_loss: the training objective that seeds the rollout adjoint.

Mirrors `monoforce/src/monoforce/losses.py:102-138` (time-weighted MSE between predicted and
ground-truth positions at the nearest predicted time stamps; optional rotation term omitted
as in every shipped caller: train.py:405-406, fit_terrain.py:57).  When the two time grids are
the same tensor the nearest-stamp search is the identity and the (N, T2, T1) distance matrix of
the reference is never built.
"""
from __future__ import annotations

import torch


def physics_loss(states_pred, states_gt, pred_ts, gt_ts, gamma=0.9, rotation_loss=False):
    if rotation_loss:
        raise NotImplementedError("rotation_loss=True is not used by any caller on the hot path")
    X = states_gt[0]
    X_pred = states_pred[0]
    same_grid = (pred_ts is gt_ts) or (pred_ts.shape == gt_ts.shape and pred_ts.data_ptr() == gt_ts.data_ptr())
    if same_grid and X_pred.shape[1] == gt_ts.shape[-1]:
        X_pred_gt_ts = X_pred
    else:
        ts_ids = torch.argmin(torch.abs(pred_ts.unsqueeze(1) - gt_ts.unsqueeze(2)), dim=2)
        X_pred_gt_ts = X_pred[torch.arange(X.shape[0], device=X.device).unsqueeze(1), ts_ids]
    time_weights = 1. / (1. + gamma * gt_ts.unsqueeze(2))
    return ((X_pred_gt_ts * time_weights - X * time_weights) ** 2).mean()
