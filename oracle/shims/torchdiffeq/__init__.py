"""Stand-in for torchdiffeq==0.2.3 (docker/requirements.txt:22), fixed-grid Euler only.

TEST INFRASTRUCTURE ONLY.  The reference calls
``odeint(f, y0_tuple, ts, method='euler', rtol=1e-3, atol=1e-3)`` (dphysics.py:510-511).
torchdiffeq's fixed-grid Euler solver (published algorithm; source not in this
container => "parity unpinned" for this dependency) advances

    y_{k+1} = y_k + (t_{k+1} - t_k) * f(t_k, y_k)

on exactly the user grid ``ts`` (no sub-stepping when ``step_size`` is not given),
returns the solution stacked on a new leading time axis with ``y[0] = y0``, and
ignores rtol/atol.
"""
import torch


def odeint(func, y0, t, method="euler", rtol=None, atol=None, **kw):
    if method != "euler":
        raise NotImplementedError("stand-in implements fixed-grid euler only")
    is_tuple = isinstance(y0, (tuple, list))
    y = tuple(y0) if is_tuple else (y0,)
    sol = [[yi] for yi in y]
    for k in range(len(t) - 1):
        t0, t1 = t[k], t[k + 1]
        dy = func(t0, y if is_tuple else y[0])
        dy = tuple(dy) if is_tuple else (dy,)
        h = t1 - t0
        y = tuple(yi + h * di for yi, di in zip(y, dy))
        for s, yi in zip(sol, y):
            s.append(yi)
    out = tuple(torch.stack(s, dim=0) for s in sol)
    return out if is_tuple else out[0]
