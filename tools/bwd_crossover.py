#!/usr/bin/env python
"""Adjoint latency per batch size for both shapes of the single-sweep kernel (one warp per trajectory / one CTA per trajectory)
-> where the MFB_BWD_WIDE_MAX_B threshold belongs.   python tools/bwd_crossover.py [--out gpurun_out/x.json]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    from monoforce_b200 import DPhysics, DPhysConfig
    dev = "cuda"
    rows = []
    for odeint in (True, False):
        cfg = DPhysConfig(robot="marv")
        cfg.use_odeint = odeint
        sim = DPhysics(cfg, device=dev)
        T = int(cfg.traj_sim_time / cfg.dt)
        g = torch.Generator().manual_seed(0)
        xg, yg = cfg.x_grid, cfg.y_grid
        z0 = (torch.exp(-(xg - 2) ** 2 / 4) * torch.exp(-yg ** 2 / 2)).to(dev)[None]
        for B in (1, 16, 64, 128, 256, 384, 512, 1024, 4096):
            ctrl = torch.stack([torch.rand(B, T, generator=g) * 2 - 1, torch.rand(B, T, generator=g) * 4 - 2], -1).to(dev)
            row = {"odeint": odeint, "B": B, "T": T}
            for name, thr in (("warp_ms", "0"), ("wide_ms", str(1 << 30))):
                os.environ["MFB_BWD_WIDE_MAX_B"] = thr
                ts = []
                for it in range(8):
                    z = z0.clone().requires_grad_(True)
                    (Xs, Xd, Rs, Om), _ = sim(z_grid=z, controls=ctrl)
                    loss = Xs.sum() + Rs.sum()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    loss.backward()
                    b.record()
                    b.synchronize()
                    if it >= 2:
                        ts.append(a.elapsed_time(b))
                ts.sort()
                row[name] = round(ts[len(ts) // 2], 4)
            rows.append(row)
            print(row, flush=True)
    if args.out:
        json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
