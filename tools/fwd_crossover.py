#!/usr/bin/env python
"""Forward-only rollout latency per batch size for both forward kernels (K1: one warp per trajectory, K1w: one CTA per
trajectory) -> where the MFB_FWD_WIDE_MAX_B threshold belongs.   python tools/fwd_crossover.py [--out gpurun_out/x.json]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


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
        z = (torch.exp(-(xg - 2) ** 2 / 4) * torch.exp(-yg ** 2 / 2)).to(dev)[None]
        for B in (1, 16, 64, 128, 256, 384, 512, 768, 1024, 2048, 4096):
            ctrl = torch.stack([torch.rand(B, T, generator=g) * 2 - 1, torch.rand(B, T, generator=g) * 4 - 2], -1).to(dev)
            row = {"odeint": odeint, "B": B, "T": T}
            for name, thr in (("warp_ms", "0"), ("wide_ms", str(1 << 30))):
                os.environ["MFB_FWD_WIDE_MAX_B"] = thr

                def fwd():
                    with torch.no_grad():
                        sim(z_grid=z, controls=ctrl)
                row[name] = round(timed(fwd), 4)
            rows.append(row)
            print(row, flush=True)
    if args.out:
        json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
