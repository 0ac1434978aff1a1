"""Depthwise conv + BN + SiLU (+ squeeze-excite pool) layer by layer at the EfficientNet-B0 shapes: time (CUDA events, L2 flushed
between runs), GB/s against the activation bytes read + written, and the fp32 FMA rate.

    python tools/dw_bench.py [--cfg4] [--only i] [--reps n]
"""
import argparse
import os
import sys
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _R)
import torch
from monoforce_b200 import ops

# (C, K, stride, spatial divisor of the input image) for the 16 MBConv blocks of B0
LAYERS = [(32, 3, 1, 2), (96, 3, 2, 2), (144, 3, 1, 4), (144, 5, 2, 4), (240, 5, 1, 8), (240, 3, 2, 8), (480, 3, 1, 16), (480, 3, 1, 16),
          (480, 5, 1, 16), (672, 5, 1, 16), (672, 5, 1, 16), (672, 5, 2, 16), (1152, 5, 1, 32), (1152, 5, 1, 32), (1152, 5, 1, 32),
          (1152, 3, 1, 32)]


def same_pad(size, k, s):
    out = -(-size // s)
    tot = max((out - 1) * s + k - size, 0)
    return tot // 2, tot - tot // 2


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg4", action="store_true")
    ap.add_argument("--only", type=int, default=-1)
    ap.add_argument("--reps", type=int, default=10)
    a = ap.parse_args()
    Hi, Wi = (512, 512) if a.cfg4 else (256, 416)
    N = 64
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    tot = 0.0
    for i, (C, K, s, div) in enumerate(LAYERS):
        if a.only >= 0 and i != a.only:
            continue
        H, W = Hi // div, Wi // div
        x = torch.randn(N, H, W, C, device="cuda").to(torch.bfloat16)
        w = torch.randn(K * K, C, device="cuda") * 0.2
        shift = torch.randn(C, device="cuda") * 0.1
        pool = torch.zeros(N, C, device="cuda")
        pad = same_pad(H, K, s)
        ts = []
        for r in range(a.reps + 2):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            y = ops.dwconv_bn_silu(x, w, shift, K, s, pad, pool)
            e1.record()
            e1.synchronize()
            if r >= 2:
                ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        us = ts[len(ts) // 2]
        byt = (x.numel() + y.numel()) * 2
        fma = y.numel() * K * K
        tot += us
        print(f"{i:2d} C={C:4d} k{K} s{s} {H:3d}x{W:3d} -> {y.shape[1]:3d}x{y.shape[2]:3d}: {us:7.1f} us  {byt / 1e6:7.1f} MB {byt / us / 1e3:6.0f} GB/s  "
              f"{fma / us / 1e6:5.1f} TFMA/s", flush=True)
    print(f"total {tot:.1f} us")


if __name__ == "__main__":
    main()
