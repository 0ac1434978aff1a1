"""Device-memory bandwidth by access mix on this GPU (context for the HBM rooflines): write-only (fill), read-only (sum), copy.
    python tools/hbm_modes.py"""
import torch

def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2] * 1e-3

n = 1 << 30                      # 4 GiB of fp32
x = torch.empty(n, dtype=torch.float32, device="cuda")
y = torch.empty(n, dtype=torch.float32, device="cuda")
x.fill_(1.0)
t = timed(lambda: x.zero_());            print(f"write-only (zero_ 4 GiB):  {4 * n / t / 1e9:7.0f} GB/s")
t = timed(lambda: x.fill_(2.0));         print(f"write-only (fill_ 4 GiB):  {4 * n / t / 1e9:7.0f} GB/s")
t = timed(lambda: x.sum());              print(f"read-only  (sum 4 GiB):    {4 * n / t / 1e9:7.0f} GB/s")
t = timed(lambda: y.copy_(x));           print(f"copy (4 GiB -> 4 GiB):     {8 * n / t / 1e9:7.0f} GB/s  (read + write bytes)")
h = x.view(torch.bfloat16)[: n]
t = timed(lambda: h.copy_(y));           print(f"fp32 -> bf16 cast:         {6 * n / t / 1e9:7.0f} GB/s  (4 B read + 2 B written per element)")
