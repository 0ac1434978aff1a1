"""Diagnostic: per-call forward / backward times over a long back-to-back run + NVML clocks / power (not the bench)."""
import sys, os, time, threading
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _R); sys.path.insert(0, os.path.join(_R, 'tests'))
import torch, pynvml
from bench import synth_inputs
from monoforce_b200 import DPhysics
from monoforce_b200.losses import physics_loss

pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
d = synth_inputs(4096, 0)
sim = DPhysics(d["cfg"], device="cuda"); sim.fused_cost = True
controls = d["controls"].cuda(); ts = d["ts"].cuda()
with torch.no_grad():
    gt, _ = sim(d["z_gt"].cuda().unsqueeze(0), controls)
z = d["z0"].cuda().unsqueeze(0).requires_grad_(True)
fr = d["fr0"].cuda().unsqueeze(0).requires_grad_(True)
samples = []; stop = False
def poll():
    while not stop:
        samples.append((time.perf_counter(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                        pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0, pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)))
        time.sleep(0.005)
for mode in ("sync_each_step", "back_to_back"):
    for _ in range(3):
        z.grad = None; fr.grad = None
        st, _ = sim(z, controls, friction=fr); physics_loss(st, gt, ts, ts, 0.9).backward()
    torch.cuda.synchronize()
    samples.clear(); stop = False
    th = threading.Thread(target=poll, daemon=True); th.start()
    sim.timings = []
    t0 = time.perf_counter()
    for it in range(40):
        z.grad = None; fr.grad = None
        st, _ = sim(z, controls, friction=fr); physics_loss(st, gt, ts, ts, 0.9).backward()
        if mode == "sync_each_step":
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    stop = True; th.join()
    f = [a.elapsed_time(b) for n, a, b in sim.timings if n == "forward"]
    bw = [a.elapsed_time(b) for n, a, b in sim.timings if n == "backward"]
    sim.timings = None
    clk = [s[1] for s in samples]; pw = [s[2] for s in samples]; rs = 0
    for s in samples: rs |= s[3]
    print(f"{mode}: wall/step {wall/40*1e3:.2f} ms  fwd mean {sum(f)/len(f):.3f} (min {min(f):.3f} max {max(f):.3f})  "
          f"bwd mean {sum(bw)/len(bw):.3f} (min {min(bw):.3f} max {max(bw):.3f})  clocks min/median/max {min(clk)}/{sorted(clk)[len(clk)//2]}/{max(clk)} MHz "
          f"power max {max(pw):.0f} W  reasons 0x{rs:x}  n_samples {len(samples)}")
    print("   bwd per step:", " ".join(f"{x:.2f}" for x in bw[:40]))
