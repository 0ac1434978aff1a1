"""BASELINE config 4 timing: TerrainEncoder (4 cams -> BEV terrain/friction) + DPhysics rollout, 16 scenes, 1 GPU.

    python tools/bench_encoder.py [--cfg4 | --default] [--scenes 16]

Prints one JSON object: encoder ms (fp32 cuDNN path vs bf16 tcgen05 path), per-layer tensor-core throughput of the
K4 convolution kernel, rollout ms, scenes/s."""
import argparse
import json
import os
import sys

_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [_R, os.path.join(_R, "tests")]
import torch  # noqa: E402
from helpers_lss import default_cfg, make_inputs  # noqa: E402
from monoforce_b200 import DPhysics, DPhysConfig, LiftSplatShoot, ops  # noqa: E402


def timed(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg4", action="store_true", help="512x512 images, 256x256 BEV (0.05 m) instead of lss_cfg.yaml")
    ap.add_argument("--scenes", type=int, default=16)
    ap.add_argument("--trajs", type=int, default=256, help="trajectories per scene for the rollout")
    args = ap.parse_args()
    grid_conf, aug_conf = default_cfg()
    if args.cfg4:
        grid_conf["xbound"] = [-6.4, 6.4, 0.05]
        grid_conf["ybound"] = [-6.4, 6.4, 0.05]
        aug_conf["final_dim"] = [512, 512]
    dev = "cuda"
    torch.manual_seed(0)
    net = LiftSplatShoot(grid_conf, aug_conf).to(dev).eval()
    B = args.scenes
    inputs = [t.to(dev) for t in make_inputs(grid_conf, aug_conf, B, 0)]
    res = {"config": "cfg4 (512x512 imgs -> 256x256 BEV)" if args.cfg4 else "lss_cfg.yaml (256x416 imgs -> 128x128 BEV)",
           "scenes": B, "cams": 4}
    with torch.no_grad():
        torch.backends.cudnn.allow_tf32 = False
        res["encoder_fp32_cudnn_ms"] = timed(lambda: net(*inputs))
        torch.backends.cudnn.allow_tf32 = True
        res["encoder_tf32_cudnn_ms"] = timed(lambda: net(*inputs))
        net.fast_inference = True
        res["encoder_fast_tcgen05_ms"] = timed(lambda: net(*inputs))
        out = net(*inputs)
        # individual K4 layers at this configuration
        layers = []
        X = int(net.nx[0])
        fH, fW = aug_conf["final_dim"][0] // 16, aug_conf["final_dim"][1] // 16
        for name, (N_, H_, W_, Cin, Cout, KS) in {
            "camencode.up1.conv0": (B * 4, fH, fW, 448, 512, 3), "camencode.up1.conv3": (B * 4, fH, fW, 512, 512, 3),
            "camencode.depthnet": (B * 4, fH, fW, 512, 128, 1),
            "bevencode.up1.conv0": (B, X // 2, X // 2, 320, 256, 3), "bevencode.up1.conv3": (B, X // 2, X // 2, 256, 256, 3),
            "bevencode.heads(3 fused)": (B, X, X, 256, 384, 3)}.items():
            x = torch.randn(N_, H_, W_, Cin, device=dev).to(torch.bfloat16)
            w = (torch.randn(Cout, KS, KS, Cin, device=dev) * 0.02).to(torch.bfloat16)
            sc, sh = torch.ones(Cout, device=dev), torch.zeros(Cout, device=dev)
            ms = timed(lambda: ops.conv_bn_act_nhwc(x, w, sc, sh, ops.ACT_GELU), n=20)
            flops = 2.0 * N_ * H_ * W_ * Cin * Cout * KS * KS
            layers.append({"layer": name, "shape": [N_, H_, W_, Cin, Cout, KS], "ms": ms, "tflops": flops / ms / 1e9})
        res["k4_layers"] = layers
        res["k4_total_ms"] = sum(l["ms"] for l in layers)
        res["k4_total_tflops"] = sum(l["tflops"] * l["ms"] for l in layers) / res["k4_total_ms"]
    # rollout on the predicted maps: `trajs` trajectories per scene, distinct map per scene
    cfg = DPhysConfig(robot="marv", grid_res=float(grid_conf["xbound"][2]))
    cfg.traj_sim_time, cfg.use_odeint = 4.0, False
    sim = DPhysics(cfg, device=dev)
    T = 400
    n_traj = B * args.trajs
    g = torch.Generator().manual_seed(0)
    ctrl = torch.stack([torch.rand(n_traj, generator=g) * 0.5 + 0.5, torch.rand(n_traj, generator=g) * 4 - 2], -1)
    ctrl = ctrl.unsqueeze(1).repeat(1, T, 1).to(dev)
    z = out["terrain"].squeeze(1).repeat_interleave(args.trajs, dim=0)
    fr = out["friction"].squeeze(1).repeat_interleave(args.trajs, dim=0)
    with torch.no_grad():
        sim.timings = []
        for _ in range(3):
            sim(z, ctrl, friction=fr)
        torch.cuda.synchronize()
        res["rollout_ms"] = min(a.elapsed_time(b) for _, a, b in sim.timings)
    res["rollout_trajectories"] = n_traj
    res["scenes_per_s_fast"] = B / ((res["encoder_fast_tcgen05_ms"] + res["rollout_ms"]) * 1e-3)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
