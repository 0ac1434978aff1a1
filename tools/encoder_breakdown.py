"""Stage-by-stage CUDA-event timing of the encoder fast path (16 scenes): where the milliseconds go.

    python tools/encoder_breakdown.py [--cfg4]
"""
import json
import os
import sys
import warnings

_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [_R, os.path.join(_R, "tests")]
import torch  # noqa: E402
from helpers_lss import default_cfg, make_inputs  # noqa: E402
from monoforce_b200 import LiftSplatShoot  # noqa: E402


def main():
    cfg4 = "--cfg4" in sys.argv
    gc, ac = default_cfg()
    if cfg4:
        gc["xbound"] = [-6.4, 6.4, 0.05]; gc["ybound"] = [-6.4, 6.4, 0.05]; ac["final_dim"] = [512, 512]
    dev = "cuda"
    torch.manual_seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        net = LiftSplatShoot(gc, ac).to(dev).eval()
    net.fast_inference = True
    x, *cal = [t.to(dev) for t in make_inputs(gc, ac, 16, 0)]
    B, N, C, H, W = x.shape
    marks = []

    def mark(name):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        marks.append((name, e))

    from monoforce_b200 import encoder_fast as EF, ops

    def run():
        marks.clear()
        with torch.no_grad():
            P = EF.prepare(net)
            mark("start")
            vox = net.cached_voxel_index(*cal)
            mark("voxel_index(cached)")
            f16, f32 = EF.trunk_endpoints(P, x.view(B * N, C, H, W))
            mark("efficientnet trunk")
            y = EF.up_block(P["cam_up"], f16, f32, 2)
            mark("cam up1 (upsample+cat+2 conv)")
            logits = EF._conv(y, P["depthnet"], pad=0)
            mark("depthnet")
            X, Y = int(net.nx[0]), int(net.nx[1])
            bev = ops.cast_bf16(ops.lift_splat_bf16(logits, vox.view(-1), B, N, net.D, net.camC, X, Y))
            mark("lift-splat + cast")
            x1, x3 = EF.bev_backbone(P, bev)
            mark("conv1 7x7 + resnet layer1-3")
            yb = EF.up_block(P["bev_up"], x1, x3, P["bev_up_scale"])
            mark("bev up1 (upsample+cat+2 conv)")
            up = ops.upsample_concat_nhwc(None, yb, (yb.shape[1] * 2, yb.shape[2] * 2), yb.shape[3])
            mark("heads upsample x2")
            hs = P["heads"]
            out = ops.conv2d_nhwc(up, hs["w"], hs["scale"], hs["shift"], hs["act"], pad=(1, 1), heads=P["head_epilogue"])
            mark("heads conv + fused 1x1 epilogue")
        return out

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    acc = {}
    n = 5
    for _ in range(n):
        run()
        torch.cuda.synchronize()
        for (n0, e0), (n1, e1) in zip(marks[:-1], marks[1:]):
            acc[n1] = acc.get(n1, 0.0) + e0.elapsed_time(e1) / n
    tot = sum(acc.values())
    with torch.no_grad():
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            net(x, *cal)
        b.record()
        torch.cuda.synchronize()
    print(json.dumps({"config": "cfg4" if cfg4 else "default", "stages_ms": acc, "sum_ms": tot, "forward_ms": a.elapsed_time(b) / n}, indent=1))


if __name__ == "__main__":
    main()
