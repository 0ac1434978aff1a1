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
from monoforce_b200 import terrain_encoder as TE  # noqa: E402


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

    def run():
        marks.clear()
        with torch.no_grad():
            mark("start")
            vox = net.cached_voxel_index(*cal)
            mark("voxel_index(cached)")
            cam = net.camencode
            feats = cam.trunk.fast_endpoints(x.view(B * N, C, H, W))
            mark("efficientnet trunk")
            y = cam.up1.fast_nhwc(feats[4], feats[3])
            mark("cam up1 (upsample+cat+2 conv)")
            logits = cam.fast_logits_from_up(y) if hasattr(cam, "fast_logits_from_up") else None
            if logits is None:
                f = TE._folded(cam, lambda: TE._fold_padded_cout(cam.depthnet))
                logits = TE.ops.conv_bn_act_nhwc(y, *f, TE.ops.ACT_NONE)[..., :cam.D + cam.C].float().contiguous()
            mark("depthnet + slice/float")
            X, Y = int(net.nx[0]), int(net.nx[1])
            bev = TE._LiftSplat.apply(logits, vox.view(-1), B, N, net.D, net.camC, X, Y).permute(0, 3, 1, 2)
            mark("lift-splat")
            be = net.bevencode
            x1, x3 = be.fast_backbone_endpoints(bev)
            mark("resnet stem + layer1-3")
            yb = be.up1.fast_nhwc(x3, x1)
            mark("bev up1 (upsample+cat+2 conv)")
            out = be.fast_heads(yb) if hasattr(be, "fast_heads") else None
            mark("heads")
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
