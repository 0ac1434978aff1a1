"""TerrainEncoder: the Lift-Splat-Shoot network that turns N camera images into BEV terrain maps.

Host-side mirror of the reference's `LiftSplatShoot` (terrain_encoder/lss.py:167-302) with the same
constructor, `forward(x, rots, trans, intrins, post_rots, post_trans)` signature, output dict
(`geom`, `terrain`, `diff`, `friction`), sub-module names and `state_dict` keys, so released
checkpoints load and `scripts/run.py:53-55,145` / `train.py:369-371,385` work unchanged.

What runs where:
  * "lift" + "splat" (`CamEncode.get_depth_feat` outer product lss.py:63-71 and `voxel_pooling`
    lss.py:238-280 incl. the sort + `QuickCumsum` trick, utils.py:155-181) are ONE fused CUDA kernel
    (csrc/lift_splat.cu): per frustum pixel, depth soft-max in registers, then depth-weighted features
    are scatter-added straight into the channels-last BEV grid.  The (B,N,C,D,fH,fW) lifted tensor
    (395 MB at B=16) and the argsort are never materialised.  Backward is a gather kernel.
  * the module path (training, or eval with gradients) runs the convolutions through torch in fp32: that is the parity
    path against the reference.  With `net.fast_inference = True` (eval mode, no_grad) the WHOLE network runs on repo
    kernels in NHWC bf16 (monoforce_b200/encoder_fast.py): every dense convolution - EfficientNet 1x1 expand / project,
    `Up` blocks, depthnet, conv1 7x7/2, ResNet-18 layer1-3, the heads incl. their 1x1 outputs - as implicit GEMMs on the
    tcgen05 tensor cores (csrc/conv_tcgen05.cuh), the depthwise / squeeze-excite / upsample / stem layers as fused
    memory-bound kernels (csrc/encoder_ops.cu).
The frustum geometry (`get_geometry`, lss.py:204-224) is a handful of tiny 3x3 ops kept in torch.
"""
from __future__ import annotations

import ctypes as C
import os

import torch
from torch import nn
from torch.nn import functional as F
from torchvision.models.resnet import resnet18

from . import _lib, ops
from .efficientnet import EfficientNet

_H_MAX = 2.0      # DPhysConfig().h_max, the default range of ScaledTanh (lss.py:15-19)


def gen_dx_bx(xbound, ybound, zbound):
    """utils.py:136-141: voxel size, first voxel centre, voxel counts."""
    rows = (xbound, ybound, zbound)
    dx = torch.Tensor([r[2] for r in rows])
    bx = torch.Tensor([r[0] + r[2] / 2.0 for r in rows])
    nx = torch.LongTensor([(r[1] - r[0]) / r[2] for r in rows])
    return dx, bx, nx


class ScaledTanh(nn.Module):
    """lss.py:17-24."""

    def __init__(self, min_val=-_H_MAX, max_val=_H_MAX):
        super().__init__()
        self.min_val, self.max_val = min_val, max_val

    def forward(self, x):
        return self.min_val + (self.max_val - self.min_val) * (torch.tanh(x) + 1) / 2


def _conv_bn_gelu(cin, cout):
    return [nn.Conv2d(cin, cout, kernel_size=3, padding=1, bias=False), nn.BatchNorm2d(cout), nn.GELU()]


class Up(nn.Module):
    """Upsample the coarse map, concatenate the skip, two 3x3 conv-BN-GELU (lss.py:27-46)."""

    def __init__(self, in_channels, out_channels, scale_factor=2):
        super().__init__()
        self.up = nn.Upsample(scale_factor=scale_factor, mode='bilinear', align_corners=True)
        self.conv = nn.Sequential(*_conv_bn_gelu(in_channels, out_channels), *_conv_bn_gelu(out_channels, out_channels))

    def forward(self, x1, x2):
        return self.conv(torch.cat([x2, self.up(x1)], dim=1))

class CamEncode(nn.Module):
    """EfficientNet-B0 trunk -> Up(320+112 -> 512) -> 1x1 `depthnet` giving D depth logits + C features
    per /16 pixel (lss.py:49-99).  `forward` returns the logits; soft-max x features happens inside the
    fused lift-splat kernel.  `get_depth_feat` keeps the reference's materialising behaviour for callers
    that want the lifted tensor."""

    def __init__(self, D, C, in_channels=3, trunk_weights=None):
        super().__init__()
        self.D, self.C = D, C
        # lss.py:55 starts from ImageNet weights (efficientnet_pytorch downloads them).  No network here: the file comes
        # from `trunk_weights` or $MFB_EFFICIENTNET_B0_WEIGHTS (an efficientnet_pytorch state_dict); without one the trunk
        # is randomly initialised and from_pretrained says so loudly.
        self.trunk = EfficientNet.from_pretrained("efficientnet-b0", in_channels=in_channels,
                                                  weights_path=trunk_weights or os.environ.get("MFB_EFFICIENTNET_B0_WEIGHTS"))
        self.up1 = Up(320 + 112, 512)
        self.depthnet = nn.Conv2d(512, self.D + self.C, kernel_size=1, padding=0)

    def get_eff_depth(self, x):
        """Trunk with the /16 (112 ch) and /32 (320 ch) endpoints merged by `up1` (lss.py:73-94)."""
        t = self.trunk
        x = t._swish(t._bn0(t._conv_stem(x)))
        feats, prev = [], x
        n_blocks = len(t._blocks)
        for i, block in enumerate(t._blocks):
            rate = t._global_params.drop_connect_rate
            if rate:
                rate *= float(i) / n_blocks
            x = block(x, drop_connect_rate=rate)
            if prev.size(2) > x.size(2):
                feats.append(prev)          # reduction_k = last map before each down-sampling
            prev = x
        feats.append(x)                      # reduction_5
        return self.up1(feats[4], feats[3])

    def get_depth_dist(self, x, eps=1e-20):
        return x.softmax(dim=1)

    def depth_logits_and_feats(self, x):
        return self.depthnet(self.get_eff_depth(x))

    def get_depth_feat(self, x):
        x = self.depth_logits_and_feats(x)
        depth = self.get_depth_dist(x[:, :self.D])
        return depth, depth.unsqueeze(1) * x[:, self.D:(self.D + self.C)].unsqueeze(2)

    def forward(self, x):
        return self.get_depth_feat(x)[1]


def _head(outC, act):
    return nn.Sequential(nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True),
                         nn.Conv2d(256, 128, kernel_size=3, padding=1, bias=False), nn.BatchNorm2d(128), nn.GELU(),
                         nn.Conv2d(128, outC, kernel_size=1, padding=0), act)


class BevEncode(nn.Module):
    """ResNet-18 stem/layers 1-3 on the BEV grid, `Up` back to /2, three x2 heads (lss.py:101-165)."""

    def __init__(self, inC, outC):
        super().__init__()
        trunk = resnet18(zero_init_residual=True)
        self.conv1 = nn.Conv2d(inC, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1, self.relu = trunk.bn1, trunk.relu
        self.layer1, self.layer2, self.layer3 = trunk.layer1, trunk.layer2, trunk.layer3
        self.up1 = Up(64 + 256, 256, scale_factor=4)
        self.up_geom = _head(outC, ScaledTanh(-1, 1))
        self.up_diff = _head(outC, nn.ReLU())
        self.up_friction = _head(outC, nn.ReLU())

    def backbone(self, x):
        x1 = self.layer1(self.relu(self.bn1(self.conv1(x))))
        return self.up1(self.layer3(self.layer2(x1)), x1)

    def forward(self, x):
        x = self.backbone(x)
        geom, diff, friction = self.up_geom(x), self.up_diff(x), self.up_friction(x)
        return {'geom': geom, 'terrain': geom - diff, 'diff': diff, 'friction': friction}

# ---------------------------------------------------------------------------------------------
# fused lift + splat
# ---------------------------------------------------------------------------------------------
def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class _LiftSplat(torch.autograd.Function):
    """logits (BN, fH, fW, D+C) fp32 channels-last rows, vox (BN*D*fH*fW,) int32 flat BEV cell or -1
    ->  bev (B, X, Y, C) fp32."""

    @staticmethod
    def forward(ctx, logits, vox, B, N, D, Cc, X, Y):
        lib = _lib.load()
        logits = logits.contiguous()
        BN, fH, fW, _ = logits.shape
        bev = torch.zeros(B, X, Y, Cc, dtype=torch.float32, device=logits.device)
        with torch.cuda.device(logits.device):
            st = torch.cuda.current_stream(logits.device).cuda_stream
            _lib.check(lib.mfb_lift_splat_forward(_p(logits), _p(vox), _p(bev), B, N, D, Cc, fH, fW, X, Y, C.c_void_p(st)),
                       "mfb_lift_splat_forward")
        ctx.save_for_backward(logits, vox)
        ctx.dims = (B, N, D, Cc, fH, fW, X, Y)
        return bev

    @staticmethod
    def backward(ctx, g_bev):
        lib = _lib.load()
        logits, vox = ctx.saved_tensors
        B, N, D, Cc, fH, fW, X, Y = ctx.dims
        g_bev = g_bev.contiguous()
        g_logits = torch.empty_like(logits)
        with torch.cuda.device(logits.device):
            st = torch.cuda.current_stream(logits.device).cuda_stream
            _lib.check(lib.mfb_lift_splat_backward(_p(logits), _p(vox), _p(g_bev), _p(g_logits), B, N, D, Cc, fH, fW, X, Y,
                                                   C.c_void_p(st)), "mfb_lift_splat_backward")
        return g_logits, None, None, None, None, None, None, None


class LiftSplatShoot(nn.Module):
    def __init__(self, grid_conf, data_aug_conf, outC=1, trunk_weights=None):
        super().__init__()
        self.grid_conf = grid_conf
        self.data_aug_conf = data_aug_conf
        dx, bx, nx = gen_dx_bx(grid_conf['xbound'], grid_conf['ybound'], grid_conf['zbound'])
        self.dx = nn.Parameter(dx, requires_grad=False)
        self.bx = nn.Parameter(bx, requires_grad=False)
        self.nx = nn.Parameter(nx, requires_grad=False)
        self.downsample = 16
        self.camC = 64
        self.frustum = self.create_frustum()
        self.D = self.frustum.shape[0]
        self.camencode = CamEncode(self.D, self.camC, trunk_weights=trunk_weights)
        self.bevencode = BevEncode(inC=self.camC, outC=outC)
        self.use_quickcumsum = True      # kept for attribute compatibility; the fused kernel needs neither path
        self.fast_inference = False      # opt-in: the whole network on repo kernels in NHWC bf16 (eval mode, no_grad)
        self.fast_graph = False          # with fast_inference: replay the launches from a CUDA graph (static shapes / calibration)

    def _graphed_forward(self, x, vox):
        """The ~110 kernel launches of the inference path captured once in a CUDA graph and replayed (static shapes, static
        calibration): removes the launch gaps that dominate at small sizes.  Re-captured when the input shape, the calibration
        (voxel index) or any weight changes.  Outputs are copies, so they stay valid across calls."""
        from . import encoder_fast
        P = encoder_fast.prepare(self)            # host-side weight folding must not happen inside the capture
        key = (tuple(x.shape), str(x.device), vox.data_ptr(), id(P))
        g = self.__dict__.get("_mfb_graph")
        if g is None or g["key"] != key:
            static_x = x.detach().float().clone()
            side = torch.cuda.Stream(device=x.device)
            side.wait_stream(torch.cuda.current_stream(x.device))
            with torch.cuda.stream(side):          # warm-up outside the capture (lazy module loading, allocator pools)
                for _ in range(2):
                    encoder_fast.forward(self, static_x, vox)
            torch.cuda.current_stream(x.device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = encoder_fast.forward(self, static_x, vox)
            g = {"key": key, "graph": graph, "x": static_x, "out": out, "P": P, "vox": vox}
            self.__dict__["_mfb_graph"] = g
        g["x"].copy_(x, non_blocking=True)
        g["graph"].replay()
        return {k: v.clone() for k, v in g["out"].items()}

    def create_frustum(self):
        """(D, fH, fW, 3) image-plane sample points (u, v, depth) - lss.py:188-202."""
        H, W = self.data_aug_conf['final_dim']
        fH, fW = H // self.downsample, W // self.downsample
        ds = torch.arange(*self.grid_conf['dbound'], dtype=torch.float).view(-1, 1, 1).expand(-1, fH, fW)
        D = ds.shape[0]
        xs = torch.linspace(0, W - 1, fW, dtype=torch.float).view(1, 1, fW).expand(D, fH, fW)
        ys = torch.linspace(0, H - 1, fH, dtype=torch.float).view(1, fH, 1).expand(D, fH, fW)
        return nn.Parameter(torch.stack((xs, ys, ds), -1), requires_grad=False)

    def get_geometry(self, rots, trans, intrins, post_rots, post_trans):
        """Ego-frame xyz of every frustum point, (B, N, D, fH, fW, 3) - lss.py:204-224."""
        B, N, _ = trans.shape
        pts = self.frustum - post_trans.view(B, N, 1, 1, 1, 3)
        pts = torch.inverse(post_rots).view(B, N, 1, 1, 1, 3, 3).matmul(pts.unsqueeze(-1))
        pts = torch.cat((pts[..., :2, :] * pts[..., 2:3, :], pts[..., 2:3, :]), 5)
        combine = rots.matmul(torch.inverse(intrins))
        pts = combine.view(B, N, 1, 1, 1, 3, 3).matmul(pts).squeeze(-1)
        return pts + trans.view(B, N, 1, 1, 1, 3)

    def voxel_index(self, geom):
        """Flat BEV cell `ix * Y + iy` of each frustum point, -1 outside the grid.

        Same arithmetic as lss.py:246-257: truncation toward zero of (p - (bx - dx/2)) / dx, then the
        0 <= i < nx test on all three axes (the z axis has a single 6.4 m voxel)."""
        idx = ((geom - (self.bx - self.dx / 2.)) / self.dx).long()
        nx = self.nx
        ok = ((idx[..., 0] >= 0) & (idx[..., 0] < nx[0]) & (idx[..., 1] >= 0) & (idx[..., 1] < nx[1]) &
              (idx[..., 2] >= 0) & (idx[..., 2] < nx[2]))
        flat = idx[..., 0] * nx[1] + idx[..., 1]
        return torch.where(ok, flat, torch.full_like(flat, -1)).to(torch.int32).contiguous()

    def cached_voxel_index(self, rots, trans, intrins, post_rots, post_trans):
        """`voxel_index(get_geometry(...))` depends only on the calibration (lss.py:204-224,246-257), which is static
        per rig: it is computed once per distinct calibration and reused, so the (B,N,D,fH,fW,3) geometry tensor (45 MB
        at 16 scenes x 512^2) is not re-materialised every forward.  Look-up: (i) the same tensors (storage + version) as
        the previous call - no device work at all; (ii) otherwise by the calibration VALUES (one small device->host read
        of B*N*33 floats), which also catches callers that rebuild the tensors for every frame."""
        cal = (rots, trans, intrins, post_rots, post_trans)
        ident = tuple((t.data_ptr(), t._version, tuple(t.shape), str(t.device)) for t in cal)
        cache = self.__dict__.setdefault("_mfb_vox_cache", {"ident": None, "by_value": {}})
        if cache["ident"] is not None and cache["ident"][0] == ident:
            return cache["ident"][1]
        key = (str(rots.device), self.frustum.shape,
               torch.cat([t.detach().reshape(-1).float() for t in cal]).cpu().numpy().tobytes())
        vox = cache["by_value"].get(key)
        if vox is None:
            if len(cache["by_value"]) >= 8:
                cache["by_value"].clear()
            with torch.no_grad():
                vox = self.voxel_index(self.get_geometry(*cal))
            cache["by_value"][key] = vox
        cache["ident"] = (ident, vox)
        return vox

    def get_cam_feats(self, x):
        """Reference-compatible lifted tensor (B, N, D, fH, fW, C) - lss.py:226-236 (materialises it)."""
        B, N, Cin, H, W = x.shape
        x = self.camencode(x.view(B * N, Cin, H, W))
        x = x.view(B, N, self.camC, self.D, H // self.downsample, W // self.downsample)
        return x.permute(0, 1, 3, 4, 5, 2)

    def get_voxels(self, x, rots, trans, intrins, post_rots, post_trans):
        B, N, Cin, H, W = x.shape
        if int(self.nx[2]) != 1:
            raise NotImplementedError("the fused lift-splat kernel assumes a single z voxel (zbound of lss_cfg.yaml)")
        if not x.is_cuda:
            raise RuntimeError("monoforce_b200.LiftSplatShoot runs on CUDA only (fused lift-splat kernel, no CPU fallback)")
        vox = self.cached_voxel_index(rots, trans, intrins, post_rots, post_trans)
        logits = self.camencode.depth_logits_and_feats(x.view(B * N, Cin, H, W)).float().permute(0, 2, 3, 1)
        X, Y = int(self.nx[0]), int(self.nx[1])
        bev = _LiftSplat.apply(logits, vox.view(-1), B, N, self.D, self.camC, X, Y)
        return bev.permute(0, 3, 1, 2)        # (B, C, X, Y) view of channels-last storage

    def _fast(self):
        return self.fast_inference and not self.training and not torch.is_grad_enabled()

    def forward(self, x, rots, trans, intrins, post_rots, post_trans):
        if self._fast():
            # inference on repo kernels only, NHWC bf16 from the images to the heads (monoforce_b200/encoder_fast.py)
            if not x.is_cuda:
                raise RuntimeError("monoforce_b200.LiftSplatShoot runs on CUDA only (sm_100a kernels, no CPU fallback)")
            if int(self.nx[2]) != 1:
                raise NotImplementedError("the fused lift-splat kernel assumes a single z voxel (zbound of lss_cfg.yaml)")
            from . import encoder_fast
            vox = self.cached_voxel_index(rots, trans, intrins, post_rots, post_trans)
            if self.fast_graph:
                return self._graphed_forward(x, vox)
            return encoder_fast.forward(self, x, vox)
        bev = self.get_voxels(x, rots, trans, intrins, post_rots, post_trans)
        return self.bevencode(bev)

    def from_pretrained(self, modelf):
        """Partial-state-dict loading like lss.py:293-302."""
        if not modelf:
            return self
        print(f'Loading pretrained {self.__class__.__name__} model from', modelf)
        state = self.state_dict()
        state.update(torch.load(modelf, map_location='cpu'))
        self.load_state_dict(state)
        return self
