"""EfficientNet-B0 trunk with the parameter layout of efficientnet_pytorch==0.7.1.

The reference's camera encoder is built on `efficientnet_pytorch.EfficientNet`
(`terrain_encoder/lss.py:9,55,73-94`; pinned in `docker/requirements.txt:29`).  That package is
NOT part of /root/reference and is not installable here, so this is a restatement of its published
architecture with identical `state_dict` keys (`_conv_stem`, `_bn0`, `_blocks.N._expand_conv`,
`_depthwise_conv`, `_se_reduce`, `_se_expand`, `_project_conv`, `_bn0/1/2`, `_conv_head`, `_bn1`,
`_fc`) so that released monoforce checkpoints (`camencode.trunk.*`) load.  Parity of the trunk
internals is therefore UNPINNED; anchors: the B0 parameter count (5,288,548) and the block
table below.

Details that matter for numerics:
  * "static same padding" computed once for `image_size=224`: stride-1 convs pad symmetrically,
    stride-2 convs pad (0,1,0,1) for k=3 and (1,2,1,2) for k=5 (left,right,top,bottom);
  * BatchNorm eps 1e-3, momentum 0.01; swish activations; squeeze-excite ratio 0.25 of the block's
    INPUT filters; drop-connect (stochastic depth) only in training mode.
"""
from __future__ import annotations

import math
import warnings
from collections import namedtuple

import torch
from torch import nn
from torch.nn import functional as F

BlockArgs = namedtuple("BlockArgs", "num_repeat kernel_size stride expand_ratio input_filters output_filters se_ratio id_skip")
GlobalParams = namedtuple("GlobalParams", "batch_norm_momentum batch_norm_epsilon dropout_rate drop_connect_rate num_classes image_size")

# efficientnet-b0: r1_k3_s11_e1_i32_o16_se0.25, r2_k3_s22_e6_i16_o24_se0.25, r2_k5_s22_e6_i24_o40_se0.25,
# r3_k3_s22_e6_i40_o80_se0.25, r3_k5_s11_e6_i80_o112_se0.25, r4_k5_s22_e6_i112_o192_se0.25, r1_k3_s11_e6_i192_o320_se0.25
B0_BLOCKS = [
    BlockArgs(1, 3, 1, 1, 32, 16, 0.25, True),
    BlockArgs(2, 3, 2, 6, 16, 24, 0.25, True),
    BlockArgs(2, 5, 2, 6, 24, 40, 0.25, True),
    BlockArgs(3, 3, 2, 6, 40, 80, 0.25, True),
    BlockArgs(3, 5, 1, 6, 80, 112, 0.25, True),
    BlockArgs(4, 5, 2, 6, 112, 192, 0.25, True),
    BlockArgs(1, 3, 1, 6, 192, 320, 0.25, True),
]
B0_GLOBAL = GlobalParams(0.99, 1e-3, 0.2, 0.2, 1000, 224)


class Swish(nn.Module):
    def forward(self, x):
        return x * torch.sigmoid(x)


def _same_pad(image_size: int, k: int, s: int):
    """Total padding a TF-style 'SAME' conv needs at `image_size`, split (before, after)."""
    out = math.ceil(image_size / s)
    total = max((out - 1) * s + (k - 1) + 1 - image_size, 0)
    return total // 2, total - total // 2


class Conv2dStaticSamePadding(nn.Conv2d):
    """Conv2d whose zero padding is fixed at construction for a nominal input size."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, image_size=None, **kwargs):
        super().__init__(in_channels, out_channels, kernel_size, stride, **kwargs)
        k, s = self.kernel_size[0], self.stride[0]
        lo, hi = _same_pad(image_size, k, s)
        self.static_padding = nn.ZeroPad2d((lo, hi, lo, hi)) if (lo or hi) else nn.Identity()

    def forward(self, x):
        x = self.static_padding(x)
        return F.conv2d(x, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)


def drop_connect(x, p, training):
    if not training:
        return x
    keep = 1 - p
    mask = torch.floor(keep + torch.rand([x.shape[0], 1, 1, 1], dtype=x.dtype, device=x.device))
    return x / keep * mask


def params_stamp(module: nn.Module):
    """Cheap fingerprint of a module's parameters / buffers: storage addresses + in-place version counters."""
    return tuple((t.data_ptr(), t._version) for t in list(module.parameters()) + list(module.buffers()))


class MBConvBlock(nn.Module):
    def __init__(self, args: BlockArgs, gp: GlobalParams, image_size: int):
        super().__init__()
        self._block_args = args
        mom, eps = 1 - gp.batch_norm_momentum, gp.batch_norm_epsilon
        inp = args.input_filters
        oup = inp * args.expand_ratio
        if args.expand_ratio != 1:
            self._expand_conv = Conv2dStaticSamePadding(inp, oup, 1, image_size=image_size, bias=False)
            self._bn0 = nn.BatchNorm2d(oup, momentum=mom, eps=eps)
        self._depthwise_conv = Conv2dStaticSamePadding(oup, oup, args.kernel_size, stride=args.stride, groups=oup,
                                                       image_size=image_size, bias=False)
        self._bn1 = nn.BatchNorm2d(oup, momentum=mom, eps=eps)
        squeezed = max(1, int(inp * args.se_ratio))
        self._se_reduce = Conv2dStaticSamePadding(oup, squeezed, 1, image_size=1)
        self._se_expand = Conv2dStaticSamePadding(squeezed, oup, 1, image_size=1)
        self._project_conv = Conv2dStaticSamePadding(oup, args.output_filters, 1, image_size=1, bias=False)
        self._bn2 = nn.BatchNorm2d(args.output_filters, momentum=mom, eps=eps)
        self._swish = Swish()

    def forward(self, inputs, drop_connect_rate=None):
        a = self._block_args
        x = inputs
        if a.expand_ratio != 1:
            x = self._swish(self._bn0(self._expand_conv(x)))
        x = self._swish(self._bn1(self._depthwise_conv(x)))
        s = F.adaptive_avg_pool2d(x, 1)
        s = self._se_expand(self._swish(self._se_reduce(s)))
        x = torch.sigmoid(s) * x
        x = self._bn2(self._project_conv(x))
        if a.id_skip and a.stride == 1 and a.input_filters == a.output_filters:
            if drop_connect_rate:
                x = drop_connect(x, drop_connect_rate, self.training)
            x = x + inputs
        return x


class EfficientNet(nn.Module):
    """`EfficientNet.from_name('efficientnet-b0')` of efficientnet_pytorch, trunk + (unused) head."""

    def __init__(self, blocks_args=None, global_params=None, in_channels=3):
        super().__init__()
        self._blocks_args = list(B0_BLOCKS if blocks_args is None else blocks_args)
        self._global_params = gp = B0_GLOBAL if global_params is None else global_params
        mom, eps = 1 - gp.batch_norm_momentum, gp.batch_norm_epsilon
        size = gp.image_size
        self._conv_stem = Conv2dStaticSamePadding(in_channels, 32, 3, stride=2, image_size=size, bias=False)
        self._bn0 = nn.BatchNorm2d(32, momentum=mom, eps=eps)
        size = math.ceil(size / 2)
        self._blocks = nn.ModuleList()
        for args in self._blocks_args:
            self._blocks.append(MBConvBlock(args, gp, size))
            size = math.ceil(size / args.stride)
            rest = args._replace(input_filters=args.output_filters, stride=1)
            for _ in range(args.num_repeat - 1):
                self._blocks.append(MBConvBlock(rest, gp, size))
        out = self._blocks_args[-1].output_filters
        self._conv_head = Conv2dStaticSamePadding(out, 1280, 1, image_size=size, bias=False)
        self._bn1 = nn.BatchNorm2d(1280, momentum=mom, eps=eps)
        self._avg_pooling = nn.AdaptiveAvgPool2d(1)
        self._dropout = nn.Dropout(gp.dropout_rate)
        self._fc = nn.Linear(1280, gp.num_classes)
        self._swish = Swish()

    @classmethod
    def from_name(cls, model_name="efficientnet-b0", in_channels=3, **_):
        assert model_name == "efficientnet-b0", "only the B0 trunk used by the reference is restated"
        return cls(in_channels=in_channels)

    @classmethod
    def from_pretrained(cls, model_name="efficientnet-b0", in_channels=3, weights_path=None, **_):
        """The package downloads ImageNet weights here; there is no network in this environment, so
        weights come from `weights_path` (a state_dict file) or stay randomly initialised."""
        model = cls.from_name(model_name, in_channels=in_channels)
        if weights_path:
            state = torch.load(weights_path, map_location="cpu")
            if in_channels != 3:       # efficientnet_pytorch re-creates the stem for other channel counts
                state = {k: v for k, v in state.items() if not k.startswith("_conv_stem.")}
            missing, unexpected = model.load_state_dict(state, strict=False)
            if unexpected or [k for k in missing if not k.startswith("_conv_stem.")]:
                raise RuntimeError(f"{weights_path} is not an efficientnet_pytorch B0 state_dict: missing {missing}, "
                                   f"unexpected {unexpected}")
        else:
            warnings.warn(
                "EfficientNet.from_pretrained: no ImageNet weights file given (pass trunk_weights=... to LiftSplatShoot / "
                "CamEncode or set MFB_EFFICIENTNET_B0_WEIGHTS to an efficientnet_pytorch 'efficientnet-b0' state_dict); "
                "the trunk is RANDOMLY initialised, unlike the reference (lss.py:55). Loading a full monoforce checkpoint "
                "afterwards (LiftSplatShoot.from_pretrained) overrides this.", RuntimeWarning, stacklevel=2)
        return model

    def extract_features(self, x):
        x = self._swish(self._bn0(self._conv_stem(x)))
        for idx, block in enumerate(self._blocks):
            rate = self._global_params.drop_connect_rate
            if rate:
                rate *= float(idx) / len(self._blocks)
            x = block(x, drop_connect_rate=rate)
        return self._swish(self._bn1(self._conv_head(x)))

    def forward(self, x):
        x = self._avg_pooling(self.extract_features(x)).flatten(1)
        return self._fc(self._dropout(x))
