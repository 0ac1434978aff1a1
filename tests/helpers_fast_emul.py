"""Torch (CPU, fp32) emulation of the encoder's inference kernels, with the SAME call signatures as monoforce_b200.ops.

TEST INFRASTRUCTURE.  Monkey-patched over `monoforce_b200.encoder_fast.ops` it lets the CPU suite check all the host logic
of the fast path (BatchNorm folding, channel bookkeeping, squeeze-excite folding into per-image weights, padding / stride /
output-size conventions, residual placement, head fusion) against the module path, and on the GPU it is the fp32 statement
of what each kernel must compute."""
import torch
import torch.nn.functional as F

ACT_NONE, ACT_RELU, ACT_GELU, ACT_SILU = 0, 1, 2, 3
HEAD_NONE, HEAD_RELU, HEAD_SCALED_TANH = 0, 1, 2


def conv_out_size(n, k, stride, pad_lo, pad_hi):
    return (n + pad_lo + pad_hi - k) // stride + 1


def _act(v, act):
    return {ACT_NONE: lambda t: t, ACT_RELU: F.relu, ACT_GELU: F.gelu, ACT_SILU: F.silu}[act](v)


def _pad_for(n, k, stride, lo, out):
    """High-side zero padding (negative = crop) that makes a conv with low-side padding `lo` produce `out` samples."""
    return (out - 1) * stride + k - n - lo


def conv2d_nhwc(x, wgt, scale, shift, act, *, stride=1, pad=(0, 0), out_hw=None, residual=None, heads=None):
    x = x.float()
    N, H, W, Cin = x.shape
    Ho, Wo = out_hw if out_hw is not None else (-(-H // stride), -(-W // stride))
    KH, KW = wgt.shape[-3], wgt.shape[-2]
    xp = F.pad(x.permute(0, 3, 1, 2), (pad[1], max(_pad_for(W, KW, stride, pad[1], Wo), 0), pad[0], max(_pad_for(H, KH, stride, pad[0], Ho), 0)))
    if wgt.dim() == 5:
        y = torch.cat([F.conv2d(xp[i:i + 1], wgt[i].float().permute(0, 3, 1, 2), stride=stride) for i in range(N)])
    else:
        y = F.conv2d(xp, wgt.float().permute(0, 3, 1, 2), stride=stride)
    y = y[:, :, :Ho, :Wo].permute(0, 2, 3, 1) * scale + shift
    if heads is not None:
        head_w, bias, acts, lo, hi = heads
        v = _act(y, act)
        G = y.shape[-1] // len(bias)
        outs = []
        for g in range(len(bias)):
            o = (v[..., g * G:(g + 1) * G] * head_w[g * G:(g + 1) * G]).sum(-1) + bias[g]
            if acts[g] == HEAD_RELU:
                o = F.relu(o)
            elif acts[g] == HEAD_SCALED_TANH:
                o = lo[g] + (hi[g] - lo[g]) * (torch.tanh(o) + 1) / 2
            outs.append(o)
        return torch.stack(outs, 1)
    if residual is not None:
        y = y + residual.float()
    return _act(y, act)


def upsample_concat_nhwc(skip, low, out_hw, c_out):
    up = F.interpolate(low.float().permute(0, 3, 1, 2), size=out_hw, mode="bilinear", align_corners=True).permute(0, 2, 3, 1)
    parts = ([skip.float()] if skip is not None else []) + [up]
    y = torch.cat(parts, -1)
    if y.shape[-1] < c_out:
        y = F.pad(y, (0, c_out - y.shape[-1]))
    return y


def stem_conv(img, w, shift, pad):
    # the kernel contracts on the tensor cores: image and weights rounded to bf16, fp32 accumulate (like every other conv of the trunk)
    lo, hi = pad
    img, w = img.to(torch.bfloat16).float(), w.to(torch.bfloat16).float()
    y = F.conv2d(F.pad(img, (lo, hi, lo, hi)), w.permute(3, 2, 0, 1), stride=2)
    return F.silu(y.permute(0, 2, 3, 1) + shift)


def dwconv_bn_silu(x, w, shift, K, stride, pad, pool=None):
    lo, hi = pad
    C_ = x.shape[-1]
    xp = F.pad(x.float().permute(0, 3, 1, 2), (lo, hi, lo, hi))
    y = F.conv2d(xp, w.view(K, K, C_).permute(2, 0, 1).unsqueeze(1), stride=stride, groups=C_)
    y = F.silu(y.permute(0, 2, 3, 1) + shift)
    if pool is not None:
        pool += y.sum((1, 2))
    return y


def se_fold(pool, inv_hw, w_reduce, b_reduce, w_expand_t, b_expand, proj_w):
    m = pool * inv_hw
    r = F.silu(m @ w_reduce.t() + b_reduce)
    s = torch.sigmoid(r @ w_expand_t + b_expand)                        # (N, C); w_expand_t (Sq, C)
    return (proj_w.float().unsqueeze(0) * s.unsqueeze(1)).view(pool.shape[0], proj_w.shape[0], 1, 1, proj_w.shape[1])


def cast_bf16(x):
    return x


def lift_splat_bf16(logits, vox, B, N, D, Cc, X, Y):
    BN, fH, fW, rs = logits.shape
    lg = logits.float()
    depth = lg[..., :D].softmax(-1)                                      # (BN,fH,fW,D)
    feats = lg[..., D:D + Cc]
    bev = torch.zeros(B, X * Y, Cc)
    v = vox.view(B, N, D, fH, fW).long()
    for b in range(B):
        for n in range(N):
            contrib = depth[b * N + n].permute(2, 0, 1).unsqueeze(-1) * feats[b * N + n].unsqueeze(0)   # (D,fH,fW,C)
            idx = v[b, n].reshape(-1)
            ok = idx >= 0
            bev[b].index_add_(0, idx[ok], contrib.reshape(-1, Cc)[ok])
    return bev.view(B, X, Y, Cc)


def terrain_postproc(geom, diff, friction, pool=1, want_terrain=True):
    t = geom - diff
    return (t if want_terrain else None), F.avg_pool2d(t, pool), F.avg_pool2d(friction, pool)
