"""CPU oracle for the lift + splat stage of the terrain encoder  --  TEST INFRASTRUCTURE, NOT PRODUCT.

Restates, with plain torch ops on the CPU, what the reference computes between the `depthnet` logits and
the BEV feature grid:
    CamEncode.get_depth_feat      terrain_encoder/lss.py:63-71    depth soft-max (x) camera features
    LiftSplatShoot.voxel_pooling  terrain_encoder/lss.py:238-280  truncating voxel index, in-grid filter,
                                  sort by rank + cumsum-trick segment sum (utils.py:144-181), scatter
A sum over all kept points of one voxel is what sort + cumsum-difference produces, so the restatement is
an index_add.  Parity status: PINNED against the unmodified reference (tests/test_encoder_oracle.py runs
both on the same inputs when /root/reference is present; tests/golden/lss_*.npz were minted from it).
"""
import torch


def lift_splat(logits, geom, dx, bx, nx, D, C):
    """logits (B*N, D+C, fH, fW); geom (B, N, D, fH, fW, 3) ego-frame points -> (B, C*Z, X, Y)."""
    B, N = geom.shape[:2]
    BN, _, fH, fW = logits.shape
    depth = logits[:, :D].softmax(dim=1)                                      # lss.py:60-61,68
    lifted = depth.unsqueeze(1) * logits[:, D:D + C].unsqueeze(2)              # lss.py:69  (BN, C, D, fH, fW)
    x = lifted.view(B, N, C, D, fH, fW).permute(0, 1, 3, 4, 5, 2).reshape(-1, C)   # lss.py:233-234, :243
    idx = ((geom - (bx - dx / 2.)) / dx).long().view(-1, 3)                    # lss.py:246-247
    batch = torch.arange(B).repeat_interleave(idx.shape[0] // B)
    kept = ((idx[:, 0] >= 0) & (idx[:, 0] < nx[0]) & (idx[:, 1] >= 0) & (idx[:, 1] < nx[1]) &
            (idx[:, 2] >= 0) & (idx[:, 2] < nx[2]))                            # lss.py:253-255
    X, Y, Z = int(nx[0]), int(nx[1]), int(nx[2])
    flat = ((batch * Z + idx[:, 2]) * X + idx[:, 0]) * Y + idx[:, 1]
    out = torch.zeros(B * Z * X * Y, C, dtype=x.dtype)
    out.index_add_(0, flat[kept], x[kept])                                     # == sort + cumsum segment sums, :260-275
    out = out.view(B, Z, X, Y, C).permute(0, 4, 1, 2, 3)                       # (B, C, Z, X, Y)
    return torch.cat(out.unbind(dim=2), 1)                                     # collapse Z, lss.py:278
