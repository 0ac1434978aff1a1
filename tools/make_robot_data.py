#!/usr/bin/env python
"""Pre-compute the robot contact-point sets the rollout kernel consumes.

Runs ONLY in the build container: it reads the reference's robot meshes
(/root/reference/monoforce/config/meshes/{marv,tradr}.obj) and applies the geometry
recipe of the reference (dphys_config.py:8-74: 0.1 m voxel-mean down-sample of all
mesh vertices, xy extents as robot size, quadrant / side tests for the driving
parts), evaluated by importing the unmodified reference through oracle/shims.

Output: monoforce_b200/data/<robot>.npz with
    points      (N,3) float32   body-frame contact points
    masks       (P,N) bool      driving-part masks in the reference's order
    part_id     (N,)  int32     index of the LAST mask containing the point, -1 if none
                                (the reference assigns cmd_vels mask by mask, so the last
                                 mask wins: dphysics.py:243-246)
    robot_size  (2,)  float32   (Lx, Ly)
This removes open3d (absent here, and on the GPU box) from the product's runtime path.
"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_import import import_reference  # noqa: E402


def main():
    _, cfgm = import_reference()
    out_dir = os.path.join(ROOT, "monoforce_b200", "data")
    os.makedirs(out_dir, exist_ok=True)
    for robot in ("marv", "tradr"):
        pts, parts, size = cfgm.robot_geometry(robot)
        pts = pts.numpy().astype(np.float32)
        masks = np.stack([m.numpy() for m in parts]).astype(bool)
        part_id = np.full(pts.shape[0], -1, dtype=np.int32)
        for i, m in enumerate(masks):
            part_id[m] = i
        size = np.asarray([float(size[0]), float(size[1])], dtype=np.float32)
        path = os.path.join(out_dir, f"{robot}.npz")
        np.savez_compressed(path, points=pts, masks=masks, part_id=part_id, robot_size=size)
        print(robot, pts.shape, masks.sum(1), size, "->", path, os.path.getsize(path), "B")


if __name__ == "__main__":
    main()
