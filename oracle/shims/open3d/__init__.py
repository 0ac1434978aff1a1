"""Stand-in for the two open3d calls the reference makes while building DPhysConfig.

TEST INFRASTRUCTURE ONLY (see oracle/README.md).  The reference imports ``open3d``
at module top (dphys_config.py:5) and uses exactly two entry points
(dphys_config.py:26-30): ``o3d.io.read_triangle_mesh(path).vertices`` and
``o3d.geometry.PointCloud().voxel_down_sample(voxel_size)``.  open3d==0.13.0
(docker/requirements.txt:17) is not installable here (no network), so this module
restates the published algorithm of ``VoxelDownSample``:

    min_bound = min(points) - voxel_size / 2
    key(p)    = floor((p - min_bound) / voxel_size)        (double precision)
    output    = mean of the points sharing a key

Open3D returns the voxels in hash-map order, which is not reproducible; we return
them sorted by key.  The order only changes fp32 summation order downstream.
Pinned by: N = 223 contact points for marv (examples/diff_physics.ipynb:217).
"""
import numpy as np


def voxel_mean_downsample(pts: np.ndarray, voxel_size: float) -> np.ndarray:
    pts = np.asarray(pts, dtype=np.float64)
    lo = pts.min(axis=0) - voxel_size * 0.5
    keys = np.floor((pts - lo) / voxel_size).astype(np.int64)
    uniq, inv = np.unique(keys, axis=0, return_inverse=True)   # lexicographic key order
    inv = inv.reshape(-1)
    out = np.zeros((len(uniq), 3), dtype=np.float64)
    np.add.at(out, inv, pts)
    cnt = np.bincount(inv, minlength=len(uniq)).astype(np.float64)
    return out / cnt[:, None]


def read_obj_vertices(path: str) -> np.ndarray:
    vs = []
    with open(path, "r") as f:
        for line in f:
            if line.startswith("v "):
                _, x, y, z = line.split()[:4]
                vs.append((float(x), float(y), float(z)))
    return np.asarray(vs, dtype=np.float64)


class _Mesh:
    def __init__(self, vertices):
        self.vertices = vertices


class _IO:
    @staticmethod
    def read_triangle_mesh(path):
        return _Mesh(read_obj_vertices(path))


class _PointCloud:
    def __init__(self):
        self.points = None

    def voxel_down_sample(self, voxel_size):
        out = _PointCloud()
        out.points = voxel_mean_downsample(np.asarray(self.points), voxel_size)
        return out


class _Geometry:
    PointCloud = _PointCloud


io = _IO()
geometry = _Geometry()
