"""Build recipe of libmonoforce_b200.so: plain nvcc, sm_100a only, in-tree output.

    python -m monoforce_b200.build [--force]

Every (kernel, scalar type, integrator variant) translation unit is compiled in its own
nvcc process (they are independent), then linked into one shared library next to this
package.  nvcc cross-compiles without a GPU; the .so travels to the GPU box with the repo
snapshot.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "build")
LIB = os.path.join(PKG, "libmonoforce_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-I", CSRC]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the CUDA toolkit is required to build monoforce_b200")
    return exe


def _units():
    units = [("c_api", "c_api.cu", []), ("lift_splat", "lift_splat.cu", []),
             ("conv_tcgen05", "conv_tcgen05.cu", []), ("physics_loss", "physics_loss.cu", []),
             ("encoder_ops", "encoder_ops.cu", []), ("dwconv_tma", "dwconv_tma.cu", [])]
    for kern in ("rollout_fwd", "rollout_bwd"):
        for tname in ("float", "double"):
            for variant in (0, 1):
                units.append((f"{kern}_{tname}_v{variant}", f"{kern}.cu",
                              [f"-DMFB_INST_T={tname}", f"-DMFB_INST_VARIANT={variant}"]))
    return units


def _source_digest() -> str:
    h = hashlib.sha256()
    for name in sorted(os.listdir(CSRC)):
        if name.endswith((".cu", ".cuh", ".h")):
            with open(os.path.join(CSRC, name), "rb") as f:
                h.update(name.encode()); h.update(f.read())
    with open(os.path.join(os.path.dirname(PKG), "include", "monoforce_b200.h"), "rb") as f:
        h.update(f.read())
    h.update(" ".join(ARCH + COMMON).encode())
    return h.hexdigest()


def _compile(unit):
    name, src, defs = unit
    out = os.path.join(OBJ, name + ".o")
    cmd = [_nvcc(), *ARCH, *COMMON, *defs, "-c", os.path.join(CSRC, src), "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {name}:\n{' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
    return out


def build(force: bool = False, verbose: bool = True) -> str:
    os.makedirs(OBJ, exist_ok=True)
    stamp = os.path.join(OBJ, "digest.txt")
    digest = _source_digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        if verbose:
            print(f"[monoforce_b200] {LIB} is up to date")
        return LIB
    units = _units()
    if verbose:
        print(f"[monoforce_b200] compiling {len(units)} translation units for sm_100a ...", flush=True)
    with cf.ThreadPoolExecutor(max_workers=min(len(units), os.cpu_count() or 4)) as ex:
        objs = list(ex.map(_compile, units))
    cmd = [_nvcc(), *ARCH, "-shared", "-o", LIB, *objs, "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(digest)
    if verbose:
        print(f"[monoforce_b200] built {LIB} ({os.path.getsize(LIB) / 1e6:.1f} MB)")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
