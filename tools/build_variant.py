#!/usr/bin/env python
"""Development helper: link an alternative libmonoforce_b200 with some translation units recompiled under extra
-D flags (kernel tuning experiments; load it with MFB_LIB_PATH).

    python tools/build_variant.py NAME UNIT[,UNIT...] -DMFB_SWEEP_MINB=4 [...]
"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from monoforce_b200 import build as B

name, units, flags = sys.argv[1], sys.argv[2].split(","), sys.argv[3:]
CSRC = os.environ.get("MFB_CSRC", B.CSRC)      # alternative source tree (e.g. a checkout of an older commit)
B.build(verbose=False)
out_dir = os.path.join(ROOT, "tools", "scratch", "variants")
os.makedirs(out_dir, exist_ok=True)
objs = []
for uname, src, defs in B._units():
    o = os.path.join(B.OBJ, uname + ".o")
    if uname in units:
        o = os.path.join(out_dir, f"{name}_{uname}.o")
        common = [c if c != B.CSRC else CSRC for c in B.COMMON]
        cmd = [B._nvcc(), *B.ARCH, *common, *defs, *flags, "-Xptxas", "-v", "-c", os.path.join(CSRC, src), "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.exit(r.stderr)
        lines = r.stderr.split("\n")
        for i, l in enumerate(lines):
            if "Compiling entry" in l and ("sweep" in l or os.environ.get("VERBOSE_ALL")):
                print(l.split("'")[1][:70], "|", lines[i + 1].strip(), "|", lines[i + 2].strip()[:40])
    objs.append(o)
lib = os.path.join(out_dir, f"libmfb_{name}.so")
subprocess.check_call([B._nvcc(), *B.ARCH, "-shared", "-o", lib, *objs, "-lcudart"])
print("built", lib)
