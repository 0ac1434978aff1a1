#!/usr/bin/env python
"""Mint the drop-in goldens: run tests/dropin_bodies.py against the UNMODIFIED reference on the CPU (build container only).

    python tests/golden/make_golden_dropin.py      # writes tests/golden/dropin_{fit_terrain,predict_states}.npz
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/monoforce/src"
SHIMS = os.path.join(ROOT, "oracle", "shims")

for which in ("fit_terrain", "predict_states"):
    out = os.path.join(HERE, f"dropin_{which}.npz")
    env = {**os.environ, "PYTHONPATH": os.pathsep.join([REF, SHIMS])}
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dropin_bodies.py"), which, out, "cpu"], env=env,
                       capture_output=True, text=True, cwd="/tmp")
    print(which, r.stdout.strip()[-200:], r.stderr.strip()[-2000:])
    assert r.returncode == 0 and "/root/reference/" in r.stdout
    print(out, os.path.getsize(out), "B")
