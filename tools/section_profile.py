#!/usr/bin/env python
"""Dynamic warp-instruction profile of a rollout kernel grouped by the OUTERMOST source line (the line of the
kernel body an inlined instruction was expanded from), summed over user-given line ranges.

    python tools/section_profile.py REP UNIT KERNEL_SUBSTR MANGLED_SUBSTR FILE 'name:lo-hi,name:lo-hi,...'
"""
import csv, io, os, re, subprocess, sys, collections, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, unit, kname, mangled, kfile, ranges = sys.argv[1:7]
steps = 4096 * 400
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks = re.split(r'(?m)^"Kernel Name",', src)
blk = next(b for b in blocks[1:] if kname in b.split("\n")[0])
rows = list(csv.reader(io.StringIO('"Kernel Name",' + blk)))
h = rows[1]; jx = {c: i for i, c in enumerate(h)}
data = [r for r in rows[2:] if len(r) == len(h)]
base = int(data[0][jx["Address"]], 16)
dyn = {int(r[jx["Address"]], 16) - base: (int(r[jx["Instructions Executed"]] or 0), int(r[jx["# Samples"]] or 0), r[jx["Source"]]) for r in data}
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "monoforce_b200", "build", unit + ".o")], cwd=tmp, capture_output=True)
cubin = os.path.join(tmp, os.listdir(tmp)[0])
dis = subprocess.run(["nvdisasm", "-gi", cubin], capture_output=True, text=True).stdout.split("\n")
sec = next(i for i, l in enumerate(dis) if l.startswith(".text.") and mangled in l)
end = next(k for k in range(sec + 1, len(dis)) if dis[k].startswith("//--------------------- ."))
outer = collections.Counter(); outer_s = collections.Counter(); ops = collections.defaultdict(collections.Counter)
cur = None
for l in dis[sec:end]:
    if "//## File" in l:
        fl = re.findall(r'"([^"]+)", line (\d+)', l)
        # outermost frame that lies in the kernel file
        cand = [(os.path.basename(f), int(n)) for f, n in fl if os.path.basename(f) == kfile]
        cur = cand[-1] if cand else (os.path.basename(fl[-1][0]), -int(fl[-1][1]))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
    if m and cur:
        off = int(m.group(1), 16)
        if off in dyn:
            outer[cur[1]] += dyn[off][0]; outer_s[cur[1]] += dyn[off][1]; ops[cur[1]][m.group(2)] += dyn[off][0]
tot = sum(outer.values()); tots = sum(outer_s.values()) or 1
print(f"total {tot / steps:.0f} instr/step")
for spec in ranges.split(","):
    name, r = spec.split(":"); lo, hi = map(int, r.split("-"))
    c = sum(v for k, v in outer.items() if lo <= k <= hi); s = sum(v for k, v in outer_s.items() if lo <= k <= hi)
    agg = collections.Counter()
    for k, v in ops.items():
        if lo <= k <= hi: agg.update(v)
    top = " ".join(f"{o}:{n / steps:.0f}" for o, n in agg.most_common(8))
    print(f"{name:>28s} {c / steps:8.1f} instr/step {c / tot:6.1%}  stall {s / tots:6.1%}   {top}")
if "-v" in sys.argv:
    for k, v in sorted(outer.items()):
        if v / steps >= 3: print(k, f"{v / steps:.1f}")
