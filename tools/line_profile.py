#!/usr/bin/env python
"""Join an ncu source-page export with nvdisasm line info: dynamic warp instructions per CUDA source line.

    python tools/line_profile.py gpurun_out/prof_rollout2.ncu-rep rollout_bwd_float_v0 rollout_bwd_kernel 20
"""
import csv, io, os, re, subprocess, sys, collections, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, unit, kname, topn = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 25
steps = 4096 * 400

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks = re.split(r'(?m)^"Kernel Name",', src)
blk = next(b for b in blocks[1:] if kname in b.split("\n")[0])
rows = list(csv.reader(io.StringIO('"Kernel Name",' + blk)))
full_name = rows[0][1]
h = rows[1]; jx = {c: i for i, c in enumerate(h)}
data = [r for r in rows[2:] if len(r) == len(h)]
base = int(data[0][jx["Address"]], 16)
dyn = {int(r[jx["Address"]], 16) - base: (int(r[jx["Instructions Executed"]] or 0), int(r[jx["# Samples"]] or 0), r[jx["Source"]]) for r in data}

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "monoforce_b200", "build", unit + ".o")], cwd=tmp, capture_output=True)
cubin = os.path.join(tmp, os.listdir(tmp)[0])
dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.split("\n")
# mangled name of the profiled kernel: match template args from the demangled ncu name
want = re.sub(r"\s", "", full_name)
secs = [i for i, l in enumerate(dis) if l.startswith(".text._ZN3mfb") and kname in l]
def demangle(m):
    return subprocess.run(["c++filt", m], capture_output=True, text=True).stdout.strip()
sec = None
for i in secs:
    m = dis[i].split(":")[0][len(".text."):]
    d = re.sub(r"\s", "", demangle(m))
    key = re.sub(r"\(int\)|\(bool\)|mfb::", "", want.split("(mfb::")[0].split("(RolloutArgs")[0])
    if re.sub(r"\(int\)|\(bool\)|mfb::", "", d.split("(mfb::")[0]) .replace("true", "1").replace("false", "0") == key.replace("void", "void"):
        sec = i
if sec is None:
    # fall back: pick the section with the same instruction count
    for i in secs:
        end = next(k for k in range(i + 1, len(dis)) if dis[k].startswith("//--------------------- ."))
        n = sum(1 for l in dis[i:end] if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l))
        if n == len(data):
            sec = i
assert sec is not None, "kernel section not found"
end = next(k for k in range(sec + 1, len(dis)) if dis[k].startswith("//--------------------- ."))
cur = ("?", 0)
per_line = collections.Counter(); per_line_samples = collections.Counter()
for l in dis[sec:end]:
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", l)
    if m:
        off = int(m.group(1), 16)
        if off in dyn:
            per_line[cur] += dyn[off][0]; per_line_samples[cur] += dyn[off][1]
tot = sum(per_line.values()); tots = sum(per_line_samples.values())
print(f"{full_name[:100]}\n total {tot / steps:.0f} instr/step, {tots} samples")
cache = {}
def text(f, n):
    p = os.path.join(ROOT, "monoforce_b200", "csrc", f)
    if p not in cache:
        cache[p] = open(p).read().split("\n") if os.path.exists(p) else []
    return cache[p][n - 1].strip()[:95] if 0 < n <= len(cache[p]) else ""
for (f, n), c in per_line.most_common(topn):
    print(f"{c / steps:7.1f} instr/step {per_line_samples[(f, n)] / tots:6.1%} stall  {f}:{n:<4d} {text(f, n)}")
