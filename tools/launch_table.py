"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` log per kernel, in launch order for one forward:
    python tools/launch_table.py gpurun_out/enc_launches.csv [skip_first_n]"""
import csv, io, re, sys
txt = open(sys.argv[1]).read()
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = [r for r in csv.DictReader(io.StringIO(txt[txt.index('"ID"'):])) if r.get("Metric Name") == "gpu__time_duration.sum"]
rows = rows[skip:]
tot = 0.0
agg = {}
for i, r in enumerate(rows):
    v = float(r["Metric Value"].replace(",", ""))
    us = v * {"ns": 1e-3, "us": 1, "ms": 1e3, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3}.get(r["Metric Unit"], 1)
    k = re.sub(r"<.*|\(.*", "", r["Kernel Name"])[-60:]
    grid = r.get("Grid Size", "")
    print(f"{i:4d} {us:9.1f} us  {k:60s} grid {grid}")
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += us
    tot += us
print("---- total %.1f us" % tot)
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{us:9.1f} us {us / tot:6.1%} x{n:3d}  {k}")
