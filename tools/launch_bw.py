"""Per-launch duration and DRAM bandwidth from an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
--csv` log (last forward only: from the last stem_conv launch on):   python tools/launch_bw.py gpurun_out/enc_launches_bw_cfg4.csv"""
import csv, io, re, sys
txt = open(sys.argv[1]).read()
rows = list(csv.DictReader(io.StringIO(txt[txt.index('"ID"'):])))
U = {"ns": 1e-3, "us": 1, "ms": 1e3, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3}
BU = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
by = {}
for r in rows:
    d = by.setdefault(int(r["ID"]), {"name": re.sub(r"<.*|\(.*", "", r["Kernel Name"])[-40:], "grid": r.get("Grid Size", "")})
    v = float(r["Metric Value"].replace(",", ""))
    if r["Metric Name"] == "gpu__time_duration.sum":
        d["us"] = v * U.get(r["Metric Unit"], 1)
    else:
        d[r["Metric Name"]] = v * BU.get(r["Metric Unit"], 1)
ls = [by[k] for k in sorted(by)]
last = max(i for i, d in enumerate(ls) if "stem_conv" in d["name"])
ls = ls[last:]
tot = sum(d["us"] for d in ls)
print(f"{len(ls)} launches, {tot:.1f} us (cold-cache, serialised)")
for i, d in enumerate(ls):
    b = d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
    print(f"{i:3d} {d['us']:8.1f} us {b / 1e6:8.1f} MB {b / d['us'] / 1e3:7.0f} GB/s  {d['name']:40s} {d['grid']}")
