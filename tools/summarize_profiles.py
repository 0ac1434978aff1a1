#!/usr/bin/env python
"""Turn gpurun_out/{prof_rollout.ncu-rep, launches.csv, bench.json} into the tracked summaries under profiles/.

    python tools/summarize_profiles.py r01
"""
import csv
import io
import json
import os
import re
import subprocess
import sys
from collections import Counter, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_sectors.sum",
    "smsp__inst_executed_op_global_red.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.max",
]


def ncu(*args):
    return subprocess.run(["ncu", *args], capture_output=True, text=True).stdout


def main():
    os.makedirs(PROF, exist_ok=True)
    rep = os.path.join(OUT, "prof_rollout.ncu-rep")
    raw = list(csv.reader(io.StringIO(ncu("-i", rep, "--page", "raw", "--csv"))))
    hdr, units = raw[0], raw[1]
    ix = {h: i for i, h in enumerate(hdr)}
    lines = [f"# ncu --set full summary ({tag}), BASELINE config 3 size: 4096 trajectories x 400 steps, marv (223 points), 256x256 shared map",
             "", "Command: `ncu --set full --clock-control none --import-source on -k regex:rollout_ -s 4 -c 2 python tools/profile_target.py 4096 3`",
             "(cold-cache, serialised launches: durations are for the kernel alone, not bench values)", ""]
    traffic = {}
    for r in raw[2:]:
        name = r[ix["Kernel Name"]]
        lines += [f"## {name}", "", "| metric | value | unit |", "|---|---|---|"]
        for k in KEYS:
            if k in ix and r[ix[k]] != "":
                lines.append(f"| {k} | {r[ix[k]]} | {units[ix[k]]} |")
        stalls = []
        for h in hdr:
            if "issue_stalled" in h and "per_issue_active" in h and r[ix[h]]:
                v = float(r[ix[h]])
                if v >= 0.1:
                    stalls.append((v, h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
        lines += ["", "warp stall reasons (cycles per issued instruction): " +
                  ", ".join(f"{n} {v:.2f}" for v, n in sorted(stalls, reverse=True)), ""]
        if "rollout_fwd" in name:
            def tobytes(key):
                v, u = float(r[ix[key]]), units[ix[key]].lower()
                return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
            traffic = {"kernel": name, "dram_bytes_read": tobytes("dram__bytes_read.sum"),
                       "dram_bytes_write": tobytes("dram__bytes_write.sum")}
            traffic["dram_bytes_per_launch"] = traffic["dram_bytes_read"] + traffic["dram_bytes_write"]
            traffic["algorithmic_bytes_per_launch"] = 4096 * 400 * 5432
    # per-opcode dynamic instruction mix from the source page
    src = ncu("-i", rep, "--page", "source", "--csv")
    blocks = re.split(r'(?m)^"Kernel Name",', src)
    for blk in blocks[1:]:
        rows = list(csv.reader(io.StringIO('"Kernel Name",' + blk)))
        kname = rows[0][1]
        h = rows[1]
        jx = {c: i for i, c in enumerate(h)}
        ops, tot = Counter(), 0
        for r in rows[2:]:
            if len(r) != len(h):
                continue
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[jx["Source"]])
            n = int(r[jx["Instructions Executed"]] or 0)
            ops[m.group(2) if m else "?"] += n
            tot += n
        lines += [f"### dynamic SASS mix: {kname}", "",
                  f"{tot} warp instructions = {tot / (4096 * 400):.0f} per trajectory-step; top opcodes per trajectory-step: " +
                  ", ".join(f"{o} {c / (4096 * 400):.0f}" for o, c in ops.most_common(18)), ""]
    open(os.path.join(PROF, f"{tag}_ncu_rollout_summary.md"), "w").write("\n".join(lines) + "\n")
    if traffic:
        json.dump(traffic, open(os.path.join(PROF, "fwd_traffic.json"), "w"), indent=1)

    # launch list of the bench command: share of each kernel in the step
    lp = os.path.join(OUT, "launches.csv")
    if os.path.exists(lp):
        txt = open(lp).read()
        start = txt.index('"ID"')
        rows = list(csv.DictReader(io.StringIO(txt[start:])))
        agg = defaultdict(lambda: [0, 0.0])
        for r in rows:
            if r.get("Metric Name") != "gpu__time_duration.sum":
                continue
            v = float(r["Metric Value"].replace(",", ""))
            u = r["Metric Unit"]
            ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6, "second": 1e9}.get(u, 1)
            k = re.sub(r"\(.*", "", r["Kernel Name"])[:90]
            agg[k][0] += 1
            agg[k][1] += ns
        total = sum(v[1] for v in agg.values())
        out = [f"# ncu launch list ({tag}): `ncu --metrics gpu__time_duration.sum --clock-control none -c 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline`",
               "", "First 200 launches of the bench process (cold-cache, serialised: compare SHARES, not absolutes).", "",
               "| kernel | launches | total ms | share |", "|---|---|---|---|"]
        for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            out.append(f"| `{k}` | {n} | {ns / 1e6:.3f} | {ns / total:.1%} |")
        open(os.path.join(PROF, f"{tag}_ncu_launch_list.md"), "w").write("\n".join(out) + "\n")
    bp = os.path.join(OUT, "bench.json")
    if os.path.exists(bp) and os.path.getsize(bp):
        open(os.path.join(PROF, f"{tag}_bench_line.json"), "w").write(open(bp).read())
    print("wrote summaries to", PROF)


if __name__ == "__main__":
    main()
