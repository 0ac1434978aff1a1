#!/usr/bin/env python
"""gpurun_out/{prof_heads.ncu-rep, enc_launches_<cfg>.csv} -> profiles/<tag>_ncu_encoder_summary.md, <tag>_encoder_launch_list_<cfg>.md

    python tools/summarize_encoder_profiles.py r02
"""
import csv, io, os, re, subprocess, sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.max"]


def ncu(*a):
    return subprocess.run(["ncu", *a], capture_output=True, text=True).stdout


def full_capture(rep, title, flops=None):
    raw = list(csv.reader(io.StringIO(ncu("-i", rep, "--page", "raw", "--csv"))))
    h, u = raw[0], raw[1]
    lines = [f"## {title}", ""]
    for r in raw[2:]:
        lines += [f"`{r[h.index('Kernel Name')]}`", "", "| metric | value | unit |", "|---|---|---|"]
        for k in KEYS:
            if k in h and r[h.index(k)] != "":
                lines.append(f"| {k} | {r[h.index(k)]} | {u[h.index(k)]} |")
        if flops:
            t = float(r[h.index("gpu__time_duration.sum")])
            unit = u[h.index("gpu__time_duration.sum")]
            ms = t * {"ms": 1, "us": 1e-3, "ns": 1e-6, "msecond": 1, "usecond": 1e-3, "nsecond": 1e-6}.get(unit, 1)
            lines.append(f"| algorithmic FLOP / duration (cold-cache, under ncu) | {flops / ms / 1e9:.0f} | TFLOP/s |")
        st = []
        for i, k in enumerate(h):
            if "issue_stalled" in k and "per_issue_active" in k and r[i] and float(r[i]) >= 0.1:
                st.append((float(r[i]), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
        lines += ["", "warp stall reasons (cycles per issued instruction): " + ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)), ""]
    return lines


def launch_list(path, title):
    txt = open(path).read()
    rows = [r for r in csv.DictReader(io.StringIO(txt[txt.index('"ID"'):])) if r.get("Metric Name") == "gpu__time_duration.sum"]
    # keep the LAST forward: from the last stem_conv launch on
    names = [r["Kernel Name"] for r in rows]
    start = max(i for i, n in enumerate(names) if "stem_conv" in n)
    rows = rows[start:]
    agg, tot = OrderedDict(), 0.0
    for r in rows:
        v = float(r["Metric Value"].replace(",", ""))
        us = v * {"ns": 1e-3, "us": 1, "ms": 1e3, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3}.get(r["Metric Unit"], 1)
        k = re.sub(r"\(.*", "", r["Kernel Name"])[:100]
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += us
        tot += us
    repo = sum(us for k, (n, us) in agg.items() if any(s in k for s in ("conv_bn_act", "enc::", "mfb::", "lift_splat")))
    lines = [f"## {title}", "", f"One forward = {len(rows)} launches, {tot / 1e3:.2f} ms summed (cold-cache, serialised under ncu: compare SHARES). "
             f"**Repo kernels: {repo / tot:.1%} of the time**, the rest are framework fills / the final subtraction.", "",
             "| kernel | launches | total us | share |", "|---|---|---|---|"]
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{k}` | {n} | {us:.1f} | {us / tot:.1%} |")
    return lines + [""]


def main():
    out = [f"# ncu evidence for the terrain encoder ({tag}), inference path on repo kernels", "",
           "Commands (under gpurun, one GPU): `ncu --set full --clock-control none --import-source on -k regex:conv_bn_act_kernel -s 103 -c 1 "
           "python tools/encoder_once.py --cfg4 1` (the last K4 launch of the second forward = the fused BEV heads) and "
           "`ncu --metrics gpu__time_duration.sum --clock-control none --csv python tools/encoder_once.py [--cfg4] 1`.", ""]
    rep = os.path.join(OUT, "prof_heads.ncu-rep")
    if os.path.exists(rep):
        out += full_capture(rep, "K4, fused BEV heads at BASELINE config 4: 16 x 256 x 256 x 256 -> 384 channels, 3x3, GELU, 1x1 heads in the epilogue "
                                 "(1.855 TFLOP)", flops=2.0 * 16 * 256 * 256 * 256 * 384 * 9)
    for cfg, title in (("cfg4", "Launch list, BASELINE config 4 (16 scenes x 4 cameras 512x512 -> 256x256 BEV)"),
                       ("default", "Launch list, lss_cfg.yaml (16 scenes x 4 cameras 256x416 -> 128x128 BEV)")):
        p = os.path.join(OUT, f"enc_launches_{cfg}.csv")
        if os.path.exists(p):
            out += launch_list(p, title)
    open(os.path.join(PROF, f"{tag}_ncu_encoder_summary.md"), "w").write("\n".join(out) + "\n")
    print("wrote", os.path.join(PROF, f"{tag}_ncu_encoder_summary.md"))


if __name__ == "__main__":
    main()
