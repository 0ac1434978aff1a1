"""Brief of one .ncu-rep: key raw metrics, stall reasons, hottest source lines.   python tools/ncu_brief.py rep [n_lines]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
nl = int(sys.argv[2]) if len(sys.argv) > 2 else 14
raw = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
h, u = raw[0], raw[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.sum", "smsp__inst_executed_pipe_xu.sum"]
for r in raw[2:]:
    print("==", r[h.index("Kernel Name")][:100])
    for k in KEYS:
        if k in h:
            print(f"  {k:70s} {r[h.index(k)]:>16s} {u[h.index(k)]}")
    st = []
    for i, k in enumerate(h):
        if "issue_stalled" in k and "per_issue_active" in k and r[i]:
            st.append((float(r[i]), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
    print("  stalls/issue:", ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = None
for i, r in enumerate(rows):
    if "Source" in r and any("Sampl" in c for c in r):
        hdr = i
        break
if hdr is not None:
    H = rows[hdr]
    si = H.index("Source")
    samp = [j for j, c in enumerate(H) if c.startswith("# Samples") or c == "Warp Stall Sampling (All Samples)" or "Sampling (All" in c]
    ie = [j for j, c in enumerate(H) if c == "Instructions Executed"]
    if samp:
        sc = samp[0]
        body = [r for r in rows[hdr + 1:] if len(r) == len(H) and r[sc].replace(",", "").replace(".", "").isdigit()]
        tot = sum(float(r[sc].replace(",", "")) for r in body) or 1
        print(f"-- hottest lines by {H[sc]} (total {tot:.0f})")
        for r in sorted(body, key=lambda r: -float(r[sc].replace(",", "")))[:nl]:
            print(f"  {float(r[sc].replace(',', '')) / tot:6.1%}  {r[ie[0]] if ie else '':>12s}  {r[si][:130]}")
