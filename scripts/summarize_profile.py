#!/usr/bin/env python
"""Turn gpurun_out/launches.csv (+ an optional .ncu-rep) into a tracked summary under profiles/."""
import collections
import csv
import subprocess
import sys

# usage: summarize_profile.py TAG [LAUNCHES_CSV] [NCU_REP]
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
rows = list(csv.reader(open(sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/launches.csv")))
hdr, agg = None, collections.OrderedDict()
for r in rows:
    if len(r) > 5 and r[0] == "ID":
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    val = float(d["Metric Value"].replace(",", ""))
    val = {"ns": val / 1e3, "us": val, "ms": val * 1e3, "s": val * 1e6}[d["Metric Unit"]]
    agg.setdefault(d["Kernel Name"][:110], []).append(val)
tot = sum(sum(v) for v in agg.values())
out = [f"# ncu launch list ({tag}): `ncu --metrics gpu__time_duration.sum --clock-control none` over the bench command",
       "", f"total device time {tot / 1e3:.2f} ms over {sum(len(v) for v in agg.values())} launches "
       "(cold-cache, serialised: compare SHARES, not absolutes)", "",
       "| share | launches | avg us | kernel |", "|---:|---:|---:|---|"]
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))[:30]:
    out.append(f"| {sum(v) / tot * 100:.1f}% | {len(v)} | {sum(v) / len(v):.1f} | `{k}` |")
open(f"profiles/{tag}_launches.md", "w").write("\n".join(out) + "\n")

rep = sys.argv[3] if len(sys.argv) > 3 else None
if rep:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h, u = rr[0], rr[1]
    keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "launch__block_size", "launch__grid_size",
            "launch__shared_mem_per_block_dynamic",
            "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
            "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
            "lts__t_sector_hit_rate.pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"]
    lines = [f"# ncu --set full capture ({tag}): {rep}", ""]
    for r in rr[2:]:
        d = dict(zip(h, r))
        for k in keys:
            if k in d:
                lines.append(f"- `{k}` [{u[h.index(k)]}] = {d[k]}")
        for k in h:
            if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and "not_issued" not in k and float(d[k]) > 0.05:
                lines.append(f"- stall `{k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}` per issue = {float(d[k]):.2f}")
        lines.append("")
    seg = subprocess.run([sys.executable, "scripts/ncu_segments.py", rep], capture_output=True, text=True).stdout
    lines += ["## executed warp instructions / stall samples per code segment (split at barriers, TMEM loads, MMA commits)", "", "```"]
    lines += [l for l in seg.splitlines() if l[:4].strip().isdigit() or l.startswith("total")] + ["```", ""]
    suffix = sys.argv[4] if len(sys.argv) > 4 else "edge_fwd_full"
    open(f"profiles/{tag}_{suffix}.md", "w").write("\n".join(lines) + "\n")
print("wrote profiles/")
