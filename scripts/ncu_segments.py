#!/usr/bin/env python
"""Per-phase executed-instruction / stall-sample breakdown of one kernel from an ncu report.

  python scripts/ncu_segments.py gpurun_out/<report>.ncu-rep

Splits the SASS stream at every barrier / TMEM load / MMA commit and prints, per segment, the executed
warp instructions, the share of stall samples and the top opcodes; then the headline counters."""
import csv
import subprocess
import sys

rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]
ia, ie, iss = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
seg = acc = sacc = tot = stot = 0
segs, ops = [], {}
for r in rows[2:]:
    s_, n, s = r[ia].strip(), int(r[ie]), int(r[iss])
    tot += n; stot += s; acc += n; sacc += s
    op = (s_.split()[1] if s_.startswith("@") else s_.split()[0]).split(".")[0]
    d = ops.setdefault(seg, {})
    d[op] = d.get(op, 0) + n
    if "BAR.SYNC" in s_ or "LDTM" in s_ or "UTCBAR" in s_:
        segs.append((seg, acc, sacc, s_[:40])); seg += 1; acc = sacc = 0
segs.append((seg, acc, sacc, "end"))
print(f"total warp instructions {tot / 1e6:.1f} M, stall samples {stot}")
for sg, a, s, t in segs:
    top = sorted(ops.get(sg, {}).items(), key=lambda kv: -kv[1])[:7]
    print(f"{sg:3d} inst={a / 1e6:7.2f}M ({a / tot * 100:4.1f}%) samples={s / max(stot, 1) * 100:4.1f}%  {t:40s}", " ".join(f"{k}:{v / 1e6:.1f}" for k, v in top))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
d = dict(zip(rr[0], rr[2]))
for k in rr[0]:
    if k in ("gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
             "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
             "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_blocks", "sm__maximum_warps_per_active_cycle_pct", "launch__waves_per_multiprocessor"):
        print(k, "=", d[k])
    if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and "not_issued" not in k:
        v = float(d[k])
        if v > 0.05:
            print("  stall", k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), f"{v:.2f}")
