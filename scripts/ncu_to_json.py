#!/usr/bin/env python
"""profiles/ncu_edge_fwd.json from `ncu --set full` reports of the edge-forward kernel (one per arithmetic mode):

  python scripts/ncu_to_json.py bf16x3=gpurun_out/prof_edge_ws_bf16x3.ncu-rep bf16=gpurun_out/prof_edge_ws_bf16.ncu-rep

bench.py reads the DRAM bytes per launch (roofline.traffic) and the pipe utilisation from this file; it records the
commit the capture was taken at, so a stale record is visible instead of silently wrong."""
import csv
import datetime
import json
import subprocess
import sys

KEYS = {"dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write", "gpu__time_duration.sum": "duration",
        "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "mufu_pipe_pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
        "smsp__inst_executed.sum": "warp_instructions", "launch__registers_per_thread": "registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
        # the shared-memory / L1 data pipe (128 B / clk / SM): LSU wavefronts (shared + global) and the tensor core's operand reads
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed": "l1_pipe_lsu_pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "l1_pipe_lsu_shared_pct",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "l1_pipe_tensor_operand_pct"}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def read(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    out = {"kernel": vals[hdr.index("Kernel Name")]}
    for k, name in KEYS.items():
        if k in hdr:
            i = hdr.index(k)
            out[name] = float(vals[i].replace(",", "")) * SCALE.get(units[i], 1.0)
    out["dram_bytes"] = out.pop("dram_read", 0.0) + out.pop("dram_write", 0.0)
    out["duration_us"] = out.pop("duration", None)
    return out


def main():
    commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    rec = {"commit": commit, "captured": datetime.date.today().isoformat(), "batch": 512,
           "how": "ncu --set full --clock-control none -k regex:edge_fwd -s 2 -c 1 python scripts/prof_edge.py <precision>",
           "kernels": {}}
    for arg in sys.argv[1:]:
        prec, rep = arg.split("=", 1)
        rec["kernels"][prec] = read(rep)
    json.dump(rec, open("profiles/ncu_edge_fwd.json", "w"), indent=1, sort_keys=True)
    print(json.dumps(rec, indent=1)[:600])


if __name__ == "__main__":
    main()
