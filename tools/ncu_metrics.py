#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` export into the small JSON bench.py reads (profiles/rNN_kernel_metrics.json).

    python tools/ncu_metrics.py gpurun_out/r01_swarm_kernel_ncu_raw.csv profiles/r01_kernel_metrics.json
"""
import csv, json, sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
h, units = rows[hdr], rows[hdr + 1]
recs = [dict(zip(h, r)) for r in rows[hdr + 2:] if len(r) == len(h)]


def f(rec, key):
    v = rec.get(key, "").replace(",", "")
    return float(v) if v else None


def to_bytes(rec, key):
    v, u = f(rec, key), units[h.index(key)].lower()
    scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    return None if v is None else v * scale


def to_us(rec, key):
    v, u = f(rec, key), units[h.index(key)].lower()
    scale = {"ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}.get(u, 1)
    return None if v is None else v * scale


out = {
    "kernel": recs[0]["Kernel Name"],
    "source": f"ncu --set full --clock-control none, {sys.argv[1].split('/')[-1]} ({len(recs)} launches)",
    "grid": recs[0].get("Grid Size"), "block": recs[0].get("Block Size"),
    "gpu_time_us": [round(to_us(r, "gpu__time_duration.sum"), 2) for r in recs],
    "dram_bytes_read": [to_bytes(r, "dram__bytes_read.sum") for r in recs],
    "dram_bytes_write": [to_bytes(r, "dram__bytes_write.sum") for r in recs],
    "registers_per_thread": f(recs[0], "launch__registers_per_thread"),
    "warp_instructions": f(recs[0], "smsp__inst_executed.sum"),
    "issue_active_pct": f(recs[0], "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "warps_active_pct": f(recs[0], "sm__warps_active.avg.pct_of_peak_sustained_active"),
    "threads_per_instruction": f(recs[0], "smsp__thread_inst_executed_per_inst_executed.ratio"),
    "pipe_alu_pct": f(recs[0], "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    "pipe_fma_pct": f(recs[0], "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    "dram_throughput_pct": f(recs[0], "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
}
n = len(recs)
out["dram_bytes_per_launch"] = (sum(out["dram_bytes_read"]) + sum(out["dram_bytes_write"])) / n
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out, indent=1))
