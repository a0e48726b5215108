#!/usr/bin/env python
"""Summarise `ncu --page raw --csv` exports into the JSON bench.py reads for its roofline (profiles/r02_kernel_metrics.json).

    python tools/ncu_metrics.py profiles/r02_kernel_metrics.json name=raw.csv [name=raw.csv ...]

``name`` is a bench.py workload (``foraging_daisy_16384``) or ``<workload>@rollout5`` for the fused decision-period
kernel.  Every number is copied from the capture: a reader can recompute bench.py's roofline fractions from the raw
CSVs committed next to the JSON.
"""
import csv, json, os, sys


def summarise(path):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    h, units = rows[hdr], rows[hdr + 1]
    recs = [dict(zip(h, r)) for r in rows[hdr + 2:] if len(r) == len(h)]
    rec = recs[0]

    def f(key):
        v = rec.get(key, "").replace(",", "")
        return float(v) if v else None

    def scaled(key, table):
        v = f(key)
        return None if v is None else v * table.get(units[h.index(key)].lower(), 1)

    byte = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}
    usec = {"ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "second": 1e6}
    stalls = {k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): round(float(v), 3)
              for k, v in rec.items() if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and v}
    rd, wr = scaled("dram__bytes_read.sum", byte), scaled("dram__bytes_write.sum", byte)
    return {
        "kernel": rec["Kernel Name"], "grid": rec.get("Grid Size"), "block": rec.get("Block Size"),
        "source": f"profiles/{os.path.basename(path)} (ncu --set full --clock-control none, 1 launch)",
        "gpu_time_us": scaled("gpu__time_duration.sum", usec),
        "warp_instructions": f("smsp__inst_executed.sum"),
        "threads_per_instruction": f("smsp__thread_inst_executed_per_inst_executed.ratio"),
        "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "warps_active_pct": f("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "sm_cycles_active_avg": f("sm__cycles_active.avg"), "sm_cycles_elapsed_avg": f("sm__cycles_elapsed.avg"),
        "pipe_alu_pct": f("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        "pipe_fma_pct": f("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        "pipe_xu_pct": f("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
        "registers_per_thread": f("launch__registers_per_thread"),
        "icache_hit_pct": f("sm__icc_request_hit_rate.pct"),
        "dram_bytes_read": rd, "dram_bytes_write": wr,
        "dram_bytes_per_launch": None if rd is None or wr is None else rd + wr,
        "dram_throughput_pct": f("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "stalls_per_issue": stalls,
    }


def main():
    out_path, pairs = sys.argv[1], [a.split("=", 1) for a in sys.argv[2:]]
    out = {"how": "tools/ncu_all.sh on a B200 (one `ncu --set full --clock-control none --import-source on` capture per "
                  "workload, after the same command exited 0 without ncu); summarised by tools/ncu_metrics.py",
           "workloads": {name: summarise(path) for name, path in pairs}}
    json.dump(out, open(out_path, "w"), indent=1)
    for name, w in out["workloads"].items():
        print(f"{name:34s} {w['gpu_time_us']:7.1f} us  {w['warp_instructions']/1e6:6.2f} M warp-instr  "
              f"{w['threads_per_instruction']:5.2f} thr/inst  issue {w['issue_active_pct']:.1f} %")


if __name__ == "__main__":
    main()
