#!/bin/bash
# Tuning aid: headline-kernel counters of the current build (run on the GPU box): tools/quick_ncu.sh [out.csv]
OUT=${1:-gpurun_out/q.csv}
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active
for r in no_instruction wait short_scoreboard branch_resolving not_selected long_scoreboard math_pipe_throttle barrier dispatch_stall mio_throttle lg_throttle membar; do M=$M,smsp__warp_issue_stalled_${r}_per_warp_active.pct; done
ncu --metrics $M --clock-control none -k regex:swarm_kernel -s 8 -c 1 --csv --log-file $OUT python bench.py --steps 10 --warmup 3 --no-others --no-cpu > /dev/null 2>&1
python - <<PY
import csv
rows=list(csv.reader(open("$OUT")))
hdr=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
h=rows[hdr]
d={}
for r in rows[hdr+1:]:
    if len(r)<len(h): continue
    rec=dict(zip(h,r))
    d[rec["Metric Name"].replace("smsp__warp_issue_stalled_","").replace("_per_warp_active.pct","")[:36]]=rec["Metric Value"]
print(d)
PY
