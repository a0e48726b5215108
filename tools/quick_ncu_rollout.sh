M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum
for r in no_instruction wait short_scoreboard branch_resolving not_selected long_scoreboard barrier; do M=$M,smsp__warp_issue_stalled_${r}_per_warp_active.pct; done
ncu --metrics $M --clock-control none --kernel-name-base mangled -k regex:ELi24ELi2E -s 4 -c 1 --csv --log-file gpurun_out/q3.csv python tools/time_rollout.py 5 foraging_daisy_16384 > /dev/null 2>&1
python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/q3.csv")))
hdr=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
h=rows[hdr]
d={}
for r in rows[hdr+1:]:
    if len(r)<len(h): continue
    rec=dict(zip(h,r))
    d[rec["Metric Name"].replace("smsp__warp_issue_stalled_","").replace("_per_warp_active.pct","")[:40]]=rec["Metric Value"]
print(d)
PY
