#!/bin/bash
# Run on the GPU box: one `ncu --set full` capture of the step kernel of every bench workload and of the fused
# decision-period kernel of the two wheel-action workloads; raw + SASS-source pages exported as CSV.
#   tools/ncu_all.sh [prefix]   ->  gpurun_out/<prefix>_<workload>[_rollout5]{.ncu-rep,_raw.csv,_sass.csv}
PFX=${1:-r02}
for WL in foraging_daisy_16384 homing_lily_4096 dirgate_dandelion_8192 sheltering_oc2_16384 xor_cyclamen_16384; do
  tools/ncu_capture.sh $WL gpurun_out/${PFX}_$WL 8 > /dev/null
  rm -f gpurun_out/${PFX}_$WL.ncu-rep     # the CSV exports carry everything quoted; gpurun_out/ is capped at 64 MiB
done
for WL in dirgate_dandelion_8192 sheltering_oc2_16384 foraging_daisy_16384 homing_lily_4096; do
  OUT=gpurun_out/${PFX}_${WL}_rollout5
  python tools/time_rollout.py 5 $WL > /dev/null 2>&1 || { echo "time_rollout failed"; exit 1; }
  ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:ELi2EEEv -s 12 -c 1 -f -o $OUT \
      python tools/time_rollout.py 5 $WL > $OUT.log 2>&1
  ncu -i $OUT.ncu-rep --page raw --csv > ${OUT}_raw.csv 2>/dev/null
  ncu -i $OUT.ncu-rep --page source --csv > ${OUT}_sass.csv 2>/dev/null
  rm -f $OUT.ncu-rep
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${PFX}_launches.csv \
    python bench.py --steps 20 --warmup 3 --no-others --no-cpu > /dev/null 2>&1
ls -la gpurun_out/${PFX}_*raw.csv gpurun_out/${PFX}_launches.csv
