#!/bin/bash
# Tuning aid (run on the GPU box): one `ncu --set full` capture of the step kernel of a bench workload, exported as
# raw-page and SASS-source-page CSVs next to the .ncu-rep.   tools/ncu_capture.sh <workload> <prefix> [skip] [extra bench args]
WL=${1:-foraging_daisy_16384}
PFX=${2:-gpurun_out/cap}
SKIP=${3:-8}
shift 3 2>/dev/null
python bench.py --steps 10 --warmup 3 --no-others --no-cpu --workload $WL "$@" > /dev/null 2>&1 || { echo "bench failed without ncu"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:swarm_kernel -s $SKIP -c 1 -f -o $PFX \
    python bench.py --steps 10 --warmup 3 --no-others --no-cpu --workload $WL "$@" > $PFX.log 2>&1
ncu -i $PFX.ncu-rep --page raw --csv > ${PFX}_raw.csv 2>/dev/null
ncu -i $PFX.ncu-rep --page source --csv > ${PFX}_sass.csv 2>/dev/null
ls -la $PFX*
