#!/bin/bash
# After `tools/ncu_all.sh r02` came back in gpurun_out/: copy the raw pages into profiles/ and rebuild the summaries.
set -e
cd "$(dirname "$0")/.."
for f in gpurun_out/r02_*_raw.csv gpurun_out/r02_launches.csv; do cp $f profiles/; done
W="foraging_daisy_16384 homing_lily_4096 dirgate_dandelion_8192 sheltering_oc2_16384 xor_cyclamen_16384"
ARGS=""
for w in $W; do ARGS="$ARGS $w=profiles/r02_${w}_raw.csv"; done
for w in dirgate_dandelion_8192 sheltering_oc2_16384 foraging_daisy_16384 homing_lily_4096; do ARGS="$ARGS $w@rollout5=profiles/r02_${w}_rollout5_raw.csv"; done
python tools/ncu_metrics.py profiles/r02_kernel_metrics.json $ARGS
python tools/ncu_lanes.py gpurun_out/r02_foraging_daisy_16384_sass.csv swarmacb-isaaclab_b200/libswarmstep.so 'swarm_kernelILi3ELb1ELi24ELi0E' 30 > profiles/r02_foraging_daisy_16384_lanes.txt 2>&1
HELPER_MAX=125 python tools/ncu_lines.py gpurun_out/r02_foraging_daisy_16384_sass.csv swarmacb-isaaclab_b200/libswarmstep.so 'swarm_kernelILi3ELb1ELi24ELi0E' > profiles/r02_foraging_daisy_16384_by_line.txt 2>&1
