#!/bin/bash
# Tuning aid (GPU box): executed warp instructions + duration of the headline step kernel for every library variant
M=gpu__time_duration.sum,smsp__inst_executed.sum,launch__registers_per_thread
for lib in swarmacb-isaaclab_b200/variants/lib_*.so; do
  SWARM_LIB_OVERRIDE=$PWD/$lib ncu --metrics $M --clock-control none -k regex:swarm_kernel -s 8 -c 1 --csv --log-file /tmp/q.csv python bench.py --steps 10 --warmup 3 --no-others --no-cpu --workload ${1:-foraging_daisy_16384} > /dev/null 2>&1
  echo "$(basename $lib) $(grep -E 'inst_executed|time_duration|registers' /tmp/q.csv | awk -F'","' '{printf "%s=%s ", $(NF-2), $NF}' | tr -d '"')"
done
