#!/bin/bash
# Tuning aid (GPU box): every library variant on the headline workload early (steps 5..55) and late (steps 1500..1550)
# in the episode - under random actions robots drift into walls and each other, the late steps do more work.
for i in 1 2; do
for lib in swarmacb-isaaclab_b200/variants/lib_*.so; do
  for w in 5 1500; do
    out=$(SWARM_LIB_OVERRIDE=$PWD/$lib python bench.py --steps 50 --warmup $w --no-others --no-cpu --workload ${1:-foraging_daisy_16384} 2>&1 | tail -1)
    echo "$(basename $lib) warmup=$w $(echo "$out" | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("ms=%.4f" % (d["ms_per_step"]))')"
  done
done
done
