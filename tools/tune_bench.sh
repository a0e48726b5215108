#!/bin/bash
# Tuning aid: time every library variant under swarmacb-isaaclab_b200/variants on the headline workload.
for lib in swarmacb-isaaclab_b200/variants/lib_*.so; do
  for wl in foraging_daisy_16384 ${EXTRA_WL}; do
    out=$(SWARM_LIB_OVERRIDE=$PWD/$lib python bench.py --steps 50 --warmup 5 --no-others --no-cpu --workload $wl 2>&1 | tail -1)
    echo "$(basename $lib) $wl $(echo "$out" | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("ms=%.4f value=%.3e" % (d["ms_per_step"], d["value"]))' 2>/dev/null || echo "FAILED: ${out:0:300}")"
  done
done
