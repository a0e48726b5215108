#!/bin/bash
# Tuning aid (GPU box): e2e agent-steps/s of the headline workload for several chunk counts of swarm_host_step
for c in 2 4 8 16 32; do
  out=$(SWARM_HOST_CHUNKS=$c python bench.py --steps 100 --warmup 10 --no-others --no-cpu 2>/dev/null | tail -1)
  echo "chunks=$c $(echo "$out" | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("e2e=%.4e ms=%.4f" % (d["e2e"]["value"], d["per_rank"]["e2e_ms_per_step"][0]))')"
done
