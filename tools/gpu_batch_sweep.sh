#!/bin/bash
# sweep of the wavefront batch size (PRT_WF_BATCH, run-time knob) on the two path-tracing workloads
for b in 4194304 8388608 16777216 33554432; do
  for wl in cbox heightfield; do
    PRT_WF_BATCH=$b python bench.py --workload $wl --steps 4 --no-cpu-baseline --e2e-steps 1 > gpurun_out/batch.json 2>/dev/null
    python -c "
import json; d=json.load(open('gpurun_out/batch.json')); print('batch $b $wl', round(d['value']), round(d['ms_per_step'],2))"
  done
done
