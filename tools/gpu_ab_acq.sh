#!/bin/bash
# A/B of two library builds on the acquisition workloads
mkdir -p gpurun_out
for so in "$@"; do
  for wl in sphere_box sphere_box:intended plate_box:intended cone_box:intended ring; do
    PRT_B200_LIB=$PWD/$so python bench.py --workload $wl --steps 20 --no-cpu-baseline --no-also > gpurun_out/ab_acq.json 2> gpurun_out/ab_acq.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/ab_acq.json")); print("%-40s %-22s Mrays/s %7.0f ms %7.2f e2e %7.0f" % ("$so", "$wl", d["value"], d["ms_per_step"], d["e2e"]["value"]))
except Exception as e:
    print("$so $wl FAILED", e); print(open("gpurun_out/ab_acq.err").read()[-800:])
PY
  done
done
