#!/bin/bash
# A/B of library builds on every mesh workload
mkdir -p gpurun_out
for so in "$@"; do
  for wl in ring cbox heightfield; do
    PRT_B200_LIB=$PWD/$so python bench.py --workload $wl --steps 5 --no-cpu-baseline --no-also --e2e-steps 1 > gpurun_out/ab_all.json 2> gpurun_out/ab_all.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/ab_all.json")); print("%-28s %-12s Mrays/s %7.0f ms %7.2f" % ("$so", "$wl", d["value"], d["ms_per_step"]), {k: round(v["ms_per_step"], 2) for k, v in d.get("kernel_classes", {}).items()})
except Exception as e:
    print("$so $wl FAILED", e); print(open("gpurun_out/ab_all.err").read()[-800:])
PY
  done
done
