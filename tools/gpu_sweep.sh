#!/bin/bash
# A/B sweep over build_variants/*.so on the two path-tracing workloads
mkdir -p gpurun_out
for so in build_variants/*.so; do
  name=$(basename $so .so)
  for wl in ${WLS:-heightfield cbox}; do
    PRT_B200_LIB=$PWD/$so timeout 300 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/sw_${name}_${wl}.json 2> gpurun_out/sw_${name}_${wl}.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/sw_${name}_${wl}.json"))
    print("%-14s %-12s Mrays/s %6.0f ms %7.2f" % ("$name", "$wl", d["value"], d["ms_per_step"]))
except Exception as e:
    print("$name $wl FAILED", e); print(open("gpurun_out/sw_${name}_${wl}.err").read()[-800:])
PY
  done
done
