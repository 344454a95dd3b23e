#!/bin/bash
# r02v: resident kernel (stash) occupancy / batching variants on cbox
mkdir -p gpurun_out
run() { # name lib
  PRT_B200_LIB=$2 timeout 300 python bench.py --workload cbox --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 2 --also none > gpurun_out/r02v_$1.json 2> gpurun_out/r02v_$1.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02v_$1.json"))
    print("%-14s Mrays/s %6.0f ms %7.2f e2e %6.0f" % ("$1", d["value"], d["ms_per_step"], d["e2e"]["value"]))
except Exception as e:
    print("$1 FAILED", e); print(open("gpurun_out/r02v_$1.err").read()[-800:])
PY
}
run base $PWD/physics-based-ray-tracing_b200/libprt_b200.so
for v in stash_m4 stash_r3; do run $v $PWD/build_variants/$v.so; done
