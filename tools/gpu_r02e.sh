#!/bin/bash
# r02e: uniform Morton scale + oversized-triangle split on the height field (A/B), state-machine resident kernel on cbox, full suite
mkdir -p gpurun_out
fmt='import sys, json
for l in sys.stdin:
    d = json.loads(l); print("%s %-40s %8.2f ms %7.0f Mrays/s maxdiff=%.3g %s" % (sys.argv[1], d["config"], d["kernel_ms"], d["mrays"], d["maxdiff"], d["classes"]))'
PRT_BIG_TRIS=0 timeout 600 python tools/hf_sweep.py --tag r02e_big0 --configs "PRT_WF_SORT=0" > gpurun_out/r02e_big0.log 2>&1
grep -E '^\{' gpurun_out/r02e_big0.log | python -c "$fmt" big0; grep bvh gpurun_out/r02e_big0.log | cut -c1-420
timeout 600 python tools/hf_sweep.py --tag r02e_big1 --configs "PRT_WF_SORT=0;PRT_WF_SORT=7;PRT_WF_SORT=7,PRT_WF_SORT_WHAT=2" > gpurun_out/r02e_big1.log 2>&1
grep -E '^\{' gpurun_out/r02e_big1.log | python -c "$fmt" big1; grep bvh gpurun_out/r02e_big1.log | cut -c1-420
run() { # name lib
  PRT_B200_LIB=$2 timeout 300 python bench.py --workload cbox --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 --also none > gpurun_out/r02e_$1.json 2> gpurun_out/r02e_$1.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02e_$1.json"))
    print("%-14s Mrays/s %6.0f ms %7.2f e2e %6.0f" % ("$1", d["value"], d["ms_per_step"], d["e2e"]["value"]))
except Exception as e:
    print("$1 FAILED", e); print(open("gpurun_out/r02e_$1.err").read()[-800:])
PY
}
run regen6 $PWD/physics-based-ray-tracing_b200/libprt_b200.so
for v in regen4 regen10 regen16; do run $v $PWD/build_variants/$v.so; done
python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r02e_pytest.log
