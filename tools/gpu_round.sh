#!/bin/bash
# round check on the GPU box: full GPU test suite, smoke, every bench workload, reference arm, launch lists
mkdir -p gpurun_out
TAG=${1:-r01}
python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/pytest_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke_$TAG.log
timeout 400 python bench.py --steps 30 --warmup 3 > gpurun_out/bench_${TAG}_sphere_box.json 2> gpurun_out/bench_${TAG}_sphere_box.err
for wl in ring cbox heightfield; do
  timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_$wl.json 2> gpurun_out/bench_${TAG}_$wl.err
done
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2>&1
for wl in sphere_box ring cbox heightfield reference; do
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${TAG}_${wl}.json"))
    print("%-12s Mrays/s %8.0f ms %8.2f e2e %8.0f frac %s launches %s cpu %s" % ("$wl", d["value"], d["ms_per_step"], d["e2e"]["value"],
          d.get("roofline", {}).get("frac"), d.get("gpu_launches"), d.get("cpu_baseline", {}).get("value")))
except Exception as e:
    print("$wl FAILED", e)
PY
done
