#!/bin/bash
# quick loop: wavefront parity + the two path-tracing bench workloads
mkdir -p gpurun_out
python -m pytest tests/test_gpu_path.py -x -q -m gpu 2>&1 | tail -5
for wl in cbox heightfield; do
    timeout 600 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/wf_${wl}.json 2> gpurun_out/wf_${wl}.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/wf_${wl}.json"))
    print("$wl Mrays/s %.0f ms %.2f e2e %.0f frac %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"]))
except Exception as e:
    print("$wl FAILED", e); print(open("gpurun_out/wf_${wl}.err").read()[-2000:])
PY
done
