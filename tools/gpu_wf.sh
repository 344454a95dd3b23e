#!/bin/bash
# wavefront vs megakernel on the GPU box: parity test, then both back ends on the cbox / height-field workloads
mkdir -p gpurun_out
python -m pytest tests/test_gpu_path.py -x -q -m gpu 2>&1 | tail -15
for mode in mega wavefront; do
  for wl in cbox heightfield; do
    PRT_PT_MODE=$mode timeout 600 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/wf_${wl}_${mode}.json 2> gpurun_out/wf_${wl}_${mode}.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/wf_${wl}_${mode}.json"))
    print("$wl $mode Mrays/s %.0f ms %.2f e2e %.0f frac %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"]))
except Exception as e:
    print("$wl $mode FAILED", e); print(open("gpurun_out/wf_${wl}_${mode}.err").read()[-2000:])
PY
  done
done
