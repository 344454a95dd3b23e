#!/bin/bash
# r02j: owner binary search in the cooperative triangle test; oversized triangles dealt as the first triangle group
mkdir -p gpurun_out
fmt='import sys, json
for l in sys.stdin:
    d = json.loads(l); print("%s %-20s %8.2f ms %7.0f Mrays/s maxdiff=%.3g %s" % (sys.argv[1], d["config"], d["kernel_ms"], d["mrays"], d["maxdiff"], d["classes"]))'
timeout 600 python tools/hf_sweep.py --tag r02j_big0 --configs "PRT_WF_SORT=0" > gpurun_out/r02j_big0.log 2>&1
grep -E '^\{' gpurun_out/r02j_big0.log | python -c "$fmt" big0
PRT_BIG_TRIS=1 timeout 600 python tools/hf_sweep.py --tag r02j_big1 --configs "PRT_WF_SORT=0" > gpurun_out/r02j_big1.log 2>&1
grep -E '^\{' gpurun_out/r02j_big1.log | python -c "$fmt" big1
PRT_PT_MODE=wavefront timeout 300 python bench.py --workload cbox --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 --also none > gpurun_out/r02j_cbox.json 2> gpurun_out/r02j_cbox.err
python -c "
import json
d = json.load(open('gpurun_out/r02j_cbox.json')); print('cbox(wavefront) Mrays/s %.0f ms %.2f' % (d['value'], d['ms_per_step']))"
python -m pytest tests/test_gpu_path.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -4
PRT_BIG_TRIS=1 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k config5 2>&1 | tail -3
