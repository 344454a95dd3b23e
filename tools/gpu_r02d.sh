#!/bin/bash
# r02d: oversized-triangle split A/B on the height field, full GPU suite, the new bench.py default line
mkdir -p gpurun_out
fmt='import sys, json
for l in sys.stdin:
    d = json.loads(l); print("%s %-40s %8.2f ms %7.0f Mrays/s maxdiff=%.3g %s" % (sys.argv[1], d["config"], d["kernel_ms"], d["mrays"], d["maxdiff"], d["classes"]))'
PRT_BIG_TRIS=0 timeout 600 python tools/hf_sweep.py --tag r02d_big0 --configs "PRT_WF_SORT=0" > gpurun_out/r02d_big0.log 2>&1
grep -E '^\{' gpurun_out/r02d_big0.log | python -c "$fmt" big0; grep bvh gpurun_out/r02d_big0.log | cut -c1-300
timeout 600 python tools/hf_sweep.py --tag r02d_big1 --configs "PRT_WF_SORT=0;PRT_WF_SORT=7;PRT_WF_SORT=7,PRT_WF_SORT_WHAT=2" > gpurun_out/r02d_big1.log 2>&1
grep -E '^\{' gpurun_out/r02d_big1.log | python -c "$fmt" big1; grep bvh gpurun_out/r02d_big1.log | cut -c1-300
python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r02d_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; tail -c 1700 gpurun_out/r02d_bench.json; tail -5 gpurun_out/r02d_bench.err
