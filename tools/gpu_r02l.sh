#!/bin/bash
# r02l: deferred triangle phase (WF_TRI_DEFER) A/B on the height field + wavefront cbox; smoke
mkdir -p gpurun_out
fmt='import sys, json
for l in sys.stdin:
    d = json.loads(l); print("%s %-20s %8.2f ms %7.0f Mrays/s maxdiff=%.3g %s" % (sys.argv[1], d["config"], d["kernel_ms"], d["mrays"], d["maxdiff"], d["classes"]))'
for v in default defer32 defer16 defer64; do
  LIB=$PWD/build_variants/$v.so; [ $v = default ] && LIB=$PWD/physics-based-ray-tracing_b200/libprt_b200.so
  PRT_B200_LIB=$LIB timeout 600 python tools/hf_sweep.py --tag r02l_$v --configs "PRT_WF_SORT=0" > gpurun_out/r02l_$v.log 2>&1
  grep -E '^\{' gpurun_out/r02l_$v.log | python -c "$fmt" $v || tail -5 gpurun_out/r02l_$v.log
  PRT_PT_MODE=wavefront PRT_B200_LIB=$LIB timeout 300 python bench.py --workload cbox --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 --also none > gpurun_out/r02l_cbox_$v.json 2> gpurun_out/r02l_cbox_$v.err
  python -c "
import json
d = json.load(open('gpurun_out/r02l_cbox_$v.json')); print('$v cbox(wavefront) Mrays/s %.0f ms %.2f' % (d['value'], d['ms_per_step']))"
done
PRT_B200_LIB=$PWD/build_variants/defer32.so python -m pytest tests/test_gpu_path.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
