#!/bin/bash
# r02i: triangle-phase peel A/B on the height field and on cbox (wavefront mode)
mkdir -p gpurun_out
fmt='import sys, json
for l in sys.stdin:
    d = json.loads(l); print("%s %-20s %8.2f ms %7.0f Mrays/s maxdiff=%.3g %s" % (sys.argv[1], d["config"], d["kernel_ms"], d["mrays"], d["maxdiff"], d["classes"]))'
for v in default nopeel peel8 peel24; do
  LIB=$PWD/build_variants/$v.so; [ $v = default ] && LIB=$PWD/physics-based-ray-tracing_b200/libprt_b200.so
  PRT_B200_LIB=$LIB timeout 600 python tools/hf_sweep.py --tag r02i_$v --configs "PRT_WF_SORT=0" > gpurun_out/r02i_$v.log 2>&1
  grep -E '^\{' gpurun_out/r02i_$v.log | python -c "$fmt" $v
  PRT_PT_MODE=wavefront PRT_B200_LIB=$LIB timeout 300 python bench.py --workload cbox --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 --also none > gpurun_out/r02i_cbox_$v.json 2> gpurun_out/r02i_cbox_$v.err
  python -c "
import json
d = json.load(open('gpurun_out/r02i_cbox_$v.json')); print('$v cbox(wavefront) Mrays/s %.0f ms %.2f' % (d['value'], d['ms_per_step']))"
done
python -m pytest tests/test_gpu_path.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -4
