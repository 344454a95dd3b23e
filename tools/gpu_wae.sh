#!/bin/bash
# validate + measure the warp-per-(angle, element) lane map (build_variants/wae1.so: mesh scenes, wae2.so: all scenes)
PRT_B200_LIB=$PWD/build_variants/wae2.so python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -m gpu -k "ring or buffer or sharded or config3 or config2" 2>&1 | tail -2
run() { PRT_B200_LIB=$PWD/$1 python bench.py --workload $2 --steps 10 --no-cpu-baseline --no-also --e2e-steps 1 > gpurun_out/wae.json 2>gpurun_out/wae.err; python -c "
import json; d=json.load(open('gpurun_out/wae.json')); print('$1 $2', round(d['value']), round(d['ms_per_step'],2))"; }
run physics-based-ray-tracing_b200/libprt_b200.so ring
run build_variants/wae1.so ring
run physics-based-ray-tracing_b200/libprt_b200.so sphere_box:intended
run build_variants/wae2.so sphere_box:intended
run build_variants/wae2.so sphere_box
