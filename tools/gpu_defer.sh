#!/bin/bash
# validate + measure the deferred-secondary acquisition loop (build_variants/defer.so) against the default library
PRT_B200_LIB=$PWD/build_variants/defer.so python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_edges.py -x -q -m gpu -k "ring or triangle or decisions or heavy or config3 or mesh or bvh" 2>&1 | tail -3
for so in physics-based-ray-tracing_b200/libprt_b200.so build_variants/defer.so; do
  PRT_B200_LIB=$PWD/$so python bench.py --workload ring --steps 10 --no-cpu-baseline --no-also > gpurun_out/defer.json 2>gpurun_out/defer.err
  python -c "
import json; d=json.load(open('gpurun_out/defer.json')); print('$so ring', round(d['value']), round(d['ms_per_step'],2), round(d['e2e']['value']))"
done
