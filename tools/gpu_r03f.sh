#!/bin/bash
# resident kernel: per-triangle box culling (default) against the plain brute-force loop
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_path.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -3
for v in base ${VARIANTS:-nocull}; do
  if [ $v = base ]; then unset PRT_B200_LIB; else export PRT_B200_LIB=$PWD/build_variants/$v.so; fi
  python bench.py --workload cbox --steps 10 --warmup 3 --no-cpu-baseline --also none > gpurun_out/r03f_$v.json 2> gpurun_out/r03f_$v.err
  python -c "
import json; d=json.loads([l for l in open('gpurun_out/r03f_$v.json') if l.startswith('{')][-1]); print('$v', round(d['value']), 'Mrays/s', round(d['ms_per_step'],3), 'ms  e2e', round(d['e2e']['value']), 'ck', d['e2e'].get('host_checksum'))" || tail -3 gpurun_out/r03f_$v.err
done
