#!/bin/bash
# ring (config 3) with the 8-wide tree in the acquisition megakernel, against the binary default
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in base ${VARIANTS:-mega8}; do
  if [ $v = base ]; then unset PRT_B200_LIB; else export PRT_B200_LIB=$PWD/build_variants/$v.so; fi
  python bench.py --workload ring --steps 10 --warmup 3 --no-cpu-baseline --also none > gpurun_out/r03d_$v.json 2> gpurun_out/r03d_$v.err
  python -c "
import json; d=json.loads([l for l in open('gpurun_out/r03d_$v.json') if l.startswith('{')][-1]); print('$v', round(d['value']), 'Mrays/s', round(d['ms_per_step'],3), 'ms  ck', d['e2e'].get('host_checksum'))" || tail -3 gpurun_out/r03d_$v.err
done
