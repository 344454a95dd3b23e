#!/bin/bash
# leaf-child size of the 8-wide tree (1 / 2 / 3 triangles) on the height field, with the popc rank select
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/hf_sweep.py --reps 3 --configs "PRT_WF_SORT=0" --tag r02z_base > gpurun_out/r02z_base.log 2>&1
for v in leaf1 leaf2; do
  PRT_B200_LIB=$PWD/build_variants/$v.so python tools/hf_sweep.py --reps 3 --configs "PRT_WF_SORT=0" --tag r02z_$v > gpurun_out/r02z_$v.log 2>&1
done
grep -h "kernel_ms\|bvh" gpurun_out/r02z_*.log | cut -c1-400
