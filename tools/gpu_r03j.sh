#!/bin/bash
# BVH8 collapse variants on the height field: kernel times + traversal counters  (VARIANTS: names with a <name>_stats twin)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in ${VARIANTS}; do
  PRT_B200_LIB=$PWD/build_variants/$v.so python tools/hf_sweep.py --reps 3 --configs "PRT_WF_SORT=0" --tag r03j_$v > gpurun_out/r03j_$v.log 2>&1
  PRT_WF_STATS=1 PRT_B200_LIB=$PWD/build_variants/${v}_stats.so python tools/hf_sweep.py --reps 1 --configs "PRT_WF_SORT=0" --tag r03j_${v}_stats > gpurun_out/r03j_${v}_stats.log 2>&1
  echo "$v $(grep -h kernel_ms gpurun_out/r03j_$v.log | cut -c28-330)"; grep -h "^bvh" gpurun_out/r03j_$v.log | grep -o "'n_nodes8': [0-9]*, 'bvh8_levels': [0-9]*"
  echo "   $(grep -h wf_stats gpurun_out/r03j_${v}_stats.log | tail -1)"
done
