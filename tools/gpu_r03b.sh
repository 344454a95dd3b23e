#!/bin/bash
# height-field size scaling (same camera, same resolution): where does the per-ray cost start to depend on the footprint?
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PRT_B200_LIB=$PWD/build_variants/epop.so
for n in 318 708 1001 1582 2237; do
  python tools/hf_sweep.py --n $n --reps 3 --configs "PRT_WF_SORT=0" --tag r03b_n$n > gpurun_out/r03b_n$n.log 2>&1
  echo "n=$n $(grep -h kernel_ms gpurun_out/r03b_n$n.log | cut -c28-330)"
done
