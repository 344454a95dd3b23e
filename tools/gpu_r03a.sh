#!/bin/bash
# latency experiments on the height field: 96-byte nodes (256-bit loads), early stack pop, prefetch combinations
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/hf_sweep.py --reps 3 --configs "PRT_WF_SORT=0" --tag r03a_base > gpurun_out/r03a_base.log 2>&1
for v in ${VARIANTS:-epop epop_top top epop_rmin10 epop_ss4 epop_top_ss4}; do
  PRT_B200_LIB=$PWD/build_variants/$v.so python tools/hf_sweep.py --reps 3 --configs "PRT_WF_SORT=0" --tag r03a_$v > gpurun_out/r03a_$v.log 2>&1
done
for v in base ${VARIANTS:-epop epop_top top epop_rmin10 epop_ss4 epop_top_ss4}; do echo "$v $(grep -h kernel_ms gpurun_out/r03a_$v.log | cut -c28-330)"; done
