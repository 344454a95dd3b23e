#!/bin/bash
# traversal counters of the height field (WF_STATS build): node steps / leaf triangles / warp iterations per ray
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PRT_B200_LIB=$PWD/build_variants/wfstats.so PRT_WF_STATS=1
python tools/hf_sweep.py --reps 1 --configs "PRT_WF_SORT=0" --tag r02w_base > gpurun_out/r02w_base.log 2>&1
PRT_BIG_TRIS=1 python tools/hf_sweep.py --reps 1 --configs "PRT_WF_SORT=0" --tag r02w_big > gpurun_out/r02w_big.log 2>&1
grep -h "wf_stats\|bvh" gpurun_out/r02w_*.log | tail -8
