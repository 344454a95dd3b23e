#!/bin/bash
# BVH8 collapse: spare slots filled with split leaves (default) against the plain collapse; traversal counters of both
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_path.py tests/test_gpu_fullsize.py tests/test_gpu_edges.py -x -q -m gpu 2>&1 | tail -2
python tools/hf_sweep.py --reps 3 --configs "PRT_WF_SORT=0" --tag r03i_base > gpurun_out/r03i_base.log 2>&1
PRT_B200_LIB=$PWD/build_variants/nofill.so python tools/hf_sweep.py --reps 3 --configs "PRT_WF_SORT=0" --tag r03i_nofill > gpurun_out/r03i_nofill.log 2>&1
for v in fill_stats nofill_stats; do
  PRT_WF_STATS=1 PRT_B200_LIB=$PWD/build_variants/$v.so python tools/hf_sweep.py --reps 1 --configs "PRT_WF_SORT=0" --tag r03i_$v > gpurun_out/r03i_$v.log 2>&1
done
for v in base nofill; do echo "$v $(grep -h kernel_ms gpurun_out/r03i_$v.log | cut -c28-330)"; grep -h "^bvh" gpurun_out/r03i_$v.log | cut -c1-330; done
for v in fill_stats nofill_stats; do echo "$v $(grep -h wf_stats gpurun_out/r03i_$v.log | tail -1)"; done
