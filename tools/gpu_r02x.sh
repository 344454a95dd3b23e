#!/bin/bash
# super root for the oversized triangles: tests, then the height field with and without the split
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
python tools/hf_sweep.py --reps 3 --configs "PRT_WF_SORT=0" --tag r02x_big > gpurun_out/r02x_big.log 2>&1
PRT_BIG_TRIS=0 python tools/hf_sweep.py --reps 3 --configs "PRT_WF_SORT=0" --tag r02x_nobig > gpurun_out/r02x_nobig.log 2>&1
grep -h "kernel_ms\|bvh" gpurun_out/r02x_*.log | cut -c1-420
