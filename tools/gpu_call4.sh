#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
cp physics-based-ray-tracing_b200/libprt_b200.so build_variants/base.so
WLS="heightfield cbox" bash tools/gpu_sweep.sh 2>&1 | tee gpurun_out/sweep4.txt
