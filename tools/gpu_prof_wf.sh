#!/bin/bash
# one full ncu capture of the wavefront closest-hit kernel (bounce 2) on the 10 M-triangle height field
mkdir -p gpurun_out
TAG=${1:-wf_hf}
CMD="python tools/prof_render.py --workload heightfield --res 3840 --spp 2 --launches 1"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_wf_trace -s 4 -c 1 -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/plain_$TAG.log; tail -2 gpurun_out/ncu_$TAG.log
