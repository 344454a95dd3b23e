#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r01h}
for wl in sphere_box ring; do
python tools/prof_acquire.py --workload $wl --launches 1 > gpurun_out/plain_${TAG}_$wl.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_acquire -s 2 -c 1 -f -o gpurun_out/prof_${TAG}_$wl python tools/prof_acquire.py --workload $wl --launches 1 > gpurun_out/ncu_${TAG}_$wl.log 2>&1
tail -1 gpurun_out/ncu_${TAG}_$wl.log
done
