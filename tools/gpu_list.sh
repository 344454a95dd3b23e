#!/bin/bash
# per-kernel launch list (durations) of one wavefront render of the height field / cbox
mkdir -p gpurun_out
TAG=${1:-list}
WL=${2:-heightfield}
if [ "$WL" = heightfield ]; then CMD="python tools/prof_render.py --workload heightfield --res 3840 --spp 2 --launches 1"; else CMD="python tools/prof_render.py --workload cbox --res 2048 --spp 4 --launches 1"; fi
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
tail -1 gpurun_out/plain_$TAG.log
