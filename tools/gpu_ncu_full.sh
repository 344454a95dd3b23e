#!/bin/bash
# (k_wf_trace launches alternate closest / shadow per bounce: skip 4 = closest hit of bounce 2, skip 5 = its shadow rays)
# `ncu --set full` captures of the dominant kernels (one launch each), after the plain command has exited 0
mkdir -p gpurun_out
TAG=${1:-r01}
run() {  # name, kernel regex, skip, command...
  local name=$1 k=$2 skip=$3; shift 3
  "$@" > gpurun_out/plain_${TAG}_$name.log 2>&1 || { echo "$name: plain run failed"; tail -3 gpurun_out/plain_${TAG}_$name.log; return; }
  ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/prof_${TAG}_$name "$@" > gpurun_out/ncu_${TAG}_$name.log 2>&1
  tail -1 gpurun_out/ncu_${TAG}_$name.log
}
run hf_closest k_wf_trace 4 python tools/prof_render.py --workload heightfield --res 3840 --spp 2 --launches 1
run hf_shadow k_wf_trace 5 python tools/prof_render.py --workload heightfield --res 3840 --spp 2 --launches 1
run hf_shade 'k_wf_shade' 2 python tools/prof_render.py --workload heightfield --res 3840 --spp 2 --launches 1
run cbox_closest k_wf_trace 2 python tools/prof_render.py --workload cbox --res 2048 --spp 4 --launches 1
run cbox_shade 'k_wf_shade' 3 python tools/prof_render.py --workload cbox --res 2048 --spp 4 --launches 1
run sphere_box k_acquire 2 python tools/prof_acquire.py --workload sphere_box --launches 1
run ring k_acquire 2 python tools/prof_acquire.py --workload ring --launches 1
ls -la gpurun_out/*.ncu-rep
