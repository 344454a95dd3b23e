#!/bin/bash
# r02n: why are the +-15 degree launches of the headline 2x slower than the 0 degree one?  ncu of the -15 degree launch
mkdir -p gpurun_out
TAG=r02n
run() {  # name, kernel regex, skip, command...
  local name=$1 k=$2 skip=$3; shift 3
  "$@" > gpurun_out/plain_${TAG}_$name.log 2>&1 || { echo "$name: plain run failed"; tail -3 gpurun_out/plain_${TAG}_$name.log; return; }
  ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/prof_${TAG}_$name "$@" > gpurun_out/ncu_${TAG}_$name.log 2>&1
  tail -1 gpurun_out/ncu_${TAG}_$name.log
}
run headline_m15 k_acquire 5 python tools/prof_acquire.py --workload sphere_box:intended --launches 2
run headline_p75 k_acquire 8 python tools/prof_acquire.py --workload sphere_box:intended --launches 2
