#!/bin/bash
# after the super root: full GPU tests, retire/refill batching variants on the height field, fresh ncu capture of the closest-hit kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python tools/hf_sweep.py --reps 3 --configs "PRT_WF_SORT=0" --tag r02y_base > gpurun_out/r02y_base.log 2>&1
for v in rmin3 rmin6 rmin10; do
  PRT_B200_LIB=$PWD/build_variants/$v.so python tools/hf_sweep.py --reps 3 --configs "PRT_WF_SORT=0" --tag r02y_$v > gpurun_out/r02y_$v.log 2>&1
done
grep -h "kernel_ms" gpurun_out/r02y_*.log | cut -c1-330
run() {  # name, kernel regex, skip, command...
  local name=$1 k=$2 skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/prof_r02y_$name "$@" > gpurun_out/ncu_r02y_$name.log 2>&1
  tail -1 gpurun_out/ncu_r02y_$name.log
}
run hf_closest 'k_wf_trace' 4 python tools/prof_render.py --workload heightfield --res 3840 --spp 2 --launches 1
run hf_shadow 'k_wf_trace' 5 python tools/prof_render.py --workload heightfield --res 3840 --spp 2 --launches 1
