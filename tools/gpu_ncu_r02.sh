#!/bin/bash
# Round-2 profile captures.  Every ncu run follows a plain run of the same command that exited 0.
#   1. launch lists (gpu__time_duration.sum per launch) of the default bench command and of the heightfield command
#   2. `ncu --set full` of one launch of every dominant kernel
mkdir -p gpurun_out
TAG=${1:-r02}
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --also none > gpurun_out/plain_${TAG}_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --also none > gpurun_out/ncu_${TAG}_bench.log 2>&1
python bench.py --workload heightfield --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --also none > gpurun_out/plain_${TAG}_bench_hf.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_${TAG}_bench_hf.csv python bench.py --workload heightfield --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --also none > gpurun_out/ncu_${TAG}_bench_hf.log 2>&1
python bench.py --workload cbox --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --also none > gpurun_out/plain_${TAG}_bench_cbox.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_${TAG}_bench_cbox.csv python bench.py --workload cbox --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --also none > gpurun_out/ncu_${TAG}_bench_cbox.log 2>&1
run() {  # name, kernel regex, skip, command...
  local name=$1 k=$2 skip=$3; shift 3
  "$@" > gpurun_out/plain_${TAG}_$name.log 2>&1 || { echo "$name: plain run failed"; tail -3 gpurun_out/plain_${TAG}_$name.log; return; }
  ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/prof_${TAG}_$name "$@" > gpurun_out/ncu_${TAG}_$name.log 2>&1
  tail -1 gpurun_out/ncu_${TAG}_$name.log
}
# k_acquire launches: 5 per acquisition (one per steering angle); skip 7 = the 0-degree angle (index 2) of the second acquisition
run headline k_acquire 7 python tools/prof_acquire.py --workload sphere_box:intended --launches 2
run sphere_box k_acquire 7 python tools/prof_acquire.py --workload sphere_box --launches 2
run ring0 k_acquire 7 python tools/prof_acquire.py --workload ring --launches 2
run ring15 k_acquire 5 python tools/prof_acquire.py --workload ring --launches 2
run cbox_resident k_render_resident 1 python tools/prof_render.py --workload cbox --res 2048 --spp 16 --launches 2
run hf_closest 'k_wf_trace' 4 python tools/prof_render.py --workload heightfield --res 3840 --spp 2 --launches 1
run hf_shadow 'k_wf_trace' 5 python tools/prof_render.py --workload heightfield --res 3840 --spp 2 --launches 1
run hf_shade 'k_wf_shade' 2 python tools/prof_render.py --workload heightfield --res 3840 --spp 2 --launches 1
PRT_WF_DEBUG=1 python tools/prof_render.py --workload heightfield --res 3840 --spp 2 --launches 1 2>&1 | grep "prt wf" > gpurun_out/wf_counts_${TAG}_hf.txt
ls -la gpurun_out/prof_${TAG}_*.ncu-rep
