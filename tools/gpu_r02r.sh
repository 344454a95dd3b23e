#!/bin/bash
# r02p: final captures after the reconvergence fix: acquisition kernels (ncu --set full), launch list, default bench line
mkdir -p gpurun_out
TAG=r02r
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --also none > gpurun_out/plain_${TAG}_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --also none > gpurun_out/ncu_${TAG}_bench.log 2>&1
run() {  # name, kernel regex, skip, command...
  local name=$1 k=$2 skip=$3; shift 3
  "$@" > gpurun_out/plain_${TAG}_$name.log 2>&1 || { echo "$name: plain run failed"; tail -3 gpurun_out/plain_${TAG}_$name.log; return; }
  ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/prof_${TAG}_$name "$@" > gpurun_out/ncu_${TAG}_$name.log 2>&1
  tail -1 gpurun_out/ncu_${TAG}_$name.log
}
run headline k_acquire 7 python tools/prof_acquire.py --workload sphere_box:intended --launches 2
run headline_m15 k_acquire 5 python tools/prof_acquire.py --workload sphere_box:intended --launches 2
run sphere_box k_acquire 7 python tools/prof_acquire.py --workload sphere_box --launches 2
run ring0 k_acquire 7 python tools/prof_acquire.py --workload ring --launches 2
run ring15 k_acquire 5 python tools/prof_acquire.py --workload ring --launches 2
python tools/per_angle_stats.py > gpurun_out/r02r_per_angle.log 2>&1
python -m pytest tests -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r02r_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r02r_smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02r_bench.json 2> gpurun_out/r02r_bench.err; tail -c 1500 gpurun_out/r02r_bench.json; echo; tail -3 gpurun_out/r02r_bench.err
