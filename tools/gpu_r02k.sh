#!/bin/bash
# r02k: per-angle ray counts, recapture of the height-field trace kernels (owner-search build), full suite, smoke, default bench, reference arm
mkdir -p gpurun_out
python tools/per_angle_stats.py > gpurun_out/r02k_per_angle.log 2>&1; tail -3 gpurun_out/r02k_per_angle.log | cut -c1-300
TAG=r02k
run() {  # name, kernel regex, skip, command...
  local name=$1 k=$2 skip=$3; shift 3
  "$@" > gpurun_out/plain_${TAG}_$name.log 2>&1 || { echo "$name: plain run failed"; tail -3 gpurun_out/plain_${TAG}_$name.log; return; }
  ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/prof_${TAG}_$name "$@" > gpurun_out/ncu_${TAG}_$name.log 2>&1
  tail -1 gpurun_out/ncu_${TAG}_$name.log
}
run hf_closest 'k_wf_trace' 4 python tools/prof_render.py --workload heightfield --res 3840 --spp 2 --launches 1
run hf_shadow 'k_wf_trace' 5 python tools/prof_render.py --workload heightfield --res 3840 --spp 2 --launches 1
python -m pytest tests -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r02k_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r02k_smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err; tail -c 1500 gpurun_out/r02k_bench.json; echo; tail -3 gpurun_out/r02k_bench.err
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02k_reference.json 2> gpurun_out/r02k_reference.err ) 2>&1 | grep real; cut -c1-400 gpurun_out/r02k_reference.json
