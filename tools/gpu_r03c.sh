#!/bin/bash
# r03c: verification of the build with the super root + new trace-kernel defaults: suite, smoke, bench, hf captures + launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=r03k
python -m pytest tests -q -m gpu 2>&1 | tail -4 | tee gpurun_out/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 1500 gpurun_out/${TAG}_bench.json; echo; tail -3 gpurun_out/${TAG}_bench.err
python tools/hf_sweep.py --reps 3 --configs "PRT_WF_SORT=0" --tag ${TAG}_hf > gpurun_out/${TAG}_hf.log 2>&1; grep -h kernel_ms gpurun_out/${TAG}_hf.log | cut -c1-330
run() {  # name, kernel regex, skip, command...
  local name=$1 k=$2 skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/prof_${TAG}_$name "$@" > gpurun_out/ncu_${TAG}_$name.log 2>&1
  tail -1 gpurun_out/ncu_${TAG}_$name.log
}
run hf_closest 'k_wf_trace' 4 python tools/prof_render.py --workload heightfield --res 3840 --spp 2 --launches 1
run hf_shadow 'k_wf_trace' 5 python tools/prof_render.py --workload heightfield --res 3840 --spp 2 --launches 1
run hf_shade 'k_wf_shade' 2 python tools/prof_render.py --workload heightfield --res 3840 --spp 2 --launches 1
python bench.py --workload heightfield --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --also none > gpurun_out/plain_${TAG}_bench_hf.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_${TAG}_bench_hf.csv python bench.py --workload heightfield --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --also none > gpurun_out/ncu_${TAG}_bench_hf.log 2>&1
PRT_WF_DEBUG=1 python tools/prof_render.py --workload heightfield --res 3840 --spp 2 --launches 1 2>&1 | grep "prt wf" > gpurun_out/wf_counts_${TAG}_hf.txt
