#!/bin/bash
# r03g: verification of the shipped build: suite, smoke, default bench line, reference arm, cbox capture + launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=r03h
python -m pytest tests -q -m gpu 2>&1 | tail -4 | tee gpurun_out/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 1500 gpurun_out/${TAG}_bench.json; echo; tail -3 gpurun_out/${TAG}_bench.err
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_reference.json 2> gpurun_out/${TAG}_reference.err ) 2>&1 | grep real; cut -c1-200 gpurun_out/${TAG}_reference.json
run() {  # name, kernel regex, skip, command...
  local name=$1 k=$2 skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/prof_${TAG}_$name "$@" > gpurun_out/ncu_${TAG}_$name.log 2>&1
  tail -1 gpurun_out/ncu_${TAG}_$name.log
}
run cbox_resident k_render_resident 1 python tools/prof_render.py --workload cbox --res 2048 --spp 16 --launches 2
python bench.py --workload cbox --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --also none > gpurun_out/plain_${TAG}_bench_cbox.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_${TAG}_bench_cbox.csv python bench.py --workload cbox --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --also none > gpurun_out/ncu_${TAG}_bench_cbox.log 2>&1
PRT_WF_DEBUG=1 python tools/prof_render.py --workload cbox --res 2048 --spp 16 --launches 1 2>&1 | tail -2
