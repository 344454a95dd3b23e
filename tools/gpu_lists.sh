#!/bin/bash
# per-bounce queue counts (PRT_WF_DEBUG), clean launch lists of the default and the cbox bench commands, then the default bench line
mkdir -p gpurun_out
PRT_WF_DEBUG=1 python tools/prof_render.py --workload cbox --res 2048 --spp 4 --launches 1 2>&1 | grep "prt wf" > gpurun_out/wf_counts_cbox.txt
PRT_WF_DEBUG=1 python tools/prof_render.py --workload heightfield --res 3840 --spp 2 --launches 1 2>&1 | grep "prt wf" > gpurun_out/wf_counts_hf.txt
cat gpurun_out/wf_counts_cbox.txt gpurun_out/wf_counts_hf.txt
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/plain_bench_g.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_g.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/ncu_bench_g.log 2>&1
python bench.py --workload cbox --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/plain_bench_cbox_g.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_bench_cbox_g.csv python bench.py --workload cbox --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_bench_cbox_g.log 2>&1
python bench.py --steps 30 > gpurun_out/bench_r01g_default.json 2> gpurun_out/bench_r01g_default.err
tail -c 300 gpurun_out/bench_r01g_default.err
