#!/bin/bash
# r02g: full suite (refit tests, tightened parity, unmodified USMain.py), ncu of the state-machine resident kernel
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -25 | tee gpurun_out/r02g_pytest.log
python tools/prof_render.py --workload cbox --res 2048 --spp 16 --launches 2 > gpurun_out/plain_r02g_res.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_render_resident -s 1 -c 1 -f -o gpurun_out/prof_r02g_cbox_resident python tools/prof_render.py --workload cbox --res 2048 --spp 16 --launches 2 > gpurun_out/ncu_r02g_res.log 2>&1
tail -2 gpurun_out/ncu_r02g_res.log
