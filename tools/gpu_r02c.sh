#!/bin/bash
# r02c: height-field sweep over the ray-sort / L2-window knobs (one process), the smem-top-level builds, ncu of the resident kernel
mkdir -p gpurun_out
timeout 900 python tools/hf_sweep.py --tag r02c_hf > gpurun_out/r02c_hf.log 2>&1; grep -E '^\{' gpurun_out/r02c_hf.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('%-44s %8.2f ms %7.0f Mrays/s same=%s %s' % (d['config'], d['kernel_ms'], d['mrays'], d['identical_film'], d['classes']))"
for v in top9 top73; do
  PRT_B200_LIB=$PWD/build_variants/$v.so timeout 600 python tools/hf_sweep.py --tag r02c_$v --configs "PRT_WF_SORT=0;PRT_WF_SORT=7" > gpurun_out/r02c_$v.log 2>&1
  grep -E '^\{' gpurun_out/r02c_$v.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('$v %-38s %8.2f ms %7.0f Mrays/s same=%s' % (d['config'], d['kernel_ms'], d['mrays'], d['identical_film']))"
done
python tools/prof_render.py --workload cbox --res 2048 --spp 16 --launches 2 > gpurun_out/plain_r02c_res.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_render_resident -s 1 -c 1 -f -o gpurun_out/prof_r02c_cbox_resident python tools/prof_render.py --workload cbox --res 2048 --spp 16 --launches 2 > gpurun_out/ncu_r02c_res.log 2>&1
tail -2 gpurun_out/ncu_r02c_res.log
