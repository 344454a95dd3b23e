#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for so in physics-based-ray-tracing_b200/libprt_b200.so build_variants/bvh2mega.so; do
  PRT_B200_LIB=$PWD/$so python bench.py --workload ring --steps 5 --no-cpu-baseline --no-also > gpurun_out/ab_ring_$(basename $so .so).json 2>gpurun_out/ab_ring.err
  PRT_PT_MODE=mega PRT_B200_LIB=$PWD/$so python bench.py --workload cbox --steps 3 --no-cpu-baseline > gpurun_out/ab_cboxmega_$(basename $so .so).json 2>gpurun_out/ab_cbox.err
  python - <<PY
import json
for w in ("ring","cboxmega"):
    d=json.load(open("gpurun_out/ab_%s_$(basename $so .so).json" % w)); print("$so", w, round(d["value"]), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]))
PY
done
python bench.py --workload cbox --steps 5 --no-cpu-baseline > gpurun_out/bench_r01d_cbox.json 2>gpurun_out/bench_r01d_cbox.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r01d_cbox.json')); print('cbox', round(d['value']), 'e2e', round(d['e2e']['value']))"
sed -i 's/^run hf_shade/#run hf_shade/; s/^run cbox_shade/#run cbox_shade/; s/^run sphere_box/#run sphere_box/; s/^run ring/#run ring/' tools/gpu_ncu_full.sh
bash tools/gpu_ncu_full.sh r01d
