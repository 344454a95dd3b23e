#!/bin/bash
# r02f: state-machine resident kernel on cbox (REGEN_MIN / occupancy variants), full suite
mkdir -p gpurun_out
run() { # name lib
  PRT_B200_LIB=$2 timeout 300 python bench.py --workload cbox --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 2 --also none > gpurun_out/r02f_$1.json 2> gpurun_out/r02f_$1.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02f_$1.json"))
    print("%-14s Mrays/s %6.0f ms %7.2f e2e %6.0f rays/path %.2f" % ("$1", d["value"], d["ms_per_step"], d["e2e"]["value"], d["rays_per_path"]))
except Exception as e:
    print("$1 FAILED", e); print(open("gpurun_out/r02f_$1.err").read()[-800:])
PY
}
run regen6 $PWD/physics-based-ray-tracing_b200/libprt_b200.so
for v in regen4 regen12 minb2 minb4; do run $v $PWD/build_variants/$v.so; done
python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r02f_pytest.log
