#!/bin/bash
# r02s: resident kernel with phase-separated iterations (shadow rays parked in a per-warp stash), A/B on cbox + parity
mkdir -p gpurun_out
run() { # name lib
  PRT_B200_LIB=$2 timeout 300 python bench.py --workload cbox --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 2 --also none > gpurun_out/r02s_$1.json 2> gpurun_out/r02s_$1.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02s_$1.json"))
    print("%-14s Mrays/s %6.0f ms %7.2f e2e %6.0f rays/path %.3f ck %.5g" % ("$1", d["value"], d["ms_per_step"], d["e2e"]["value"], d["rays_per_path"], d["e2e"]["host_checksum"]))
except Exception as e:
    print("$1 FAILED", e); print(open("gpurun_out/r02s_$1.err").read()[-800:])
PY
}
run base $PWD/physics-based-ray-tracing_b200/libprt_b200.so
for v in stash stash_r12 stash_m2; do run $v $PWD/build_variants/$v.so; done
PRT_B200_LIB=$PWD/build_variants/stash.so python -m pytest tests/test_gpu_path.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -4
PRT_B200_LIB=$PWD/build_variants/stash.so python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
