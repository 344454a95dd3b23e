#!/bin/bash
# r02b: reference-fixture report on CUDA, path-tracer parity with the resident kernel, cbox A/B (wavefront vs resident at 3 occupancies)
mkdir -p gpurun_out
python tools/ref_fixture_report.py > gpurun_out/r02b_report.log 2>&1
python -m pytest tests/test_gpu_path.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r02b_pytest.log
run() { # name env lib
  PRT_PT_MODE=$2 PRT_B200_LIB=$3 timeout 300 python bench.py --workload cbox --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r02b_$1.json 2> gpurun_out/r02b_$1.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02b_$1.json"))
    print("%-14s Mrays/s %6.0f ms %7.2f e2e %6.0f" % ("$1", d["value"], d["ms_per_step"], d["e2e"]["value"]))
except Exception as e:
    print("$1 FAILED", e); print(open("gpurun_out/r02b_$1.err").read()[-800:])
PY
}
run wavefront wavefront $PWD/physics-based-ray-tracing_b200/libprt_b200.so
run mega mega $PWD/physics-based-ray-tracing_b200/libprt_b200.so
run res3 resident $PWD/physics-based-ray-tracing_b200/libprt_b200.so
run res2 resident $PWD/build_variants/res2.so
run res4 resident $PWD/build_variants/res4.so
