#!/bin/bash
# r02q: phase-separated warp iterations (PRT_ACQ_DEFER=2) for analytic scenes, A/B
mkdir -p gpurun_out
run() { # name workload lib
  PRT_B200_LIB=$3 timeout 300 python bench.py --workload $2 --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 --also none > gpurun_out/r02q_$1.json 2> gpurun_out/r02q_$1.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02q_$1.json"))
    print("%-34s Mrays/s %6.0f ms %7.3f e2e %6.0f ck %.6g kernel_ms %.3f" % ("$1", d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("device_checksum", 0), d["kernel_ms"]))
except Exception as e:
    print("$1 FAILED", e); print(open("gpurun_out/r02q_$1.err").read()[-800:])
PY
}
L=$PWD/physics-based-ray-tracing_b200/libprt_b200.so
for wl in sphere_box:intended sphere_box plate_box:intended cone_box:intended; do
  run ${wl}_base $wl $L
  run ${wl}_defer2 $wl $PWD/build_variants/defer2.so
done
PRT_B200_LIB=$PWD/build_variants/defer2.so python -m pytest tests -x -q -m gpu -k "acquire or fixtures or parity or edges" 2>&1 | tail -3
