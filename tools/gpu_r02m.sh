#!/bin/bash
# r02m: echo cache (shared-memory combining of deposits) A/B on the acquisition workloads; full suite
mkdir -p gpurun_out
run() { # name workload lib
  PRT_B200_LIB=$3 timeout 300 python bench.py --workload $2 --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 --also none > gpurun_out/r02m_$1.json 2> gpurun_out/r02m_$1.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02m_$1.json"))
    print("%-22s Mrays/s %6.0f ms %7.3f e2e %6.0f dep/path %.3f ck %.6g kernel_ms %.3f" % ("$1", d["value"], d["ms_per_step"], d["e2e"]["value"], d["deposits_per_path"], d["device_checksum"], d["kernel_ms"]))
except Exception as e:
    print("$1 FAILED", e); print(open("gpurun_out/r02m_$1.err").read()[-800:])
PY
}
L=$PWD/physics-based-ray-tracing_b200/libprt_b200.so
for wl in sphere_box:intended ring plate_box:intended cone_box:intended; do
  run ${wl}_cache2k $wl $L
  run ${wl}_cache0 $wl $PWD/build_variants/cache0.so
  run ${wl}_cache1k $wl $PWD/build_variants/cache1k.so
done
python -m pytest tests -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r02m_pytest.log
