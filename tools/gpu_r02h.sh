#!/bin/bash
# r02h: mesh acquisition kernel A/B on the ring: segment-at-a-time (PRT_ACQ_SM=0) vs per-lane state machine (=1), parity of the latter
mkdir -p gpurun_out
run() { # name sm lib
  PRT_ACQ_SM=$2 PRT_B200_LIB=$3 timeout 300 python bench.py --workload ring --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 2 --also none > gpurun_out/r02h_$1.json 2> gpurun_out/r02h_$1.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02h_$1.json"))
    print("%-14s Mrays/s %6.0f ms %7.2f e2e %6.0f seg/path %.3f dep/path %.3f ck %.6g" % ("$1", d["value"], d["ms_per_step"], d["e2e"]["value"], d["segments_per_path"], d["deposits_per_path"], d["device_checksum"]))
except Exception as e:
    print("$1 FAILED", e); print(open("gpurun_out/r02h_$1.err").read()[-800:])
PY
}
L=$PWD/physics-based-ray-tracing_b200/libprt_b200.so
run sm0 0 $L
run sm1 1 $L
for v in smb4 smb16 smn1 smn3; do run $v 1 $PWD/build_variants/$v.so; done
PRT_ACQ_SM=1 python -m pytest tests -x -q -m gpu -k "ring or mesh or acquire or fixtures or smoke or edges" 2>&1 | tail -5 | tee gpurun_out/r02h_pytest.log
