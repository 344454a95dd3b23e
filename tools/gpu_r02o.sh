#!/bin/bash
# r02o: warp-reconvergence fix in k_acquire / k_render_path: acquisition workloads, cbox in mega mode, launch list, full suite
mkdir -p gpurun_out
run() { # name workload
  timeout 300 python bench.py --workload $2 --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 --also none > gpurun_out/r02o_$1.json 2> gpurun_out/r02o_$1.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02o_$1.json"))
    print("%-22s Mrays/s %6.0f ms %7.3f e2e %6.0f ck %.6g kernel_ms %.3f frac %s" % ("$1", d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("device_checksum", 0), d["kernel_ms"], d["roofline"].get("frac")))
except Exception as e:
    print("$1 FAILED", e); print(open("gpurun_out/r02o_$1.err").read()[-800:])
PY
}
for wl in sphere_box:intended sphere_box ring plate_box:intended cone_box:intended sphere_floating:intended; do run $wl $wl; done
PRT_PT_MODE=mega timeout 300 python bench.py --workload cbox --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 --also none > gpurun_out/r02o_cbox_mega.json 2> gpurun_out/r02o_cbox_mega.err
python -c "
import json
d = json.load(open('gpurun_out/r02o_cbox_mega.json')); print('cbox(mega) Mrays/s %.0f ms %.2f' % (d['value'], d['ms_per_step']))"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --also none > gpurun_out/plain_r02o_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02o_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --also none > gpurun_out/ncu_r02o_bench.log 2>&1
python tools/launch_summary.py gpurun_out/launches_r02o_bench.csv | head -3
python -m pytest tests -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r02o_pytest.log
