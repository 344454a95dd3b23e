#!/bin/bash
# the default bench line (headline + also) at N GPUs exactly as the driver launches it, plus N=1 for the ratio
N=${1:-2}
mkdir -p gpurun_out
show() {
python - <<PY
import json
try:
    d = json.loads([l for l in open("$1") if l.startswith("{")][-1])
    print("$2 Mrays/s %.0f ms %.3f e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
    for k, v in d.get("also", {}).items():
        print("   %-12s %s" % (k, {a: v.get(a) for a in ("value", "e2e", "ms")} if "error" not in v else v))
except Exception as e:
    print("$2 FAILED", e); print(open("$1".replace(".json", ".err")).read()[-1500:])
PY
}
[ -n "$SKIP_N1" ] || { timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/scale2_n1.json 2> gpurun_out/scale2_n1.err; show gpurun_out/scale2_n1.json N=1; }
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/scale2_n$N.json 2> gpurun_out/scale2_n$N.err; show gpurun_out/scale2_n$N.json N=$N
