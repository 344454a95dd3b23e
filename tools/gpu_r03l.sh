#!/bin/bash
# NCCL stream priority A/B at N GPUs (headline + also), driver-style launch
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for hp in 1 0; do
  PRT_NCCL_HIPRIO=$hp timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$hp bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r03l_n${N}_hp$hp.json 2> gpurun_out/r03l_n${N}_hp$hp.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/r03l_n${N}_hp$hp.json") if l.startswith("{")][-1])
    print("hiprio=$hp N=$N Mrays/s %.0f ms %.3f e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
    for k, v in d.get("also", {}).items():
        print("   %-12s %s" % (k, {a: v.get(a) for a in ("value", "e2e", "ms")}))
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/r03l_n${N}_hp$hp.err").read()[-1500:])
PY
done
