#!/bin/bash
# multi-GPU bench exactly as the driver launches it: torchrun, one rank per GPU, NCCL
N=${1:-2}
mkdir -p gpurun_out
for wl in sphere_box cbox heightfield; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --workload $wl > gpurun_out/scale_${wl}_n$N.json 2> gpurun_out/scale_${wl}_n$N.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/scale_${wl}_n$N.json") if l.startswith("{")][-1])
    print("$wl N=$N Mrays/s %.0f ms %.2f e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), "also cbox %.0f" % d["also"]["cbox"]["value"] if "also" in d else "")
except Exception as e:
    print("$wl N=$N FAILED", e); print(open("gpurun_out/scale_${wl}_n$N.err").read()[-1500:])
PY
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 | tail -1 | cut -c1-200
