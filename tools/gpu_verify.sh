#!/bin/bash
# what the driver runs at round end, on the shipped in-tree build: GPU suite, smoke(), default bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-verify}
python -m pytest tests -q -m gpu 2>&1 | tail -3 | tee gpurun_out/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/${TAG}_bench.json") if l.startswith("{")][-1])
print("headline %.0f Mrays/s  e2e %.0f  ms %.3f  frac %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"]))
for k, v in d["also"].items(): print("  %-12s %s" % (k, {a: v.get(a) for a in ("value", "e2e", "ms", "frac")}))
PY
