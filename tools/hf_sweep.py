#!/usr/bin/env python
"""GPU box: one process, the 10 M-triangle height field built once, then the wavefront render under different run-time
knobs (PRT_WF_SORT*, PRT_L2_PERSIST_MB are read at every launch).  Prints kernel ms per configuration and per class."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from prt_b200 import mi_compat as mi, scenes
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=2237)
ap.add_argument("--spp", type=int, default=2)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--configs", default="")
ap.add_argument("--tag", default="hf")
a = ap.parse_args()
desc = scenes.heightfield_scene(a.n, (3840, 2160), a.spp)
scene = mi.Scene(desc)
rp = scene.integrator().render_params(scene)
dev = scene.device()
print("bvh", dev.bvh_stats, flush=True)
default = ["PRT_WF_SORT=0", "PRT_WF_SORT=6", "PRT_WF_SORT=7", "PRT_WF_SORT=8", "PRT_WF_SORT=9", "PRT_WF_SORT=7,PRT_WF_SORT_MODE=1",
           "PRT_WF_SORT=7,PRT_WF_SORT_WHAT=1", "PRT_WF_SORT=7,PRT_WF_SORT_WHAT=2", "PRT_WF_SORT=0,PRT_L2_PERSIST_MB=64",
           "PRT_WF_SORT=0,PRT_L2_PERSIST_MB=100", "PRT_WF_SORT=7,PRT_L2_PERSIST_MB=100"]
configs = [c for c in a.configs.split(";") if c] or default
keys = ("PRT_WF_SORT", "PRT_WF_SORT_MODE", "PRT_WF_SORT_WHAT", "PRT_L2_PERSIST_MB")
ref = None
rows = []
for cfg in configs:
    for k in keys:
        os.environ.pop(k, None)
    for kv in cfg.split(","):
        k, v = kv.split("=")
        os.environ[k] = v
    film, st = dev.render_path(rp, seed=1, spp=a.spp)           # warm-up (allocations)
    ms = []
    dev.ctx.profile_begin()
    for r in range(a.reps):
        film, st = dev.render_path(rp, seed=1, spp=a.spp)
        ms.append(st["kernel_ms"])
    cls = dev.ctx.profile_read()
    img = np.array(film)
    if ref is None:
        ref = img
    same = bool(np.array_equal(img, ref))
    row = dict(config=cfg, kernel_ms=float(np.median(ms)), mrays=st["rays"] / (np.median(ms) * 1e-3) / 1e6, identical_film=same,
               maxdiff=float(np.abs(img - ref).max()), classes={k: round(v["ms"] / a.reps, 3) for k, v in cls.items() if v["launches"]})
    rows.append(row)
    print(json.dumps(row), flush=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", f"{a.tag}_sweep.json"), "w"), indent=1)
