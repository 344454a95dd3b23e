#!/usr/bin/env python
"""GPU box: rays / segments / paths of every steering-angle launch of an acquisition workload at bench size (the ray count
of the launch an ncu capture profiled: profiles/traffic.json needs warp instructions PER RAY)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import workload_desc, C2_SPP
from prt_b200 import mi_compat as mi
out = {}
for wl in sys.argv[1:] or ["sphere_box:intended", "sphere_box", "ring"]:
    desc, label = workload_desc(wl)
    scene = mi.Scene(desc)
    dev = scene.device()
    p = scene.integrator().acq_params(scene)
    dv = torch.device("cuda", 0)
    buf = torch.zeros((p.n_angles, p.n_elements, p.time_samples), dtype=torch.float32, device=dv)
    st = torch.zeros(8, dtype=torch.int64, device=dv)
    rows = []
    for a in range(p.n_angles):
        st.zero_()
        dev.acquire_dev(p, buf.data_ptr(), 0, st.data_ptr(), torch.cuda.current_stream().cuda_stream, seed=1, spp=C2_SPP, angle_first=a, angle_count=1)
        h = st.cpu().numpy()
        rows.append(dict(angle=float(p.angles_deg[a]), paths=int(h[0]), segments=int(h[1]), rays=int(h[2]), deposits=int(h[3])))
    out[wl] = rows
    print(wl, rows)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "per_angle_stats.json"), "w"), indent=1)
