"""Small driver for ncu: a few launches of the path-tracing kernel on cbox or the height field."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from prt_b200 import mi_compat as mi, scenes
ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cbox")
ap.add_argument("--res", type=int, default=1024)
ap.add_argument("--spp", type=int, default=4)
ap.add_argument("--n", type=int, default=2237)
ap.add_argument("--launches", type=int, default=3)
a = ap.parse_args()
desc = scenes.cbox_scene(a.res, a.spp) if a.workload == "cbox" else scenes.heightfield_scene(a.n, (a.res, a.res * 9 // 16), a.spp)
scene = mi.Scene(desc)
rp = scene.integrator().render_params(scene)
dev = scene.device()
print(dev.bvh_stats)
for k in range(a.launches):
    film, st = dev.render_path(rp, seed=k, spp=a.spp)
print(st)
