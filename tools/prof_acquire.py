"""Small driver for ncu: a few launches of the acquisition kernel on one workload."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import workload_desc, C2_SPP
from prt_b200.engine import DeviceScene
from prt_b200.scene import AcqParams
ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="sphere_box")
ap.add_argument("--spp", type=int, default=C2_SPP)
ap.add_argument("--launches", type=int, default=4)
a = ap.parse_args()
desc, label = workload_desc(a.workload)
p = AcqParams.from_props(desc.integrator, desc.sensor)
ds = DeviceScene(desc)
for k in range(a.launches):
    _, _, st = ds.acquire(p, seed=k, spp=a.spp)
print(label, st)
