"""Where does the end-to-end time go?  Wall clock per public call vs the library's own event timings."""
import os, sys, time, io, contextlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from bench import workload_desc, C2_SPP
from prt_b200 import mi_compat as mi, scenes

def probe_acq(name):
    desc, label = workload_desc(name)
    scene = mi.Scene(desc)
    integ = scene.integrator()
    integ.samples_per_element = C2_SPP
    rows = []
    for k in range(8):
        integ.seed = k
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            integ.simulate_acquisition_parallel(scene)
        t1 = time.perf_counter()
        s = float(integ.channel_buf.ravel()[::997].sum())
        t2 = time.perf_counter()
        st = integ.last_stats
        rows.append((1e3 * (t1 - t0), st["total_ms"], st["kernel_ms"], 1e3 * (t2 - t1)))
    print(name, "wall / lib total / lib kernel / checksum ms:", [tuple(round(x, 2) for x in r) for r in rows])

def probe_pt(name, res, spp):
    desc = scenes.cbox_scene(res, spp) if name == "cbox" else scenes.heightfield_scene(2237, (3840, 2160), spp)
    scene = mi.Scene(desc)
    integ = scene.integrator()
    rows = []
    for k in range(5):
        t0 = time.perf_counter()
        img = integ.render(scene, seed=k, spp=spp)
        t1 = time.perf_counter()
        m = float(img[::7, ::7].mean())
        t2 = time.perf_counter()
        st = integ.last_stats
        rows.append((1e3 * (t1 - t0), st["total_ms"], st["kernel_ms"], 1e3 * (t2 - t1)))
    print(name, "wall / lib total / lib kernel / checksum ms:", [tuple(round(x, 2) for x in r) for r in rows])

probe_acq("sphere_box")
probe_acq("sphere_box:intended")
probe_pt("cbox", 2048, 16)
if "--hf" in sys.argv:
    probe_pt("heightfield", 0, 2)
