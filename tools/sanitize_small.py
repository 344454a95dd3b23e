#!/usr/bin/env python
"""Tiny invocations of every kernel family, for `compute-sanitizer --tool memcheck|racecheck|initcheck python tools/sanitize_small.py`."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from prt_b200 import mi_compat as mi, scenes
from prt_b200.engine import DeviceScene, ultra_bsdf_sample, directivity_weights
from prt_b200.scene import AcqParams
from prt_b200.transforms import Transform4f

for name in ("Plate_Box", "Sphere_Box"):
    desc = scenes.ultrasound_scene(name, "intended")
    p = AcqParams.from_props(desc.integrator, desc.sensor)
    dev = DeviceScene(desc)
    buf, tx, st = dev.acquire(p, seed=1, spp=8)
    print(name, st["paths"], st["segments"], float(np.abs(buf).sum()))
ring = scenes.test_ring_scene()
p = AcqParams.from_props(ring.integrator, ring.sensor)
rdev = DeviceScene(ring)
buf, tx, st = rdev.acquire(p, seed=1, spp=64)          # >= 64 samples per pair: the warp-per-(angle, element) lane map
print("ring", st)
rec = rdev.acquire_trace(p, np.arange(64, dtype=np.uint64), seed=1, spp=4)
sc = mi.Scene(ring)
params = mi.traverse(sc)
sc.device()
params["ring.to_world"] = Transform4f().translate([0.001, 0, 0]) @ Transform4f(ring.shapes[0].to_world)
params.update()
o = np.random.default_rng(1).uniform(-0.05, 0.05, (256, 3)).astype(np.float32)
d = np.tile(np.array([[0, 0, 1]], np.float32), (256, 1))
print("refit hits", int((sc.device().trace_closest(o, d)["prim"] >= 0).sum()), int(sc.device().trace_occluded(o, d).sum()))
for mode in ("resident", "wavefront", "mega"):
    os.environ["PRT_PT_MODE"] = mode
    c = mi.Scene(scenes.cbox_scene(32, 4))
    film, fst = c.device().render_path(c.integrator().render_params(c), seed=2, spp=4)
    print("cbox", mode, fst["rays"], float(film[..., :3].sum()))
os.environ["PRT_PT_MODE"] = "wavefront"
for knob in ({}, {"PRT_WF_SORT": "7"}):
    os.environ.update(knob)
    h = mi.Scene(scenes.heightfield_scene(40, (32, 18), 2))
    film, fst = h.device().render_path(h.integrator().render_params(h), seed=3, spp=2)
    print("heightfield", knob, fst["rays"], fst["launches"])
os.environ.pop("PRT_WF_SORT", None)
os.environ.pop("PRT_PT_MODE", None)
d = scenes.usmain_scene_dict()
d["integrator"]["samples_per_element"] = 4
us = mi.load_dict(d)
integ = us.integrator()
lam = integ.sound_speed / integ.frequency
img = integ.render_bmode(us, np.arange(-0.01, 0.01, lam), np.arange(0.001, 0.03, lam / 2))
print("bmode", img.shape, float(img.max()))
n = 64
g = np.random.default_rng(0)
w = g.normal(size=(n, 3)).astype(np.float32)
print("bsdf", ultra_bsdf_sample(w, w, w, 7.8, 0.5, g.random(n), g.random(n))[1][:2])
print("dir", directivity_weights(np.eye(4), w, w, w, 24.0, 30.0, 320.0)[0][:2])
