#!/usr/bin/env python
"""GPU box: how far the CUDA path is from the reference-Python fixtures (tests/golden/ref_*.npz), as numbers (no asserts).
Writes gpurun_out/ref_fixture_report.json; the bars of tests/test_ref_fixtures.py's -m gpu tests are set from it."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import orc_py  # noqa: E402
import test_ref_fixtures as T  # noqa: E402
import make_ref_fixtures as M  # noqa: E402
from prt_b200.engine import DeviceScene, ultra_bsdf_sample  # noqa: E402

fx = {k: np.load(os.path.join(T.GOLDEN, f"ref_{k}.npz")) for k in ("bsdf", "directivity", "segments")}
scenes = dict(M.fixture_scenes())
out = {}
b = fx["bsdf"]
d, pdf, amp, rf = ultra_bsdf_sample(b["wi"], b["ng"], b["ns"], b["impedance"], b["roughness"], b["s1"], b["s2"])
out["bsdf"] = T._check_bsdf(b, d, pdf, amp, rf, "cuda", None)
seg = fx["segments"]
devs = {}


def trace(desc, p, idx, seed, spp):
    dev = devs.setdefault(id(desc), DeviceScene(desc))
    return dev.acquire_trace(p, idx, seed=seed, spp=spp)


def acquire(desc, p, seed, spp, run):
    dev = devs.setdefault(id(desc), DeviceScene(desc))
    buf, tx, _ = dev.acquire(p, seed=seed, spp=spp, sample_offset=run, sample_stride=spp)
    return np.array(buf), np.array(tx)


for mode in "DP":
    out["segments_" + mode] = T._compare_segments(seg, seg["scene_names"], scenes, mode, T._flags(orc_py, mode), trace, "cuda", None)
    out["buffers_" + mode] = T._compare_buffers(seg, seg["scene_names"], scenes, mode, T._flags(orc_py, mode), acquire, "cuda", None, None)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "ref_fixture_report.json"), "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out)[:3000])
