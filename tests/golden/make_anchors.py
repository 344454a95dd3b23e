#!/usr/bin/env python3
"""Generates tests/golden/anchors.json: the analytic known-answer vectors that pin the oracle.

The reference ships no tests or golden vectors for this path and cannot run here (mitsuba / drjit are
not installable), so these anchors are DERIVED, in binary64 and closed form, from the formulas the
reference states (file:line given per entry) -- independently of oracle/orc.c and of the product's
transforms.py (this script carries its own 4x4 helpers).  One external KAT is included: the PCG32
reference vector (pcg32_srandom(42, 54), O'Neill's pcg32-demo).

Run:  python tests/golden/make_anchors.py   (rewrites anchors.json; committed with its output)
"""
import json
import math
import os

import numpy as np


def translate(v):
    m = np.eye(4); m[:3, 3] = v; return m


def scale(v):
    v = np.broadcast_to(np.asarray(v, dtype=float), (3,)); return np.diag([v[0], v[1], v[2], 1.0])


def rotate(axis, deg):
    a = np.asarray(axis, dtype=float); a = a / np.linalg.norm(a)
    t = math.radians(deg); c, s = math.cos(t), math.sin(t)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    m = np.eye(4); m[:3, :3] = c * np.eye(3) + s * K + (1 - c) * np.outer(a, a); return m


def compose(ops, order):
    m = np.eye(4)
    for op in ops:
        m = op @ m if order == "mitsuba" else m @ op
    return m


def primary_ray(a_deg, e, n_e=64, pitch=0.00012):
    th = math.radians(a_deg)
    xe = pitch * (e - (n_e - 1) / 2)          # CustomIntegrator.py:84
    return np.array([xe, 0.0, 0.0]), np.array([math.sin(th), 0.0, math.cos(th)])


def hit_sphere(o, d, c, r):
    oc = o - c
    A, B, C = d @ d, 2 * (oc @ d), oc @ oc - r * r
    disc = B * B - 4 * A * C
    if disc < 0:
        return None
    q = -0.5 * (B + math.copysign(math.sqrt(disc), B))
    t0, t1 = sorted((q / A, C / q))
    if t1 < 0:
        return None
    return t1 if t0 < 0 else t0


def hit_rect(o, d, M):
    Mi = np.linalg.inv(M)
    ol = Mi[:3, :3] @ o + Mi[:3, 3]
    dl = Mi[:3, :3] @ d
    t = -ol[2] / dl[2]
    x, y = ol[0] + t * dl[0], ol[1] + t * dl[1]
    if t < 0 or abs(x) > 1 or abs(y) > 1:
        return None
    return t, x, y


ANGLES = [-15.0, -7.5, 0.0, 7.5, 15.0]                     # MitsubaScenes/Sphere_Box.xml:14
out = {"_about": "analytic f64 anchors; see make_anchors.py. PARITY UNPINNED: derived, not reference output."}

# element positions / transmit delays -- CustomIntegrator.py:84,87
out["elem_x"] = {str(e): 0.00012 * (e - 31.5) for e in (0, 31, 32, 63)}
out["tx_delay_c1480"] = {f"{a},{e}": (0.00012 * (e - 31.5) * math.sin(math.radians(ANGLES[a]))) / 1480.0
                         for a, e in ((0, 0), (0, 63), (2, 10), (4, 63), (1, 20))}

# sphere primary hits -- Sphere_*.xml:36-45 under both composition rules
pairs = [(2, 31), (2, 32), (0, 0), (4, 63), (1, 20)]
for order in ("mitsuba", "intended"):
    M = compose([translate([0, 0, 0.08]), scale(0.06)], order)
    c, r = M[:3, 3], np.linalg.norm(M[:3, 0])
    out[f"sphere_{order}"] = {"center": c.tolist(), "radius": r,
                              "t": {f"{a},{e}": hit_sphere(*primary_ray(ANGLES[a], e), c, r) for a, e in pairs}}

# plate primary hits -- Plate_Box.xml:36-46
for order in ("mitsuba", "intended"):
    M = compose([translate([0, 0, 0.05]), rotate([0, 1, 0], 45), scale([0.17, 0.17, 0.02])], order)
    res = {}
    for a, e in pairs:
        h = hit_rect(*primary_ray(ANGLES[a], e), M)
        res[f"{a},{e}"] = None if h is None else {"t": h[0], "u": h[1], "v": h[2]}
    out[f"plate_{order}"] = {"to_world": M.tolist(), "hits": res}

# USMain.py:67-71 flat_plate = T(0,0,0.05) @ R_y(45) @ S(.17,.17,.14), angles linspace(-15,15,5)
M = translate([0, 0, 0.05]) @ rotate([0, 1, 0], 45) @ scale([0.17, 0.17, 0.14])
out["usmain_plate"] = {"to_world": M.tolist(),
                       "t": {f"{a},{e}": hit_rect(*primary_ray(ANGLES[a], e), M)[0] for a, e in pairs}}

# UltraBSDF constants -- CustomBSDF.py:105-124,142 with Z = 7.8, medium 1.2
Z1, Z2 = 7.8, 1.2
ratio = Z1 / Z2
Ar0 = (Z1 - Z2) / (Z1 + Z2)
out["ultra_bsdf"] = {"snells_ratio": ratio, "tir_angle_deg": math.degrees(math.asin(1 / ratio)), "Ar_normal": Ar0,
                     "At_normal": 1 - Ar0, "p_reflect_normal": Ar0 * Ar0}
# a fully worked sample at normal incidence on a +z surface, s1 = 0.5 (disk centre): the sampled micro-normal is
# +z, flipped to m = -z by CB:100, so cwm = wi.m = -1, cTr = cTt = 1 and, LITERALLY as CB:130-131 write it,
#   refl  = wi + 2 cwm m              = (0,0,1) + 2(-1)(0,0,-1) = (0,0,3)      (Q8: not a mirror direction)
#   trans = ratio refl + (ratio cTr - cTt) m = 6.5 (0,0,3) + 5.5 (0,0,-1) = (0,0,14)
refl = [0.0, 0.0, 3.0]
trans = [0.0, 0.0, ratio * 3.0 - (ratio - 1.0)]
out["ultra_bsdf"]["normal_incidence"] = {"wi": [0.0, 0.0, 1.0], "n": [0.0, 0.0, 1.0], "s1": 0.5, "refl_dir": refl,
                                         "pdf_reflect": 0.25, "trans_dir": trans,
                                         "pdf_trans": ratio ** 2 * abs(trans[2] * -1.0) / (1.0 * abs(trans[2]))}

# attenuation per metre -- CustomIntegrator.py:162
out["atten_per_metre"] = {"xml": math.exp(-0.1 * 3e6 * 1e-6 / 8.686), "usmain": math.exp(-0.2 * 5e6 * 1e-6 / 8.686)}

# concentric-disk quirk -- CustomBSDF.py:48 with a scalar sample
out["disk_scalar"] = {str(s): [(2 * s - 1) / math.sqrt(2)] * 2 for s in (0.0, 0.25, 0.5, 0.9)}

# GGX polar sampling -- sampling_test.py:18
out["ggx_cos_theta"] = {f"{xi},{al}": math.sqrt((1 - xi) / (1 + (al * al - 1) * xi)) for xi, al in ((0.1, 0.5), (0.5, 0.9), (0.9, 0.2))}


# PCG32 external KAT (pcg32-demo, seed 42 / stream 54) + TEA / float conversion computed with python ints
def pcg32(state, inc):
    old = state
    state = (old * 0x5851f42d4c957f2d + inc) & (2 ** 64 - 1)
    xs = (((old >> 18) ^ old) >> 27) & 0xffffffff
    rot = old >> 59
    return state, ((xs >> rot) | (xs << ((-rot) & 31))) & 0xffffffff


def pcg32_seed(initstate, initseq):
    inc = ((initseq << 1) | 1) & (2 ** 64 - 1)
    state, _ = pcg32(0, inc)
    state = (state + initstate) & (2 ** 64 - 1)
    state, _ = pcg32(state, inc)
    return state, inc


def tea(v0, v1, rounds=4):
    s = 0
    for _ in range(rounds):
        s = (s + 0x9e3779b9) & 0xffffffff
        v0 = (v0 + ((((v1 << 4) & 0xffffffff) + 0xa341316c) ^ (v1 + s) ^ ((v1 >> 5) + 0xc8013ea4))) & 0xffffffff
        v1 = (v1 + ((((v0 << 4) & 0xffffffff) + 0xad90777d) ^ (v0 + s) ^ ((v0 >> 5) + 0x7e95761e))) & 0xffffffff
    return v0, v1


st, inc = pcg32_seed(42, 54)
vals = []
for _ in range(6):
    st, v = pcg32(st, inc)
    vals.append(v)
assert vals == [0xa15c02b7, 0x7b47f409, 0xba1d3330, 0x83d2f293, 0xbfa4784b, 0xcbed606e], [hex(v) for v in vals]
out["pcg32_demo_42_54"] = vals
out["tea32"] = {f"{a},{b}": list(tea(a, b)) for a, b in ((0, 0), (1, 1), (0, 12345), (7, 0xffffffff))}
streams = {}
for seed, path in ((0, 0), (0, 1), (9, 123456), (3, 2 ** 32 + 5)):
    v0, v1 = tea((seed + (path >> 32)) & 0xffffffff, path & 0xffffffff)
    st, inc = pcg32_seed(v0, v1)
    u = []
    for _ in range(4):
        st, v = pcg32(st, inc)
        u.append(v)
    streams[f"{seed},{path}"] = u
out["path_streams_u32"] = streams

# CustomSensor.put_data smoke vectors -- CustomSensor.py:81-96
out["custom_sensor_put_data"] = {"props": {"number_of_elements": 5, "pitch": 1.0, "sample_rate": 10.0, "time_samples": 20},
                                 "rays": [{"o": [-2.0, 0, 0], "d": [0, 0, -1], "time": 1.0, "amp": 1.0},
                                          {"o": [0.0, 0, 0], "d": [0, 0, -1], "time": 1.5, "amp": 2.0},
                                          {"o": [2.0, 0, 0], "d": [0, 0.8, -1], "time": 0.5, "amp": 1.0},
                                          {"o": [10.0, 0, 0], "d": [0, 0, -1], "time": 1.0, "amp": 3.0}],
                                 "expected_nonzero": [[0, 10, 1.0], [2, 15, 2.0], [4, 5, 1.0 / math.sqrt(1.64)]]}

path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "anchors.json")
with open(path, "w") as fh:
    json.dump(out, fh, indent=1, sort_keys=True)
print("wrote", path)
