"""CPU: pins the oracle (oracle/orc.c) against tests/golden/anchors.json -- analytic f64 known answers derived
from the reference's formulas by tests/golden/make_anchors.py, plus the external PCG32 KAT.  The reference
ships no golden vectors of its own for this path ("parity unpinned")."""
import json
import math
import os

import numpy as np
import pytest

from prt_b200 import scenes
from prt_b200.scene import AcqParams

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "anchors.json")) as fh:
    A = json.load(fh)

ANGLES = [-15.0, -7.5, 0.0, 7.5, 15.0]


def _ray(a, e):
    th = math.radians(ANGLES[a])
    return [0.00012 * (e - 31.5), 0.0, 0.0], [math.sin(th), 0.0, math.cos(th)]


def test_golden_file_is_reproducible(tmp_path):
    """The committed anchors are exactly what the committed generator writes."""
    import subprocess, sys, shutil
    gen = os.path.join(HERE, "golden", "make_anchors.py")
    work = tmp_path / "golden"
    work.mkdir()
    shutil.copy(gen, work / "make_anchors.py")
    subprocess.run([sys.executable, str(work / "make_anchors.py")], check=True, capture_output=True)
    with open(work / "anchors.json") as fh:
        assert json.load(fh) == A


def test_pcg32_external_kat(orc):
    import ctypes as C
    L = orc.lib()
    st, inc = C.c_uint64(), C.c_uint64()
    L.orc_pcg32_seed(42, 54, C.byref(st), C.byref(inc))
    got = [orc.next_u32(st, inc) for _ in range(6)]
    assert got == A["pcg32_demo_42_54"]


def test_tea_and_path_streams(orc):
    import ctypes as C
    L = orc.lib()
    for key, exp in A["tea32"].items():
        a, b = (int(x) for x in key.split(","))
        o0, o1 = C.c_uint32(), C.c_uint32()
        L.orc_sample_tea_32(a, b, 4, C.byref(o0), C.byref(o1))
        assert [o0.value, o1.value] == exp
    for key, exp in A["path_streams_u32"].items():
        seed, path = (int(x) for x in key.split(","))
        st, inc = orc.path_rng(seed, path)
        assert [orc.next_u32(st, inc) for _ in range(4)] == exp
    st, inc = orc.path_rng(0, 0)
    u = [orc.next_f32(st, inc) for _ in range(1000)]
    assert 0.0 <= min(u) and max(u) < 1.0


@pytest.mark.parametrize("order", ["mitsuba", "intended"])
def test_sphere_primary_hits(orc, order):
    desc = scenes.ultrasound_scene("Sphere_Floating", order)
    sc = orc.OracleScene(desc)
    exp = A[f"sphere_{order}"]
    M = desc.shapes[0].to_world
    assert np.allclose(M[:3, 3], exp["center"], atol=1e-15) and abs(np.linalg.norm(M[:3, 0]) - exp["radius"]) < 1e-15
    for key, t in exp["t"].items():
        a, e = (int(x) for x in key.split(","))
        o, d = _ray(a, e)
        r64 = sc.trace_closest([o], [d], prec=64)
        r32 = sc.trace_closest([o], [d], prec=32)
        assert r64["prim"][0] == 0 and abs(r64["t"][0] - t) <= 1e-12 * t
        assert r32["prim"][0] == 0 and abs(r32["t"][0] - t) <= 1e-5 * t


@pytest.mark.parametrize("order", ["mitsuba", "intended"])
def test_plate_primary_hits(orc, order):
    desc = scenes.ultrasound_scene("Plane_Floating", order)
    sc = orc.OracleScene(desc)
    exp = A[f"plate_{order}"]
    assert np.allclose(desc.shapes[0].to_world, exp["to_world"], atol=1e-15)
    for key, h in exp["hits"].items():
        a, e = (int(x) for x in key.split(","))
        o, d = _ray(a, e)
        for prec, tol in ((64, 1e-12), (32, 1e-5)):
            r = sc.trace_closest([o], [d], prec=prec)
            if h is None:
                assert r["prim"][0] == -1
            else:
                assert r["prim"][0] == 0 and abs(r["t"][0] - h["t"]) <= tol * h["t"]


def test_usmain_plate_hits(orc):
    from prt_b200.scene import load_dict_desc
    desc = load_dict_desc(scenes.usmain_scene_dict())
    sc = orc.OracleScene(desc)
    assert np.allclose(desc.shapes[0].to_world, A["usmain_plate"]["to_world"], atol=1e-15)
    for key, t in A["usmain_plate"]["t"].items():
        a, e = (int(x) for x in key.split(","))
        o, d = _ray(a, e)
        r = sc.trace_closest([o], [d], prec=64)
        assert r["shape"][0] == 0 and abs(r["t"][0] - t) <= 1e-12 * t


def test_tx_delays_and_element_positions(orc):
    desc = scenes.ultrasound_scene("Plane_Floating", "mitsuba")
    p = AcqParams.from_props(desc.integrator, desc.sensor)
    _, tx, st = orc.OracleScene(desc).acquire(p, seed=0, spp=1, prec=64, n_threads=1)
    for key, v in A["tx_delay_c1480"].items():
        a, e = (int(x) for x in key.split(","))
        assert abs(tx[a, e] - v) <= 1e-15 + 1e-12 * abs(v)
    for e, x in A["elem_x"].items():
        assert abs(p.pitch * (int(e) - (p.n_elements - 1) / 2) - x) < 1e-18
    assert st["paths"] == 320


def test_ultra_bsdf_constants(orc):
    u = A["ultra_bsdf"]
    ni = u["normal_incidence"]
    # s1 = 0.5 -> disk centre -> micro-normal along -wi; s2 below / above Ar^2 picks reflect / transmit
    d, pdf, amp, refl = orc.ultra_bsdf(ni["wi"], ni["n"], ni["n"], 7.8, 0.5, ni["s1"], u["p_reflect_normal"] - 1e-6, prec=64)
    assert refl and np.allclose(d, ni["refl_dir"], atol=1e-12) and abs(pdf - ni["pdf_reflect"]) < 1e-12
    assert abs(amp - u["Ar_normal"]) < 1e-12
    d, pdf, amp, refl = orc.ultra_bsdf(ni["wi"], ni["n"], ni["n"], 7.8, 0.5, ni["s1"], u["p_reflect_normal"] + 1e-6, prec=64)
    assert (not refl) and np.allclose(d, ni["trans_dir"], atol=1e-9) and abs(pdf - ni["pdf_trans"]) < 1e-9
    assert abs(amp - u["At_normal"]) < 1e-12
    # total internal reflection beyond asin(1/6.5): a tilted micro-normal cannot be forced, but a tilted wi can
    ang = math.radians(u["tir_angle_deg"] + 1.0)
    wi = [math.sin(ang), 0.0, math.cos(ang)]
    _, _, amp, refl = orc.ultra_bsdf(wi, [0, 0, 1], [0, 0, 1], 7.8, 1e-4, 0.5, 0.999999, prec=64)
    assert refl                                     # sq < 0 forces reflection whatever s2 is (CB:137-145)
    ang = math.radians(u["tir_angle_deg"] - 1.0)
    wi = [math.sin(ang), 0.0, math.cos(ang)]
    _, _, _, refl = orc.ultra_bsdf(wi, [0, 0, 1], [0, 0, 1], 7.8, 1e-4, 0.5, 0.999999, prec=64)
    assert not refl


def test_attenuation_anchor(orc):
    """One segment of exactly known length: atten after the segment = exp(-alpha f 1e-6 d / 8.686)."""
    desc = scenes.ultrasound_scene("Plane_Floating", "intended")
    p = AcqParams.from_props(desc.integrator, desc.sensor, max_depth=1)
    sc = orc.OracleScene(desc)
    rec = sc.acquire_trace(p, [2 * 64 + 31], seed=0, spp=1, prec=64)[0, 0]
    assert rec["valid"] == 1
    per_metre = A["atten_per_metre"]["xml"]
    # recorded atten is post-RR (divided by rr = min(|atten*amp|, 1)) or 0 when killed
    if rec["survive"]:
        rr = min(abs(per_metre ** rec["t"] * rec["amp"]), 1.0)
        assert abs(rec["atten"] * rr - per_metre ** rec["t"]) < 1e-12


def test_disk_quirk_and_ggx_formula():
    from prt_b200 import mi_compat as mi
    for s, exp in A["disk_scalar"].items():
        d = mi.warp.square_to_uniform_disk_concentric(mi.Float(float(s)))
        assert np.allclose(np.asarray(d).reshape(-1), exp, atol=1e-6)
    for key, v in A["ggx_cos_theta"].items():
        xi, al = (float(x) for x in key.split(","))
        assert abs(math.sqrt((1 - xi) / (1 + (al * al - 1) * xi)) - v) < 1e-15


def test_custom_sensor_put_data_vectors():
    """/root/reference/CustomSensor.py:81-96 smoke vectors."""
    from prt_b200 import mi_compat as mi
    from prt_b200.scene import Properties
    from prt_b200 import plugins  # noqa: F401
    import CustomSensor
    g = A["custom_sensor_put_data"]
    s = CustomSensor.CustomSensor(Properties("custom", g["props"]))
    for r in g["rays"]:
        s.put_data(mi.Ray3f(o=r["o"], d=r["d"], time=r["time"]), r["amp"])
    buf = s.channel_data()
    nz = np.argwhere(buf != 0)
    assert len(nz) == len(g["expected_nonzero"])
    for i, k, v in g["expected_nonzero"]:
        assert abs(buf[i, k] - v) < 1e-6


def test_oracle_bvh_equals_bruteforce(orc):
    desc = scenes.test_ring_scene()
    rng = np.random.default_rng(0)
    o = rng.uniform((-0.07, -0.03, -0.02), (0.07, 0.03, 0.15), size=(4000, 3))
    d = rng.normal(size=(4000, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    a = orc.OracleScene(desc, use_bvh=True).trace_closest(o, d, prec=32)
    b = orc.OracleScene(desc, use_bvh=False).trace_closest(o, d, prec=32)
    assert np.array_equal(a["prim"], b["prim"]) and np.array_equal(a["t"], b["t"])


def test_oracle_sharding_and_determinism(orc):
    desc = scenes.ultrasound_scene("Plate_Box", "intended")
    p = AcqParams.from_props(desc.integrator, desc.sensor)
    sc = orc.OracleScene(desc)
    full, _, st = sc.acquire(p, seed=3, spp=8, prec=32, n_threads=2)
    parts = [sc.acquire(p, seed=3, spp=8, s_offset=g, s_stride=2, prec=32, n_threads=1) for g in range(2)]
    assert np.allclose(parts[0][0] + parts[1][0], full, rtol=1e-12, atol=1e-18)
    assert parts[0][2]["paths"] + parts[1][2]["paths"] == st["paths"] == 320 * 8
    again, _, _ = sc.acquire(p, seed=3, spp=8, prec=32, n_threads=4)
    assert np.allclose(again, full, rtol=1e-12, atol=1e-18)


def test_oracle_quirks(orc):
    desc = scenes.ultrasound_scene("Sphere_Floating", "intended")
    p = AcqParams.from_props(desc.integrator, desc.sensor)
    sc = orc.OracleScene(desc)
    _, _, st = sc.acquire(p, seed=1, spp=4, prec=32)
    p1 = AcqParams.from_props(desc.integrator, desc.sensor, quirk_flags=orc.QF_SINGLE_BOUNCE)
    _, _, st1 = sc.acquire(p1, seed=1, spp=4, prec=32)
    assert st1["segments"] == st1["paths"] and st["segments"] > st["paths"]
    # Q1: array inside the sphere under Mitsuba's transform rule -> every connection occluded -> buffer == 0
    dm = scenes.ultrasound_scene("Sphere_Floating", "mitsuba")
    b, _, s = orc.OracleScene(dm).acquire(AcqParams.from_props(dm.integrator, dm.sensor), seed=1, spp=4, prec=32)
    assert s["deposits"] == 0 and np.abs(b).max() == 0
    # ... unless the connection ray is stopped at the receive element
    pf = AcqParams.from_props(dm.integrator, dm.sensor, quirk_flags=orc.QF_CONNECT_TO_TARGET)
    b, _, s = orc.OracleScene(dm).acquire(pf, seed=1, spp=4, prec=32)
    assert s["deposits"] > 0


@pytest.mark.parametrize("name,order", [("Plate_Box", "intended"), ("Sphere_Box", "intended"), ("Plane_Floating", "mitsuba")])
def test_c_oracle_equals_python_transliteration(orc, name, order):
    """oracle/orc.c (C, f64) vs oracle/pyref.py (pure Python, f64, written independently from the reference's
    _trace_single_ray): same injected PCG32 streams -> the same buffer, bin for bin."""
    import pyref
    desc = scenes.ultrasound_scene(name, order)
    p = AcqParams.from_props(desc.integrator, desc.sensor)
    buf_c, tx_c, st_c = orc.OracleScene(desc).acquire(p, seed=4, spp=2, prec=64, n_threads=2)
    buf_p, tx_p, st_p = pyref.acquire(pyref.shapes_from_desc(desc), p, seed=4, spp=2)
    assert st_c["paths"] == st_p["paths"] == 640
    assert st_c["segments"] == st_p["segments"] and st_c["rays"] == st_p["rays"] and st_c["deposits"] == st_p["deposits"]
    assert np.allclose(tx_c, tx_p, rtol=1e-13, atol=0)
    assert np.array_equal(buf_c != 0, buf_p != 0)
    assert np.allclose(buf_c, buf_p, rtol=1e-6, atol=1e-15)


@pytest.mark.parametrize("max_depth,rho,rr_depth", [(1, 0.5, 1000), (2, 0.5, 1000), (6, 0.5, 1000), (6, 0.8, 2), (12, 0.3, 3)])
def test_path_tracer_oracle_furnace_closed_form(max_depth, rho, rr_depth):
    """Closed-form pin of oracle/orc_pt.inl (row a14 has no reference code): in a closed box whose walls all emit 1 and
    reflect with albedo rho, every pixel is sum_{i < max_depth} rho^i -- emitter sampling, BSDF sampling, MIS and Russian
    roulette must cancel to exactly that.  max_depth = 1 is exact (emission only); the others are Monte-Carlo means."""
    import orc_py
    from prt_b200 import mi_compat as mi
    desc = scenes.furnace_scene(16, 64, max_depth, rho, rr_depth)
    sc = mi.Scene(desc)
    rp = sc.integrator().render_params(sc)
    film, st = orc_py.render_path(orc_py.OracleScene(desc), rp, seed=1, spp=64, prec=32)
    img = film[..., :3] / film[..., 3:]
    expect = sum(rho ** i for i in range(max_depth))
    assert st["misses"] == 0
    if max_depth == 1:
        assert np.allclose(img, 1.0, atol=1e-6)
    else:
        assert abs(img.mean() - expect) <= 4e-3 * expect, (img.mean(), expect)
        assert np.abs(img.mean((0, 1)) - expect).max() <= 6e-3 * expect          # per channel


def test_path_tracer_oracle_furnace_with_specular_spheres():
    """A glass and a mirror sphere inside the furnace absorb nothing: radiance stays 1 / (1 - rho) everywhere, on and
    through the spheres too.  Pins the dielectric (Fresnel split, refraction, eta^2 bookkeeping) and conductor BSDFs of the
    oracle's path tracer to a closed form."""
    import orc_py
    from prt_b200 import mi_compat as mi
    desc = scenes.furnace_scene(24, 64, max_depth=40, rho=0.5, rr_depth=5, spheres=True)
    sc = mi.Scene(desc)
    rp = sc.integrator().render_params(sc)
    osc = orc_py.OracleScene(desc)
    film, st = orc_py.render_path(osc, rp, seed=1, spp=64, prec=32)
    img = (film[..., :3] / film[..., 3:]).mean(-1)
    assert st["misses"] == 0
    assert abs(img.mean() - 2.0) <= 0.01
    # which pixels look at a sphere: with max_depth = 1 only directly visible emission counts -- walls 1, spheres 0
    d1 = scenes.furnace_scene(24, 16, max_depth=1, rho=0.5, spheres=True)
    s1 = mi.Scene(d1)
    f1, _ = orc_py.render_path(orc_py.OracleScene(d1), s1.integrator().render_params(s1), seed=2, spp=16, prec=32)
    direct = (f1[..., :3] / f1[..., 3:]).mean(-1)
    on_sphere, on_wall = direct < 0.02, direct > 0.98
    assert on_sphere.sum() > 150 and on_wall.sum() > 200
    assert abs(img[on_sphere].mean() - 2.0) <= 0.04 and abs(img[on_wall].mean() - 2.0) <= 0.01


def test_pulse_shape_oracle_equals_the_prototype_echo_sum():
    """oracle/pyref.pulse_shape (checker of the f4 kernel) against the literal per-echo sum of RayTracingV0.py:193-201."""
    import pyref
    fs, fc, sigma = 50e6, 3e6, 2e-7
    echoes = [(10, 1.0), (200, -2.0), (399, 0.5)]
    ch = np.zeros(400)
    for i, a in echoes:
        ch[i] = a
    t = np.arange(400) / fs
    lit = sum(a * np.sin(2 * np.pi * fc * (t - i / fs)) * np.exp(-((t - i / fs) ** 2) / sigma ** 2) for i, a in echoes)
    assert np.abs(pyref.pulse_shape(ch, fs, fc, sigma) - lit).max() < 2e-7


def test_pinned_pool_never_hands_out_a_live_buffer():
    """engine.Context.pinned_array: a result the caller still references (directly or through a derived view) is
    never reused (host logic; allocator faked, no GPU)."""
    import ctypes as C
    from prt_b200 import engine

    class FakeL:
        def __init__(self):
            self.keep = []

        def prt_host_alloc(self, h, n, out):
            b = (C.c_byte * n)()
            self.keep.append(b)
            out._obj.value = C.addressof(b)
            return 0

    ctx = object.__new__(engine.Context)
    ctx.L, ctx.h = FakeL(), None
    a = ctx.pinned_array((2, 2), np.float32)
    b = ctx.pinned_array((2, 2), np.float32)
    assert a.ctypes.data != b.ctypes.data
    pa = a.ctypes.data
    del a
    c = ctx.pinned_array((2, 2), np.float32)
    assert c.ctypes.data == pa                       # released -> reused
    r = c.reshape(4)[1:]
    del c
    d = ctx.pinned_array((2, 2), np.float32)
    assert d.ctypes.data != pa                       # a derived view keeps it busy
    del r
    assert ctx.pinned_array((2, 2), np.float32).ctypes.data == pa


def test_das_restatement_focuses_a_point_echo():
    """oracle/pyref.py::das_beamform (the numpy restatement of what USMain.py:175-200 asks ultraspy for; the checker of
    prt_das_beamform): channel data holding, per element, one linearly interpolated unit echo at the two-way time of flight
    of a point scatterer must focus AT that pixel, with the closed-form value sum_e ((1 - w_e)^2 + w_e^2) / n_angles."""
    import pyref
    fs, c, pitch, n_e, T = 50e6, 1540.0, 3e-4, 16, 2000
    x0, z0 = 0.0006, 0.0100
    xe = pitch * (np.arange(n_e) - (n_e - 1) / 2)
    ch = np.zeros((1, n_e, T))
    expect = 0.0
    for e in range(n_e):
        s = (z0 / c + np.sqrt((x0 - xe[e]) ** 2 + z0 ** 2) / c) * fs
        i0, w = int(np.floor(s)), s - np.floor(s)
        ch[0, e, i0], ch[0, e, i0 + 1] = 1 - w, w
        expect += (1 - w) ** 2 + w ** 2
    x = x0 + 1e-4 * np.arange(-6, 7)
    z = z0 + 5e-5 * np.arange(-10, 11)
    img = pyref.das_beamform(ch, np.array([0.0]), x, z, fs, c, pitch, 0.0, 0.0)
    ix, iz = np.unravel_index(np.argmax(img), img.shape)
    assert (ix, iz) == (6, 10)
    assert abs(img[6, 10] - expect) < 1e-9
    # an f-number that excludes the outer elements lowers the focus value by exactly their terms
    img2 = pyref.das_beamform(ch, np.array([0.0]), x, z, fs, c, pitch, 0.0, 4.0)
    keep = np.abs(x0 - xe) * 2 * 4.0 <= z0
    assert 0 < keep.sum() < n_e and img2[6, 10] < img[6, 10]
