"""GPU edge cases of the hot path (through the C ABI): empty and minimal scenes, ragged sizes, shards past the end,
degenerate geometry, and the loud-failure paths.  The reference has no tests; these are the cases its
code would meet at the boundary: a scene without shapes (every `scene.ray_intersect` misses, CustomIntegrator.py:146-147),
one element / one angle / one sample, a mesh of a single triangle, zero-area triangles, and rays that start on a surface."""
import copy

import numpy as np
import pytest

from prt_b200 import scenes
from prt_b200.scene import AcqParams, MaterialDesc, SceneDesc, ShapeDesc

pytestmark = pytest.mark.gpu


def _base(name="Plate_Box"):
    desc = scenes.ultrasound_scene(name, "intended")
    return desc, AcqParams.from_props(desc.integrator, desc.sensor)


def _mesh_scene(v, idx, base):
    d = SceneDesc(shapes=[ShapeDesc(kind="mesh", to_world=np.eye(4), material=0, id="m", v=np.asarray(v, dtype=np.float64),
                                    idx=np.asarray(idx, dtype=np.uint32))],
                  materials=[MaterialDesc(kind="ultra", params=np.array([7.8, 0.5, 0, 0, 0, 0, 0, 0.0]), emission=np.zeros(3))],
                  integrator=base.integrator, sensor=base.sensor)
    return d


def test_empty_scene_every_ray_misses(orc):
    from prt_b200.engine import DeviceScene
    base, p = _base()
    empty = SceneDesc(shapes=[], materials=copy.deepcopy(base.materials), integrator=base.integrator, sensor=base.sensor)
    ds = DeviceScene(empty)
    assert ds.bvh_stats["n_triangles"] == 0 and ds.bvh_stats["n_primitives"] == 0
    buf, tx, st = ds.acquire(p, seed=0, spp=7)
    n = p.n_angles * p.n_elements * 7
    assert st["paths"] == n and st["rays"] == n and st["misses"] == n and st["segments"] == 0 and st["deposits"] == 0
    assert not np.asarray(buf).any()
    ob, otx, ost = orc.OracleScene(empty).acquire(p, seed=0, spp=7, prec=32)
    assert ost["paths"] == n and ost["misses"] == n and np.allclose(tx, otx, rtol=1e-6, atol=1e-12)
    g = ds.trace_closest(np.zeros((5, 3), np.float32), np.tile(np.float32([0, 0, 1]), (5, 1)))
    assert (g["prim"] < 0).all() and not ds.trace_occluded(np.zeros((5, 3), np.float32), np.tile(np.float32([0, 0, 1]), (5, 1))).any()


@pytest.mark.parametrize("n_e,n_a,T,spp", [(1, 1, 1, 1), (3, 2, 17, 5), (64, 1, 10000, 1), (257, 3, 999, 2)])
def test_ragged_acquisition_sizes_match_oracle(orc, n_e, n_a, T, spp):
    """Sizes that are not multiples of a warp / CTA / tile, down to a single path and a single time bin."""
    from prt_b200.engine import DeviceScene
    base, _ = _base("Plate_Box")
    p = AcqParams.from_props(base.integrator, base.sensor, n_elements=n_e, angles_deg=np.linspace(-10, 10, n_a) if n_a > 1 else np.array([0.0]),
                             time_samples=T)
    ds, oc = DeviceScene(base), orc.OracleScene(base)
    gb, gtx, gs = ds.acquire(p, seed=5, spp=spp)
    ob, otx, os_ = oc.acquire(p, seed=5, spp=spp, prec=32)
    assert gb.shape == (n_a, n_e, T) and gs["paths"] == os_["paths"] == n_a * n_e * spp
    assert abs(gs["segments"] - os_["segments"]) <= max(2, 0.02 * os_["segments"])
    assert np.allclose(gtx, otx, rtol=1e-6, atol=1e-12)
    scale = max(np.abs(ob).max(), 1e-30)
    # few paths: compare bin by bin where the oracle deposited, allowing the +-1 bin of an fp32 time on a bin edge
    if gs["deposits"] and gs["segments"] == os_["segments"]:
        assert abs(float(np.asarray(gb, dtype=np.float64).sum()) - float(ob.sum())) <= 2e-3 * max(np.abs(ob).sum(), scale)


def test_shard_past_the_end_and_strides(orc):
    from prt_b200.engine import DeviceScene
    base, p = _base()
    ds = DeviceScene(base)
    buf, _, st = ds.acquire(p, seed=1, spp=3, sample_offset=3, sample_stride=4)     # offset == spp: this shard is empty
    assert st["paths"] == 0 and st["rays"] == 0 and not np.asarray(buf).any()
    a, _, sa = ds.acquire(p, seed=1, spp=5, sample_offset=4, sample_stride=1000)     # one sample
    assert sa["paths"] == p.n_angles * p.n_elements
    b, _, sb = orc.OracleScene(base).acquire(p, seed=1, spp=5, s_offset=4, s_stride=1000, prec=32)
    assert sb["paths"] == sa["paths"] and abs(sa["segments"] - sb["segments"]) <= 3


def test_single_and_degenerate_triangles(orc):
    """One triangle (the BVH8 builder's n == 1 branch), two triangles, and zero-area triangles mixed in."""
    from prt_b200.engine import DeviceScene
    base, p = _base("Plane_Floating")
    tri = [[-0.02, -0.02, 0.03], [0.03, -0.02, 0.03], [-0.02, 0.03, 0.035]]
    rng = np.random.default_rng(0)
    o = rng.uniform((-0.03, -0.03, 0.0), (0.03, 0.03, 0.01), size=(4000, 3)).astype(np.float32)
    d = rng.normal(size=(4000, 3)); d[:, 2] = np.abs(d[:, 2]) + 1.0
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    for v, idx in [(tri, [[0, 1, 2]]),
                   (tri + [[0.03, 0.03, 0.04]], [[0, 1, 2], [1, 3, 2]]),
                   (tri + [[0.03, 0.03, 0.04], [0.0, 0.0, 0.02]], [[0, 1, 2], [1, 3, 2], [4, 4, 4], [0, 0, 1], [0, 1, 1]])]:
        desc = _mesh_scene(v, idx, base)
        ds = DeviceScene(desc)
        assert ds.bvh_stats["n_triangles"] == len(idx) and ds.bvh_stats["n_nodes8"] >= 1
        g = ds.trace_closest(o, d)
        c = orc.OracleScene(desc, use_bvh=False).trace_closest(o, d, prec=32)
        assert (g["prim"] >= 0).sum() > 100
        assert ((g["prim"] >= 0) == (c["prim"] >= 0)).mean() > 0.999
        both = (g["prim"] >= 0) & (c["prim"] >= 0)
        assert (g["prim"][both] == c["prim"][both]).mean() > 0.999
        assert np.allclose(g["t"][both], c["t"][both], rtol=2e-5, atol=1e-7)
        assert (g["prim"][both] < 2).all()                      # a zero-area triangle is never reported
        gb, _, gs = ds.acquire(p, seed=2, spp=8)
        ob, _, os_ = orc.OracleScene(desc).acquire(p, seed=2, spp=8, prec=32)
        assert gs["paths"] == os_["paths"] and abs(gs["segments"] - os_["segments"]) <= max(3, 0.01 * os_["segments"])


def test_rays_starting_on_a_surface_and_parallel_to_it():
    from prt_b200.engine import DeviceScene
    base, _ = _base("Plate_Box")
    ds = DeviceScene(base)
    g0 = ds.trace_closest(np.float32([[0, 0, 0]]), np.float32([[0, 0, 1]]))
    assert g0["prim"][0] >= 0
    p = g0["p"][0]
    # from the hit point along the surface: no self-hit at t = 0 with a parallel direction (NaN-free slab / plane tests)
    t1 = np.cross(g0["ng"][0], [0.3, 0.5, 0.8]); t1 /= np.linalg.norm(t1)
    g1 = ds.trace_closest(p[None].astype(np.float32), t1[None].astype(np.float32))
    assert np.isfinite(g1["t"]).all() or (g1["prim"] < 0).all()
    assert (g1["prim"] != g0["prim"]).all() or (g1["t"] > 1e-6).all()
    # zero-length and non-finite directions never hang or crash the query kernels
    bad = np.float32([[0, 0, 0], [np.nan, 0, 1], [np.inf, 0, 0]])
    gb = ds.trace_closest(np.zeros((3, 3), np.float32), bad)
    assert gb["prim"].shape == (3,)


def test_loud_failures():
    import ctypes as C
    from prt_b200 import capi
    from prt_b200.capi import PrtError
    from prt_b200.engine import DeviceScene, das_beamform, pulse_shape
    base, p = _base()
    ds = DeviceScene(base)
    with pytest.raises(PrtError):
        ds.acquire(p, seed=0, spp=0)                                        # spp_total must be > 0
    bad = copy.copy(p)
    bad.time_samples = 0
    with pytest.raises(PrtError):
        ds.acquire(bad, seed=0, spp=1)
    with pytest.raises(PrtError):
        ds.set_material_param(99, 0, 1.0)
    with pytest.raises(PrtError):                                            # not a light-transport scene
        from prt_b200 import mi_compat as mi
        sc = mi.Scene(base)
        rp = capi.RenderParamsC()
        rp.width = rp.height = 8
        rp.fov_deg, rp.max_depth, rp.rr_depth = 40.0, 3, 5
        for i, v in enumerate(np.eye(4).reshape(16)):
            rp.to_world[i] = v
        sc.device().render_path(rp, seed=0, spp=1)
    with pytest.raises(PrtError):
        pulse_shape(np.zeros((2, 100), np.float32), 50e6, 3e6, sigma_s=1e-3)   # pulse longer than the kernel supports
    with pytest.raises((PrtError, ValueError)):
        das_beamform(np.zeros((2, 4, 100), np.float32), [0.0], np.linspace(-1e-3, 1e-3, 4), np.linspace(1e-3, 2e-3, 4), 50e6, 1540.0, 3e-4)
    L = capi.load()
    assert L.prt_scene_commit(None, None) != 0 and L.prt_last_error()


@pytest.mark.parametrize("which", ["ring", "heightfield", "heightfield_wall", "analytic"])
def test_transform_update_refits_instead_of_rebuilding(orc, which, monkeypatch):
    """prt_scene_set_shape_transform (params['<shape>.to_world'] = T; params.update()): the moved scene must answer ray
    queries like a scene BUILT with the new transform (same primitive ids, t within 1e-5) and like the oracle; for meshes
    the update is a refit of the kept topology (same node count, no re-sort), and moving the shape back restores the
    original answers."""
    from prt_b200 import mi_compat as mi
    from prt_b200.engine import DeviceScene
    from prt_b200.transforms import Transform4f
    import copy
    if which == "ring":
        desc, sid = scenes.test_ring_scene(), "ring"
    elif which == "heightfield":
        desc, sid = scenes.heightfield_scene(120, (32, 18), 1), None
    elif which == "heightfield_wall":       # the moved shape consists of OVERSIZED triangles (outside the LBVH, under the super root)
        desc, sid = scenes.heightfield_scene(120, (32, 18), 1), "back"
    else:
        desc, sid = scenes.ultrasound_scene("Sphere_Box", "intended"), None
    scene = mi.Scene(desc)
    dev = scene.device()
    si = desc.shape_index(sid) if sid else 0
    shape = desc.shapes[si]
    old = np.array(shape.to_world, dtype=np.float64)
    move = (Transform4f().translate([0.004, -0.003, 0.006]) @ Transform4f().rotate([0.3, 1.0, 0.2], 17.0)).matrix
    new = move @ old
    rng = np.random.default_rng(9)
    lo, hi = (np.array(dev.bvh_stats["scene_lo"]), np.array(dev.bvh_stats["scene_hi"])) if desc.n_triangles() else (np.full(3, -0.1), np.full(3, 0.2))
    ext = hi - lo
    o = rng.uniform(lo - 0.2 * ext, hi + 0.2 * ext, size=(20000, 3)).astype(np.float32)
    d = rng.normal(size=(20000, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    d = d.astype(np.float32)
    before = dev.trace_closest(o, d)
    nodes_before = dev.bvh_stats["n_nodes"]
    params = mi.traverse(scene)
    params[f"{shape.id}.to_world"] = Transform4f(new)
    params.update()
    moved = dev.trace_closest(o, d)
    desc2 = copy.deepcopy(desc)
    desc2.shapes[si].to_world = new
    fresh = DeviceScene(desc2).trace_closest(o, d)
    orac = orc.OracleScene(desc2).trace_closest(o, d, prec=32)
    assert (moved["prim"] != before["prim"]).mean() > 0.01            # the move changed the answers
    for ref, bar in ((fresh, 0.9995), (orac, 0.998)):
        same = moved["prim"] == ref["prim"]
        assert same.mean() >= bar, same.mean()
        hit = same & (moved["prim"] >= 0)
        cosv = np.abs(np.sum(d.astype(np.float64) * np.asarray(ref["ng"], dtype=np.float64), axis=1))
        ok = hit & (cosv > 0.05)
        assert np.all(np.abs(moved["t"][ok].astype(np.float64) - ref["t"][ok]) <= 2e-5 * np.abs(ref["t"][ok]) + 5e-7)
        assert np.allclose(moved["ns"][ok], ref["ns"][ok], atol=2e-4)
    assert (dev.trace_occluded(o, d) == (moved["prim"] >= 0)).mean() > 0.999
    if desc.n_triangles():
        assert dev.bvh_stats["n_nodes"] == nodes_before                    # same topology: refitted, not rebuilt
    if which.startswith("heightfield"):
        # the 8-wide tree (re-derived from the refitted binary tree, the oversized box triangles under its super root) must
        # give the wavefront path tracer the image a freshly built scene gives it
        assert dev.bvh_stats["n_oversized"] >= 10       # walls and ceiling (the light quad is small at this size)
        monkeypatch.setenv("PRT_PT_MODE", "wavefront")
        rp = scene.integrator().render_params(scene)
        f_moved, st_moved = dev.render_path(rp, seed=3, spp=4)
        f_fresh, st_fresh = DeviceScene(desc2).render_path(rp, seed=3, spp=4)
        assert st_moved["paths"] == st_fresh["paths"] and abs(st_moved["segments"] - st_fresh["segments"]) <= 2e-3 * st_fresh["segments"]
        a, b = np.array(f_moved)[..., :3], np.array(f_fresh)[..., :3]
        assert np.mean((a - b) ** 2) <= 1e-3 * np.mean(b ** 2)
    params[f"{shape.id}.to_world"] = Transform4f(old)
    params.update()
    back = dev.trace_closest(o, d)
    assert (back["prim"] == before["prim"]).mean() > 0.9995
