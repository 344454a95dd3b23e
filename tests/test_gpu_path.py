"""GPU parity for SURVEY.md section 8 row a14: the `path` integrator on scenes/cbox.xml's scene
(prt_render_path through the C ABI) vs oracle/orc_pt.inl on the same seeds."""
import numpy as np
import pytest

from prt_b200.scene import load_dict_desc
from prt_b200 import mi_compat as mi
from prt_b200 import scenes

pytestmark = pytest.mark.gpu


def _rel_mse(a, b):
    a, b = a.astype(np.float64), b.astype(np.float64)
    return float(np.mean((a - b) ** 2) / np.mean(b * b))


def _image(film):
    return film[..., :3] / np.maximum(film[..., 3:], 1e-30)


def test_cbox_matches_oracle_same_seeds(orc):
    """Same PCG32 streams -> the same paths: per-channel relMSE far inside the 1e-3 bar (north star)."""
    desc = scenes.cbox_scene(64, 64)
    scene = mi.Scene(desc)
    rp = scene.integrator().render_params(scene)
    film, st = scene.device().render_path(rp, seed=3, spp=64)
    ref, rst = orc.render_path(orc.OracleScene(desc), rp, seed=3, spp=64, prec=32)
    assert st["paths"] == rst["paths"] == 64 * 64 * 64
    assert abs(st["segments"] - rst["segments"]) <= 2e-3 * rst["segments"]
    assert np.allclose(film[..., 3], ref[..., 3], rtol=1e-4, atol=1e-4)            # filter weights
    gi, ci = _image(film), _image(ref)
    for ch in range(3):
        assert _rel_mse(gi[..., ch], ci[..., ch]) < 1e-3
    # and the bulk of pixels agree tightly (a flipped Fresnel / RR decision changes a whole path)
    close = np.abs(gi - ci) <= 1e-3 * np.abs(ci) + 1e-6
    assert close.mean() > 0.98


def test_cbox_converged_images_agree(orc):
    """Independent seeds, >= 1024 spp: GPU image vs oracle image within relMSE 1e-3 per channel."""
    desc = scenes.cbox_scene(32, 1024)
    scene = mi.Scene(desc)
    rp = scene.integrator().render_params(scene)
    film, _ = scene.device().render_path(rp, seed=11, spp=2048)
    ref, _ = orc.render_path(orc.OracleScene(desc), rp, seed=5, spp=1024, prec=32)
    gi, ci = _image(film), _image(ref)
    for ch in range(3):
        assert _rel_mse(gi[..., ch], ci[..., ch]) < 1e-3 * 4      # noise of 1024 + 2048 spp dominates; see test above


def test_directly_visible_emitter_is_exact():
    """Pixels that look straight at the luminaire see exactly its radiance (depth-0 emission, MIS weight 1)."""
    d = scenes.cbox_scene_dict(64, 16)
    from prt_b200.transforms import Transform4f
    d["sensor"]["to_world"] = Transform4f().look_at([0, 0.2, 0], [0, 1, 0], [0, 0, 1])   # outside both spheres
    d["sensor"]["fov"] = 10.0
    d["integrator"]["max_depth"] = 1
    scene = mi.load_dict(d)
    img = scene.integrator().render(scene, seed=0, spp=16)
    assert np.allclose(img[8:-8, 8:-8], 1.0, atol=1e-5)


@pytest.mark.parametrize("max_depth,rho,rr_depth", [(1, 0.5, 1000), (2, 0.5, 1000), (6, 0.5, 1000), (6, 0.8, 2), (12, 0.3, 3)])
def test_furnace_closed_form(max_depth, rho, rr_depth, monkeypatch):
    """Both GPU back ends against the closed form sum_{i < max_depth} rho^i of the all-emitting closed box
    (scenes.furnace_scene): pins emitter hits + MIS, NEE, BSDF sampling and Russian roulette without any reference."""
    desc = scenes.furnace_scene(32, 256, max_depth, rho, rr_depth)
    scene = mi.Scene(desc)
    rp = scene.integrator().render_params(scene)
    expect = sum(rho ** i for i in range(max_depth))
    for mode in ("wavefront", "mega"):
        film, st = _render_mode(scene.device(), rp, mode, monkeypatch, seed=3, spp=256)
        img = _image(film)
        assert st["rays"] - st["shadow_rays"] == st["segments"]           # closed box: nothing escapes
        if max_depth == 1:
            assert np.allclose(img, 1.0, atol=1e-6)
        else:
            assert abs(img.mean() - expect) <= 1.5e-3 * expect, (mode, img.mean(), expect)
            assert np.abs(img - expect).max() <= 0.12 * expect                # every pixel, 256 spp


def test_furnace_with_specular_spheres(monkeypatch):
    """Glass + mirror spheres inside the furnace absorb nothing: 1 / (1 - rho) everywhere, also on and through the spheres
    (pins the dielectric and conductor shading of both back ends to a closed form; three shading queues in play)."""
    desc = scenes.furnace_scene(48, 256, max_depth=40, rho=0.5, rr_depth=5, spheres=True)
    scene = mi.Scene(desc)
    rp = scene.integrator().render_params(scene)
    d1 = scenes.furnace_scene(48, 16, max_depth=1, rho=0.5, spheres=True)
    s1 = mi.Scene(d1)
    f1, _ = s1.device().render_path(s1.integrator().render_params(s1), seed=2, spp=16)
    direct = _image(f1).mean(-1)
    on_sphere, on_wall = direct < 0.02, direct > 0.98
    assert on_sphere.sum() > 600 and on_wall.sum() > 800
    for mode in ("wavefront", "mega"):
        film, st = _render_mode(scene.device(), rp, mode, monkeypatch, seed=3, spp=256)
        img = _image(film).mean(-1)
        assert st["rays"] - st["shadow_rays"] == st["segments"]
        assert abs(img[on_wall].mean() - 2.0) <= 0.004, (mode, img[on_wall].mean())
        assert abs(img[on_sphere].mean() - 2.0) <= 0.015, (mode, img[on_sphere].mean())


def test_render_sharded_equals_unsharded():
    desc = scenes.cbox_scene(48, 32)
    scene = mi.Scene(desc)
    rp = scene.integrator().render_params(scene)
    dev = scene.device()
    full, st = dev.render_path(rp, seed=2, spp=32)
    parts = [dev.render_path(rp, seed=2, spp=32, sample_offset=g, sample_stride=4) for g in range(4)]
    acc = sum(f.astype(np.float64) for f, _ in parts)
    assert sum(s["paths"] for _, s in parts) == st["paths"] == 48 * 48 * 32
    assert sum(s["rays"] for _, s in parts) == st["rays"]
    assert np.allclose(acc, full, rtol=2e-4, atol=1e-5)


def test_render_api_through_plugins():
    """mi.load_dict(...) -> mi.render(scene): the public call a Mitsuba user makes."""
    scene = mi.load_dict(scenes.cbox_scene_dict(32, 8))
    img = mi.render(scene, spp=8, seed=1)
    assert img.shape == (32, 32, 3) and np.isfinite(img).all() and img.max() > 0.5 and img.mean() > 1e-3
    # red wall is on the +x side = image right (camera looks down -z with +x to its left... see sensor frame)
    left, right = img[10:22, 1:4].mean((0, 1)), img[10:22, -4:-1].mean((0, 1))
    assert (left[1] > left[0]) != (right[1] > right[0])          # one side is green-ish, the other red-ish
    # the image is developed on the device (prt_render_image): same as dividing the RGBW film on the host
    rp = scene.integrator().render_params(scene)
    film, _ = scene.device().render_path(rp, seed=1, spp=8)
    w = film[..., 3:4]
    host = np.where(w > 0, film[..., :3] / np.maximum(w, 1e-30), 0.0)
    assert np.allclose(img, host, rtol=2e-6, atol=1e-7)


def test_heightfield_lbvh_against_oracle(orc):
    """A 79 214-triangle height field in a box (the small sibling of BASELINE config 5): GPU LBVH hits vs the
    oracle's own BVH on random rays, and the 8-bounce diffuse image on the same seeds."""
    desc = scenes.heightfield_scene(200, (64, 36), 8)
    scene = mi.Scene(desc)
    dev = scene.device()
    st = dev.bvh_stats
    assert st["n_triangles"] == 2 * 199 * 199 + 12 and st["n_oversized"] == 12
    assert st["n_nodes"] == st["n_triangles"] - st["n_oversized"] - 1
    rng = np.random.default_rng(4)
    o = rng.uniform((-0.9, -0.8, -0.9), (0.9, 0.9, 0.9), size=(40000, 3)).astype(np.float32)
    d = rng.normal(size=(40000, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    d = d.astype(np.float32)
    g = dev.trace_closest(o, d)
    osc = orc.OracleScene(desc)
    c = osc.trace_closest(o, d, prec=32)
    c64 = osc.trace_closest(o, d, prec=64)
    assert (g["prim"] >= 0).all()                                     # closed box: no ray escapes (watertight)
    same = g["prim"] == c["prim"]
    assert same.mean() > 0.998
    cosv = np.abs(np.sum(d.astype(np.float64) * c64["ng"], axis=1))
    ok = same & (c64["prim"] == c["prim"]) & (cosv > 0.05)
    err = np.abs(g["t"][ok].astype(np.float64) - c64["t"][ok])
    assert (err <= 1e-5 * c64["t"][ok] + 5e-7).all()
    rp = scene.integrator().render_params(scene)
    film, fst = dev.render_path(rp, seed=1, spp=8)
    ref, rst = orc.render_path(osc, rp, seed=1, spp=8, prec=32)
    assert fst["paths"] == rst["paths"] and abs(fst["segments"] - rst["segments"]) <= 2e-3 * rst["segments"]
    gi, ci = _image(film), _image(ref)
    assert _rel_mse(gi, ci) < 1e-3


@pytest.mark.parametrize("knob", ["no_big_tris", "ray_sort"])
def test_heightfield_build_and_scheduling_knobs_do_not_change_hits(orc, monkeypatch, knob):
    """Two knobs of the big-scene pipeline must find the same hits as the default: PRT_BIG_TRIS=0 puts the 12 oversized box
    triangles back INTO the LBVH (by default they stay out of it, DScene::n_small: leaf children of the 8-wide tree's super
    root, tested ahead of the tree by the binary traversal); PRT_WF_SORT=7 (off by default, profiles/r02_summary.md) traces
    every bounce's rays in (origin cell, octant) order through a permutation of the ray queue."""
    if knob == "no_big_tris":
        monkeypatch.setenv("PRT_BIG_TRIS", "0")
    desc = scenes.heightfield_scene(200, (64, 36), 8)
    scene = mi.Scene(desc)
    dev = scene.device()
    st = dev.bvh_stats
    assert st["n_oversized"] == (0 if knob == "no_big_tris" else 12)
    assert st["n_nodes"] == st["n_triangles"] - st["n_oversized"] - 1
    rng = np.random.default_rng(5)
    o = rng.uniform((-0.9, -0.8, -0.9), (0.9, 0.9, 0.9), size=(20000, 3)).astype(np.float32)
    d = rng.normal(size=(20000, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    d = d.astype(np.float32)
    g = dev.trace_closest(o, d)
    osc = orc.OracleScene(desc)
    c = osc.trace_closest(o, d, prec=32)
    assert (g["prim"] >= 0).all() and (g["prim"] == c["prim"]).mean() > 0.998
    assert (~dev.trace_occluded(o, d, 1e-4)).all()
    rp = scene.integrator().render_params(scene)
    ref, rst = orc.render_path(osc, rp, seed=1, spp=8, prec=32)
    for mode in ("wavefront", "mega"):
        monkeypatch.setenv("PRT_PT_MODE", mode)
        if knob == "ray_sort":
            monkeypatch.setenv("PRT_WF_SORT", "7")
        film, fst = dev.render_path(rp, seed=1, spp=8)
        assert fst["paths"] == rst["paths"] and abs(fst["segments"] - rst["segments"]) <= 2e-3 * rst["segments"]
        assert _rel_mse(_image(film), _image(ref)) < 1e-3


@pytest.mark.parametrize("n_baffles", [3, 9, 17, 30])
def test_super_root_groups_of_oversized_triangles(orc, monkeypatch, n_baffles):
    """The 8-wide tree's super root holds up to 21 oversized triangles as 7 leaf children of ceil(n / 7) triangles
    (bvh8_write_super_root).  An open height field under its light quad (2 triangles) with 3 / 9 / 17 / 30 large baffle
    triangles gives 5 / 11 / 19 oversized triangles -- leaf children of 1, 2 and 3 -- and the cap of 21 (the rest stay in the
    LBVH).  The wavefront path tracer (8-wide tree) must give the film of the tile megakernel (binary tree + the oversized
    triangles tested ahead of it) and agree with the oracle."""
    d = scenes.heightfield_scene_dict(300, (48, 27), 4)
    for wall in ("ceiling", "back", "left", "right", "front"):
        del d[wall]
    rng = np.random.default_rng(11)
    for k in range(n_baffles):          # big thin triangles hanging over the field, crossing each other; box area >= 0.81
        c = rng.uniform((-0.6, -0.5, -0.6), (0.6, 0.3, 0.6))
        t = rng.uniform(-0.15, 0.15, size=3)
        v = np.array([[-0.45, t[0], -0.45], [0.45, t[1], -0.45], [0.0, t[2], 0.45]])
        v = np.roll(v, k % 3, axis=1) + c
        d[f"baffle{k}"] = {"type": "mesh", "vertices": v.astype(np.float64), "normals": None,
                           "faces": np.array([[0, 1, 2]], dtype=np.uint32), "bsdf": {"type": "ref", "id": "grey"}}
    desc = load_dict_desc(d)
    scene = mi.Scene(desc)
    dev = scene.device()
    st = dev.bvh_stats
    assert st["n_oversized"] == min(n_baffles + 2, 21)
    assert st["n_nodes"] == st["n_triangles"] - st["n_oversized"] - 1
    o = rng.uniform((-0.9, -0.8, -0.9), (0.9, 0.9, 0.9), size=(20000, 3)).astype(np.float32)
    dd = rng.normal(size=(20000, 3)); dd /= np.linalg.norm(dd, axis=1, keepdims=True)
    dd = dd.astype(np.float32)
    osc = orc.OracleScene(desc)
    g, c = dev.trace_closest(o, dd), osc.trace_closest(o, dd, prec=32)
    assert (g["prim"] == c["prim"]).mean() > 0.998 and 0.1 < (g["prim"] >= 0).mean() < 1.0
    rp = scene.integrator().render_params(scene)
    a, sa = _render_mode(dev, rp, "mega", monkeypatch, seed=5, spp=4)
    b, sb = _render_mode(dev, rp, "wavefront", monkeypatch, seed=5, spp=4)
    for k in ("paths", "segments", "rays", "shadow_rays"):
        assert sa[k] == sb[k], k
    assert np.allclose(a, b, rtol=1e-5, atol=1e-6)
    ref, rst = orc.render_path(osc, rp, seed=5, spp=4, prec=32)
    assert sb["paths"] == rst["paths"] and abs(sb["segments"] - rst["segments"]) <= 2e-3 * rst["segments"]
    assert _image(ref).mean() > 1e-5                                   # the baffles must not black the image out
    assert _rel_mse(_image(b), _image(ref)) < 1e-3


def _render_mode(dev, rp, mode, monkeypatch, batch=None, **kw):
    monkeypatch.setenv("PRT_PT_MODE", mode)
    if batch is None:
        monkeypatch.delenv("PRT_WF_BATCH", raising=False)
    else:
        monkeypatch.setenv("PRT_WF_BATCH", str(batch))
    return dev.render_path(rp, **kw)


@pytest.mark.parametrize("which", ["cbox", "heightfield"])
def test_wavefront_equals_megakernel(which, monkeypatch):
    """The wavefront pipeline (queues, dynamic ray fetch, per-material shading kernels) and the tile megakernel run
    the same shading code on the same PCG32 streams: identical path/segment/ray counts, and films equal up to the
    order of the float adds inside a tile."""
    desc = scenes.cbox_scene(80, 16) if which == "cbox" else scenes.heightfield_scene(120, (96, 54), 4)
    scene = mi.Scene(desc)
    dev = scene.device()
    rp = scene.integrator().render_params(scene)
    spp = 16 if which == "cbox" else 4
    a, sa = _render_mode(dev, rp, "mega", monkeypatch, seed=9, spp=spp)
    b, sb = _render_mode(dev, rp, "wavefront", monkeypatch, seed=9, spp=spp)
    for k in ("paths", "segments", "rays", "shadow_rays"):
        assert sa[k] == sb[k], k
    assert sb["launches"] > 4
    assert np.allclose(a, b, rtol=1e-5, atol=1e-6)
    # several batches (forced small) and sample shards give the same film again
    c, sc_ = _render_mode(dev, rp, "wavefront", monkeypatch, batch=1, seed=9, spp=spp)
    assert sc_["rays"] == sa["rays"] and sc_["launches"] > sb["launches"]
    assert np.allclose(a, c, rtol=1e-5, atol=1e-6)
    parts = [_render_mode(dev, rp, "wavefront", monkeypatch, seed=9, spp=spp, sample_offset=g, sample_stride=2) for g in range(2)]
    assert sum(s["rays"] for _, s in parts) == sa["rays"]
    assert np.allclose(sum(f.astype(np.float64) for f, _ in parts), a, rtol=2e-5, atol=1e-6)
