"""GPU parity at BASELINE.json's FULL sizes, through properties that do not need the oracle to trace the whole job:

* strided sub-sampling: the oracle traces every 4096th (8192nd ...) sample of the full-size job -- the SAME PCG32
  streams, because a path's stream is indexed by (angle, element, sample) and spp_total, not by the shard --
  and must reproduce the GPU's shard of the full-size run;
* counting identities (rays = 2 segments + misses, deposits <= segments, paths = n_a n_e spp);
* shard additivity (sum over sample shards == the unsharded run);
* execution-model independence (wavefront pipeline == tile megakernel on the same seeds).

Sizes: config 2 = 512*512*256 paths on each of the six MitsubaScenes, config 3 = 1024*1024*1024 paths on TestRing,
config 4 = cbox at 2048 x 2048 (one 16-spp step of the 4096), config 5 = the 9 999 392-triangle height field at
3840 x 2160.  The GPU side of each takes milliseconds to a second.
"""
import numpy as np
import pytest

from prt_b200 import mi_compat as mi
from prt_b200 import scenes
from prt_b200.scene import AcqParams

pytestmark = pytest.mark.gpu

C2_SPP = 209716            # ceil(512*512*256 / (5*64))
C3_SPP = 3355444           # ceil(1024*1024*1024 / (5*64))


def _rel_mse(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.mean((a - b) ** 2) / max(np.mean(b * b), 1e-300))


def _check_counts(st, n_ae, spp):
    assert st["paths"] == n_ae * spp
    assert st["rays"] == 2 * st["segments"] + st["misses"]
    assert st["deposits"] <= st["segments"]
    assert st["segments"] >= st["paths"] - st["misses"]


@pytest.mark.parametrize("name", scenes.MITSUBA_SCENES)
def test_config2_full_size_properties(orc, name):
    from prt_b200.engine import DeviceScene
    for order in ("mitsuba", "intended"):
        desc = scenes.ultrasound_scene(name, order)
        p = AcqParams.from_props(desc.integrator, desc.sensor)
        n_ae = p.n_angles * p.n_elements
        ds = DeviceScene(desc)
        full, tx, st = ds.acquire(p, seed=0, spp=C2_SPP)
        full = np.array(full)
        assert st["paths"] == 67109120
        _check_counts(st, n_ae, C2_SPP)
        # shard additivity at full size: 8 GPUs' worth of sample shards
        acc = np.zeros(full.shape, dtype=np.float64)
        tot = dict(paths=0, segments=0, rays=0, deposits=0)
        for g in range(8):
            b, _, s = ds.acquire(p, seed=0, spp=C2_SPP, sample_offset=g, sample_stride=8)
            acc += b
            for k in tot:
                tot[k] += s[k]
        assert all(tot[k] == st[k] for k in tot), (tot, st)
        scale = np.abs(full).max()
        if scale > 0:
            assert np.abs(acc - full).max() <= 2e-4 * scale
        # the oracle on every 4096th sample of THIS job (same streams) vs the GPU's identical shard
        stride = 4096
        gb, _, gs = ds.acquire(p, seed=0, spp=C2_SPP, sample_offset=5, sample_stride=stride)
        ob, _, os_ = orc.OracleScene(desc).acquire(p, seed=0, spp=C2_SPP, s_offset=5, s_stride=stride, prec=32)
        assert gs["paths"] == os_["paths"] == n_ae * ((C2_SPP - 5 + stride - 1) // stride)
        assert abs(gs["segments"] - os_["segments"]) <= 3e-3 * os_["segments"] + 2
        assert abs(gs["deposits"] - os_["deposits"]) <= 5e-3 * os_["deposits"] + 2
        if name.startswith(("Plate", "Plane")) and np.abs(ob).max() > 0:
            # flat targets: light-tailed estimator, the shard's buffers agree to rounding (see test_gpu_parity.py for
            # why curved targets are compared on counts and on the clipped bulk only)
            assert _rel_mse(gb, ob) < 1e-3

def test_config3_ring_full_size(orc):
    """1024*1024*1024 paths through the GPU LBVH of TestRing.obj; oracle on every 65536th sample of the same job."""
    from prt_b200.engine import DeviceScene
    desc = scenes.test_ring_scene()
    p = AcqParams.from_props(desc.integrator, desc.sensor)
    n_ae = p.n_angles * p.n_elements
    ds = DeviceScene(desc)
    full, _, st = ds.acquire(p, seed=0, spp=C3_SPP)
    assert st["paths"] == n_ae * C3_SPP >= 1024 ** 3
    _check_counts(st, n_ae, C3_SPP)
    assert np.isfinite(np.asarray(full)).all()
    stride = 65536
    gb, _, gs = ds.acquire(p, seed=0, spp=C3_SPP, sample_offset=17, sample_stride=stride)
    ob, _, os_ = orc.OracleScene(desc).acquire(p, seed=0, spp=C3_SPP, s_offset=17, s_stride=stride, prec=32)
    assert gs["paths"] == os_["paths"]
    assert abs(gs["segments"] - os_["segments"]) <= 3e-3 * os_["segments"] + 5
    assert abs(gs["deposits"] - os_["deposits"]) <= 1e-2 * os_["deposits"] + 5
    # curved target -> heavy-tailed estimator: compare the bulk (clipped at the oracle's 90th percentile), as in
    # test_gpu_parity.py::test_acquire_buffer_heavy_tailed_scene
    nz = ob[ob != 0]
    if nz.size:
        clip = float(np.quantile(np.abs(nz), 0.9))
        assert _rel_mse(np.clip(gb, -clip, clip), np.clip(ob, -clip, clip)) < 5e-3


def test_config4_cbox_full_resolution():
    """scenes/cbox.xml at 2048 x 2048, one 16-spp step: wavefront == megakernel on the same seeds; sample shards add up;
    the mean radiance matches a 256 x 256 render of the same scene (resolution-independent)."""
    import os
    desc = scenes.cbox_scene(2048, 16)
    scene = mi.Scene(desc)
    rp = scene.integrator().render_params(scene)
    dev = scene.device()
    film, st = dev.render_path(rp, seed=4, spp=16)
    film = np.array(film)
    assert st["paths"] == 2048 * 2048 * 16 and st["rays"] - st["shadow_rays"] >= st["segments"] >= st["paths"]
    assert np.isfinite(film).all() and film[..., 3].min() > 0
    os.environ["PRT_PT_MODE"] = "mega"
    try:
        mega, mst = dev.render_path(rp, seed=4, spp=16)
        mega = np.array(mega)
    finally:
        del os.environ["PRT_PT_MODE"]
    assert mst["paths"] == st["paths"] and abs(mst["rays"] - st["rays"]) <= 1e-4 * st["rays"]
    img_w, img_m = film[..., :3] / film[..., 3:], mega[..., :3] / mega[..., 3:]
    for ch in range(3):
        assert _rel_mse(img_w[..., ch], img_m[..., ch]) < 1e-4
    # oracle leg at the full resolution: one of the job's 16 sample planes (sample 3 of every pixel: the SAME PCG32 streams the
    # full job uses for it), 4.2 M paths on the host
    import orc_py
    shard, sst = dev.render_path(rp, seed=4, spp=16, sample_offset=3, sample_stride=16)
    oshard, ost = orc_py.render_path(orc_py.OracleScene(desc), rp, seed=4, spp=16, s_offset=3, s_stride=16, prec=32)
    assert sst["paths"] == ost["paths"] == 2048 * 2048 and abs(sst["segments"] - ost["segments"]) <= 1e-3 * ost["segments"]
    gi, oi = np.array(shard), np.asarray(oshard)
    assert np.abs(gi[..., 3] - oi[..., 3]).max() <= 1e-4                       # filter weights: identical sample positions
    # 8 x 8 block means (a single sample per pixel is all noise; the clipped block means are what a 1-spp plane pins)
    blk = lambda f: np.minimum(f[..., :3] / np.maximum(f[..., 3:], 1e-30), 2.0).reshape(256, 8, 256, 8, 3).mean((1, 3))
    bg, bo = blk(gi), blk(oi)
    for ch in range(3):
        assert _rel_mse(bg[..., ch], bo[..., ch]) < 1e-3
    # shards (offset g, stride 4) add up to the unsharded film
    acc = np.zeros(film.shape, dtype=np.float64)
    n_rays = 0
    for g in range(4):
        f, s = dev.render_path(rp, seed=4, spp=16, sample_offset=g, sample_stride=4)
        acc += f
        n_rays += s["rays"]
    assert n_rays == st["rays"]
    assert np.abs(acc - film).max() <= 2e-4 * np.abs(film).max()
    # resolution independence of the mean image
    small = scenes.cbox_scene(256, 256)
    s2 = mi.Scene(small)
    f2, _ = s2.device().render_path(s2.integrator().render_params(s2), seed=9, spp=256)
    # (8 x 8 block averages of the big image against the small one; clipped at twice the emitter radiance, because the
    # glass sphere's caustic paths are rare and bright: the plain mean of either image is dominated by a few fireflies)
    big = (film[..., :3] / film[..., 3:]).reshape(256, 8, 256, 8, 3).mean((1, 3))
    m_big = np.minimum(big, 2.0).mean((0, 1))
    m_small = np.minimum(f2[..., :3] / f2[..., 3:], 2.0).mean((0, 1))
    assert np.allclose(m_big, m_small, rtol=1e-2), (m_big, m_small)


def test_config5_heightfield_full_size(orc):
    """The 9 999 392-triangle height field at 3840 x 2160: GPU LBVH / BVH8 hits against the oracle's own BVH on random
    rays, closed-box invariants of the 8-bounce diffuse path (no ray escapes), wavefront == megakernel."""
    import os
    desc = scenes.heightfield_scene(2237, (3840, 2160), 2)
    assert desc.n_triangles() == 2 * 2236 * 2236 + 12 == 9999404
    scene = mi.Scene(desc)
    dev = scene.device()
    bs = dev.bvh_stats
    assert bs["n_triangles"] == 9999404 and bs["n_nodes8"] > 0
    assert bs["n_nodes"] == bs["n_triangles"] - bs["n_oversized"] - 1
    rng = np.random.default_rng(4)
    n = 20000
    o = rng.uniform((-0.9, -0.8, -0.9), (0.9, 0.9, 0.9), size=(n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    g = dev.trace_closest(o, d)
    c = orc.OracleScene(desc).trace_closest(o, d, prec=32)
    hit_g, hit_c = g["prim"] >= 0, c["prim"] >= 0
    assert (hit_g == hit_c).mean() > 0.999
    both = hit_g & hit_c
    assert (g["prim"][both] == c["prim"][both]).mean() > 0.998          # ties on shared edges may pick the neighbour
    same = both.copy()
    same[both] = g["prim"][both] == c["prim"][both]
    assert np.allclose(g["t"][same], c["t"][same], rtol=2e-5, atol=2e-6)
    # the render: every path stays in the closed box (segments / path = max_depth unless it ends on the luminaire)
    rp = scene.integrator().render_params(scene)
    film, st = dev.render_path(rp, seed=1, spp=2)
    film = np.array(film)
    assert st["paths"] == 3840 * 2160 * 2
    closest = st["rays"] - st["shadow_rays"]
    # the box is closed, but the floor mesh and the walls do not share vertices: a path spawned within its ray-epsilon
    # of that seam can start outside the wall.  Measured 3e-6 of the rays; anything more is a leaking BVH.
    assert 0 <= closest - st["segments"] <= 1e-5 * closest, "rays escape the closed box"
    assert 8.5 * st["paths"] < st["segments"] <= 9 * st["paths"]
    assert np.isfinite(film).all()
    os.environ["PRT_PT_MODE"] = "mega"
    try:
        mega, mst = dev.render_path(rp, seed=1, spp=2)
        mega = np.array(mega)
    finally:
        del os.environ["PRT_PT_MODE"]
    assert abs(mst["segments"] - st["segments"]) <= 1e-4 * st["segments"]
    iw, im = film[..., :3] / np.maximum(film[..., 3:], 1e-30), mega[..., :3] / np.maximum(mega[..., 3:], 1e-30)
    # same seeds, different BVH (8-wide compressed vs binary) and scheduling: identical up to edge ties
    close = np.abs(iw - im) <= 1e-3 * np.abs(im) + 1e-6
    assert close.mean() > 0.995
    assert abs(iw.mean() - im.mean()) <= 1e-3 * im.mean()
