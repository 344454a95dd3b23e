#!/usr/bin/env python
"""Golden fixtures produced by the REFERENCE'S OWN PYTHON -- test infrastructure.

This script loads /root/reference/CustomBSDF.py and /root/reference/CustomIntegrator.py UNMODIFIED (by file
path, under the module names ``ref_CustomBSDF`` / ``ref_CustomIntegrator``) on top of the repository's
``mitsuba`` / ``drjit`` stand-ins (prt_b200.shims; the real wheels are not installable here) and records
what the reference's code computes:

  A  ``UltraBSDF.sample``                     CustomBSDF.py:87-175 (+ ``_ggx_sample`` :30-61, ``ggx_pdf`` :64-83)
  B  ``directivity_weight_i`` / ``_o``        CustomIntegrator.py:114-135 -- the nested functions' own bytecode,
                                              re-bound to a sensor transform (no line of it is restated here)
  C  ``simulate_acquisition``        ("D")    CustomIntegrator.py:60-232, per-segment records
  D  ``simulate_acquisition_parallel`` ("P")  CustomIntegrator.py:235-405, per-ray records (the reference as
                                              literally written performs ONE segment per ray)

What is and is not pinned by these fixtures
-------------------------------------------
Pinned: every arithmetic statement of the two reference files on the path (rows a2, a5-a11 of SURVEY.md 8(a)).
NOT pinned (still "[MEM]", restated from memory of Mitsuba 3): everything the reference obtains from the
``mitsuba`` wheel -- ``scene.ray_intersect`` (served here by the CPU oracle's intersector in binary32, so no
GPU is involved), ``si.spawn_ray``'s offset, ``Frame3f``/``si.sh_frame``/``si.wi`` conventions and
``warp.square_to_uniform_disk_concentric`` (prt_b200.mi_compat).

The reference is unseeded (``np.random.uniform`` / ``np.random.default_rng()``, CustomIntegrator.py:153,173,
174,219,283); the harness injects the uniforms: ray (a, e) of run s draws the PCG32 stream of path index
``(a*n_e + e)*n_runs + s`` (the RNG contract of SURVEY.md 8(d)), in the reference's own call order, so the
oracle / CUDA decision traces of those same paths can be compared record by record.

Scheduling notes (harness-level; the reference's code is untouched):
  * "P" runs on a ThreadPoolExecutor; the harness substitutes a serial executor that hands the reference's
    ``worker`` ONE (angle, element) at a time.  As written, a block is 16 rays and the UnboundLocalError of
    ``survive`` (CustomIntegrator.py:365-376) aborts the REST of the block after the first ray that survives
    Russian roulette (the exception is swallowed by the never-consumed ``pool.map``); feeding single-ray blocks
    gives every ray its (one) segment, which is what SURVEY.md Appendix A.1#8 / PRT_QF_SINGLE_BOUNCE describe.
  * ``os.cpu_count`` is left alone; ``tqdm`` output is silenced.

Usage:  python tests/golden/make_ref_fixtures.py [--out tests/golden] [--quick]
Writes  ref_bsdf.npz, ref_directivity.npz, ref_segments.npz  (+ ref_fixtures.json with sizes / hashes).
"""
from __future__ import annotations

import argparse
import contextlib
import hashlib
import importlib.util
import io
import json
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFERENCE = os.environ.get("PRT_REFERENCE_DIR", "/root/reference")
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

_mods = {}


def reference_modules():
    """(ref_CustomBSDF, ref_CustomIntegrator, mi, dr): the reference's files executed unmodified on the shims."""
    if _mods:
        return _mods["bsdf"], _mods["integ"], _mods["mi"], _mods["dr"]
    from prt_b200 import shims
    shims.install(force=("mitsuba", "drjit"))
    import drjit as dr
    import mitsuba as mi
    assert "shims" in (mi.__file__ or "") and "shims" in (dr.__file__ or ""), "a real mitsuba/drjit is installed: use it"
    out = {}
    for key, fn in (("bsdf", "CustomBSDF.py"), ("integ", "CustomIntegrator.py")):
        path = os.path.join(REFERENCE, fn)
        spec = importlib.util.spec_from_file_location("ref_" + fn[:-3], path)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        out[key] = m
    _mods.update(out, mi=mi, dr=dr)
    return out["bsdf"], out["integ"], mi, dr


def sha16(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()[:16]


# -------------------------------------------------------------------------------------------------
# A. UltraBSDF.sample
# -------------------------------------------------------------------------------------------------
def _unit(v):
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def bsdf_inputs(n, seed=20261018):
    """n random (wi_local, n_geo, n_shading, impedance, roughness, s1, s2) tuples, all binary32.
    wi covers both hemispheres; the shading normal equals the geometric one for 2/3 of the cases and is a
    perturbed one otherwise (meshes with vertex normals); impedance spans no-TIR .. always-TIR ratios."""
    g = np.random.default_rng(seed)
    wi = _unit(g.normal(size=(n, 3)))
    k = n // 4                                                  # near-normal incidence: the non-TIR branch (Z=7.8: < 8.85 deg)
    wi[:k] = _unit(np.stack([g.normal(size=k) * 0.08, g.normal(size=k) * 0.08, np.where(g.random(k) < 0.5, 1.0, -1.0)], -1))
    ng = _unit(g.normal(size=(n, 3)))
    ns = ng.copy()
    pert = g.random(n) < 1 / 3
    ns[pert] = _unit(ng[pert] + 0.2 * g.normal(size=(int(pert.sum()), 3)))
    Z = np.where(g.random(n) < 0.5, 7.8, g.uniform(0.3, 9.0, n))
    alpha = np.where(g.random(n) < 0.3, g.choice([0.5, 0.7, 0.9], n), g.uniform(0.02, 1.0, n))
    s1, s2 = g.random(n), g.random(n)
    edge = g.random(n) < 0.02                                   # the disk centre / rim (CustomBSDF.py:48,52,55)
    s1[edge] = g.choice([0.0, 0.5, 1.0 - 2.0 ** -24], int(edge.sum()))
    f = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    return dict(wi=f(wi), ng=f(ng), ns=f(ns), impedance=f(Z), roughness=f(alpha), s1=f(s1), s2=f(s2))


class _SI:
    """The members of SurfaceInteraction3f that UltraBSDF.sample reads (CustomBSDF.py:90-95,165)."""

    def __init__(self, mi, wi, ng, ns):
        self.wi, self.n = mi.Vector3f(wi), mi.Normal3f(ng)
        self.sh_frame = mi.Frame3f(ns)

    def to_local(self, v):
        return self.sh_frame.to_local(v)

    def to_world(self, v):
        return self.sh_frame.to_world(v)


def run_bsdf(inp):
    refb, _, mi, dr = reference_modules()
    n = inp["wi"].shape[0]
    wo = np.zeros((n, 3), np.float32)
    chosen = np.zeros((n, 3), np.float32)
    pdf, amp = np.zeros(n, np.float32), np.zeros(n, np.float32)
    comp, stype = np.zeros(n, np.int32), np.zeros(n, np.int32)
    ctx = mi.BSDFContext()
    for i in range(n):
        props = mi.Properties("ultrasound_bsdf", {"impedance": float(inp["impedance"][i]), "roughness": float(inp["roughness"][i])})
        b = refb.UltraBSDF(props)
        si = _SI(mi, inp["wi"][i], inp["ng"][i], inp["ns"][i])
        bs, a = b.sample(ctx, si, dr.full(mi.Float, inp["s1"][i]), dr.full(mi.Float, inp["s2"][i]), True)
        wo[i] = np.asarray(bs.wo).reshape(3)
        chosen[i] = np.asarray(si.to_world(bs.wo)).reshape(3)      # what CustomIntegrator.py:205,358 does with bs.wo
        pdf[i], amp[i] = np.asarray(bs.pdf).reshape(-1)[0], np.asarray(a).reshape(-1)[0]
        comp[i], stype[i] = int(np.asarray(bs.sampled_component).reshape(-1)[0]), int(np.asarray(bs.sampled_type).reshape(-1)[0])
    return dict(wo=wo, dir=chosen, pdf=pdf, amp=amp, component=comp, sampled_type=stype)


# -------------------------------------------------------------------------------------------------
# B. directivity weights: the nested functions' own code objects
# -------------------------------------------------------------------------------------------------
def _nested(fn, name, **cells):
    """Re-bind the nested function ``name`` of ``fn`` (its code object, untouched) to closure cells."""
    for c in fn.__code__.co_consts:
        if isinstance(c, types.CodeType) and c.co_name == name:
            closure = tuple(types.CellType(cells[v]) for v in c.co_freevars)
            return types.FunctionType(c, fn.__globals__, name, None, closure)
    raise LookupError(name)


def run_directivity(n, seed=7):
    _, refi, mi, dr = reference_modules()
    from prt_b200.transforms import Transform4f
    g = np.random.default_rng(seed)
    out = {}
    poses = [Transform4f(), Transform4f().look_at([0.01, 0.0, 0.0], [0.0, 0.02, 0.05], [0, 1, 0])]
    T = np.zeros((n, 16))
    d = _unit(g.normal(size=(n, 3)))
    d[: n // 2, 2] = -np.abs(d[: n // 2, 2]) - 1.5      # half of them head back towards the array (inside the cones)
    d = _unit(d).astype(np.float32)
    nrm = _unit(g.normal(size=(n, 3))).astype(np.float32)
    am = g.choice([10.0, 24.0], n)
    ac = np.where(am == 10.0, 20.0, 30.0)
    N = g.choice([320.0, 3200.0], n)
    wi_D, wi_P, wo_D = np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros(n, np.float32)
    for i in range(n):
        pose = poses[i % 2]
        T[i] = pose.matrix.reshape(16)
        fD = _nested(refi.UltraIntegrator.simulate_acquisition, "directivity_weight_i", sensor_transform=mi.Transform4f(pose.matrix))
        oD = _nested(refi.UltraIntegrator.simulate_acquisition, "directivity_weight_o")
        # the closure of the threaded version lives one level deeper (inside _trace_single_ray)
        tsr = [c for c in refi.UltraIntegrator.simulate_acquisition_parallel.__code__.co_consts
               if isinstance(c, types.CodeType) and c.co_name == "_trace_single_ray"][0]
        holder = types.SimpleNamespace(__code__=tsr, __globals__=refi.UltraIntegrator.simulate_acquisition_parallel.__globals__)
        fP = _nested(holder, "directivity_weight_i", sensor_T=mi.Transform4f(pose.matrix))
        wi_D[i] = np.asarray(fD(mi.Vector3f(d[i]), dr.deg2rad(am[i]), dr.deg2rad(ac[i]))).reshape(-1)[0]
        wi_P[i] = np.asarray(fP(mi.Vector3f(d[i]), dr.deg2rad(am[i]), dr.deg2rad(ac[i]))).reshape(-1)[0]
        wo_D[i] = np.asarray(oD(mi.Vector3f(d[i]), mi.Vector3f(nrm[i]), N[i])).reshape(-1)[0]
    out.update(sensor_to_world=T, sec_dir=d, normal=nrm, main_beam_deg=am, cutoff_deg=ac, num_rays=N,
               w_i_D=wi_D, w_i_P=wi_P, w_o=wo_D)
    return out


# -------------------------------------------------------------------------------------------------
# C / D. the two acquisition loops on an oracle-backed scene
# -------------------------------------------------------------------------------------------------
SEG_FIELDS = [("scene", "i4"), ("run", "i4"), ("a", "i4"), ("e", "i4"), ("seg", "i4"), ("valid", "i4"), ("prim", "i4"),
              ("shape", "i4"), ("recv", "i4"), ("visible", "i4"), ("reflect", "i4"), ("k", "i4"), ("deposited", "i4"),
              ("active_after", "i4"), ("t", "f8"), ("press", "f8"), ("amp", "f8"), ("atten", "f8"), ("pdf", "f8"),
              ("a_resp", "f8"), ("geo_len", "f8"), ("dir", "f8", (3,)), ("u", "f4", (4,))]
SEG_DTYPE = np.dtype(SEG_FIELDS)


class _Stream:
    """The PCG32 stream of one path (oracle/orc.c: orc_path_rng + orc_pcg32_next_f32), handed to the reference
    through the numpy.random surface it calls."""

    def __init__(self, orc_py, seed, path):
        self.orc, (self.st, self.inc) = orc_py, orc_py.path_rng(seed, path)
        self.log = []

    def next(self):
        u = np.float32(self.orc.next_f32(self.st, self.inc))
        self.log.append(u)
        return u

    # numpy Generator surface used by _trace_single_ray (CustomIntegrator.py:319,337,365)
    def integers(self, lo, hi, dtype=np.int64):
        u = self.next()
        self.recv = min(int(np.floor(u * np.float32(hi - lo))), hi - lo - 1) + lo
        return dtype(self.recv)

    def random(self, size=None, dtype=np.float64):
        if size is None:
            return float(self.next())
        return np.array([self.next() for _ in range(int(size))], dtype=dtype)


class _LoggingBSDF:
    """Forwards to the reference's UltraBSDF.sample and keeps what went in and came out."""

    def __init__(self, inner, harness):
        self._inner, self._h = inner, harness

    def sample(self, ctx, si, sample1, sample2, active=True):
        bs, a = self._inner.sample(ctx, si, sample1, sample2, active)
        self._h.cur.update(pdf=float(np.asarray(bs.pdf).reshape(-1)[0]), a_resp=float(np.asarray(a).reshape(-1)[0]),
                           reflect=int(np.asarray(bs.sampled_component).reshape(-1)[0] == 0))
        return bs, a


class HarnessScene:
    """The slice of ``mi.Scene`` the reference's integrator touches (``sensors()[0].transform``,
    ``ray_intersect``), with the intersections served by the CPU oracle (binary32) instead of Mitsuba/Embree."""

    def __init__(self, desc, seed, n_runs):
        import orc_py
        refb, refi, mi, dr = reference_modules()
        self.mi, self.orc_py, self.desc = mi, orc_py, desc
        self.oracle = orc_py.OracleScene(desc)
        self.seed, self.n_runs = seed, n_runs
        import prt_b200.plugins  # noqa: F401  (puts the plugin directory on sys.path)
        import CustomSensor as repo_sensor            # UltraSensor only survives in the reference's stale .pyc
        self._sensors = [repo_sensor.UltraSensor(desc.sensor)]
        self.integrator = refi.UltraIntegrator(desc.integrator)
        self.bsdfs = []
        for m in desc.materials:
            props = mi.Properties("ultrasound_bsdf", {"impedance": float(m.params[0]), "roughness": float(m.params[1])})
            self.bsdfs.append(_LoggingBSDF(refb.UltraBSDF(props), self))
        self.n_e = int(self.integrator.n_elements)
        self.mode, self.run, self.ray_no, self.stream = "D", 0, -1, None
        self.records, self.cur, self.q = [], {}, 0

    # --- what the reference calls -----------------------------------------------------------------
    def sensors(self):
        if self.mode == "D":                       # CustomIntegrator.py:101: once per (angle, element) ray
            self._begin_ray()
        return self._sensors

    def _shape_bsdf(self, shape_index):
        return self.bsdfs[self.desc.shapes[shape_index].material]

    def ray_intersect(self, ray, active=True):
        mi = self.mi
        o = np.asarray(ray.o, dtype=np.float32).reshape(-1, 3)
        d = np.asarray(ray.d, dtype=np.float32).reshape(-1, 3)
        res = self.oracle.trace_closest(o, d, None, prec=32)
        hit = bool(np.isfinite(res["t"][0]))
        if not hit:                                # a well-defined (if meaningless) record so masked arithmetic stays finite
            res["p"][:] = o + d
            res["ng"][:] = res["ns"][:] = (0, 0, 1)
            res["sh_s"][:] = (1, 0, 0)
            res["wi"][:] = (0, 0, 1)
            res["shape"][:] = 0
        si = mi.SurfaceInteraction3f(self, ray, {k: (v.astype(np.float32) if v.dtype == np.float64 else v) for k, v in res.items()})
        if self.q % 2 == 0:                        # closest hit of the segment (CustomIntegrator.py:146 / 309)
            self.cur = dict(valid=int(hit), prim=int(res["prim"][0]), shape=int(res["shape"][0]) if hit else -1,
                            t=float(res["t"][0]) if hit else float("inf"))
        else:                                      # the connection query (:159 / :324)
            self.cur["visible"] = int(not hit)
        self.q += 1
        return si

    # --- bookkeeping ------------------------------------------------------------------------------
    def _begin_ray(self):
        self.ray_no += 1
        path = self.ray_no * self.n_runs + self.run
        self.stream = _Stream(self.orc_py, self.seed, path)
        self.q, self.seg = 0, 0

    def uniform(self, lo=0.0, hi=1.0):            # np.random.uniform stand-in for "D" (:153,173,174,219)
        return float(self.stream.next())

    def default_rng(self, *a):                     # np.random.default_rng stand-in for "P" (:283): once per ray
        self._begin_ray()
        return self.stream


def _record(h, scene_id, extra):
    a, e = divmod(h.ray_no, h.n_e)
    r = np.zeros((), SEG_DTYPE)
    r["scene"], r["run"], r["a"], r["e"], r["seg"] = scene_id, h.run, a, e, h.seg
    for k in ("valid", "prim", "shape", "visible", "reflect"):
        r[k] = h.cur.get(k, -1)
    for k in ("t", "pdf", "a_resp"):
        r[k] = h.cur.get(k, np.nan)
    for k, v in extra.items():
        r[k] = v
    u = h.stream.log[4 * h.seg:4 * h.seg + 4]
    r["u"][:len(u)] = u
    h.records.append(r)
    h.seg += 1


def run_D(h: HarnessScene, scene_id, run):
    """simulate_acquisition: records from the loop state after every body() and from the scatter_reduce call."""
    _, refi, mi, dr = reference_modules()
    h.mode, h.run, h.ray_no = "D", run, -1
    integ = h.integrator
    T, n_e = int(integ.time_samples), h.n_e
    pending = {}

    def scatter_reduce(op, target, value, index, active=True):
        flat = int(np.asarray(index).reshape(-1)[0])
        pending.update(press=float(np.asarray(value).reshape(-1)[0]), k=flat % T, recv=(flat // T) % n_e,
                       deposited=int(bool(np.all(np.asarray(active)))))
        if pending["deposited"]:
            np.asarray(target)[flat] += np.float32(pending["press"])

    def while_loop(state, cond, body, **kw):
        while bool(np.all(np.asarray(cond(*state)))):
            state = body(*state)
            amp, atten, tof, geo_len, depth, ray, active = state
            _record(h, scene_id, dict(pending, amp=float(np.asarray(amp).reshape(-1)[0]), atten=float(np.asarray(atten).reshape(-1)[0]),
                                      geo_len=float(np.asarray(geo_len).reshape(-1)[0]), dir=np.asarray(ray.d, dtype=np.float64).reshape(3),
                                      active_after=int(bool(np.all(np.asarray(active))))))
        return state

    saved = (dr.scatter_reduce, dr.while_loop, np.random.uniform)
    dr.scatter_reduce, dr.while_loop, np.random.uniform = scatter_reduce, while_loop, h.uniform
    try:
        with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings(), np.errstate(all="ignore"):
            warnings.simplefilter("ignore")
            integ.simulate_acquisition(h)
    finally:
        dr.scatter_reduce, dr.while_loop, np.random.uniform = saved
    return np.asarray(integ.channel_buf, dtype=np.float32).copy(), np.asarray(integ.transmission_delays_buf, dtype=np.float32).copy()


class _SerialPool:
    """ThreadPoolExecutor stand-in: one (angle, element) per call of the reference's worker, in job order, with the
    worker's exception swallowed exactly as the never-consumed ``pool.map`` swallows it (CustomIntegrator.py:398-399)."""
    harness = None
    scene_id = 0

    def __init__(self, max_workers=None):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def map(self, worker, blocks):
        h = self.harness
        T = int(h.integrator.time_samples)
        for block in blocks:
            for job in block:
                before = h.ray_no
                err = None
                try:
                    worker([job])
                except UnboundLocalError as ex:          # `survive` (:376) -- the one exception the reference raises here
                    err = ex
                if h.ray_no == before:
                    raise RuntimeError("the reference's worker did not start a ray")
                if h.cur.get("valid"):
                    a, e = job
                    recv = h.stream.recv
                    row = h.integrator.channel_buf[a, recv]
                    diff = row - h.snap[a, recv]
                    nz = np.flatnonzero(diff)
                    extra = dict(recv=recv, deposited=int(nz.size > 0), active_after=0 if err is None else 1,
                                 k=int(nz[0]) if nz.size else -1, press=float(diff[nz[0]]) if nz.size else np.nan,
                                 amp=np.nan, atten=np.nan, geo_len=h.cur["t"])
                    h.snap[a, recv] = row
                    _record(h, self.scene_id, extra)
                else:
                    _record(h, self.scene_id, dict(recv=-1, deposited=0, active_after=0, k=-1, press=np.nan, amp=np.nan,
                                                   atten=np.nan, geo_len=0.0))
        return []


def run_P(h: HarnessScene, scene_id, run):
    """simulate_acquisition_parallel as written: per-ray records (hit, receive element, visibility, BSDF outputs,
    deposit read back from the channel buffer)."""
    _, refi, mi, dr = reference_modules()
    h.mode, h.run, h.ray_no = "P", run, -1
    integ = h.integrator
    h.snap = np.zeros((int(integ.n_angles), h.n_e, int(integ.time_samples)), np.float32)

    _SerialPool.harness, _SerialPool.scene_id = h, scene_id
    saved = (refi.ThreadPoolExecutor, np.random.default_rng, refi.tqdm.tqdm)
    refi.ThreadPoolExecutor, np.random.default_rng = _SerialPool, h.default_rng
    refi.tqdm.tqdm = lambda *a, **k: types.SimpleNamespace(update=lambda n: None, close=lambda: None)
    try:
        with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings(), np.errstate(all="ignore"):
            warnings.simplefilter("ignore")
            integ.simulate_acquisition_parallel(h)
    finally:
        refi.ThreadPoolExecutor, np.random.default_rng, refi.tqdm.tqdm = saved
    return np.asarray(integ.channel_buf, dtype=np.float32).copy(), np.asarray(integ.transmission_delays_buf, dtype=np.float32).copy()


def fixture_scenes():
    """(name, SceneDesc) in fixture order: the six MitsubaScenes under the intended transform order (the array looks
    AT the targets), Sphere_Box / Plate_Box under Mitsuba's rule as well, the driver's dict scene, and the ring mesh."""
    from prt_b200 import scenes
    out = []
    for name in scenes.MITSUBA_SCENES:
        out.append((f"{name}:intended", scenes.ultrasound_scene(name, "intended")))
    for name in ("Sphere_Box", "Plate_Box"):
        out.append((f"{name}:mitsuba", scenes.ultrasound_scene(name, "mitsuba")))
    from prt_b200.scene import load_dict_desc
    out.append(("usmain", load_dict_desc(scenes.usmain_scene_dict(), None)))
    out.append(("ring", scenes.test_ring_scene()))
    return out


SEED = 1234
N_RUNS = 4


def run_segments(quick=False, only=None):
    recs, bufs = [], {}
    names = []
    for sid, (name, desc) in enumerate(fixture_scenes()):
        names.append(name)
        if only is not None and name not in only:
            continue
        for mode in ("D", "P"):
            for run in range(1 if quick else N_RUNS):
                h = HarnessScene(desc, SEED, N_RUNS)
                buf, tx = (run_D if mode == "D" else run_P)(h, sid, run)
                r = np.array(h.records, dtype=SEG_DTYPE)
                recs.append((mode, r))
                nz = np.flatnonzero(buf.reshape(-1))
                bufs[f"{mode}_{sid}_{run}_idx"] = nz.astype(np.int64)
                bufs[f"{mode}_{sid}_{run}_val"] = buf.reshape(-1)[nz]
                if run == 0:
                    bufs[f"{mode}_{sid}_tx"] = tx.reshape(-1)
    out = dict(bufs)
    for mode in ("D", "P"):
        rs = [r for m, r in recs if m == mode]
        out[f"seg_{mode}"] = np.concatenate(rs) if rs else np.zeros(0, SEG_DTYPE)
    out["scene_names"] = np.array(names)
    out["seed"], out["n_runs"] = np.int64(SEED), np.int64(N_RUNS)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=HERE)
    ap.add_argument("--quick", action="store_true", help="small fixture (for the regeneration self-check)")
    ap.add_argument("--n-bsdf", type=int, default=12000)
    ap.add_argument("--n-dir", type=int, default=4000)
    args = ap.parse_args()
    if not os.path.isdir(REFERENCE):
        sys.exit(f"{REFERENCE} not found: the fixtures can only be (re)generated where the reference is mounted")
    os.makedirs(args.out, exist_ok=True)
    nb, nd = (300, 200) if args.quick else (args.n_bsdf, args.n_dir)
    inp = bsdf_inputs(nb)
    np.savez_compressed(os.path.join(args.out, "ref_bsdf.npz"), **inp, **run_bsdf(inp))
    np.savez_compressed(os.path.join(args.out, "ref_directivity.npz"), **run_directivity(nd))
    np.savez_compressed(os.path.join(args.out, "ref_segments.npz"), **run_segments(args.quick))
    meta = {"generator": "tests/golden/make_ref_fixtures.py", "reference_files": {
        fn: sha16(os.path.join(REFERENCE, fn)) for fn in ("CustomBSDF.py", "CustomIntegrator.py")},
        "n_bsdf": nb, "n_directivity": nd, "seed": SEED, "n_runs": N_RUNS, "numpy": np.__version__}
    with open(os.path.join(args.out, "ref_fixtures.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print(json.dumps(meta))


if __name__ == "__main__":
    main()
