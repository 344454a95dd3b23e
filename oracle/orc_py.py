"""ctypes binding of the CPU oracle (oracle/liborc.so) -- TEST INFRASTRUCTURE.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` / ``--impl reference`` legs
may import this module.  The product package never does (its CUDA path fails loudly instead of
falling back).  What the oracle is pinned on (and what stays [MEM]): see oracle/orc.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liborc.so")

PRIM_KINDS = {"sphere": 0, "rectangle": 1, "cone": 2, "disk": 3, "cylinder": 4}
MAT_KINDS = {"ultra": 0, "diffuse": 1, "dielectric": 2, "conductor": 3, "null": 4}

QF_CLAMP_TIDX = 1 << 0
QF_TOF_LAST_SEGMENT = 1 << 1
QF_SINGLE_BOUNCE = 1 << 2
QF_RR_NO_ABS = 1 << 3
QF_CONNECT_TO_TARGET = 1 << 4


class AcqParamsC(C.Structure):
    _fields_ = [("n_angles", C.c_int32), ("n_elements", C.c_int32), ("time_samples", C.c_int32),
                ("max_depth", C.c_int32), ("pitch", C.c_double), ("fs", C.c_double),
                ("sound_speed", C.c_double), ("frequency", C.c_double), ("attenuation", C.c_double),
                ("main_beam_deg", C.c_double), ("cutoff_deg", C.c_double), ("max_path_len", C.c_double),
                ("sensor_to_world", C.c_double * 16), ("quirk_flags", C.c_uint32), ("_pad", C.c_uint32),
                ("angles_deg", C.POINTER(C.c_double))]


class StatsC(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("paths", "segments", "rays", "deposits", "misses",
                                           "nodes_visited", "tris_tested")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class SegRecordC(C.Structure):
    _fields_ = [("valid", C.c_int32), ("prim", C.c_int32), ("shape", C.c_int32), ("recv", C.c_int32),
                ("visible", C.c_int32), ("reflect", C.c_int32), ("k", C.c_int32), ("survive", C.c_int32),
                ("t", C.c_double), ("total_time", C.c_double), ("press", C.c_double), ("amp", C.c_double),
                ("atten", C.c_double), ("dir", C.c_double * 3)]


class RenderParamsC(C.Structure):
    _fields_ = [("to_world", C.c_double * 16), ("fov_deg", C.c_double), ("near_clip", C.c_double),
                ("far_clip", C.c_double), ("width", C.c_int32), ("height", C.c_int32), ("max_depth", C.c_int32),
                ("rr_depth", C.c_int32), ("rfilter", C.c_int32), ("_pad", C.c_int32)]


def make_render_params(rp) -> "RenderParamsC":
    """rp: any object with the fields of prt_render_params (e.g. prt_b200.capi.RenderParamsC)."""
    o = RenderParamsC()
    for i in range(16):
        o.to_world[i] = rp.to_world[i]
    for f in ("fov_deg", "near_clip", "far_clip", "width", "height", "max_depth", "rr_depth", "rfilter"):
        setattr(o, f, getattr(rp, f))
    return o


SEG_DTYPE = np.dtype([("valid", "i4"), ("prim", "i4"), ("shape", "i4"), ("recv", "i4"), ("visible", "i4"),
                      ("reflect", "i4"), ("k", "i4"), ("survive", "i4"), ("t", "f8"), ("total_time", "f8"),
                      ("press", "f8"), ("amp", "f8"), ("atten", "f8"), ("dir", "f8", (3,))])
assert SEG_DTYPE.itemsize == C.sizeof(SegRecordC)

_lib = None


def build(force: bool = False) -> str:
    """Compile oracle/liborc.so with the committed Makefile (gcc, a few seconds)."""
    src_m = max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("orc.c", "orc_impl.inl", "orc.h", "orc_pt.inl")
                if os.path.exists(os.path.join(_HERE, f)))
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < src_m:
        subprocess.run(["make", "-C", _HERE, "-B", "liborc.so"], check=True, capture_output=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        dp, ip, u8p, u64p = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_uint8), C.POINTER(C.c_uint64)
        L.orc_scene_create.restype = C.c_void_p
        L.orc_scene_destroy.argtypes = [C.c_void_p]
        L.orc_scene_add_material.argtypes = [C.c_void_p, C.c_int, dp, dp]
        L.orc_scene_add_prim.argtypes = [C.c_void_p, C.c_int, dp, C.c_int, C.c_int]
        L.orc_scene_add_mesh.argtypes = [C.c_void_p, dp, C.c_uint32, dp, C.POINTER(C.c_uint32), C.c_uint32, dp,
                                         C.c_int, C.c_int]
        L.orc_scene_set_material_param.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double]
        L.orc_scene_commit.argtypes = [C.c_void_p, C.c_int]
        L.orc_scene_counts.argtypes = [C.c_void_p, ip, ip, ip]
        L.orc_trace_closest.argtypes = [C.c_void_p, C.c_int, dp, dp, dp, C.c_uint64, dp, ip, ip, dp, dp, dp, dp]
        L.orc_trace_closest_frame.argtypes = [C.c_void_p, C.c_int, dp, dp, dp, C.c_uint64, dp, ip, ip, dp, dp, dp, dp, dp]
        L.orc_trace_occluded.argtypes = [C.c_void_p, C.c_int, dp, dp, dp, C.c_uint64, u8p]
        L.orc_ultra_bsdf.argtypes = [C.c_int, dp, dp, dp, C.c_double, C.c_double, C.c_double, C.c_double, dp, dp, dp, ip]
        L.orc_directivity.argtypes = [C.c_int, dp, dp, dp, dp, C.c_double, C.c_double, C.c_double, dp, dp]
        L.orc_ultra_bsdf_n.argtypes = [C.c_int, C.c_uint64, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp, ip]
        L.orc_directivity_n.argtypes = [C.c_int, C.c_uint64, dp, dp, dp, dp, dp, dp, dp, dp, dp]
        L.orc_acquire.argtypes = [C.c_void_p, C.c_int, C.POINTER(AcqParamsC), C.c_uint64, C.c_uint32, C.c_uint32,
                                  C.c_uint32, dp, dp, C.POINTER(StatsC), C.c_int]
        L.orc_acquire_trace.argtypes = [C.c_void_p, C.c_int, C.POINTER(AcqParamsC), C.c_uint64, C.c_uint32, u64p,
                                        C.c_uint64, C.c_void_p]
        L.orc_render_path.argtypes = [C.c_void_p, C.c_int, C.POINTER(RenderParamsC), C.c_uint64, C.c_uint32, C.c_uint32,
                                      C.c_uint32, dp, C.POINTER(StatsC), u64p, C.c_int]
        L.orc_sample_tea_32.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.orc_pcg32_seed.argtypes = [C.c_uint64, C.c_uint64, u64p, u64p]
        L.orc_pcg32_next_u32.argtypes = [u64p, C.c_uint64]
        L.orc_pcg32_next_u32.restype = C.c_uint32
        L.orc_pcg32_next_f32.argtypes = [u64p, C.c_uint64]
        L.orc_pcg32_next_f32.restype = C.c_float
        L.orc_path_rng.argtypes = [C.c_uint64, C.c_uint64, u64p, u64p]
        _lib = L
    return _lib


def _dptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def make_params(p) -> AcqParamsC:
    """p: prt_b200.scene.AcqParams (duck-typed).  Keeps the angles array alive on the struct."""
    s = AcqParamsC()
    s.n_angles, s.n_elements, s.time_samples, s.max_depth = p.n_angles, p.n_elements, p.time_samples, p.max_depth
    s.pitch, s.fs, s.sound_speed, s.frequency, s.attenuation = p.pitch, p.fs, p.sound_speed, p.frequency, p.attenuation
    s.main_beam_deg, s.cutoff_deg, s.max_path_len = p.main_beam_deg, p.cutoff_deg, p.max_path_len
    m = _f64(p.sensor_to_world).reshape(16)
    for i in range(16):
        s.sensor_to_world[i] = m[i]
    s.quirk_flags = int(p.quirk_flags)
    ang = _f64(p.angles_deg).reshape(-1)
    s._angles_keepalive = ang
    s.angles_deg = ang.ctypes.data_as(C.POINTER(C.c_double))
    return s


class OracleScene:
    """Oracle-side counterpart of the product's DeviceScene, built from the same SceneDesc."""

    def __init__(self, desc=None, use_bvh: bool = True):
        self.L = lib()
        self.h = C.c_void_p(self.L.orc_scene_create())
        self.n_materials = 0
        if desc is not None:
            for m in desc.materials:
                self.add_material(m.kind, m.params, m.emission)
            for s in desc.shapes:
                if s.kind == "mesh":
                    self.add_mesh(s.v, s.vn, s.idx, s.to_world, s.material, s.flip_normals)
                else:
                    self.add_prim(s.kind, s.to_world, s.material, s.flip_normals)
            self.commit(use_bvh)

    def __del__(self):
        try:
            if self.h:
                self.L.orc_scene_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def add_material(self, kind, params, emission=None) -> int:
        p = np.zeros(8)
        p[:len(params)] = params
        e = _f64(emission if emission is not None else np.zeros(3))
        r = self.L.orc_scene_add_material(self.h, MAT_KINDS[kind] if isinstance(kind, str) else kind, _dptr(p), _dptr(e))
        self.n_materials += 1
        return r

    def add_prim(self, kind, to_world, material, flip=False) -> int:
        m = _f64(to_world).reshape(16)
        r = self.L.orc_scene_add_prim(self.h, PRIM_KINDS[kind] if isinstance(kind, str) else kind, _dptr(m), material, int(flip))
        if r < 0:
            raise RuntimeError(f"orc_scene_add_prim failed: {r}")
        return r

    def add_mesh(self, v, vn, idx, to_world, material, flip=False) -> int:
        v = _f64(v).reshape(-1, 3)
        vn = None if vn is None else _f64(vn).reshape(-1, 3)
        idx = np.ascontiguousarray(idx, dtype=np.uint32).reshape(-1, 3)
        m = _f64(to_world).reshape(16)
        r = self.L.orc_scene_add_mesh(self.h, _dptr(v), v.shape[0], _dptr(vn), idx.ctypes.data_as(C.POINTER(C.c_uint32)),
                                      idx.shape[0], _dptr(m), material, int(flip))
        if r < 0:
            raise RuntimeError(f"orc_scene_add_mesh failed: {r}")
        return r

    def set_material_param(self, material: int, index: int, value: float):
        if self.L.orc_scene_set_material_param(self.h, material, index, float(value)):
            raise RuntimeError("orc_scene_set_material_param failed")

    def commit(self, use_bvh: bool = True):
        self.L.orc_scene_commit(self.h, int(use_bvh))

    def counts(self):
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        self.L.orc_scene_counts(self.h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def trace_closest(self, o, d, tmax=None, prec: int = 32):
        o, d = _f64(o).reshape(-1, 3), _f64(d).reshape(-1, 3)
        n = o.shape[0]
        tm = None if tmax is None else _f64(np.broadcast_to(tmax, (n,)))
        t = np.empty(n)
        prim = np.empty(n, dtype=np.int32)
        shape = np.empty(n, dtype=np.int32)
        p, ng, ns, wi, fs = (np.zeros((n, 3)) for _ in range(5))
        ip = C.POINTER(C.c_int32)
        rc = self.L.orc_trace_closest_frame(self.h, prec, _dptr(o), _dptr(d), _dptr(tm), n, _dptr(t), prim.ctypes.data_as(ip),
                                            shape.ctypes.data_as(ip), _dptr(p), _dptr(ng), _dptr(ns), _dptr(wi), _dptr(fs))
        if rc:
            raise RuntimeError(f"orc_trace_closest failed: {rc}")
        return dict(t=t, prim=prim, shape=shape, p=p, ng=ng, ns=ns, wi=wi, sh_s=fs)

    def trace_occluded(self, o, d, tmax=None, prec: int = 32):
        o, d = _f64(o).reshape(-1, 3), _f64(d).reshape(-1, 3)
        n = o.shape[0]
        tm = None if tmax is None else _f64(np.broadcast_to(tmax, (n,)))
        hit = np.empty(n, dtype=np.uint8)
        self.L.orc_trace_occluded(self.h, prec, _dptr(o), _dptr(d), _dptr(tm), n, hit.ctypes.data_as(C.POINTER(C.c_uint8)))
        return hit.astype(bool)

    def acquire(self, params, seed=0, spp=1, s_offset=0, s_stride=1, prec=32, n_threads=None):
        ps = make_params(params)
        buf = np.zeros((params.n_angles, params.n_elements, params.time_samples))
        tx = np.zeros((params.n_angles, params.n_elements))
        st = StatsC()
        if n_threads is None:
            n_threads = os.cpu_count() or 1
        rc = self.L.orc_acquire(self.h, prec, C.byref(ps), seed, spp, s_offset, s_stride, _dptr(buf), _dptr(tx),
                                C.byref(st), n_threads)
        if rc:
            raise RuntimeError(f"orc_acquire failed: {rc}")
        return buf, tx, st.as_dict()

    def acquire_trace(self, params, path_idx, seed=0, spp=1, prec=32) -> np.ndarray:
        ps = make_params(params)
        idx = np.ascontiguousarray(path_idx, dtype=np.uint64).reshape(-1)
        rec = np.zeros((idx.size, params.max_depth), dtype=SEG_DTYPE)
        rc = self.L.orc_acquire_trace(self.h, prec, C.byref(ps), seed, spp, idx.ctypes.data_as(C.POINTER(C.c_uint64)),
                                      idx.size, rec.ctypes.data_as(C.c_void_p))
        if rc:
            raise RuntimeError(f"orc_acquire_trace failed: {rc}")
        return rec


def render_path(scene: "OracleScene", rp, seed=0, spp=1, s_offset=0, s_stride=1, prec=32, n_threads=None):
    """film [H,W,4] float64 = (sum w R, sum w G, sum w B, sum w); stats incl. shadow_rays."""
    L = lib()
    p = make_render_params(rp)
    film = np.zeros((p.height, p.width, 4))
    st = StatsC()
    sh = C.c_uint64()
    rc = L.orc_render_path(scene.h, prec, C.byref(p), seed, spp, s_offset, s_stride, _dptr(film), C.byref(st), C.byref(sh),
                           n_threads or (os.cpu_count() or 1))
    if rc:
        raise RuntimeError(f"orc_render_path failed: {rc}")
    d = st.as_dict()
    d["shadow_rays"] = int(sh.value)
    return film, d


def ultra_bsdf(wi, ng, ns, impedance, roughness, s1, s2, prec=32):
    L = lib()
    wi, ng, ns = _f64(wi), _f64(ng), _f64(ns)
    d = np.zeros(3)
    pdf, amp, rf = C.c_double(), C.c_double(), C.c_int32()
    L.orc_ultra_bsdf(prec, _dptr(wi), _dptr(ng), _dptr(ns), impedance, roughness, s1, s2, _dptr(d), C.byref(pdf),
                     C.byref(amp), C.byref(rf))
    return d, pdf.value, amp.value, bool(rf.value)


def ultra_bsdf_n(wi, ng, ns, impedance, roughness, s1, s2, prec=32):
    """Batched ultra_bsdf: arrays of n -> (dir [n,3], pdf [n], amp [n], reflect [n] bool)."""
    wi, ng, ns = _f64(wi).reshape(-1, 3), _f64(ng).reshape(-1, 3), _f64(ns).reshape(-1, 3)
    n = wi.shape[0]
    z, r, a, b = (_f64(np.broadcast_to(x, (n,))) for x in (impedance, roughness, s1, s2))
    d, pdf, amp, rf = np.zeros((n, 3)), np.zeros(n), np.zeros(n), np.zeros(n, dtype=np.int32)
    rc = lib().orc_ultra_bsdf_n(prec, n, _dptr(wi), _dptr(ng), _dptr(ns), _dptr(z), _dptr(r), _dptr(a), _dptr(b), _dptr(d),
                                _dptr(pdf), _dptr(amp), rf.ctypes.data_as(C.POINTER(C.c_int32)))
    if rc:
        raise RuntimeError("orc_ultra_bsdf_n failed")
    return d, pdf, amp, rf.astype(bool)


def directivity_n(sensor_to_world, sec_dir, ray_dir, normal, main_beam_deg, cutoff_deg, num_rays, prec=32):
    sec, rd, nr = _f64(sec_dir).reshape(-1, 3), _f64(ray_dir).reshape(-1, 3), _f64(normal).reshape(-1, 3)
    n = sec.shape[0]
    T = _f64(np.broadcast_to(_f64(sensor_to_world).reshape(-1, 16), (n, 16)))
    am, ac, nn = (_f64(np.broadcast_to(x, (n,))) for x in (main_beam_deg, cutoff_deg, num_rays))
    wi, wo = np.zeros(n), np.zeros(n)
    if lib().orc_directivity_n(prec, n, _dptr(T), _dptr(sec), _dptr(rd), _dptr(nr), _dptr(am), _dptr(ac), _dptr(nn), _dptr(wi), _dptr(wo)):
        raise RuntimeError("orc_directivity_n failed")
    return wi, wo


def directivity(sensor_to_world, sec_dir, ray_dir, normal, main_beam_deg, cutoff_deg, num_rays, prec=32):
    """(w_i, w_o) of CustomIntegrator.py:114-135 for one connection direction / one (ray direction, normal) pair."""
    wi, wo = C.c_double(), C.c_double()
    rc = lib().orc_directivity(prec, _dptr(_f64(sensor_to_world).reshape(16)), _dptr(_f64(sec_dir)), _dptr(_f64(ray_dir)),
                               _dptr(_f64(normal)), float(main_beam_deg), float(cutoff_deg), float(num_rays), C.byref(wi), C.byref(wo))
    if rc:
        raise RuntimeError("orc_directivity failed")
    return wi.value, wo.value


def path_rng(seed: int, path: int):
    L = lib()
    st, inc = C.c_uint64(), C.c_uint64()
    L.orc_path_rng(seed, path, C.byref(st), C.byref(inc))
    return st, inc


def next_f32(st, inc) -> float:
    return float(lib().orc_pcg32_next_f32(C.byref(st), inc))


def next_u32(st, inc) -> int:
    return int(lib().orc_pcg32_next_u32(C.byref(st), inc))
