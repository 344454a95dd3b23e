"""ctypes binding of ``libprt_b200.so`` (include/prt_b200.h) -- the only way Python reaches the GPU.

There is deliberately no fallback: if the shared library is missing or no CUDA device is visible,
every compute call raises :class:`PrtError`.  The library is built in-tree by
``__graft_entry__.build()`` (``make -C physics-based-ray-tracing_b200/csrc``).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PRT_B200_LIB") or os.path.join(_HERE, "libprt_b200.so")   # override: A/B runs of kernel builds

PRIM_KINDS = {"sphere": 0, "rectangle": 1, "cone": 2, "disk": 3, "cylinder": 4}
MAT_KINDS = {"ultra": 0, "diffuse": 1, "dielectric": 2, "conductor": 3, "null": 4}

MAX_VARIANTS = 16
QF_CLAMP_TIDX = 1 << 0
QF_TOF_LAST_SEGMENT = 1 << 1
QF_SINGLE_BOUNCE = 1 << 2
QF_RR_NO_ABS = 1 << 3
QF_CONNECT_TO_TARGET = 1 << 4

# every symbol include/prt_b200.h declares (tests check the library exports all of them)
EXPORTS = [
    "prt_last_error", "prt_version", "prt_device_count", "prt_create", "prt_destroy", "prt_device_info", "prt_profile_begin", "prt_profile_read", "prt_host_alloc", "prt_host_free",
    "prt_scene_create", "prt_scene_destroy", "prt_scene_add_material", "prt_scene_set_material_param", "prt_scene_set_shape_transform", "prt_scene_get_stats",
    "prt_scene_add_primitive", "prt_scene_add_mesh", "prt_scene_commit", "prt_trace_closest", "prt_trace_occluded",
    "prt_ultra_bsdf_sample", "prt_directivity_weights", "prt_acquire", "prt_acquire_dev", "prt_acquire_dev_angles", "prt_acquire_variants", "prt_acquire_trace", "prt_render_path",
    "prt_render_path_dev", "prt_render_image", "prt_film_develop_dev", "prt_das_beamform", "prt_envelope", "prt_us_render", "prt_us_postprocess_dev", "prt_pulse_shape", "prt_pulse_shape_dev",
]


class PrtError(RuntimeError):
    pass


class AcqParamsC(C.Structure):
    _fields_ = [("n_angles", C.c_int32), ("n_elements", C.c_int32), ("time_samples", C.c_int32),
                ("max_depth", C.c_int32), ("pitch", C.c_double), ("fs", C.c_double),
                ("sound_speed", C.c_double), ("frequency", C.c_double), ("attenuation", C.c_double),
                ("main_beam_deg", C.c_double), ("cutoff_deg", C.c_double), ("max_path_len", C.c_double),
                ("sensor_to_world", C.c_double * 16), ("quirk_flags", C.c_uint32), ("_pad", C.c_uint32),
                ("angles_deg", C.POINTER(C.c_double))]


class AcqStatsC(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("segments", C.c_uint64), ("rays", C.c_uint64), ("deposits", C.c_uint64),
                ("misses", C.c_uint64), ("kernel_ms", C.c_float), ("total_ms", C.c_float), ("launches", C.c_uint32),
                ("_pad", C.c_uint32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_ if n != "_pad"}


class BvhStatsC(C.Structure):
    _fields_ = [("n_primitives", C.c_uint32), ("n_triangles", C.c_uint32), ("n_nodes", C.c_uint32),
                ("max_leaf_size", C.c_uint32), ("build_ms", C.c_float), ("sah_cost", C.c_float),
                ("scene_lo", C.c_float * 3), ("scene_hi", C.c_float * 3), ("device_bytes", C.c_uint64),
                ("n_nodes8", C.c_uint32), ("bvh8_levels", C.c_uint32), ("bvh8_build_ms", C.c_float), ("n_oversized", C.c_uint32)]

    def as_dict(self):
        return dict(n_primitives=self.n_primitives, n_triangles=self.n_triangles, n_nodes=self.n_nodes,
                    max_leaf_size=self.max_leaf_size, build_ms=self.build_ms, sah_cost=self.sah_cost,
                    scene_lo=list(self.scene_lo), scene_hi=list(self.scene_hi), device_bytes=self.device_bytes,
                    n_nodes8=self.n_nodes8, bvh8_levels=self.bvh8_levels, bvh8_build_ms=self.bvh8_build_ms,
                    n_oversized=self.n_oversized)


KERNEL_CLASSES = ("generate", "trace_closest", "trace_shadow", "shade", "film", "acquire", "megakernel", "other")


class KernelTimesC(C.Structure):
    _fields_ = [("ms", C.c_double * 8), ("launches", C.c_uint32 * 8)]

    def as_dict(self):
        return {n: {"ms": self.ms[i], "launches": int(self.launches[i])} for i, n in enumerate(KERNEL_CLASSES) if self.launches[i]}


class RenderParamsC(C.Structure):
    _fields_ = [("to_world", C.c_double * 16), ("fov_deg", C.c_double), ("near_clip", C.c_double),
                ("far_clip", C.c_double), ("width", C.c_int32), ("height", C.c_int32), ("max_depth", C.c_int32),
                ("rr_depth", C.c_int32), ("rfilter", C.c_int32), ("_pad", C.c_int32)]


class RenderStatsC(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("segments", C.c_uint64), ("rays", C.c_uint64), ("shadow_rays", C.c_uint64),
                ("kernel_ms", C.c_float), ("total_ms", C.c_float), ("launches", C.c_uint32), ("_pad", C.c_uint32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_ if n != "_pad"}


class DasParamsC(C.Structure):
    _fields_ = [("n_angles", C.c_int32), ("n_elements", C.c_int32), ("time_samples", C.c_int32), ("nx", C.c_int32),
                ("nz", C.c_int32), ("fs", C.c_double), ("sound_speed", C.c_double), ("pitch", C.c_double),
                ("t0", C.c_double), ("f_number", C.c_double)]


class UsRenderParamsC(C.Structure):
    _fields_ = [("nx", C.c_int32), ("nz", C.c_int32), ("t0", C.c_double), ("f_number", C.c_double), ("shape_pulse", C.c_int32),
                ("_pad", C.c_int32), ("wave_cycles", C.c_double), ("dynamic_range_db", C.c_double)]


SEG_DTYPE = np.dtype([("valid", "i4"), ("prim", "i4"), ("shape", "i4"), ("recv", "i4"), ("visible", "i4"),
                      ("reflect", "i4"), ("k", "i4"), ("survive", "i4"), ("t", "f4"), ("total_time", "f4"),
                      ("press", "f4"), ("amp", "f4"), ("atten", "f4"), ("dir", "f4", (3,))])

_lib = None


def load():
    """Load libprt_b200.so; raises PrtError (never falls back) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PrtError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback for the CUDA path)")
    L = C.CDLL(LIB_PATH)
    vp, dp, fp = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_float)
    ip, u8p, u64p, u32p = C.POINTER(C.c_int32), C.POINTER(C.c_uint8), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)
    L.prt_last_error.restype = C.c_char_p
    L.prt_version.restype = C.c_char_p
    L.prt_device_count.argtypes = [C.POINTER(C.c_int)]
    L.prt_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.prt_destroy.argtypes = [vp]
    L.prt_device_info.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), u64p]
    L.prt_profile_begin.argtypes = [vp]
    L.prt_profile_read.argtypes = [vp, C.POINTER(KernelTimesC)]
    L.prt_host_alloc.argtypes = [vp, C.c_uint64, C.POINTER(vp)]
    L.prt_host_free.argtypes = [vp, vp]
    L.prt_scene_create.argtypes = [vp, C.POINTER(vp)]
    L.prt_scene_destroy.argtypes = [vp]
    L.prt_scene_add_material.argtypes = [vp, C.c_int, dp, dp, C.POINTER(C.c_int)]
    L.prt_scene_set_material_param.argtypes = [vp, C.c_int, C.c_int, C.c_double]
    L.prt_scene_set_shape_transform.argtypes = [vp, C.c_int, dp]
    L.prt_scene_get_stats.argtypes = [vp, C.POINTER(BvhStatsC)]
    L.prt_scene_add_primitive.argtypes = [vp, C.c_int, dp, C.c_int, C.c_int, C.POINTER(C.c_int)]
    L.prt_scene_add_mesh.argtypes = [vp, dp, C.c_uint32, dp, u32p, C.c_uint32, dp, C.c_int, C.c_int, C.POINTER(C.c_int)]
    L.prt_scene_commit.argtypes = [vp, C.POINTER(BvhStatsC)]
    L.prt_trace_closest.argtypes = [vp, fp, fp, fp, C.c_uint64, fp, ip, ip, fp, fp, fp, fp, fp]
    L.prt_trace_occluded.argtypes = [vp, fp, fp, fp, C.c_uint64, u8p]
    L.prt_ultra_bsdf_sample.argtypes = [vp, C.c_uint64, fp, fp, fp, fp, fp, fp, fp, fp, fp, fp, ip]
    L.prt_directivity_weights.argtypes = [vp, C.c_uint64, dp, fp, fp, fp, C.c_double, C.c_double, C.c_double, fp, fp]
    L.prt_acquire.argtypes = [vp, C.POINTER(AcqParamsC), C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, fp, fp,
                              C.POINTER(AcqStatsC)]
    L.prt_acquire_dev.argtypes = [vp, C.POINTER(AcqParamsC), C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, vp, vp, vp, vp]
    L.prt_acquire_dev_angles.argtypes = [vp, C.POINTER(AcqParamsC), C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int32, C.c_int32,
                                         vp, vp, vp, vp]
    L.prt_acquire_variants.argtypes = [vp, C.POINTER(AcqParamsC), C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_int,
                                       dp, C.c_uint32, fp, fp, C.POINTER(AcqStatsC)]
    L.prt_acquire_trace.argtypes = [vp, C.POINTER(AcqParamsC), C.c_uint64, C.c_uint32, u64p, C.c_uint64, vp]
    L.prt_render_path.argtypes = [vp, C.POINTER(RenderParamsC), C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, fp,
                                  C.POINTER(RenderStatsC)]
    L.prt_render_path_dev.argtypes = [vp, C.POINTER(RenderParamsC), C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, vp, vp, vp]
    L.prt_render_image.argtypes = L.prt_render_path.argtypes
    L.prt_film_develop_dev.argtypes = [vp, vp, C.c_uint64, vp, vp]
    L.prt_das_beamform.argtypes = [vp, C.POINTER(DasParamsC), fp, fp, dp, fp, fp, fp, fp]
    L.prt_envelope.argtypes = [vp, fp, C.c_int32, C.c_int32, fp]
    L.prt_us_render.argtypes = [vp, C.POINTER(AcqParamsC), C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(UsRenderParamsC),
                                fp, fp, fp, fp, C.POINTER(AcqStatsC)]
    L.prt_us_postprocess_dev.argtypes = [vp, C.POINTER(AcqParamsC), C.POINTER(UsRenderParamsC), fp, fp, vp, vp, fp, fp]
    L.prt_pulse_shape.argtypes = [vp, fp, C.c_uint64, C.c_int32, C.c_double, C.c_double, C.c_double, fp]
    L.prt_pulse_shape_dev.argtypes = [vp, vp, C.c_uint64, C.c_int32, C.c_double, C.c_double, C.c_double, vp, vp]
    _lib = L
    return L


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().prt_last_error().decode("utf-8", "replace")
        raise PrtError(f"{what or 'libprt_b200'} failed ({rc}): {msg}")


def fptr(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_float))


def dptr(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def make_acq_params(p) -> AcqParamsC:
    """p: prt_b200.scene.AcqParams."""
    s = AcqParamsC()
    s.n_angles, s.n_elements, s.time_samples, s.max_depth = p.n_angles, p.n_elements, p.time_samples, p.max_depth
    s.pitch, s.fs, s.sound_speed, s.frequency, s.attenuation = p.pitch, p.fs, p.sound_speed, p.frequency, p.attenuation
    s.main_beam_deg, s.cutoff_deg, s.max_path_len = p.main_beam_deg, p.cutoff_deg, p.max_path_len
    m = np.ascontiguousarray(p.sensor_to_world, dtype=np.float64).reshape(16)
    for i in range(16):
        s.sensor_to_world[i] = m[i]
    s.quirk_flags = int(p.quirk_flags)
    ang = np.ascontiguousarray(p.angles_deg, dtype=np.float64).reshape(-1)
    s._keepalive = ang
    s.angles_deg = ang.ctypes.data_as(C.POINTER(C.c_double))
    return s
