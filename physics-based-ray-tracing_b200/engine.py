"""Device engine: owns a ``prt_context`` + ``prt_scene`` and exposes the hot path to the plugin layer.

``DeviceScene`` is what ``mi.load_dict`` / ``mi.load_file`` hand back underneath the Mitsuba-style
``Scene`` object: the scene uploaded once to device-resident SoA buffers with the LBVH built on the GPU,
then ``trace_closest`` (== ``scene.ray_intersect``, /root/reference/CustomIntegrator.py:146,309),
``acquire`` (== ``UltraIntegrator.simulate_acquisition*``, :60-405) and ``render_path``.
All compute goes through the C ABI (capi.py); nothing here computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import Dict, Optional

import numpy as np

from . import capi
from .capi import PrtError, check, dptr, fptr
from .scene import AcqParams, SceneDesc

_contexts: Dict[int, "Context"] = {}
_ctx_lock = threading.Lock()


class Context:
    """One per (process, device)."""

    def __init__(self, device: int = 0):
        self.L = capi.load()
        self.device = device
        h = C.c_void_p()
        check(self.L.prt_create(device, C.byref(h)), "prt_create")
        self.h = h
        sm, maj, mnr, mem = C.c_int(), C.c_int(), C.c_int(), C.c_uint64()
        check(self.L.prt_device_info(h, C.byref(sm), C.byref(maj), C.byref(mnr), C.byref(mem)))
        self.sm_count, self.cc, self.global_mem = sm.value, (maj.value, mnr.value), mem.value

    def profile_begin(self):
        """Start per-kernel-class event timing inside the library (bench.py's roofline leg)."""
        check(self.L.prt_profile_begin(self.h), "prt_profile_begin")

    def profile_read(self) -> dict:
        kt = capi.KernelTimesC()
        check(self.L.prt_profile_read(self.h, C.byref(kt)), "prt_profile_read")
        return kt.as_dict()

    def pinned_array(self, shape, dtype=np.float32) -> np.ndarray:
        """A page-locked numpy array (prt_host_alloc) for a result: D2H goes straight into it -- no staging memcpy,
        no page faults on a fresh allocation.  Buffers are pooled per shape and a buffer is handed out again only
        once NO numpy view of it is alive any more: every view (and every view derived from one: numpy collapses
        ``.base`` chains onto the pool's array) holds a reference to the pool's array, so its refcount tells.  A
        result the caller still holds is therefore never overwritten; the pool grows instead."""
        import sys
        key = (tuple(shape), np.dtype(dtype).str)
        pool = self.__dict__.setdefault("_pinned", {}).setdefault(key, [])
        for k in range(len(pool)):
            if sys.getrefcount(pool[k][0]) == pool[k][1]:
                return pool[k][0].reshape(shape)
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        ptr = C.c_void_p()
        check(self.L.prt_host_alloc(self.h, nbytes, C.byref(ptr)), "prt_host_alloc")
        raw = (C.c_byte * nbytes).from_address(ptr.value)
        pool.append([np.frombuffer(raw, dtype=dtype), 0])     # the array every hand-out's .base collapses onto
        k = len(pool) - 1
        pool[k][1] = sys.getrefcount(pool[k][0])          # idle refcount, measured the same way as above
        return pool[k][0].reshape(shape)

    @staticmethod
    def get(device: Optional[int] = None) -> "Context":
        if device is None:
            import os
            device = int(os.environ.get("LOCAL_RANK", "0")) if "PRT_DEVICE" not in os.environ else int(os.environ["PRT_DEVICE"])
            n = C.c_int()
            L = capi.load()
            check(L.prt_device_count(C.byref(n)), "prt_device_count")
            if n.value == 0:
                raise PrtError("no CUDA device visible; the acquisition / render path has no CPU fallback")
            device %= n.value
        with _ctx_lock:
            c = _contexts.get(device)
            if c is None:
                c = Context(device)
                _contexts[device] = c
            return c


class DeviceScene:
    def __init__(self, desc: SceneDesc, device: Optional[int] = None, context: Optional[Context] = None):
        self.ctx = context or Context.get(device)
        self.L = self.ctx.L
        self.desc = desc
        h = C.c_void_p()
        check(self.L.prt_scene_create(self.ctx.h, C.byref(h)), "prt_scene_create")
        self.h = h
        out = C.c_int()
        for m in desc.materials:
            p = np.zeros(8)
            p[:len(m.params)] = m.params
            e = np.ascontiguousarray(m.emission, dtype=np.float64)
            check(self.L.prt_scene_add_material(h, capi.MAT_KINDS[m.kind], dptr(p), dptr(e), C.byref(out)))
        for s in desc.shapes:
            tw = np.ascontiguousarray(s.to_world, dtype=np.float64).reshape(16)
            if s.kind == "mesh":
                v = np.ascontiguousarray(s.v, dtype=np.float64).reshape(-1, 3)
                vn = None if s.vn is None else np.ascontiguousarray(s.vn, dtype=np.float64).reshape(-1, 3)
                idx = np.ascontiguousarray(s.idx, dtype=np.uint32).reshape(-1, 3)
                check(self.L.prt_scene_add_mesh(h, dptr(v), v.shape[0], dptr(vn), idx.ctypes.data_as(C.POINTER(C.c_uint32)),
                                                idx.shape[0], dptr(tw), s.material, int(s.flip_normals), C.byref(out)),
                      "prt_scene_add_mesh")
            else:
                check(self.L.prt_scene_add_primitive(h, capi.PRIM_KINDS[s.kind], dptr(tw), s.material, int(s.flip_normals),
                                                     C.byref(out)), "prt_scene_add_primitive")
        st = capi.BvhStatsC()
        check(self.L.prt_scene_commit(h, C.byref(st)), "prt_scene_commit")
        self.bvh_stats = st.as_dict()

    def close(self):
        if getattr(self, "h", None):
            self.L.prt_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- mi.traverse(scene)[...] = v; params.update() ------------------------------------------
    def set_material_param(self, material: int, index: int, value: float):
        check(self.L.prt_scene_set_material_param(self.h, material, index, float(value)), "prt_scene_set_material_param")

    def set_shape_transform(self, shape_index: int, to_world):
        """Move one shape of the committed scene (refit, no rebuild: prt_scene_set_shape_transform); refreshes bvh_stats."""
        m = np.ascontiguousarray(getattr(to_world, "matrix", to_world), dtype=np.float64).reshape(16)
        check(self.L.prt_scene_set_shape_transform(self.h, int(shape_index), dptr(m)), "prt_scene_set_shape_transform")
        self.desc.shapes[shape_index].to_world = m.reshape(4, 4).copy()
        st = capi.BvhStatsC()
        check(self.L.prt_scene_get_stats(self.h, C.byref(st)), "prt_scene_get_stats")
        self.bvh_stats = st.as_dict()

    # -- scene.ray_intersect ----------------------------------------------------------------------
    def trace_closest(self, o, d, tmax=None):
        o = np.ascontiguousarray(o, dtype=np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(d, dtype=np.float32).reshape(-1, 3)
        n = o.shape[0]
        tm = None if tmax is None else np.ascontiguousarray(np.broadcast_to(tmax, (n,)), dtype=np.float32)
        t = np.empty(n, dtype=np.float32)
        prim = np.empty(n, dtype=np.int32)
        shape = np.empty(n, dtype=np.int32)
        p, ng, ns, wi, sh_s = (np.zeros((n, 3), dtype=np.float32) for _ in range(5))
        ip = C.POINTER(C.c_int32)
        check(self.L.prt_trace_closest(self.h, fptr(o), fptr(d), fptr(tm), n, fptr(t), prim.ctypes.data_as(ip),
                                       shape.ctypes.data_as(ip), fptr(p), fptr(ng), fptr(ns), fptr(wi), fptr(sh_s)),
              "prt_trace_closest")
        return dict(t=t, prim=prim, shape=shape, p=p, ng=ng, ns=ns, wi=wi, sh_s=sh_s)

    def trace_occluded(self, o, d, tmax=None):
        o = np.ascontiguousarray(o, dtype=np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(d, dtype=np.float32).reshape(-1, 3)
        n = o.shape[0]
        tm = None if tmax is None else np.ascontiguousarray(np.broadcast_to(tmax, (n,)), dtype=np.float32)
        hit = np.empty(n, dtype=np.uint8)
        check(self.L.prt_trace_occluded(self.h, fptr(o), fptr(d), fptr(tm), n, hit.ctypes.data_as(C.POINTER(C.c_uint8))),
              "prt_trace_occluded")
        return hit.astype(bool)

    # -- simulate_acquisition* ---------------------------------------------------------------------
    def acquire(self, params: AcqParams, seed: int = 0, spp: int = 1, sample_offset: int = 0, sample_stride: int = 1):
        """Host-buffer entry point: returns (channel_buf [n_a,n_e,T] f32, tx_delays [n_a,n_e] f32, stats)."""
        ps = capi.make_acq_params(params)
        buf = self.ctx.pinned_array((params.n_angles, params.n_elements, params.time_samples), np.float32)
        tx = np.empty((params.n_angles, params.n_elements), dtype=np.float32)
        st = capi.AcqStatsC()
        check(self.L.prt_acquire(self.h, C.byref(ps), seed, spp, sample_offset, sample_stride, fptr(buf), fptr(tx),
                                 C.byref(st)), "prt_acquire")
        return buf, tx, st.as_dict()

    def acquire_variants(self, params: AcqParams, materials, index: int, values, seed: int = 0, spp: int = 1,
                         sample_offset: int = 0, sample_stride: int = 1):
        """The finite-difference loop of USMain.py:262-289 in one call: one acquisition per value of ONE ultrasound_bsdf
        parameter (index 0 impedance, 1 roughness) of ``materials`` (an id or several), common random numbers.
        Returns (bufs [V,n_a,n_e,T], tx, [stats])."""
        mask = 0
        for m in np.atleast_1d(materials):
            mask |= 1 << int(m)
        vals = np.ascontiguousarray(values, dtype=np.float64).reshape(-1)
        if not 1 <= vals.size <= capi.MAX_VARIANTS:
            raise ValueError(f"acquire_variants: 1..{capi.MAX_VARIANTS} values")
        ps = capi.make_acq_params(params)
        bufs = self.ctx.pinned_array((vals.size, params.n_angles, params.n_elements, params.time_samples), np.float32)
        tx = np.empty((params.n_angles, params.n_elements), dtype=np.float32)
        st = (capi.AcqStatsC * vals.size)()
        check(self.L.prt_acquire_variants(self.h, C.byref(ps), seed, spp, sample_offset, sample_stride, mask, int(index),
                                          dptr(vals), vals.size, fptr(bufs), fptr(tx), st), "prt_acquire_variants")
        return bufs, tx, [x.as_dict() for x in st]

    def us_render(self, params: AcqParams, x, z, seed: int = 0, spp: int = 1, t0: float = 0.0, f_number: float = 1.0,
                  dynamic_range: float = 60.0, shape_pulse: bool = False, wave_cycles: float = 5.0, want_envelope: bool = True):
        """The driver's us_render() (USMain.py:92-224) in one library call, channel data resident on the device.
        Returns (display_image [nz, nx] in [0, 1], envelope [nx, nz] or None, stats)."""
        ps = capi.make_acq_params(params)
        xs = np.ascontiguousarray(x, dtype=np.float32).reshape(-1)
        zs = np.ascontiguousarray(z, dtype=np.float32).reshape(-1)
        u = capi.UsRenderParamsC(xs.size, zs.size, float(t0), float(f_number), int(bool(shape_pulse)), 0, float(wave_cycles),
                                 float(dynamic_range))
        img = self.ctx.pinned_array((zs.size, xs.size), np.float32)
        env = self.ctx.pinned_array((xs.size, zs.size), np.float32) if want_envelope else None
        st = capi.AcqStatsC()
        check(self.L.prt_us_render(self.h, C.byref(ps), seed, spp, 0, 1, C.byref(u), fptr(xs), fptr(zs), fptr(img), fptr(env),
                                   C.byref(st)), "prt_us_render")
        return img, env, st.as_dict()

    def us_postprocess_dev(self, params: AcqParams, channel_ptr: int, x, z, stream: int = 0, t0: float = 0.0, f_number: float = 1.0,
                           dynamic_range: float = 60.0, shape_pulse: bool = False, wave_cycles: float = 5.0, want_envelope: bool = True):
        """us_render() minus the acquisition on a DEVICE channel buffer (prt_us_postprocess_dev): what a sample-sharded
        multi-GPU acquisition feeds its all-reduced buffer to.  Returns (display_image [nz, nx], envelope [nx, nz] or None)."""
        ps = capi.make_acq_params(params)
        xs = np.ascontiguousarray(x, dtype=np.float32).reshape(-1)
        zs = np.ascontiguousarray(z, dtype=np.float32).reshape(-1)
        u = capi.UsRenderParamsC(xs.size, zs.size, float(t0), float(f_number), int(bool(shape_pulse)), 0, float(wave_cycles),
                                 float(dynamic_range))
        img = self.ctx.pinned_array((zs.size, xs.size), np.float32)
        env = self.ctx.pinned_array((xs.size, zs.size), np.float32) if want_envelope else None
        check(self.L.prt_us_postprocess_dev(self.ctx.h, C.byref(ps), C.byref(u), fptr(xs), fptr(zs), C.c_void_p(channel_ptr),
                                            C.c_void_p(stream or None), fptr(img), fptr(env)), "prt_us_postprocess_dev")
        return img, env

    def acquire_dev(self, params: AcqParams, buf_ptr: int, tx_ptr: int = 0, stats_ptr: int = 0, stream: int = 0,
                    seed: int = 0, spp: int = 1, sample_offset: int = 0, sample_stride: int = 1, angle_first: int = 0,
                    angle_count: Optional[int] = None, ps=None):
        """Device-buffer entry point (accumulates into ``buf_ptr`` on ``stream``, asynchronous); optionally only the
        steering angles [angle_first, angle_first + angle_count).  ``ps``: a prebuilt capi.make_acq_params(params)."""
        if ps is None:
            ps = capi.make_acq_params(params)
        n = params.n_angles - angle_first if angle_count is None else angle_count
        check(self.L.prt_acquire_dev_angles(self.h, C.byref(ps), seed, spp, sample_offset, sample_stride, angle_first, n,
                                            C.c_void_p(buf_ptr), C.c_void_p(tx_ptr or None), C.c_void_p(stats_ptr or None),
                                            C.c_void_p(stream or None)), "prt_acquire_dev")

    def acquire_trace(self, params: AcqParams, path_idx, seed: int = 0, spp: int = 1) -> np.ndarray:
        ps = capi.make_acq_params(params)
        idx = np.ascontiguousarray(path_idx, dtype=np.uint64).reshape(-1)
        rec = np.zeros((idx.size, max(params.max_depth, 1)), dtype=capi.SEG_DTYPE)
        check(self.L.prt_acquire_trace(self.h, C.byref(ps), seed, spp, idx.ctypes.data_as(C.POINTER(C.c_uint64)), idx.size,
                                       rec.ctypes.data_as(C.c_void_p)), "prt_acquire_trace")
        return rec

    # -- mi.render with the `path` integrator ------------------------------------------------------
    def render_path(self, rp: "capi.RenderParamsC", seed: int = 0, spp: int = 1, sample_offset: int = 0, sample_stride: int = 1):
        film = self.ctx.pinned_array((rp.height, rp.width, 4), np.float32)
        st = capi.RenderStatsC()
        check(self.L.prt_render_path(self.h, C.byref(rp), seed, spp, sample_offset, sample_stride, fptr(film), C.byref(st)),
              "prt_render_path")
        return film, st.as_dict()

    def render_image(self, rp: "capi.RenderParamsC", seed: int = 0, spp: int = 1, sample_offset: int = 0, sample_stride: int = 1):
        """mi.render(scene): the film is developed on the device; returns (image [H,W,3] f32 in page-locked memory, stats)."""
        img = self.ctx.pinned_array((rp.height, rp.width, 3), np.float32)
        st = capi.RenderStatsC()
        check(self.L.prt_render_image(self.h, C.byref(rp), seed, spp, sample_offset, sample_stride, fptr(img), C.byref(st)),
              "prt_render_image")
        return img, st.as_dict()

    def develop_dev(self, film_ptr: int, n_pixels: int, rgb_ptr: int, stream: int = 0):
        check(self.L.prt_film_develop_dev(self.ctx.h, C.c_void_p(film_ptr), n_pixels, C.c_void_p(rgb_ptr), C.c_void_p(stream or None)),
              "prt_film_develop_dev")

    def render_path_dev(self, rp: "capi.RenderParamsC", film_ptr: int, stats_ptr: int = 0, stream: int = 0, seed: int = 0,
                        spp: int = 1, sample_offset: int = 0, sample_stride: int = 1):
        check(self.L.prt_render_path_dev(self.h, C.byref(rp), seed, spp, sample_offset, sample_stride, C.c_void_p(film_ptr),
                                         C.c_void_p(stats_ptr or None), C.c_void_p(stream or None)), "prt_render_path_dev")


def ultra_bsdf_sample(wi, ng, ns, impedance, roughness, s1, s2, context: Optional[Context] = None):
    """Batched UltraBSDF.sample on the GPU (/root/reference/CustomBSDF.py:87-175)."""
    ctx = context or Context.get()
    wi = np.ascontiguousarray(wi, dtype=np.float32).reshape(-1, 3)
    n = wi.shape[0]
    ng = np.ascontiguousarray(np.broadcast_to(ng, (n, 3)), dtype=np.float32)
    ns = np.ascontiguousarray(np.broadcast_to(ns, (n, 3)), dtype=np.float32)
    z = np.ascontiguousarray(np.broadcast_to(impedance, (n,)), dtype=np.float32)
    r = np.ascontiguousarray(np.broadcast_to(roughness, (n,)), dtype=np.float32)
    a = np.ascontiguousarray(np.broadcast_to(s1, (n,)), dtype=np.float32)
    b = np.ascontiguousarray(np.broadcast_to(s2, (n,)), dtype=np.float32)
    d = np.empty((n, 3), dtype=np.float32)
    pdf = np.empty(n, dtype=np.float32)
    amp = np.empty(n, dtype=np.float32)
    rf = np.empty(n, dtype=np.int32)
    check(ctx.L.prt_ultra_bsdf_sample(ctx.h, n, fptr(wi), fptr(ng), fptr(ns), fptr(z), fptr(r), fptr(a), fptr(b), fptr(d),
                                      fptr(pdf), fptr(amp), rf.ctypes.data_as(C.POINTER(C.c_int32))), "prt_ultra_bsdf_sample")
    return d, pdf, amp, rf.astype(bool)


def directivity_weights(sensor_to_world, sec_dir, ray_dir, normal, main_beam_deg, cutoff_deg, num_rays,
                        context: Optional[Context] = None):
    """Batched directivity_weight_i / directivity_weight_o on the GPU (/root/reference/CustomIntegrator.py:114-135)."""
    ctx = context or Context.get()
    sec = np.ascontiguousarray(sec_dir, dtype=np.float32).reshape(-1, 3)
    n = sec.shape[0]
    rd = np.ascontiguousarray(np.broadcast_to(ray_dir, (n, 3)), dtype=np.float32)
    nr = np.ascontiguousarray(np.broadcast_to(normal, (n, 3)), dtype=np.float32)
    T = np.ascontiguousarray(sensor_to_world, dtype=np.float64).reshape(16)
    w_i, w_o = np.empty(n, dtype=np.float32), np.empty(n, dtype=np.float32)
    check(ctx.L.prt_directivity_weights(ctx.h, n, dptr(T), fptr(sec), fptr(rd), fptr(nr), float(main_beam_deg), float(cutoff_deg),
                                        float(num_rays), fptr(w_i), fptr(w_o)), "prt_directivity_weights")
    return w_i, w_o


def das_beamform(channel, angles_deg, x, z, fs, sound_speed, pitch, t0=0.0, f_number=0.0, tx_delays=None,
                 context: Optional[Context] = None):
    """Plane-wave delay-and-sum + envelope on the GPU (replaces ultraspy's DelayAndSum, USMain.py:175-208).
    channel [n_a, n_e, T] f32 -> (rf [nx, nz], envelope [nx, nz])."""
    ctx = context or Context.get()
    ch = np.ascontiguousarray(channel, dtype=np.float32)
    n_a, n_e, T = ch.shape
    xs = np.ascontiguousarray(x, dtype=np.float32).reshape(-1)
    zs = np.ascontiguousarray(z, dtype=np.float32).reshape(-1)
    ang = np.ascontiguousarray(angles_deg, dtype=np.float64).reshape(-1)
    if ang.size != n_a:
        raise ValueError("das_beamform: one angle per transmission expected")
    p = capi.DasParamsC(n_a, n_e, T, xs.size, zs.size, float(fs), float(sound_speed), float(pitch), float(t0), float(f_number))
    txd = None if tx_delays is None else np.ascontiguousarray(tx_delays, dtype=np.float32)
    rf = ctx.pinned_array((xs.size, zs.size), np.float32)      # page-locked: the two D2H copies need no staging
    env = ctx.pinned_array((xs.size, zs.size), np.float32)
    check(ctx.L.prt_das_beamform(ctx.h, C.byref(p), fptr(ch), fptr(txd), dptr(ang), fptr(xs), fptr(zs), fptr(rf), fptr(env)),
          "prt_das_beamform")
    return rf, env


def envelope(rf, context: Optional[Context] = None):
    ctx = context or Context.get()
    r = np.ascontiguousarray(rf, dtype=np.float32)
    env = np.empty_like(r)
    check(ctx.L.prt_envelope(ctx.h, fptr(r), r.shape[0], r.shape[1], fptr(env)), "prt_envelope")
    return env


def pulse_shape(channel, fs, fc, sigma_s=None, wave_cycles=None, context: Optional[Context] = None):
    """Band-limit delta echoes with the Gaussian-modulated tone burst of /root/reference/RayTracingV0.py:185-204 on the
    GPU.  channel [..., T] f32 -> same shape.  ``sigma_s`` (seconds) or ``wave_cycles`` (sigma = cycles / (4 fc))."""
    ctx = context or Context.get()
    if sigma_s is None:
        if wave_cycles is None:
            raise ValueError("pulse_shape: give sigma_s or wave_cycles")
        sigma_s = float(wave_cycles) / (4.0 * float(fc))
    ch = np.ascontiguousarray(channel, dtype=np.float32)
    T = ch.shape[-1]
    rows = ch.size // T
    out = ctx.pinned_array(ch.shape, np.float32)
    check(ctx.L.prt_pulse_shape(ctx.h, fptr(ch), rows, T, float(fs), float(fc), float(sigma_s), fptr(out)), "prt_pulse_shape")
    return out
