"""The slice of the Mitsuba 3 Python API the reference touches, re-hosted on the
B200 engine.  ``import mitsuba as mi`` resolves to this module through shims/mitsuba when the real
Mitsuba wheel is not installed (it is not, here or on the GPU box), so that
/root/reference/USMain.py and TestScene.py run unchanged:

    mi.set_variant(...)                         USMain.py:12, TestScene.py:3
    mi.register_{integrator,sensor,emitter,bsdf} USMain.py:15-24
    mi.ScalarTransform4f().look_at/translate/rotate/scale, @   USMain.py:53-57,69-71
    mi.load_dict / mi.load_file / mi.traverse   USMain.py:257-259
    scene.integrator() / scene.sensors() / scene.ray_intersect   USMain.py:95, CustomIntegrator.py:101,146

Array types are thin numpy subclasses (float32, like the `llvm_ad_mono` variant's Float); everything
that traces rays goes to the GPU through prt_b200.engine -- nothing here intersects geometry on the CPU.
"""
from __future__ import annotations

import math
import os
from typing import Any, Callable, Dict, List, Optional

import numpy as np

from . import scene as _scene
from .scene import Properties  # noqa: F401  (re-exported as mi.Properties)
from .transforms import Transform4f

ScalarTransform4f = Transform4f
ScalarTransform4d = Transform4f
Transform4d = Transform4f

_VARIANTS = {"scalar_rgb", "scalar_mono", "llvm_ad_rgb", "llvm_ad_mono", "llvm_rgb", "llvm_mono", "cuda_ad_rgb",
             "cuda_ad_mono", "cuda_rgb", "cuda_mono", "scalar_spectral", "llvm_ad_spectral", "cuda_ad_spectral"}
_variant: Optional[str] = None


def set_variant(*names: str) -> None:
    """Every variant maps onto the one sm_100a backend (fp32 arithmetic, as *_ad_mono's Float)."""
    global _variant
    for n in names:
        if n in _VARIANTS:
            _variant = n
            return
    raise AttributeError(f"set_variant(): requested variant(s) {names!r} are not available")


def variant() -> Optional[str]:
    return _variant


def variants() -> List[str]:
    return sorted(_VARIANTS)


# ------------------------------------------------------------------------------------------------
# array types
# ------------------------------------------------------------------------------------------------
class _Arr(np.ndarray):
    _dtype = np.float32
    _kind: Optional[str] = None

    def __new__(cls, *args):
        if len(args) == 0:
            a = np.zeros(1, dtype=cls._dtype)
        elif len(args) == 1:
            v = args[0]
            if hasattr(v, "numpy") and not isinstance(v, np.ndarray):
                v = v.numpy()
            a = np.atleast_1d(np.array(v, dtype=cls._dtype))
        else:
            a = np.array(args, dtype=cls._dtype)
        return a.view(cls)

    def numpy(self) -> np.ndarray:
        return np.asarray(self)

    @property
    def _a(self):
        return np.asarray(self)


class Float(_Arr):
    pass


class Float32(Float):
    pass


class Float64(_Arr):
    _dtype = np.float64


class UInt32(_Arr):
    _dtype = np.uint32


class Int32(_Arr):
    _dtype = np.int32


class Bool(_Arr):
    _dtype = np.bool_


class Color1f(Float):
    pass


class UnpolarizedSpectrum(Float):
    pass


class _Vec3(np.ndarray):
    """[..., 3] float32 with .x/.y/.z; broadcasting constructor like mi.Vector3f(x, 0, z)."""
    _kind = "vector"

    def __new__(cls, *args):
        if len(args) == 0:
            a = np.zeros(3, dtype=np.float32)
        elif len(args) == 1:
            v = args[0]
            a = np.array(v.numpy() if hasattr(v, "numpy") and not isinstance(v, np.ndarray) else v, dtype=np.float32)
            if a.ndim == 0:
                a = np.repeat(a, 3)
        else:
            comps = np.broadcast_arrays(*[np.asarray(c, dtype=np.float32) for c in args])
            a = np.stack(comps, axis=-1)
        if a.shape[-1] != 3:
            raise ValueError("expected 3 components")
        return np.ascontiguousarray(a, dtype=np.float32).view(cls)

    @property
    def _a(self):
        return np.asarray(self)

    def numpy(self):
        return np.asarray(self)

    x = property(lambda s: np.asarray(s)[..., 0].view(Float) if np.asarray(s).ndim > 1 else Float(np.asarray(s)[0]),
                 lambda s, v: np.asarray(s).__setitem__((Ellipsis, 0), v))
    y = property(lambda s: np.asarray(s)[..., 1].view(Float) if np.asarray(s).ndim > 1 else Float(np.asarray(s)[1]),
                 lambda s, v: np.asarray(s).__setitem__((Ellipsis, 1), v))
    z = property(lambda s: np.asarray(s)[..., 2].view(Float) if np.asarray(s).ndim > 1 else Float(np.asarray(s)[2]),
                 lambda s, v: np.asarray(s).__setitem__((Ellipsis, 2), v))


class Vector3f(_Vec3):
    _kind = "vector"


class Point3f(_Vec3):
    _kind = "point"


class Normal3f(_Vec3):
    _kind = "normal"


class Color3f(_Vec3):
    _kind = "vector"


ScalarPoint3f = Point3f
ScalarVector3f = Vector3f


class Vector2f(np.ndarray):
    def __new__(cls, *args):
        if len(args) == 1:
            a = np.array(args[0], dtype=np.float32)
            if a.ndim == 0 or a.shape[-1] != 2:
                a = np.stack([a, a], axis=-1)   # scalar broadcast
        else:
            a = np.stack(np.broadcast_arrays(*[np.asarray(c, dtype=np.float32) for c in args]), axis=-1)
        return a.view(cls)

    x = property(lambda s: np.asarray(s)[..., 0], lambda s, v: np.asarray(s).__setitem__((Ellipsis, 0), v))
    y = property(lambda s: np.asarray(s)[..., 1], lambda s, v: np.asarray(s).__setitem__((Ellipsis, 1), v))


Point2f = Vector2f


class Ray3f:
    def __init__(self, o=None, d=None, maxt=None, time=0.0, wavelengths=None):
        self.o = Point3f(o if o is not None else [0, 0, 0])
        self.d = Vector3f(d if d is not None else [0, 0, 1])
        self.maxt = np.float32(np.finfo(np.float32).max) if maxt is None else maxt
        self.time = Float(time)
        self.wavelengths = wavelengths

    def __call__(self, t):
        return Point3f(np.asarray(self.o) + np.asarray(t, dtype=np.float32)[..., None] * np.asarray(self.d))


class RayDifferential3f(Ray3f):
    pass


def _coordinate_system(n: np.ndarray):
    n = np.asarray(n, dtype=np.float32)
    sign = np.copysign(np.float32(1), n[..., 2])
    a = -1.0 / (sign + n[..., 2])
    b = n[..., 0] * n[..., 1] * a
    s = np.stack([(n[..., 0] * n[..., 0] * a) * sign + 1.0, b * sign, -n[..., 0] * sign], -1)
    t = np.stack([b, n[..., 1] * (n[..., 1] * a) + sign, -n[..., 1]], -1)
    return s.astype(np.float32), t.astype(np.float32)


class Frame3f:
    """SURVEY.md C.4."""

    def __init__(self, *args):
        if len(args) == 1:
            self.n = Vector3f(args[0])
            s, t = _coordinate_system(np.asarray(self.n))
            self.s, self.t = Vector3f(s), Vector3f(t)
        elif len(args) == 3:
            self.s, self.t, self.n = Vector3f(args[0]), Vector3f(args[1]), Vector3f(args[2])
        else:
            self.s, self.t, self.n = Vector3f(1, 0, 0), Vector3f(0, 1, 0), Vector3f(0, 0, 1)

    def to_local(self, v):
        v = np.asarray(v, dtype=np.float32)
        return Vector3f(np.stack([np.sum(v * np.asarray(self.s), -1), np.sum(v * np.asarray(self.t), -1),
                                  np.sum(v * np.asarray(self.n), -1)], -1))

    def to_world(self, v):
        v = np.asarray(v, dtype=np.float32)
        return Vector3f(np.asarray(self.s) * v[..., 0:1] + np.asarray(self.t) * v[..., 1:2] + np.asarray(self.n) * v[..., 2:3])


class _Flags(int):
    def __pos__(self):
        return int(self)

    def __or__(self, o):
        return _Flags(int(self) | int(o))


class BSDFFlags:
    Empty = _Flags(0x0)
    Null = _Flags(0x1)
    DiffuseReflection = _Flags(0x2)
    DiffuseTransmission = _Flags(0x4)
    GlossyReflection = _Flags(0x8)
    GlossyTransmission = _Flags(0x10)
    DeltaReflection = _Flags(0x20)
    DeltaTransmission = _Flags(0x40)
    Anisotropic = _Flags(0x1000)
    SpatiallyVarying = _Flags(0x2000)
    NonSymmetric = _Flags(0x4000)
    FrontSide = _Flags(0x8000)
    BackSide = _Flags(0x10000)


class EmitterFlags:
    Empty = _Flags(0x0)
    DeltaPosition = _Flags(0x1)
    DeltaDirection = _Flags(0x2)
    Infinite = _Flags(0x4)
    Surface = _Flags(0x8)
    SpatiallyVarying = _Flags(0x10)


class ParamFlags:
    Differentiable = _Flags(0x0)
    NonDifferentiable = _Flags(0x1)
    Discontinuous = _Flags(0x2)


class BSDFContext:
    def __init__(self, mode=None, type_mask=0x1FF, component=0xFFFFFFFF):
        self.mode, self.type_mask, self.component = mode, type_mask, component


class BSDFSample3f:
    def __init__(self):
        self.wo = Vector3f(0, 0, 0)
        self.pdf = Float(0)
        self.eta = Float(1)
        self.sampled_type = UInt32(0)
        self.sampled_component = UInt32(0)


class PositionSample3f:
    def __init__(self):
        self.p, self.n, self.uv, self.time, self.pdf, self.delta = Point3f(), Normal3f(0, 0, 1), Vector2f(0, 0), Float(0), Float(0), False


class warp:
    @staticmethod
    def square_to_uniform_disk_concentric(sample):
        """SURVEY.md C.5; a scalar Float broadcasts to (s, s)."""
        u = Vector2f(sample)
        x, y = 2.0 * u.x - 1.0, 2.0 * u.y - 1.0
        is_zero = (x == 0) & (y == 0)
        q = np.abs(x) < np.abs(y)
        r = np.where(q, y, x)
        rp = np.where(q, x, y)
        with np.errstate(divide="ignore", invalid="ignore"):
            phi = np.float32(0.25 * math.pi) * rp / r
        phi = np.where(q, np.float32(0.5 * math.pi) - phi, phi)
        phi = np.where(is_zero, 0.0, phi)
        return Vector2f(r * np.cos(phi), r * np.sin(phi))

    @staticmethod
    def square_to_uniform_hemisphere(sample):
        u = Vector2f(sample)
        p = warp.square_to_uniform_disk_concentric(u)
        z = 1.0 - (p.x * p.x + p.y * p.y)
        s = np.sqrt(z + 1.0)
        return Vector3f(s * p.x, s * p.y, z)

    @staticmethod
    def square_to_cosine_hemisphere(sample):
        p = warp.square_to_uniform_disk_concentric(sample)
        z = np.sqrt(np.maximum(1.0 - p.x * p.x - p.y * p.y, 0.0))
        return Vector3f(p.x, p.y, z)


# ------------------------------------------------------------------------------------------------
# plugin base classes + registry
# ------------------------------------------------------------------------------------------------
class Object:
    def __init__(self, props: Optional[Properties] = None):
        self._props = props if props is not None else Properties()

    def id(self) -> str:
        return self._props.id() if hasattr(self._props, "id") else ""

    def traverse(self, callback):
        pass

    def parameters_changed(self, keys=None):
        pass


class Integrator(Object):
    pass


class SamplingIntegrator(Integrator):
    pass


class Sensor(Object):
    pass


class Emitter(Object):
    pass


class BSDF(Object):
    pass


class Shape(Object):
    pass


_registry: Dict[str, Dict[str, Callable]] = {"integrator": {}, "sensor": {}, "emitter": {}, "bsdf": {}}


def register_integrator(name: str, factory: Callable) -> None:
    _registry["integrator"][name] = factory


def register_sensor(name: str, factory: Callable) -> None:
    _registry["sensor"][name] = factory


def register_emitter(name: str, factory: Callable) -> None:
    _registry["emitter"][name] = factory


def register_bsdf(name: str, factory: Callable) -> None:
    _registry["bsdf"][name] = factory


def _ensure_builtin_plugins() -> None:
    """The XML scenes name `ultrasound_*` (and `ultraray`) plugins without any script having registered
    them; supply the repo's own classes under those names unless the caller registered something else."""
    from . import plugins  # noqa: F401  (puts the reference-named modules on sys.path)
    import CustomBSDF
    import CustomEmmitter
    import CustomIntegrator
    import CustomSensor
    _registry["integrator"].setdefault("ultrasound_integrator", CustomIntegrator.UltraIntegrator)
    _registry["integrator"].setdefault("path", CustomIntegrator.PathIntegrator)
    _registry["sensor"].setdefault("ultrasound_sensor", CustomSensor.UltraSensor)
    _registry["sensor"].setdefault("perspective", CustomSensor.PerspectiveSensor)
    _registry["emitter"].setdefault("ultrasound_emitter", CustomEmmitter.CustomEmitter)
    _registry["emitter"].setdefault("ultraray", CustomEmmitter.CustomEmitter)
    _registry["bsdf"].setdefault("ultrasound_bsdf", CustomBSDF.UltraBSDF)


# ------------------------------------------------------------------------------------------------
# scene objects
# ------------------------------------------------------------------------------------------------
class SurfaceInteraction3f:
    """What scene.ray_intersect returns: the members the reference reads."""

    def __init__(self, scene, ray, res):
        self._scene, self._ray = scene, ray
        self.t = Float(res["t"])
        self.p = Point3f(res["p"])
        self.n = Normal3f(res["ng"])
        s = res["sh_s"]
        self.sh_frame = Frame3f(s, np.cross(res["ns"], s), res["ns"])
        self.wi = Vector3f(res["wi"])
        self.prim_index = UInt32(np.maximum(res["prim"], 0))
        self.shape_index = res["shape"]
        self.time = ray.time
        self.wavelengths = ray.wavelengths

    def is_valid(self):
        return Bool(np.isfinite(np.asarray(self.t)))

    def to_world(self, v):
        return self.sh_frame.to_world(v)

    def to_local(self, v):
        return self.sh_frame.to_local(v)

    def bsdf(self, ray=None):
        idx = np.asarray(self.shape_index).reshape(-1)
        return self._scene._shape_bsdf(int(idx[0]) if idx.size and idx[0] >= 0 else 0)

    def spawn_ray(self, d):
        """SURVEY.md C.3."""
        p, n, d = np.asarray(self.p, dtype=np.float32), np.asarray(self.n, dtype=np.float32), np.asarray(d, dtype=np.float32)
        mag = (1.0 + np.max(np.abs(p), axis=-1)) * np.float32(1500.0 * 2.0 ** -24)
        mag = np.copysign(mag, np.sum(n * d, -1)).astype(np.float32)
        return Ray3f(p + n * mag[..., None], d, None, self.time, self.wavelengths)


class Scene:
    def __init__(self, desc: _scene.SceneDesc):
        _ensure_builtin_plugins()
        self.desc = desc
        self._device_scene = None
        integ = desc.integrator
        self._integrator = None
        if integ is not None:
            fac = _registry["integrator"].get(integ.plugin_name())
            if fac is None:
                raise RuntimeError(f"integrator plugin {integ.plugin_name()!r} is not registered")
            self._integrator = fac(integ)
        self._sensors = []
        if desc.sensor is not None:
            fac = _registry["sensor"].get(desc.sensor.plugin_name())
            if fac is None:
                raise RuntimeError(f"sensor plugin {desc.sensor.plugin_name()!r} is not registered")
            sensor = fac(desc.sensor)
            sensor._film, sensor._sampler, sensor._rfilter = desc.film, desc.sampler, desc.rfilter
            self._sensors.append(sensor)
        self._emitters = []
        for s in desc.shapes:
            if s.emitter is not None and s.emitter.plugin_name() in _registry["emitter"] and s.emitter.plugin_name() != "ultraray":
                try:
                    self._emitters.append(_registry["emitter"][s.emitter.plugin_name()](s.emitter))
                except Exception:
                    pass

    # -- Mitsuba surface ----------------------------------------------------------------------------
    def integrator(self):
        return self._integrator

    def sensors(self):
        return self._sensors

    def emitters(self):
        return self._emitters

    def shapes(self):
        return list(self.desc.shapes)

    def _shape_bsdf(self, shape_index: int):
        mat = self.desc.materials[self.desc.shapes[shape_index].material]
        return mat.plugin if mat.plugin is not None else mat

    # -- engine ------------------------------------------------------------------------------------
    def device(self):
        """Upload once (SoA buffers + GPU LBVH); reused by every acquisition / render / ray query."""
        if self._device_scene is None:
            from .engine import DeviceScene
            self._device_scene = DeviceScene(self.desc)
        return self._device_scene

    def ray_intersect(self, ray: Ray3f, active=True) -> SurfaceInteraction3f:
        o = np.asarray(ray.o, dtype=np.float32).reshape(-1, 3)
        d = np.asarray(ray.d, dtype=np.float32).reshape(-1, 3)
        n = max(o.shape[0], d.shape[0])
        o, d = np.broadcast_to(o, (n, 3)), np.broadcast_to(d, (n, 3))
        res = self.device().trace_closest(o, d, None)
        return SurfaceInteraction3f(self, ray, res)

    def ray_test(self, ray: Ray3f, active=True):
        o = np.asarray(ray.o, dtype=np.float32).reshape(-1, 3)
        d = np.asarray(ray.d, dtype=np.float32).reshape(-1, 3)
        return Bool(self.device().trace_occluded(o, d, None))


class SceneParameters(dict):
    """``mi.traverse(scene)``: a mapping key -> value with ``update()`` pushing edits to the device scene
    without a rebuild (prt_scene_set_material_param).  Keys: ``<shape id>.bsdf.{impedance,roughness}`` and
    ``<integrator>.pitch``.  The reference driver writes ``'shape.bsdf.roughness'`` although its shapes are
    called flat_plate / wall_back (USMain.py:264; real Mitsuba would raise KeyError): that key is accepted
    as an alias for EVERY shape carrying an ultrasound_bsdf."""

    def __init__(self, scene: Scene):
        super().__init__()
        self._scene = scene
        self._dirty = set()
        self._targets: Dict[str, list] = {}
        alias: Dict[str, list] = {}
        for si, sh in enumerate(scene.desc.shapes):
            mat = scene.desc.materials[sh.material]
            if mat.kind == "ultra":
                for pi, pname in enumerate(("impedance", "roughness")):
                    key = f"{sh.id}.bsdf.{pname}"
                    dict.__setitem__(self, key, float(mat.params[pi]))
                    self._targets[key] = [(sh.material, pi)]
                    alias.setdefault(f"shape.bsdf.{pname}", []).append((sh.material, pi))
        for k, v in alias.items():
            self._targets[k] = v
        self._shape_keys = {}
        for si, sh in enumerate(scene.desc.shapes):
            key = f"{sh.id}.to_world"
            dict.__setitem__(self, key, Transform4f(sh.to_world))
            self._shape_keys[key] = si
        integ = scene.integrator()
        if integ is not None and hasattr(integ, "pitch"):
            dict.__setitem__(self, "integrator.pitch", integ.pitch)

    def __setitem__(self, key, value):
        if key not in self._targets and key != "integrator.pitch" and key not in self._shape_keys:
            raise KeyError(key)
        dict.__setitem__(self, key, value)
        self._dirty.add(key)

    def __contains__(self, key):
        return dict.__contains__(self, key) or key in self._targets

    def update(self, values=None):
        if values:
            for k, v in dict(values).items():
                self[k] = v
        changed = sorted(self._dirty)
        for key in changed:
            val = dict.__getitem__(self, key)
            if key == "integrator.pitch":
                self._scene.integrator().pitch = float(val)
                continue
            if key in self._shape_keys:        # '<shape id>.to_world': refit-only update of the device scene
                si = self._shape_keys[key]
                m = np.array(getattr(val, "matrix", val), dtype=np.float64).reshape(4, 4)
                if self._scene._device_scene is not None:
                    self._scene._device_scene.set_shape_transform(si, m)
                else:
                    self._scene.desc.shapes[si].to_world = m
                continue
            v = float(np.asarray(val.numpy() if hasattr(val, "numpy") else val).reshape(-1)[0])
            for material, index in self._targets[key]:
                m = self._scene.desc.materials[material]
                m.params[index] = v
                if m.plugin is not None:
                    setattr(m.plugin, ("impedance", "roughness")[index], Float(v))
                if self._scene._device_scene is not None:
                    self._scene._device_scene.set_material_param(material, index, v)
        self._dirty.clear()
        return changed

    def keep(self, keys):
        pass


def load_dict(d: Dict[str, Any], parallel: bool = True) -> Scene:
    _ensure_builtin_plugins()
    if d.get("type") != "scene":
        raise ValueError("load_dict: only whole scenes ({'type': 'scene', ...}) are supported")
    return Scene(_scene.load_dict_desc(d, _registry))


def load_file(path: str, update_scene: bool = False, parallel: bool = True, transform_order: Optional[str] = None, **kwargs) -> Scene:
    _ensure_builtin_plugins()
    order = transform_order or os.environ.get("PRT_TRANSFORM_ORDER", "mitsuba")
    return Scene(_scene.load_xml(path, transform_order=order, registry=_registry, **kwargs))


def traverse(obj) -> SceneParameters:
    if isinstance(obj, Scene):
        return SceneParameters(obj)
    raise TypeError("traverse(): expected a Scene")


def render(scene: Scene, params=None, sensor=0, integrator=None, seed: int = 0, seed_grad: int = 0, spp: int = 0, spp_grad: int = 0):
    integ = integrator or scene.integrator()
    if not hasattr(integ, "render"):
        raise RuntimeError("render(): the scene's integrator has no render() (use simulate_acquisition*)")
    return integ.render(scene, sensor=sensor, seed=seed, spp=spp)


class _Ad:
    class Adam:  # referenced only inside a dead string of the driver (USMain.py:306)
        def __init__(self, lr=0.05, **kw):
            raise NotImplementedError("automatic differentiation is outside the hot path")


ad = _Ad()
