"""Backend-agnostic scene description + Mitsuba-style XML / dict loaders.

This is the host half of "scene upload": it turns the reference's scene sources
  * the dict of /root/reference/USMain.py:26-90 (``mi.load_dict``),
  * /root/reference/MitsubaScenes/*.xml  (``<float_array>``, ``<shape type="cone">``,
    ``<rotate axis="x,y,z" angle=...>`` -- none of which stock Mitsuba accepts),
  * /root/reference/scenes/cbox.xml (``<default>`` / ``$var``, ``<ref>``, ``obj`` shapes, the
    unregistered ``ultraray`` emitter at :64)
into a flat :class:`SceneDesc` (analytic primitives, triangle meshes, materials, sensor,
integrator) that both the CUDA engine and the test oracle consume.  Nothing here touches a GPU.
"""
from __future__ import annotations

import math
import os
import re
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional

import numpy as np

from .meshio import load_mesh
from .transforms import Transform4f, apply_xml_ops

# analytic primitive kinds; numeric values are the C-ABI's (include/prt_b200.h)
PRIM_KINDS = {"sphere": 0, "rectangle": 1, "cone": 2, "disk": 3, "cylinder": 4}
MAT_KINDS = {"ultra": 0, "diffuse": 1, "dielectric": 2, "conductor": 3, "null": 4}

IOR_TABLE = {"vacuum": 1.0, "air": 1.000277, "water": 1.3330, "bk7": 1.5046, "diamond": 2.419,
             "acrylic glass": 1.49, "polypropylene": 1.49, "pyrex": 1.470, "silicone oil": 1.52045}


class Properties:
    """The slice of ``mi.Properties`` the reference's constructors use: ``get(name, default)``
    (CustomIntegrator.py:16-42), ``has_property`` / ``[]`` (CustomBSDF.py:13-18), ``id()``
    (CustomEmmitter.py:28)."""

    def __init__(self, plugin_name: str = "", values: Optional[Dict[str, Any]] = None, id: str = ""):
        self._plugin = plugin_name
        self._id = id
        self._values: Dict[str, Any] = dict(values or {})

    def get(self, name, default=None):
        return self._values.get(name, default)

    def has_property(self, name) -> bool:
        return name in self._values

    def __contains__(self, name) -> bool:
        return name in self._values

    def __getitem__(self, name):
        return self._values[name]

    def __setitem__(self, name, value):
        self._values[name] = value

    def id(self) -> str:
        return self._id

    def set_id(self, v: str):
        self._id = v

    def plugin_name(self) -> str:
        return self._plugin

    def property_names(self):
        return list(self._values.keys())

    def keys(self):
        return self._values.keys()

    def __repr__(self):
        return f"Properties[{self._plugin!r}, id={self._id!r}, {self._values!r}]"


@dataclass
class MaterialDesc:
    kind: str                       # key of MAT_KINDS
    params: np.ndarray              # [8] f64
    emission: np.ndarray            # [3] f64 (area-emitter radiance attached to the shape)
    id: str = ""
    props: Optional[Properties] = None
    plugin: Any = None              # live BSDF plugin object (e.g. CustomBSDF.UltraBSDF), if any


@dataclass
class ShapeDesc:
    kind: str                       # 'sphere' | 'rectangle' | 'cone' | 'disk' | 'cylinder' | 'mesh'
    to_world: np.ndarray            # [4,4] f64
    material: int
    flip_normals: bool = False
    id: str = ""
    v: Optional[np.ndarray] = None  # mesh: object-space vertices [nv,3]
    vn: Optional[np.ndarray] = None
    idx: Optional[np.ndarray] = None
    emitter: Optional[Properties] = None


@dataclass
class SceneDesc:
    shapes: List[ShapeDesc] = field(default_factory=list)
    materials: List[MaterialDesc] = field(default_factory=list)
    integrator: Optional[Properties] = None
    sensor: Optional[Properties] = None
    film: Optional[Properties] = None
    sampler: Optional[Properties] = None
    rfilter: Optional[Properties] = None
    source: str = ""

    def n_triangles(self) -> int:
        return int(sum(s.idx.shape[0] for s in self.shapes if s.kind == "mesh"))

    def n_analytic(self) -> int:
        return sum(1 for s in self.shapes if s.kind != "mesh")

    def shape_index(self, shape_id: str) -> int:
        for i, s in enumerate(self.shapes):
            if s.id == shape_id:
                return i
        raise KeyError(shape_id)


# ------------------------------------------------------------------------------------------------
# value helpers
# ------------------------------------------------------------------------------------------------
def _to_numpy(x):
    if hasattr(x, "numpy") and callable(x.numpy):
        return np.asarray(x.numpy(), dtype=np.float64)
    return np.asarray(x, dtype=np.float64)


def _floats(text: str) -> List[float]:
    return [float(t) for t in re.split(r"[\s,]+", text.strip()) if t]


def _rgb(value) -> np.ndarray:
    if isinstance(value, str):
        a = np.array(_floats(value), dtype=np.float64)
    elif isinstance(value, dict):
        a = _to_numpy(value.get("value", 0.5)).reshape(-1)
    else:
        a = _to_numpy(value).reshape(-1)
    if a.size == 1:
        a = np.repeat(a, 3)
    return a[:3].copy()


def _ior(value, default: str) -> float:
    if value is None:
        value = default
    if isinstance(value, str):
        try:
            return float(value)
        except ValueError:
            return IOR_TABLE[value]
    return float(value)


def _matrix_of(x) -> np.ndarray:
    if x is None:
        return np.eye(4)
    if isinstance(x, Transform4f):
        return x.matrix.copy()
    if hasattr(x, "matrix"):
        return np.array(_to_numpy(x.matrix), dtype=np.float64).reshape(4, 4)
    return np.array(x, dtype=np.float64).reshape(4, 4)


def make_material(plugin: str, props: Properties, registry=None) -> MaterialDesc:
    """Map a BSDF plugin name + properties to a MaterialDesc."""
    p = np.zeros(8)
    obj = None
    if plugin in ("twosided",):
        inner = props.get("_nested_bsdf")
        if inner is None:
            raise ValueError("twosided bsdf without a nested bsdf")
        return inner
    if plugin == "ultrasound_bsdf" or (registry and plugin in registry.get("bsdf", {}) and
                                       getattr(registry["bsdf"][plugin], "_prt_material_kind", "") == "ultra"):
        # CustomBSDF.py:12-18 defaults
        p[0] = float(props.get("impedance", 1.54))
        p[1] = float(props.get("roughness", 0.5))
        kind = "ultra"
        cls = registry["bsdf"].get(plugin) if registry else None
        if cls is not None:
            obj = cls(props)
    elif plugin == "diffuse":
        p[:3] = _rgb(props.get("reflectance", 0.5))
        kind = "diffuse"
    elif plugin in ("dielectric", "thindielectric", "roughdielectric"):
        p[0] = _ior(props.get("int_ior"), "bk7")
        p[1] = _ior(props.get("ext_ior"), "air")
        kind = "dielectric"
    elif plugin in ("conductor", "roughconductor"):
        p[:3] = _rgb(props.get("specular_reflectance", 1.0))
        kind = "conductor"
    elif plugin == "null":
        kind = "null"
    else:
        raise ValueError(f"unsupported bsdf plugin {plugin!r}")
    return MaterialDesc(kind=kind, params=p, emission=np.zeros(3), id=props.id(), props=props, plugin=obj)


def _emitter_radiance(em: Properties) -> np.ndarray:
    """`area` emitter radiance; the unregistered `ultraray` block of scenes/cbox.xml:64-84 is mapped
    to an area emitter of radiance = its `intensity` rgb (builder decision)."""
    if em.plugin_name() == "area":
        return _rgb(em.get("radiance", 1.0))
    if em.has_property("intensity"):
        return _rgb(em.get("intensity"))
    if em.has_property("radiance"):
        return _rgb(em.get("radiance"))
    return np.ones(3)


# ------------------------------------------------------------------------------------------------
# XML
# ------------------------------------------------------------------------------------------------
class _XmlLoader:
    def __init__(self, path: str, params: Dict[str, Any], transform_order: str, registry):
        self.path = path
        self.dir = os.path.dirname(os.path.abspath(path))
        self.params = {k: str(v) for k, v in params.items()}
        self.order = transform_order
        self.registry = registry
        self.desc = SceneDesc(source=path)
        self.named_bsdf: Dict[str, int] = {}

    def sub(self, text: Optional[str]) -> str:
        if text is None:
            return ""

        def rep(m):
            key = m.group(1)
            if key not in self.params:
                raise KeyError(f"{self.path}: undefined parameter ${key}")
            return self.params[key]

        return re.sub(r"\$(\w+)", rep, text)

    def parse_transform(self, node) -> Transform4f:
        ops = []
        for ch in node:
            a = {k: self.sub(v) for k, v in ch.attrib.items()}
            tag = ch.tag
            if tag == "translate":
                if "value" in a:
                    ops.append(Transform4f().translate(_floats(a["value"])))
                else:
                    ops.append(Transform4f().translate([float(a.get("x", 0)), float(a.get("y", 0)), float(a.get("z", 0))]))
            elif tag == "scale":
                if "value" in a:
                    ops.append(Transform4f().scale(_floats(a["value"])))
                else:
                    ops.append(Transform4f().scale([float(a.get("x", 1)), float(a.get("y", 1)), float(a.get("z", 1))]))
            elif tag == "rotate":
                if "axis" in a:   # non-stock spelling used by MitsubaScenes/*.xml (Sphere_Box.xml:50)
                    axis = _floats(a["axis"])
                else:             # stock: <rotate y="1" angle="..."/>
                    axis = [float(a.get("x", 0)), float(a.get("y", 0)), float(a.get("z", 0))]
                ops.append(Transform4f().rotate(axis, float(a["angle"])))
            elif tag == "lookat":
                ops.append(Transform4f().look_at(_floats(a["origin"]), _floats(a["target"]),
                                                 _floats(a.get("up", "0,1,0"))))
            elif tag == "matrix":
                vals = _floats(a["value"])
                if len(vals) == 9:
                    m = np.eye(4)
                    m[:3, :3] = np.array(vals).reshape(3, 3)
                else:
                    m = np.array(vals).reshape(4, 4)
                ops.append(Transform4f(m))
            else:
                raise ValueError(f"{self.path}: unsupported transform op <{tag}>")
        return apply_xml_ops(ops, self.order)

    def parse_props(self, node, plugin: str) -> Properties:
        props = Properties(plugin, id=node.attrib.get("id", ""))
        for ch in node:
            name = ch.attrib.get("name", "")
            tag = ch.tag
            val = self.sub(ch.attrib.get("value"))
            if tag == "float":
                props[name] = float(val)
            elif tag == "integer":
                props[name] = int(float(val))
            elif tag == "boolean":
                props[name] = val.strip().lower() == "true"
            elif tag == "string":
                props[name] = val
            elif tag in ("rgb", "spectrum", "color"):
                props[name] = np.array(_floats(val))
            elif tag in ("point", "vector"):
                if "value" in ch.attrib:
                    props[name] = np.array(_floats(val))
                else:
                    props[name] = np.array([float(self.sub(ch.attrib.get(k, "0"))) for k in "xyz"])
            elif tag == "float_array":      # non-stock tag, MitsubaScenes/Sphere_Box.xml:14
                props[name] = np.array(_floats(val))
            elif tag == "transform":
                props[name] = self.parse_transform(ch)
            elif tag in ("film", "sampler", "rfilter", "bsdf", "emitter", "ref", "texture", "medium"):
                pass  # handled by the caller
            else:
                raise ValueError(f"{self.path}: unsupported property tag <{tag}>")
        return props

    def parse_bsdf(self, node) -> int:
        plugin = self.sub(node.attrib["type"])
        props = self.parse_props(node, plugin)
        for ch in node:
            if ch.tag == "bsdf":     # twosided / bumpmap wrappers: use the inner bsdf
                inner = self.parse_bsdf(ch)
                if node.attrib.get("id"):
                    self.named_bsdf[node.attrib["id"]] = inner
                return inner
        mat = make_material(plugin, props, self.registry)
        self.desc.materials.append(mat)
        index = len(self.desc.materials) - 1
        if node.attrib.get("id"):
            self.named_bsdf[node.attrib["id"]] = index
        return index

    def default_material(self) -> int:
        props = Properties("diffuse")
        self.desc.materials.append(make_material("diffuse", props))
        return len(self.desc.materials) - 1

    def parse_shape(self, node):
        plugin = self.sub(node.attrib["type"])
        props = self.parse_props(node, plugin)
        material = None
        emitter = None
        for ch in node:
            if ch.tag == "bsdf":
                material = self.parse_bsdf(ch)
            elif ch.tag == "ref":
                rid = self.sub(ch.attrib["id"])
                if rid in self.named_bsdf:
                    material = self.named_bsdf[rid]
                else:
                    raise KeyError(f"{self.path}: unresolved <ref id={rid!r}>")
            elif ch.tag == "emitter":
                emitter = self.parse_props(ch, self.sub(ch.attrib["type"]))
        if material is None:
            material = self.default_material()
        if emitter is not None:
            # an emissive shape needs its own material slot (radiance is stored per material)
            base = self.desc.materials[material]
            mat = MaterialDesc(kind=base.kind, params=base.params.copy(), emission=_emitter_radiance(emitter),
                               id=base.id + "+emitter", props=base.props, plugin=base.plugin)
            self.desc.materials.append(mat)
            material = len(self.desc.materials) - 1
        to_world = _matrix_of(props.get("to_world"))
        flip = bool(props.get("flip_normals", False))
        sid = node.attrib.get("id", "") or f"shape{len(self.desc.shapes)}"
        self.desc.shapes.append(_make_shape(plugin, props, to_world, material, flip, sid, emitter, self.dir))

    def load(self) -> SceneDesc:
        root = ET.parse(self.path).getroot()
        for ch in root:
            if ch.tag == "default":
                self.params.setdefault(ch.attrib["name"], ch.attrib["value"])
        for ch in root:
            tag = ch.tag
            if tag == "default":
                continue
            if tag == "integrator":
                plugin = self.sub(ch.attrib["type"])
                self.desc.integrator = self.parse_props(ch, plugin)
            elif tag == "sensor":
                plugin = self.sub(ch.attrib["type"])
                self.desc.sensor = self.parse_props(ch, plugin)
                for sub in ch:
                    if sub.tag == "film":
                        self.desc.film = self.parse_props(sub, self.sub(sub.attrib["type"]))
                        for f2 in sub:
                            if f2.tag == "rfilter":
                                self.desc.rfilter = self.parse_props(f2, self.sub(f2.attrib["type"]))
                    elif sub.tag == "sampler":
                        self.desc.sampler = self.parse_props(sub, self.sub(sub.attrib["type"]))
            elif tag == "bsdf":
                self.parse_bsdf(ch)
            elif tag == "shape":
                self.parse_shape(ch)
            elif tag in ("emitter", "texture", "medium", "include", "alias"):
                continue  # environment emitters / textures are outside the hot path
            else:
                raise ValueError(f"{self.path}: unsupported top-level tag <{tag}>")
        return self.desc


def _make_shape(plugin, props, to_world, material, flip, sid, emitter, base_dir) -> ShapeDesc:
    if plugin in ("obj", "ply"):
        fn = props["filename"]
        if not os.path.isabs(fn):
            fn = os.path.join(base_dir, fn)
        v, vn, idx = load_mesh(fn)
        if props.get("face_normals", False):
            vn = None
        return ShapeDesc("mesh", to_world, material, flip, sid, v, vn, idx, emitter)
    if plugin == "mesh":  # in-memory mesh handed over by the dict loader
        return ShapeDesc("mesh", to_world, material, flip, sid, _to_numpy(props["vertices"]).reshape(-1, 3),
                         None if props.get("normals") is None else _to_numpy(props["normals"]).reshape(-1, 3),
                         np.asarray(props["faces"], dtype=np.uint32).reshape(-1, 3), emitter)
    if plugin == "sphere":
        # mitsuba sphere: optional `center` / `radius` compose with to_world
        c = props.get("center")
        r = props.get("radius")
        m = to_world
        if c is not None or r is not None:
            loc = Transform4f().translate(_to_numpy(c) if c is not None else [0, 0, 0]).scale(float(r) if r is not None else 1.0)
            m = to_world @ loc.matrix
        return ShapeDesc("sphere", m, material, flip, sid, emitter=emitter)
    if plugin == "cylinder":
        p0 = _to_numpy(props.get("p0", [0, 0, 0])).reshape(3)
        p1 = _to_numpy(props.get("p1", [0, 0, 1])).reshape(3)
        r = float(props.get("radius", 1.0))
        axis = p1 - p0
        length = float(np.linalg.norm(axis))
        n = axis / length
        # frame with +z -> axis
        sgn = math.copysign(1.0, n[2])
        a = -1.0 / (sgn + n[2])
        b = n[0] * n[1] * a
        s = np.array([(n[0] * n[0] * a) * sgn + 1.0, b * sgn, -n[0] * sgn])
        t = np.array([b, n[1] * n[1] * a + sgn, -n[1]])
        loc = np.eye(4)
        loc[:3, 0], loc[:3, 1], loc[:3, 2], loc[:3, 3] = s * r, t * r, n * length, p0
        return ShapeDesc("cylinder", to_world @ loc, material, flip, sid, emitter=emitter)
    if plugin in PRIM_KINDS:
        return ShapeDesc(plugin, to_world, material, flip, sid, emitter=emitter)
    raise ValueError(f"unsupported shape plugin {plugin!r}")


def load_xml(path: str, transform_order: str = "mitsuba", registry=None, **params) -> SceneDesc:
    """``mi.load_file`` equivalent.  ``transform_order``: 'mitsuba' (document order, left-multiplied;
    what Mitsuba renders) or 'intended' (T @ R @ S)."""
    return _XmlLoader(path, params, transform_order, registry).load()


# ------------------------------------------------------------------------------------------------
# dict (mi.load_dict)
# ------------------------------------------------------------------------------------------------
_SHAPE_PLUGINS = set(PRIM_KINDS) | {"obj", "ply", "mesh"}
_BSDF_PLUGINS = {"ultrasound_bsdf", "diffuse", "dielectric", "thindielectric", "roughdielectric", "conductor",
                 "roughconductor", "twosided", "null"}


def _props_from_dict(d: Dict[str, Any], id: str = "") -> Properties:
    props = Properties(d.get("type", ""), id=id)
    for k, v in d.items():
        if k == "type" or isinstance(v, dict):
            continue
        props[k] = v
    return props


def load_dict_desc(d: Dict[str, Any], registry=None, base_dir: str = ".") -> SceneDesc:
    """``mi.load_dict`` equivalent for the reference's scene dict (/root/reference/USMain.py:26-90)."""
    if d.get("type") != "scene":
        raise ValueError("load_dict: top-level dict must have type 'scene'")
    desc = SceneDesc(source="<dict>")
    named: Dict[str, int] = {}
    reg_bsdf = set((registry or {}).get("bsdf", {}))
    reg_int = set((registry or {}).get("integrator", {}))
    reg_sens = set((registry or {}).get("sensor", {}))

    def add_bsdf(bd: Dict[str, Any], bid: str) -> int:
        if bd.get("type") == "ref":
            return named[bd["id"]]
        props = _props_from_dict(bd, bid)
        for k, v in bd.items():
            if isinstance(v, dict) and v.get("type") in (_BSDF_PLUGINS | reg_bsdf):
                inner = add_bsdf(v, k)
                if bd["type"] == "twosided":
                    return inner
        desc.materials.append(make_material(bd["type"], props, registry))
        return len(desc.materials) - 1

    # named bsdfs first so refs resolve regardless of dict order
    for key, val in d.items():
        if isinstance(val, dict) and val.get("type") in (_BSDF_PLUGINS | reg_bsdf):
            named[key] = add_bsdf(val, key)
    for key, val in d.items():
        if not isinstance(val, dict):
            continue
        t = val.get("type")
        if t in (_BSDF_PLUGINS | reg_bsdf):
            continue
        if t in _SHAPE_PLUGINS:
            props = _props_from_dict(val, key)
            material = None
            emitter = None
            for k2, v2 in val.items():
                if not isinstance(v2, dict):
                    continue
                t2 = v2.get("type")
                if t2 == "ref":
                    material = named[v2["id"]]
                elif t2 in (_BSDF_PLUGINS | reg_bsdf):
                    material = add_bsdf(v2, f"{key}.{k2}")
                elif t2 in ("area", "ultraray", "ultrasound_emitter"):
                    emitter = _props_from_dict(v2, k2)
            if material is None:
                desc.materials.append(make_material("diffuse", Properties("diffuse")))
                material = len(desc.materials) - 1
            if emitter is not None:
                base = desc.materials[material]
                desc.materials.append(MaterialDesc(base.kind, base.params.copy(), _emitter_radiance(emitter),
                                                   base.id + "+emitter", base.props, base.plugin))
                material = len(desc.materials) - 1
            desc.shapes.append(_make_shape(t, props, _matrix_of(val.get("to_world")), material,
                                           bool(val.get("flip_normals", False)), key, emitter, base_dir))
        elif key == "integrator" or t in reg_int or t in ("path", "direct", "ultrasound_integrator"):
            desc.integrator = _props_from_dict(val, key)
        elif key == "sensor" or t in reg_sens or t in ("perspective", "ultrasound_sensor"):
            desc.sensor = _props_from_dict(val, key)
            for k2, v2 in val.items():
                if isinstance(v2, dict):
                    t2 = v2.get("type", "")
                    if k2 == "film" or t2 == "hdrfilm":
                        desc.film = _props_from_dict(v2, k2)
                        for k3, v3 in v2.items():
                            if isinstance(v3, dict):
                                desc.rfilter = _props_from_dict(v3, k3)
                    elif k2 == "sampler" or t2 == "independent":
                        desc.sampler = _props_from_dict(v2, k2)
        elif t in ("area", "constant", "envmap", "point"):
            continue
        else:
            raise ValueError(f"load_dict: unsupported entry {key!r} of type {t!r}")
    return desc


# ------------------------------------------------------------------------------------------------
# acquisition parameters shared by engine + tests
# ------------------------------------------------------------------------------------------------
@dataclass
class AcqParams:
    """Scalar inputs of simulate_acquisition* (CustomIntegrator.py:13-48 names / defaults)."""
    n_elements: int = 128
    pitch: float = 0.00035
    angles_deg: np.ndarray = field(default_factory=lambda: np.linspace(-30.0, 30.0, 25))
    time_samples: int = 3000
    max_depth: int = 2
    fs: float = 50e6
    sound_speed: float = 1540.0
    frequency: float = 5e6
    attenuation: float = 0.5
    main_beam_deg: float = 10.0
    cutoff_deg: float = 20.0
    max_path_len: float = 0.2       # hard-coded in the reference, CustomIntegrator.py:141,307
    sensor_to_world: np.ndarray = field(default_factory=lambda: np.eye(4))
    quirk_flags: int = 0

    @property
    def n_angles(self) -> int:
        return int(len(self.angles_deg))

    @classmethod
    def from_props(cls, integ: Properties, sensor: Optional[Properties] = None, **over) -> "AcqParams":
        g = integ.get
        angles = g("angles", None)
        angles = np.linspace(-30.0, 30.0, 25) if angles is None else _to_numpy(angles).reshape(-1)
        p = cls(n_elements=int(g("n_elements", 128)), pitch=float(g("pitch", 0.00035)), angles_deg=angles,
                time_samples=int(g("time_samples", 3000)), max_depth=int(g("max_depth", 2)),
                fs=float(g("sampling_rate", 50e6)), sound_speed=float(g("sound_speed", 1540)),
                frequency=float(g("frequency", 5e6)), attenuation=float(g("attenuation", 0.5)),
                main_beam_deg=float(g("main_beam_angle", 10)), cutoff_deg=float(g("cutoff_angle", 20)))
        if sensor is not None:
            p.sensor_to_world = _matrix_of(sensor.get("to_world"))
        for k, v in over.items():
            setattr(p, k, v)
        return p
