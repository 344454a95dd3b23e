"""4x4 affine transforms with Mitsuba 3's ``ScalarTransform4f`` call surface.

The reference builds its poses with ``mi.ScalarTransform4f().look_at(...)``,
``.translate(...) @ .rotate(axis, deg) @ .scale(...)`` (/root/reference/USMain.py:53-57,69-71,81-83)
and the sensor wraps ``props['to_world'].matrix`` into ``mi.Transform4f``
(UltraSensor.__init__, recovered pyc lines 22-26).  Semantics follow
SURVEY.md Appendix C.1: ``A @ B`` is the ordinary matrix product, points are affine (w = 1),
vectors ignore translation, normals use the inverse transpose.
"""
from __future__ import annotations

import math
from typing import Iterable, Sequence

import numpy as np


def _vec3(v) -> np.ndarray:
    if isinstance(v, (int, float)):
        return np.array([v, v, v], dtype=np.float64)
    a = np.asarray(v, dtype=np.float64).reshape(-1)
    if a.size == 1:
        return np.repeat(a, 3)
    if a.size != 3:
        raise ValueError(f"expected a 3-vector, got {v!r}")
    return a


class Transform4f:
    """Row-major 4x4 double matrix; immutable value semantics like Mitsuba's transform."""

    __slots__ = ("matrix",)

    def __init__(self, matrix=None):
        if matrix is None:
            m = np.eye(4, dtype=np.float64)
        elif isinstance(matrix, Transform4f):
            m = matrix.matrix.copy()
        else:
            m = np.array(matrix, dtype=np.float64).reshape(4, 4)
        self.matrix = m

    # -- constructors; callable on the class or on an instance (``T().translate(v)`` composes
    #    on the right exactly as Mitsuba's chained form does) ---------------------------------
    def _compose(self, other: "Transform4f") -> "Transform4f":
        return Transform4f(self.matrix @ other.matrix)

    def translate(self, v) -> "Transform4f":
        m = np.eye(4)
        m[:3, 3] = _vec3(v)
        return self._compose(Transform4f(m))

    def scale(self, v) -> "Transform4f":
        m = np.eye(4)
        s = _vec3(v)
        m[0, 0], m[1, 1], m[2, 2] = s
        return self._compose(Transform4f(m))

    def rotate(self, axis, angle: float) -> "Transform4f":
        """Rotation by ``angle`` DEGREES about ``axis`` (Rodrigues), as Mitsuba's rotate()."""
        a = _vec3(axis)
        n = np.linalg.norm(a)
        if n == 0:
            raise ValueError("rotate: zero axis")
        x, y, z = a / n
        r = math.radians(float(angle))
        s, c = math.sin(r), math.cos(r)
        m = np.eye(4)
        m[0, 0] = x * x + (1 - x * x) * c
        m[0, 1] = x * y * (1 - c) - z * s
        m[0, 2] = x * z * (1 - c) + y * s
        m[1, 0] = x * y * (1 - c) + z * s
        m[1, 1] = y * y + (1 - y * y) * c
        m[1, 2] = y * z * (1 - c) - x * s
        m[2, 0] = x * z * (1 - c) - y * s
        m[2, 1] = y * z * (1 - c) + x * s
        m[2, 2] = z * z + (1 - z * z) * c
        return self._compose(Transform4f(m))

    def look_at(self, origin, target, up) -> "Transform4f":
        """columns [left | new_up | dir | origin]."""
        o, t, u = _vec3(origin), _vec3(target), _vec3(up)
        d = t - o
        d = d / np.linalg.norm(d)
        left = np.cross(u, d)
        ln = np.linalg.norm(left)
        if ln == 0:
            raise ValueError("look_at: up is parallel to the viewing direction")
        left = left / ln
        new_up = np.cross(d, left)
        m = np.eye(4)
        m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = left, new_up, d, o
        return self._compose(Transform4f(m))

    def perspective(self, fov: float, near: float, far: float) -> "Transform4f":
        recip = 1.0 / (far - near)
        cot = 1.0 / math.tan(math.radians(fov * 0.5))
        m = np.zeros((4, 4))
        m[0, 0] = cot
        m[1, 1] = cot
        m[2, 2] = far * recip
        m[2, 3] = -near * far * recip
        m[3, 2] = 1.0
        return self._compose(Transform4f(m))

    # -- algebra ---------------------------------------------------------------------------
    def __matmul__(self, other):
        if isinstance(other, Transform4f):
            return self._compose(other)
        kind = getattr(other, "_kind", None)
        a = np.asarray(getattr(other, "_a", other), dtype=np.float64)
        if kind == "vector":
            out = self.transform_vector(a)
        elif kind == "normal":
            out = self.transform_normal(a)
        else:
            out = self.transform_point(a)
        if kind is not None:
            return type(other)(out)
        return out

    def inverse(self) -> "Transform4f":
        return Transform4f(np.linalg.inv(self.matrix))

    def transform_point(self, p) -> np.ndarray:
        p = np.asarray(p, dtype=np.float64)
        return p @ self.matrix[:3, :3].T + self.matrix[:3, 3]

    transform_affine = transform_point

    def transform_vector(self, v) -> np.ndarray:
        return np.asarray(v, dtype=np.float64) @ self.matrix[:3, :3].T

    def transform_normal(self, n) -> np.ndarray:
        inv = np.linalg.inv(self.matrix[:3, :3])
        return np.asarray(n, dtype=np.float64) @ inv

    def __eq__(self, other):
        return isinstance(other, Transform4f) and np.array_equal(self.matrix, other.matrix)

    def __repr__(self):
        return f"Transform4f(\n{self.matrix}\n)"

    def flat16(self) -> np.ndarray:
        return np.ascontiguousarray(self.matrix, dtype=np.float64).reshape(16)


ScalarTransform4f = Transform4f


def apply_xml_ops(ops: Iterable[Transform4f], order: str = "mitsuba") -> Transform4f:
    """Compose the children of an XML ``<transform>`` element.

    ``order='mitsuba'``: each op is LEFT-multiplied in document order (M <- Op @ M), so
    ``<translate/><rotate/><scale/>`` yields S @ R @ T.  ``order='intended'``:
    right-multiplied (T @ R @ S), the composition the reference's own dict scene uses
    (/root/reference/USMain.py:69-71) and evidently what the XML author meant (Appendix D).
    """
    m = Transform4f()
    for op in ops:
        if order == "mitsuba":
            m = op @ m
        elif order == "intended":
            m = m @ op
        else:
            raise ValueError(f"unknown transform order {order!r}")
    return m
