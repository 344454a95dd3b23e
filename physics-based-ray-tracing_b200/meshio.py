"""OBJ / PLY triangle-mesh readers (host side of scene upload).

Restates what Mitsuba's ``obj`` / ``ply`` shape plugins do for the reference's assets
(SURVEY.md Appendix C.2): polygons are fan-triangulated ``(0,1,2),(0,2,3)`` (the cbox quads,
/root/reference/scenes/meshes/cbox_luminaire.obj:5), vertices are de-duplicated per
(position, normal) key so ``f v//vn`` corners keep their own normals
(/root/reference/TestRing/TestRing.obj:1014), meshes without normals get face normals at
intersection time.  Returns plain numpy arrays: ``v [nv,3] f64``, ``vn [nv,3] f64 | None``,
``idx [nt,3] u32``.
"""
from __future__ import annotations

import os
import struct
from typing import Optional, Tuple

import numpy as np

Mesh = Tuple[np.ndarray, Optional[np.ndarray], np.ndarray]


def load_obj(path: str) -> Mesh:
    pos, nrm = [], []
    key_to_index = {}
    out_v, out_n, tris = [], [], []
    any_normal = False
    with open(path, "r", errors="replace") as fh:
        for line in fh:
            if not line or line[0] == "#":
                continue
            tok = line.split()
            if not tok:
                continue
            tag = tok[0]
            if tag == "v":
                pos.append((float(tok[1]), float(tok[2]), float(tok[3])))
            elif tag == "vn":
                nrm.append((float(tok[1]), float(tok[2]), float(tok[3])))
            elif tag == "f":
                corner = []
                for c in tok[1:]:
                    parts = c.split("/")
                    vi = int(parts[0])
                    vi = vi - 1 if vi > 0 else len(pos) + vi
                    ni = -1
                    if len(parts) >= 3 and parts[2] != "":
                        ni = int(parts[2])
                        ni = ni - 1 if ni > 0 else len(nrm) + ni
                        any_normal = True
                    key = (vi, ni)
                    j = key_to_index.get(key)
                    if j is None:
                        j = len(out_v)
                        key_to_index[key] = j
                        out_v.append(pos[vi])
                        out_n.append(nrm[ni] if ni >= 0 else (0.0, 0.0, 0.0))
                    corner.append(j)
                for k in range(1, len(corner) - 1):
                    tris.append((corner[0], corner[k], corner[k + 1]))
    v = np.asarray(out_v, dtype=np.float64).reshape(-1, 3)
    vn = np.asarray(out_n, dtype=np.float64).reshape(-1, 3) if any_normal else None
    idx = np.asarray(tris, dtype=np.uint32).reshape(-1, 3)
    return v, vn, idx


_PLY_TYPES = {
    "char": "b", "int8": "b", "uchar": "B", "uint8": "B", "short": "h", "int16": "h",
    "ushort": "H", "uint16": "H", "int": "i", "int32": "i", "uint": "I", "uint32": "I",
    "float": "f", "float32": "f", "double": "d", "float64": "d",
}


def load_ply(path: str) -> Mesh:
    with open(path, "rb") as fh:
        data = fh.read()
    end = data.find(b"end_header")
    if end < 0 or not data.startswith(b"ply"):
        raise ValueError(f"{path}: not a PLY file")
    nl = data.find(b"\n", end)
    header = data[:end].decode("ascii", "replace").splitlines()
    body = data[nl + 1:]
    fmt = None
    elements = []  # (name, count, [(kind, name, types...)])
    for line in header:
        tok = line.split()
        if not tok:
            continue
        if tok[0] == "format":
            fmt = tok[1]
        elif tok[0] == "element":
            elements.append((tok[1], int(tok[2]), []))
        elif tok[0] == "property":
            if tok[1] == "list":
                elements[-1][2].append(("list", tok[4], tok[2], tok[3]))
            else:
                elements[-1][2].append(("scalar", tok[2], tok[1]))
    if fmt not in ("ascii", "binary_little_endian", "binary_big_endian"):
        raise ValueError(f"{path}: unsupported PLY format {fmt!r}")
    verts = None
    faces = []
    if fmt == "ascii":
        toks = body.split()
        cur = 0
        for name, count, props in elements:
            if name == "vertex":
                ncol = len(props)
                arr = np.array(toks[cur:cur + ncol * count], dtype=np.float64).reshape(count, ncol)
                cur += ncol * count
                verts = {p[1]: arr[:, i] for i, p in enumerate(props)}
            elif name == "face":
                for _ in range(count):
                    for p in props:
                        if p[0] == "list":
                            n = int(toks[cur]); cur += 1
                            ids = [int(x) for x in toks[cur:cur + n]]; cur += n
                            if p[1] in ("vertex_indices", "vertex_index"):
                                for k in range(1, n - 1):
                                    faces.append((ids[0], ids[k], ids[k + 1]))
                        else:
                            cur += 1
            else:
                for _ in range(count):
                    for p in props:
                        if p[0] == "list":
                            n = int(toks[cur]); cur += 1 + n
                        else:
                            cur += 1
    else:
        en = "<" if fmt == "binary_little_endian" else ">"
        off = 0
        for name, count, props in elements:
            all_scalar = all(p[0] == "scalar" for p in props)
            if all_scalar:
                dt = np.dtype([(p[1], en + _PLY_TYPES[p[2]]) for p in props])
                arr = np.frombuffer(body, dtype=dt, count=count, offset=off)
                off += dt.itemsize * count
                if name == "vertex":
                    verts = {p[1]: arr[p[1]].astype(np.float64) for p in props}
                continue
            # fast path: a single list property with a constant count (the common triangle soup)
            if name == "face" and len(props) == 1 and props[0][0] == "list":
                ct, it = _PLY_TYPES[props[0][2]], _PLY_TYPES[props[0][3]]
                csz, isz = struct.calcsize(ct), struct.calcsize(it)
                n0 = struct.unpack_from(en + ct, body, off)[0]
                rec = csz + n0 * isz
                if off + rec * count <= len(body):
                    dt = np.dtype([("n", en + ct), ("i", en + it, (n0,))])
                    arr = np.frombuffer(body, dtype=dt, count=count, offset=off)
                    if np.all(arr["n"] == n0):
                        ids = arr["i"].astype(np.int64)
                        for k in range(1, n0 - 1):
                            faces.append(np.stack([ids[:, 0], ids[:, k], ids[:, k + 1]], axis=1))
                        off += rec * count
                        continue
            for _ in range(count):
                for p in props:
                    if p[0] == "list":
                        ct, it = _PLY_TYPES[p[2]], _PLY_TYPES[p[3]]
                        n = struct.unpack_from(en + ct, body, off)[0]
                        off += struct.calcsize(ct)
                        ids = struct.unpack_from(en + str(n) + it, body, off)
                        off += struct.calcsize(it) * n
                        if name == "face" and p[1] in ("vertex_indices", "vertex_index"):
                            for k in range(1, n - 1):
                                faces.append((ids[0], ids[k], ids[k + 1]))
                    else:
                        off += struct.calcsize(_PLY_TYPES[p[2]])
    if verts is None:
        raise ValueError(f"{path}: no vertex element")
    v = np.stack([verts["x"], verts["y"], verts["z"]], axis=1).astype(np.float64)
    vn = None
    if all(k in verts for k in ("nx", "ny", "nz")):
        vn = np.stack([verts["nx"], verts["ny"], verts["nz"]], axis=1).astype(np.float64)
    if faces and isinstance(faces[0], np.ndarray):
        idx = np.concatenate(faces, axis=0).astype(np.uint32)
    else:
        idx = np.asarray(faces, dtype=np.uint32).reshape(-1, 3)
    return v, vn, idx


def load_mesh(path: str) -> Mesh:
    ext = os.path.splitext(path)[1].lower()
    if ext == ".obj":
        return load_obj(path)
    if ext == ".ply":
        return load_ply(path)
    raise ValueError(f"unsupported mesh format: {path}")


def heightfield_mesh(n: int, seed: int = 1234) -> Mesh:
    """BASELINE.json config 5: an n x n-vertex height field over [-1,1]^2,
    z = 0.05 sin(37x) cos(41y) + 0.02 fbm(x,y), 2 (n-1)^2 triangles (n = 2237 -> 9 999 392)."""
    xs = np.linspace(-1.0, 1.0, n)
    x, y = np.meshgrid(xs, xs, indexing="xy")
    z = 0.05 * np.sin(37.0 * x) * np.cos(41.0 * y)
    rng = np.random.Generator(np.random.PCG64(seed))
    amp, freq = 0.02, 4.0
    for _ in range(4):  # value-noise fbm on a hashed lattice
        g = int(freq) + 2
        lat = rng.random((g, g))
        fx, fy = (x + 1.0) * 0.5 * (g - 2), (y + 1.0) * 0.5 * (g - 2)
        ix, iy = np.minimum(fx.astype(np.int64), g - 2), np.minimum(fy.astype(np.int64), g - 2)
        tx, ty = fx - ix, fy - iy
        tx, ty = tx * tx * (3 - 2 * tx), ty * ty * (3 - 2 * ty)
        v00, v10 = lat[iy, ix], lat[iy, ix + 1]
        v01, v11 = lat[iy + 1, ix], lat[iy + 1, ix + 1]
        z = z + amp * ((v00 * (1 - tx) + v10 * tx) * (1 - ty) + (v01 * (1 - tx) + v11 * tx) * ty - 0.5)
        amp *= 0.5
        freq *= 2.0
    v = np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1)
    i = np.arange(n - 1)
    jj, ii = np.meshgrid(i, i, indexing="ij")
    a = (jj * n + ii).ravel()
    idx = np.concatenate([np.stack([a, a + 1, a + n + 1], 1), np.stack([a, a + n + 1, a + n], 1)], 0)
    return v, None, idx.astype(np.uint32)
