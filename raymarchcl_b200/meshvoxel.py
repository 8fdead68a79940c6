"""Host mirror of the reference's mesh point-splat voxeliser (src/thi/ng/raymarchcl/meshvoxel.clj).

* ``load_mesh``     <- ``load-mesh`` (:12-14): binary STL -> the mesh's (unique) vertices
* ``mesh_scale``    <- ``mesh-scale`` (:16-23)
* ``voxelize``      <- ``voxelize`` (:60-69)
* ``voxelize_ks``   <- ``voxelize-ks`` (:45-58)

These numpy versions only PRODUCE INPUTS on the host (like generators.py); the product path for
large clouds is ``Renderer.voxelize_points`` -> ``rm_voxelize_points`` (CUDA, rm_generate.cu),
which leaves the volume resident on the device. ``voxelize-scatter`` (:25-43) draws from an
unseeded ``(rand)`` and ``make-heatmap`` (:71-83) needs an image library; neither is mirrored.
"""
from __future__ import annotations

import struct
from typing import Tuple

import numpy as np


def load_mesh(path: str) -> np.ndarray:
    """Vertices of a binary STL file, float32 [n, 3], duplicates removed (a thi.ng mesh holds each
    vertex once; the splat is idempotent, so the order does not matter)."""
    with open(path, "rb") as f:
        data = f.read()
    if len(data) < 84:
        raise ValueError(f"{path}: not a binary STL file (shorter than its header)")
    (ntri,) = struct.unpack_from("<I", data, 80)
    if len(data) < 84 + 50 * ntri:
        raise ValueError(f"{path}: truncated binary STL ({ntri} facets declared)")
    rec = np.frombuffer(data, dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]), count=ntri, offset=84)
    return np.unique(rec["v"].reshape(-1, 3), axis=0).astype(np.float32)


def save_stl(path: str, triangles: np.ndarray) -> None:
    """Write float32 triangles [n, 3, 3] as a binary STL (normals left zero) -- test helper."""
    tri = np.ascontiguousarray(triangles, dtype="<f4").reshape(-1, 3, 3)
    rec = np.zeros(len(tri), dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]))
    rec["v"] = tri
    with open(path, "wb") as f:
        f.write(b"raymarchcl_b200 binary STL".ljust(80, b" "))
        f.write(struct.pack("<I", len(tri)))
        f.write(rec.tobytes())


def mesh_scale(vertices: np.ndarray, res: int) -> Tuple[np.ndarray, np.ndarray, float]:
    """(p, off, s) of ``mesh-scale``: a vertex v maps to ``off + (v - p) * s`` (fp64)."""
    v = np.asarray(vertices, dtype=np.float32).astype(np.float64).reshape(-1, 3)
    p = v.min(axis=0)
    size = v.max(axis=0) - p
    md = size.max()
    with np.errstate(divide="ignore", invalid="ignore"):
        off = (0.5 * float(res)) * (1.0 - size / md)
        s = float(res) / md
    return p, off, float(s)


def _grid_coords(vertices: np.ndarray, res: int) -> np.ndarray:
    p, off, s = mesh_scale(vertices, res)
    v = np.asarray(vertices, dtype=np.float32).astype(np.float64).reshape(-1, 3)
    with np.errstate(invalid="ignore"):
        c = off + (v - p) * s
    c = np.where(np.isnan(c), 0.0, c)                      # (int NaN) = 0
    return np.trunc(np.clip(c, -2147483648.0, 2147483647.0)).astype(np.int64)  # (map int ..)


def voxelize(vertices: np.ndarray, res: int) -> np.ndarray:
    """One voxel of value 255 per vertex; vertices outside the grid are dropped. uint8 [res, res, res] (z, y, x)."""
    c = _grid_coords(vertices, res)
    vol = np.zeros((res, res, res), dtype=np.uint8)
    ok = ((c >= 0) & (c < res)).all(axis=1)
    c = c[ok]
    vol[c[:, 2], c[:, 1], c[:, 0]] = 255
    return vol


def voxelize_ks(vertices: np.ndarray, res: int, ks: int) -> np.ndarray:
    """A (2ks+1)^3 cube of value 255 around every vertex, clamped to the grid."""
    c = _grid_coords(vertices, res)
    vol = np.zeros((res, res, res), dtype=np.uint8)
    for dz in range(-ks, ks + 1):
        for dy in range(-ks, ks + 1):
            for dx in range(-ks, ks + 1):
                q = c + np.array([dx, dy, dz])
                ok = ((q >= 0) & (q < res)).all(axis=1)
                q = q[ok]
                vol[q[:, 2], q[:, 1], q[:, 0]] = 255
    return vol
