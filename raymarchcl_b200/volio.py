"""``.vox`` volume files: the on-disk format on the input side of the render op.

Mirrors /root/reference/src/thi/ng/raymarchcl/io.clj: ``save-volume`` (:9-17) and ``load-volume``
(:19-33). Layout: magic ``"VOXEL"`` (5 bytes), resx, resy, resz as big-endian int32
(``DataOutputStream.writeInt``), one byte element size (always 1), then resx*resy*resz raw bytes,
x fastest (index z*rx*ry + y*rx + x, generators.clj:37).
"""
from __future__ import annotations

import struct

import numpy as np

MAGIC = b"VOXEL"
HEADER_BYTES = 18


def save_volume(path: str, voxels: np.ndarray, res=None) -> None:
    """Write ``voxels`` (uint8[rz,ry,rx] or flat with ``res``) as a .vox file (io.clj:9-17).

    The reference writes ``res`` three times (cubic volumes only); non-cubic shapes are written
    with their true extents, which ``load_volume`` and the reference's reader both accept.
    """
    v = np.ascontiguousarray(voxels, dtype=np.uint8)
    if v.ndim == 3:
        rz, ry, rx = v.shape
    else:
        r = int(res)
        rx = ry = rz = r
        if v.size != r * r * r:
            raise ValueError(f"flat volume of {v.size} bytes does not match res {r}^3")
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack(">iiiB", rx, ry, rz, 1))
        f.write(v.tobytes())


def load_volume(path: str) -> np.ndarray:
    """Read a .vox file -> uint8[rz, ry, rx] (io.clj:19-33). The reference ignores the magic and
    the element-size byte; here a wrong magic, element size != 1 or a short file is an error."""
    with open(path, "rb") as f:
        head = f.read(HEADER_BYTES)
        if len(head) != HEADER_BYTES or head[:5] != MAGIC:
            raise ValueError(f"{path}: not a VOXEL file")
        rx, ry, rz, esize = struct.unpack(">iiiB", head[5:])
        if esize != 1 or min(rx, ry, rz) <= 0:
            raise ValueError(f"{path}: unsupported header res=({rx},{ry},{rz}) element size {esize}")
        data = np.fromfile(f, dtype=np.uint8, count=rx * ry * rz)
    if data.size != rx * ry * rz:
        raise ValueError(f"{path}: truncated, expected {rx * ry * rz} voxel bytes, got {data.size}")
    return data.reshape(rz, ry, rx)
