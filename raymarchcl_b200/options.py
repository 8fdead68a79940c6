"""Render options: the host-side parameter surface of the render op.

Mirrors the reference's Clojure host for this path:

* ``PRESETS``           <- ``materials/presets``           (/root/reference/src/thi/ng/raymarchcl/materials.clj:3-76)
* ``render_options``    <- ``core/render-options``         (core.clj:28-74)
* ``encode_render_opts``<- ``sg/encode`` of ``TRenderOpts`` (core.clj:101-105; struct text renderer.cl:35-78)
* ``compute_eyepos``    <- ``core/compute-eyepos``         (core.clj:150-152)

The encoded 544-byte blob is the ENTIRE parameter surface the device code sees; its layout
follows OpenCL alignment rules (float3/int3 occupy 16 bytes) and is cross-checked in the tests
against ``offsetof`` in the compiled reference text. thi.ng/structgen 0.2.1 (the reference's
encoder) is not in the reference tree, so the layout is pinned by the kernel's own struct text.
"""
from __future__ import annotations

import math
import struct
from typing import Any, Dict, Iterable, List, Mapping, Optional, Sequence

import numpy as np

OPTS_BYTES = 544
TABLE_ENTRIES = 0x4000  # renderer.cl:143 hard-codes the 0x3fff mask; core.clj:138 allocates 0x4000

# (name, byte offset, kind) -- kind: f3 float3(16B) | i4 int4 | i2 int2 | f float | i int | u8 uchar
#                                    f4x4 float4[4] | mat4 TMaterial[4]
OPTS_FIELDS = [
    ("eyePos", 0, "f3"), ("targetPos", 16, "f3"), ("up", 32, "f3"), ("voxelBounds", 48, "f3"),
    ("voxelBounds2", 64, "f3"), ("voxelBoundsMin", 80, "f3"), ("voxelBoundsMax", 96, "f3"),
    ("invVoxelScale", 112, "f3"), ("skyColor1", 128, "f3"), ("skyColor2", 144, "f3"),
    ("voxelRes", 160, "i4"), ("resolution", 176, "i2"), ("invAspect", 184, "f"), ("time", 188, "f"),
    ("fov", 192, "f"), ("maxIter", 196, "i"), ("maxVoxelIter", 200, "i"), ("maxDist", 204, "f"),
    ("startDist", 208, "f"), ("eps", 212, "f"), ("aoIter", 216, "i"), ("aoStepDist", 220, "f"),
    ("aoAmp", 224, "f"), ("voxelSize", 228, "f"), ("groundY", 232, "f"), ("shadowIter", 236, "i"),
    ("reflectIter", 240, "i"), ("shadowBias", 244, "f"), ("lightScatter", 248, "f"),
    ("minLightAtt", 252, "f"), ("gamma", 256, "f"), ("exposure", 260, "f"), ("dof", 264, "f"),
    ("frameBlend", 268, "f"), ("fogPow", 272, "f"), ("flareAmp", 276, "f"), ("mcTableLength", 280, "i"),
    ("isoVal", 284, "u8"), ("numLights", 285, "u8"), ("lightPos", 288, "f4x4"),
    ("lightColor", 352, "f4x4"), ("materials", 416, "mat4"),
]
OPTS_OFFSETS = {name: off for name, off, _ in OPTS_FIELDS}

# materials.clj:3-76, values verbatim (they are inputs to the device path)
PRESETS: Dict[str, Dict[str, Any]] = {
    "orange-stripes": {
        "lightColor": [[28, 18, 8, 0], [8, 18, 28, 0]],
        "lightPos": [[-2, 0, -2, 0], [2, 0, 2, 0]],
        "materials": [
            {"albedo": [1.0, 1.0, 1.0, 1.0], "r0": 0.1, "smoothness": 0.9},
            {"albedo": [4.9, 0.9, 0.05, 1.0], "r0": 0.01, "smoothness": 0.5},
            {"albedo": [1.9, 1.9, 1.9, 1.0], "r0": 0.01, "smoothness": 0.4},
            {"albedo": [0.9, 0.9, 0.9, 1.0], "r0": 0.8, "smoothness": 0.1},
        ],
        "numLights": 2, "aoAmp": 0.25, "reflectIter": 1,
    },
    "metal": {
        "lightColor": [[28, 18, 8, 0], [16, 36, 56, 0]],
        "lightPos": [[0, 2, 0, 0], [3, 0, 3, 0]],
        "materials": [
            {"albedo": [0.01, 0.01, 0.01, 1.0], "r0": 0.1, "smoothness": 0.5},
            {"albedo": [1.9, 1.9, 1.9, 1.0], "r0": 0.1, "smoothness": 0.5},
            {"albedo": [0.25, 0.27, 0.5, 1.0], "r0": 0.7, "smoothness": 0.1},
            {"albedo": [1.0, 1.0, 1.0, 1.0], "r0": 0.2, "smoothness": 0.1},
        ],
        "numLights": 2, "aoAmp": 0.25, "reflectIter": 3,
    },
    "metal2": {
        "lightColor": [[28, 18, 8, 0], [8, 18, 28, 0]],
        "lightPos": [[-2, 0, -2, 0], [2, 0, 2, 0]],
        "materials": [
            {"albedo": [0.0, 0.0, 0.0, 1.0], "r0": 0.1, "smoothness": 0.9},
            {"albedo": [1.0, 1.01, 1.075, 1.0], "r0": 0.4, "smoothness": 0.7},
            {"albedo": [1.9, 1.9, 1.9, 1.0], "r0": 0.4, "smoothness": 0.5},
            {"albedo": [0.9, 0.9, 0.9, 1.0], "r0": 0.75, "smoothness": 0.2},
        ],
        "numLights": 2, "aoAmp": 0.25, "reflectIter": 3,
    },
    "ao": {
        "lightColor": [[50, 50, 50, 0]],
        "materials": [{"albedo": [1.0, 1.0, 1.0, 1.0], "r0": 0.0, "smoothness": 1.0}] * 4,
        "numLights": 1, "aoAmp": 0.25, "reflectIter": 0,
    },
}


def compute_eyepos(theta_deg: float, dist: float, y: float) -> List[float]:
    """Orbit camera: (0, y, dist) rotated about +Y by theta (core.clj:150-152).

    thi.ng/geom 0.0.803 (not in the reference tree) supplies rotate-y; the sign convention used
    here is x = d sin(theta), z = d cos(theta). It only moves the bench camera;
    eyePos is an input of the device path.
    """
    t = math.radians(theta_deg)
    return [dist * math.sin(t), y, dist * math.cos(t)]


def _pad_lights(rows: Sequence[Sequence[float]]) -> List[List[float]]:
    out = [list(map(float, r)) + [0.0] * (4 - len(r)) for r in rows]
    while len(out) < 4:
        out.append([0.0, 0.0, 0.0, 0.0])
    return out[:4]


def render_options(opts: Mapping[str, Any]) -> Dict[str, Any]:
    """Field map of TRenderOpts for one pass (core.clj:28-74).

    Honoured caller keys (core.clj:29): width height vres t iter eyepos mat fov dof targetpos gamma
    groundY voxelSize. Everything else is the fixed default, then overridden by the material preset
    (unknown ``mat`` falls back to ``ao``, core.clj:74).
    """
    width, height = int(opts["width"]), int(opts["height"])
    vres = opts["vres"]
    vres = [int(vres)] * 3 if isinstance(vres, (int, np.integer)) else [int(v) for v in vres]
    it = opts.get("iter", 1)
    eps, clip = 0.005, 0.99

    def _or(key, default):
        v = opts.get(key)
        return default if v is None else v

    fields: Dict[str, Any] = {
        "aoAmp": 0.2, "aoIter": 5, "aoStepDist": 0.05,
        "dof": _or("dof", 0.001),
        "eps": eps, "exposure": 3.5,
        "eyePos": list(_or("eyepos", [2, 0, 2])),
        "flareAmp": 0.015, "fogPow": 0.05,
        "fov": math.radians(_or("fov", 90)),
        "frameBlend": 1.0 / it,
        "gamma": _or("gamma", 1.5),
        "groundY": _or("groundY", 1.05),
        "invAspect": float(np.float32(height / width)),
        "invVoxelScale": [0.5, 0.5, 0.5],
        "isoVal": 32,
        "lightColor": [[50, 50, 50]],
        "lightPos": [[-2, 0, -2, 0], [2, 0, 2, 0]],
        "lightScatter": 0.2, "maxDist": 30, "maxIter": 128, "maxVoxelIter": 192, "minLightAtt": 0.0,
        "numLights": 2, "reflectIter": 0,
        "resolution": [width, height],
        "shadowBias": 0.1, "shadowIter": 128,
        "skyColor1": [1.8, 1.8, 1.9], "skyColor2": [0.1, 0.1, 0.1],
        "startDist": 0.0,
        "targetPos": list(_or("targetpos", [0, -0.15, 0])),
        "time": _or("t", 0.0),
        "up": [0, 1, 0],
        "voxelBounds": [1, 1, 1], "voxelBounds2": [2, 2, 2],
        "voxelBoundsMax": [clip, clip, clip], "voxelBoundsMin": [-clip, -clip, -clip],
        "voxelRes": vres + [vres[0] * vres[1]],
        "voxelSize": _or("voxelSize", 1.0 / vres[0]),
    }
    preset = PRESETS.get(str(opts.get("mat")).lstrip(":"), PRESETS["ao"])
    fields.update({k: v for k, v in preset.items()})
    return fields


def encode_render_opts(fields: Mapping[str, Any]) -> bytes:
    """Pack a field map into the 544-byte TRenderOpts blob (little-endian, OpenCL layout)."""
    buf = bytearray(OPTS_BYTES)
    for name, off, kind in OPTS_FIELDS:
        v = fields.get(name)
        if v is None:
            continue  # e.g. mcTableLength: never set by the host, never read by the kernel
        if kind == "f3":
            struct.pack_into("<3f", buf, off, *[float(x) for x in list(v)[:3]])
        elif kind == "i4":
            struct.pack_into("<4i", buf, off, *[int(x) for x in v])
        elif kind == "i2":
            struct.pack_into("<2i", buf, off, *[int(x) for x in v])
        elif kind == "f":
            struct.pack_into("<f", buf, off, float(v))
        elif kind == "i":
            struct.pack_into("<i", buf, off, int(v))
        elif kind == "u8":
            struct.pack_into("<B", buf, off, int(v) & 0xFF)
        elif kind == "f4x4":
            for i, row in enumerate(_pad_lights(v)):
                struct.pack_into("<4f", buf, off + 16 * i, *row)
        elif kind == "mat4":
            for i, m in enumerate(list(v)[:4]):
                alb = [float(x) for x in m["albedo"]] + [0.0] * (4 - len(m["albedo"]))
                struct.pack_into("<4f", buf, off + 32 * i, *alb[:4])
                struct.pack_into("<2f", buf, off + 32 * i + 16, float(m["r0"]), float(m["smoothness"]))
        else:  # pragma: no cover
            raise ValueError(kind)
    return bytes(buf)


def decode_render_opts(blob: bytes) -> Dict[str, Any]:
    """Inverse of :func:`encode_render_opts` (used by tests and error messages)."""
    if len(blob) != OPTS_BYTES:
        raise ValueError(f"TRenderOpts blob must be {OPTS_BYTES} bytes, got {len(blob)}")
    out: Dict[str, Any] = {}
    for name, off, kind in OPTS_FIELDS:
        if kind == "f3":
            out[name] = list(struct.unpack_from("<3f", blob, off))
        elif kind == "i4":
            out[name] = list(struct.unpack_from("<4i", blob, off))
        elif kind == "i2":
            out[name] = list(struct.unpack_from("<2i", blob, off))
        elif kind == "f":
            out[name] = struct.unpack_from("<f", blob, off)[0]
        elif kind == "i":
            out[name] = struct.unpack_from("<i", blob, off)[0]
        elif kind == "u8":
            out[name] = blob[off]
        elif kind == "f4x4":
            out[name] = [list(struct.unpack_from("<4f", blob, off + 16 * i)) for i in range(4)]
        elif kind == "mat4":
            out[name] = [
                {"albedo": list(struct.unpack_from("<4f", blob, off + 32 * i)),
                 "r0": struct.unpack_from("<f", blob, off + 32 * i + 16)[0],
                 "smoothness": struct.unpack_from("<f", blob, off + 32 * i + 20)[0]}
                for i in range(4)]
    return out


def make_render_option_buffers(n: int, opts: Mapping[str, Any], t_step: float = 0.333) -> List[bytes]:
    """One encoded blob per pass, pass i at time i*0.333 (core.clj:99-106).

    ``update-render-option-buffer`` (core.clj:108-117) uses 0.3333; pass ``t_step=0.3333`` for it.
    """
    base = dict(opts)
    base["iter"] = n if opts.get("iter") is None else opts["iter"]
    return [encode_render_opts(render_options({**base, "t": i * t_step})) for i in range(n)]
