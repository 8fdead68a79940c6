"""Input generators of the render op (host side, numpy).

Mirrors /root/reference/src/thi/ng/raymarchcl/generators.clj:

* ``generate_scatter_offsets`` <- ``generate-scatter-offsets`` (:8-16)  -- the mcSamples table
* ``make_gyroid_volume``       <- ``gyroid`` / ``make-gyroid-volume`` (:18-42)
* ``make_terrain``             <- ``make-terrain`` (:44-60)

plus ``make_blob_volume``, the declared synthetic stand-in for the bunny / dragon voxelisations
(no mesh asset exists in the reference tree), built the way
``meshvoxel/voxelize-ks`` (meshvoxel.clj:45-58) splats mesh vertices with a cubic kernel.

The reference seeds ``java.util.Random`` from ``System/nanoTime`` (:10), so its tables are not
reproducible; here the same generator (48-bit LCG, ``nextDouble`` = (next(26)<<27 + next(27)) / 2^53)
is restated with an explicit seed so a JVM host can produce bit-identical tables.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np

_LCG_A = np.uint64(0x5DEECE66D)
_LCG_C = np.uint64(0xB)
_MASK48 = np.uint64((1 << 48) - 1)


def java_random_next_doubles(seed: int, count: int) -> np.ndarray:
    """``count`` successive ``java.util.Random(seed).nextDouble()`` values (vectorised jump-ahead)."""
    n = 2 * count  # two next() calls per double
    with np.errstate(over="ignore"):
        s0 = (np.uint64(seed & ((1 << 64) - 1)) ^ _LCG_A) & _MASK48
        a_pow = np.empty(n + 1, dtype=np.uint64)
        a_pow[0] = 1
        a_pow[1:] = _LCG_A
        a_pow = np.multiply.accumulate(a_pow)            # a^k mod 2^64
        c_sum = np.zeros(n + 1, dtype=np.uint64)
        c_sum[1:] = np.add.accumulate(a_pow[:-1] * _LCG_C)  # c * (a^0 + ... + a^(k-1))
        states = (a_pow[1:] * s0 + c_sum[1:]) & _MASK48     # state after k = 1..n steps
    hi = (states[0::2] >> np.uint64(48 - 26)).astype(np.int64)
    lo = (states[1::2] >> np.uint64(48 - 27)).astype(np.int64)
    return ((hi << 27) + lo).astype(np.float64) * (1.0 / float(1 << 53))


def generate_scatter_offsets(num: int = 0x4000, seed: int = 0) -> np.ndarray:
    """``num`` random unit 4-vectors, flat float32[num*4] (generators.clj:8-16).

    Each component is ``float(2*nextDouble-1)``; the vector is scaled by 1/sqrt(x2+y2+z2+w2)
    computed in double and stored as float32 (the reference's ``:float`` CL buffer).
    """
    d = java_random_next_doubles(seed, 4 * num).reshape(num, 4)
    v = (2.0 * d - 1.0).astype(np.float32).astype(np.float64)
    m = 1.0 / np.sqrt(v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1] + v[:, 2] * v[:, 2] + v[:, 3] * v[:, 3])
    return (v * m[:, None]).astype(np.float32).reshape(-1)


def _res3(vres) -> Sequence[int]:
    return [int(vres)] * 3 if isinstance(vres, (int, np.integer)) else [int(v) for v in vres]


def make_gyroid_volume(vres) -> np.ndarray:
    """Sliced, striped gyroid shell volume, uint8[rz, ry, rx] (generators.clj:18-42).

    ``g = |cos x sin z + cos y sin x + cos z sin y| - 1`` at ``p*scl + (0.3875,0,0)``,
    ``scl = 0.01*512/rx``; only slabs with ``(z & 63) >= 32`` are filled: ``|0.2-g| < 0.05`` ->
    64 where ``(x & 63) < 32`` else 128; otherwise ``g > 0.35`` -> 255. Values {0,64,128,255}.
    """
    rx, ry, rz = _res3(vres)
    scl = 0.01 * (512.0 / rx)
    x = np.arange(rx, dtype=np.float64) * scl + 0.3875
    y = np.arange(ry, dtype=np.float64) * scl
    cx, sx = np.cos(x)[None, :], np.sin(x)[None, :]
    cy, sy = np.cos(y)[:, None], np.sin(y)[:, None]
    stripe = np.where((np.arange(rx) & 0x3F) < 32, 64, 128).astype(np.uint8)[None, :]
    vol = np.zeros((rz, ry, rx), dtype=np.uint8)
    for iz in range(rz):
        if (iz & 0x3F) < 32:
            continue
        z = iz * scl
        g = np.abs(cx * np.sin(z) + cy * sx + np.cos(z) * sy) - 1.0
        shell = np.abs(0.2 - g) < 0.05
        vol[iz] = np.where(shell, stripe, np.where(g > 0.35, 255, 0)).astype(np.uint8)
    return vol


def make_terrain(vres) -> np.ndarray:
    """Walls + bumpy pillars test volume, uint8[rz, ry, rx] (generators.clj:44-60)."""
    rx, ry, rz = _res3(vres)
    vol = np.zeros((rz, ry, rx), dtype=np.uint8)
    ymax = int(ry * 0.666)
    for z in range(4):
        vol[z, :ymax, :] = 64                      # idx = z*rxy + y*rx + x
        vol[:rx, :ymax, rx - z - 1] = 64           # idx = x*rxy + y*rx + (rx-z-1)
    xs = np.arange(rx)
    zs = np.arange(rz)
    dx = 16 - (xs % 32)
    dz = 16 - (zs % 32)
    r = dx[None, :] ** 2 + dz[:, None] ** 2
    h = (ry * (0.25 + 0.125 * (np.sin(zs * 0.02)[:, None] * np.cos(xs * 0.03)[None, :]))).astype(np.int64)
    yy = np.arange(ry)[None, :, None]
    fill = (r <= 121)[:, None, :] & (yy <= h[:, None, :])
    vol[fill] = 255
    return vol


def make_blob_volume(vres, n_points: int = 0, ks: int = 1, lobes: int = 5, thin: bool = False,
                     seed: int = 7) -> np.ndarray:
    """Closed blobby surface, point-splatted with a (2ks+1)^3 kernel of value 255.

    Declared synthetic STAND-IN for the Stanford bunny (512^3) / dragon (1024^3, ``thin=True``
    adds high-frequency ridges so the 5.3-voxel march step of the reference matters) configs:
    no mesh asset exists in the reference tree. Vertices lie on r(theta,phi) = R(1 + sum of a few
    spherical lobes); they are mapped into the grid like ``mesh-scale`` does (meshvoxel.clj:16-23)
    and splatted like ``voxelize-ks`` (meshvoxel.clj:45-58). Deterministic.
    """
    rx, ry, rz = _res3(vres)
    res = rx
    if n_points <= 0:
        n_points = int(14 * res * res)  # ~ surface area in voxels x oversampling -> closed shell
    rng = np.random.default_rng(seed)
    i = np.arange(n_points, dtype=np.float64) + 0.5
    phi = np.arccos(1.0 - 2.0 * i / n_points)
    theta = np.pi * (1.0 + 5.0 ** 0.5) * i
    r = np.ones(n_points)
    for k in range(lobes):
        f1, f2 = int(rng.integers(1, 5)), int(rng.integers(1, 5))
        r += 0.11 * np.sin(f1 * theta + rng.uniform(0, 6.28)) * np.sin(f2 * phi + rng.uniform(0, 6.28))
    if thin:
        r += 0.035 * np.sin(37.0 * theta) * np.sin(29.0 * phi) + 0.02 * np.sin(91.0 * phi)
    px = r * np.sin(phi) * np.cos(theta)
    py = r * np.cos(phi)
    pz = r * np.sin(phi) * np.sin(theta)
    pts = np.stack([px, py, pz], axis=1)
    lo = pts.min(axis=0)
    size = pts.max(axis=0) - lo
    md = size.max()
    off = 0.5 * res * (1.0 - size / md)
    q = (off + (pts - lo) * ((res - 2 * ks - 2) / md) + ks + 1).astype(np.int64)
    vol = np.zeros((rz, ry, rx), dtype=np.uint8)
    flat = vol.reshape(-1)
    for dz in range(-ks, ks + 1):
        for dy in range(-ks, ks + 1):
            for dx in range(-ks, ks + 1):
                x = np.clip(q[:, 0] + dx, 0, rx - 1)
                y = np.clip(q[:, 1] + dy, 0, ry - 1)
                z = np.clip(q[:, 2] + dz, 0, rz - 1)
                flat[(z * ry + y) * rx + x] = 255
    return vol
