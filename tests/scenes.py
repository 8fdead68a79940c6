"""Seeded scenes shared by the tests, the golden generator, smoke() and bench.py.

Camera = the reference's test-render defaults (core.clj:156-168): theta 135, dist 2.25, eye y 0.35,
target (0,-0.4,0); pass i at t = 0.333*i with scatter table seed 1000+i.
"""
from __future__ import annotations

import functools

import numpy as np

from raymarchcl_b200 import (compute_eyepos, generate_scatter_offsets, make_blob_volume,
                             make_gyroid_volume, make_render_option_buffers, make_terrain)


@functools.lru_cache(maxsize=8)
def _volume(kind: str, vres: int) -> np.ndarray:
    if kind == "gyroid":
        return make_gyroid_volume(vres)
    if kind == "terrain":
        return make_terrain(vres)
    if kind == "blob":
        return make_blob_volume(vres, ks=1)
    if kind == "dragon":
        return make_blob_volume(vres, ks=1, thin=True)
    if kind == "empty":
        return np.zeros((vres, vres, vres), dtype=np.uint8)
    if kind == "full":
        return np.full((vres, vres, vres), 255, dtype=np.uint8)
    raise ValueError(kind)


@functools.lru_cache(maxsize=128)
def _table(seed: int) -> np.ndarray:
    return generate_scatter_offsets(0x4000, seed)


def build_scene(vres=64, width=64, height=48, iters=1, mat="metal", dof=0.001, volume="gyroid",
                theta=135.0, dist=2.25, seed0=1000, **extra):
    """Returns (volume uint8[rz,ry,rx], [opts blobs], [scatter tables])."""
    vol = _volume(volume, int(vres))
    args = dict(width=width, height=height, vres=vres, iter=iters, mat=mat, dof=dof,
                eyepos=compute_eyepos(theta, dist, 0.35), targetpos=[0, -0.4, 0], **extra)
    opts = make_render_option_buffers(iters, args)
    mcs = [_table(seed0 + i) for i in range(iters)]
    return vol, opts, mcs


# name -> build_scene kwargs; small enough for the CPU checkers to finish in well under a second
GOLDEN_SCENES = {
    "c1_ao_64": dict(vres=64, width=64, height=64, iters=1, mat="ao"),
    "metal_64": dict(vres=64, width=64, height=36, iters=2, mat="metal"),
    "metal2_96": dict(vres=96, width=48, height=32, iters=2, mat="metal2", dof=0.025),
    "stripes_128": dict(vres=128, width=48, height=32, iters=1, mat="orange-stripes", theta=-45.0),
    "terrain_64": dict(vres=64, width=48, height=32, iters=1, mat="metal", volume="terrain"),
}
