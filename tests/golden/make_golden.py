#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the REFERENCE'S OWN kernel text (oracle/_ref/libref_strict.so,
i.e. /root/reference/resources/renderer.cl compiled through oracle/clshim.h, strict fp32).

Run in the build container (where /root/reference exists):  python tests/golden/make_golden.py
The fixtures let machines without the reference (the GPU box) pin the oracle and the CUDA path.

frames.npz : per scene of tests/scenes.py:GOLDEN_SCENES -> fp32 RGBA accumulator, ARGB words,
             work counters {inner steps, occupancy taps, outer iterations}, sha1 of the inputs
kat.npz    : known-answer vectors of the path's building blocks: intersectsBox (:153-161),
             voxelLookup (:163-170), voxelNormal/voxelNormalSmooth (:180-203), distanceToScene
             (:209-237), raymarch (:239-257), initRenderState+cameraRayLookat (:456-476)
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import build_ref, refso  # noqa: E402
from tests.scenes import GOLDEN_SCENES, build_scene  # noqa: E402


def inputs_digest(vol, opts, mcs) -> str:
    h = hashlib.sha1()
    h.update(vol.tobytes())
    for o, m in zip(opts, mcs):
        h.update(o)
        h.update(m.tobytes())
    return h.hexdigest()


def main():
    if not build_ref.build(verbose=False):
        raise SystemExit("reference not available; cannot generate golden vectors")
    ref = refso.load("ref_strict")
    frames = {}
    for name, kw in GOLDEN_SCENES.items():
        vol, opts, mcs = build_scene(**kw)
        px, cnt = ref.render_frame(vol, mcs, opts, kw["width"], kw["height"])
        frames[name + "/accum"] = px
        frames[name + "/argb"] = ref.tonemap(px, opts[0])
        frames[name + "/counters"] = cnt
        frames[name + "/digest"] = np.array(inputs_digest(vol, opts, mcs))
        print(name, "counters", cnt, "mean", float(px[..., :3].mean()))
    np.savez_compressed(os.path.join(HERE, "frames.npz"), **frames)

    rng = np.random.default_rng(20261017)
    vol, opts, mcs = build_scene(vres=64, width=64, height=48, iters=1, mat="metal")
    o = opts[0]
    kat = {}
    # intersectsBox: random rays + axis-parallel (zero component -> inf) + origins inside the box
    n = 256
    P = rng.uniform(-2.5, 2.5, (n, 3)).astype(np.float32)
    D = rng.normal(size=(n, 3)).astype(np.float32)
    D /= np.linalg.norm(D, axis=1, keepdims=True).astype(np.float32)
    D[:24, 0] = 0.0
    D[8:32, 1] = 0.0
    P[32:96] = rng.uniform(-0.98, 0.98, (64, 3)).astype(np.float32)
    bmin = np.full(3, -0.99, np.float32)
    bmax = np.full(3, 0.99, np.float32)
    kat["box/p"], kat["box/d"] = P, D
    kat["box/out"] = np.array([ref.intersects_box(bmin, bmax, P[i], D[i]) for i in range(n)], np.float32)
    # voxelLookup incl. the truncation quirk for p in (-1/res, 0) and the upper edge
    Q = rng.uniform(-0.05, 1.05, (n, 3)).astype(np.float32)
    Q[:8] = np.float32(-0.5 / 64)
    Q[8:16] = np.float32(1.0)
    Q[16:24] = np.nextafter(np.float32(1.0), np.float32(0.0))
    kat["lookup/p"] = Q
    kat["lookup/out"] = np.array([ref.voxel_lookup(vol, o, Q[i]) for i in range(n)], np.int32)
    # normals at solid voxels, border voxels and random cells
    solid = np.argwhere(vol > 32)
    pick = solid[rng.choice(len(solid), 96, replace=False)][:, ::-1].astype(np.int32)  # (x,y,z)
    cells = np.concatenate([pick, rng.integers(-1, 65, (32, 3)).astype(np.int32),
                            np.array([[0, 0, 0], [63, 63, 63], [0, 63, 32], [-1, 5, 5]], np.int32)])
    kat["normal/q"] = cells
    kat["normal/six"] = np.stack([ref.voxel_normal(vol, o, q, False) for q in cells])
    kat["normal/smooth"] = np.stack([ref.voxel_normal(vol, o, q, True) for q in cells])
    # distanceToScene / raymarch from camera-like and interior origins
    RO = np.concatenate([rng.uniform(-2.2, 2.2, (96, 3)), rng.uniform(-0.9, 0.9, (96, 3))]).astype(np.float32)
    RD = rng.normal(size=(192, 3)).astype(np.float32)
    RD /= np.linalg.norm(RD, axis=1, keepdims=True).astype(np.float32)
    RD[:96] = (-RO[:96] / np.linalg.norm(RO[:96], axis=1, keepdims=True) + 0.15 * RD[:96]).astype(np.float32)
    RD /= np.linalg.norm(RD, axis=1, keepdims=True).astype(np.float32)
    kat["scene/ro"], kat["scene/rd"] = RO, RD
    kat["scene/dist192s"] = np.stack([ref.distance_to_scene(vol, o, RO[i], RD[i], 192, True) for i in range(192)])
    kat["scene/dist96"] = np.stack([ref.distance_to_scene(vol, o, RO[i], RD[i], 96, False) for i in range(192)])
    kat["scene/march_s"] = np.stack([ref.raymarch(vol, o, RO[i], RD[i], 30.0, 128, True) for i in range(192)])
    kat["scene/march"] = np.stack([ref.raymarch(vol, o, RO[i], RD[i], 2.5, 128, False) for i in range(192)])
    ids = rng.integers(0, 64 * 48, 128).astype(np.int32)
    kat["camera/id"] = ids
    kat["camera/out"] = np.stack([ref.camera_ray(o, mcs[0], int(i)) for i in ids])
    kat["opts_offsets"] = ref.opts_offsets()
    np.savez_compressed(os.path.join(HERE, "kat.npz"), **kat)
    for f in ("frames.npz", "kat.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
