"""The oracle (oracle/rm_oracle.c) pinned against the reference: (a) the committed golden vectors
generated from the reference's own kernel text, (b) when oracle/_ref is built, the reference itself
on further seeded scenes. Bar: BIT-identical fp32 accumulators, ARGB words and work counters."""
import hashlib
import os

import numpy as np
import pytest

from tests.scenes import GOLDEN_SCENES, build_scene

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _digest(vol, opts, mcs):
    h = hashlib.sha1()
    h.update(vol.tobytes())
    for o, m in zip(opts, mcs):
        h.update(o)
        h.update(m.tobytes())
    return h.hexdigest()


@pytest.fixture(scope="module")
def frames():
    return np.load(os.path.join(GOLD, "frames.npz"))


@pytest.fixture(scope="module")
def kat():
    return np.load(os.path.join(GOLD, "kat.npz"))


@pytest.mark.parametrize("name", sorted(GOLDEN_SCENES))
def test_oracle_matches_golden_frames(oracle, frames, name):
    kw = GOLDEN_SCENES[name]
    vol, opts, mcs = build_scene(**kw)
    assert _digest(vol, opts, mcs) == str(frames[name + "/digest"]), "input generators drifted from the fixtures"
    px, cnt = oracle.render_frame(vol, mcs, opts, kw["width"], kw["height"])
    assert np.array_equal(px.view(np.uint32), frames[name + "/accum"].view(np.uint32))
    assert np.array_equal(cnt, frames[name + "/counters"])
    assert np.array_equal(oracle.tonemap(px, opts[0]), frames[name + "/argb"])


def _kat_scene():
    return build_scene(vres=64, width=64, height=48, iters=1, mat="metal")


def _same(a, b):
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    return np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_oracle_kat_intersects_box(oracle, kat):
    bmin, bmax = np.full(3, -0.99, np.float32), np.full(3, 0.99, np.float32)
    out = [oracle.intersects_box(bmin, bmax, p, d) for p, d in zip(kat["box/p"], kat["box/d"])]
    assert _same(out, kat["box/out"])


def test_oracle_kat_voxel_lookup(oracle, kat):
    vol, opts, _ = _kat_scene()
    out = [oracle.voxel_lookup(vol, opts[0], p) for p in kat["lookup/p"]]
    assert np.array_equal(np.array(out, np.int32), kat["lookup/out"])
    # truncation toward zero: p in (-1/res, 0) addresses cell 0, it is NOT outside (renderer.cl:165)
    assert kat["lookup/out"][0] >= 0


def test_oracle_kat_normals(oracle, kat):
    vol, opts, _ = _kat_scene()
    six = np.stack([oracle.voxel_normal(vol, opts[0], q, False) for q in kat["normal/q"]])
    smooth = np.stack([oracle.voxel_normal(vol, opts[0], q, True) for q in kat["normal/q"]])
    assert _same(six, kat["normal/six"])
    assert _same(smooth, kat["normal/smooth"])


def test_oracle_kat_scene_distance_and_raymarch(oracle, kat):
    vol, opts, _ = _kat_scene()
    o = opts[0]
    ro, rd = kat["scene/ro"], kat["scene/rd"]
    assert _same(np.stack([oracle.distance_to_scene(vol, o, p, d, 192, True) for p, d in zip(ro, rd)]), kat["scene/dist192s"])
    assert _same(np.stack([oracle.distance_to_scene(vol, o, p, d, 96, False) for p, d in zip(ro, rd)]), kat["scene/dist96"])
    assert _same(np.stack([oracle.raymarch(vol, o, p, d, 30.0, 128, True) for p, d in zip(ro, rd)]), kat["scene/march_s"])
    assert _same(np.stack([oracle.raymarch(vol, o, p, d, 2.5, 128, False) for p, d in zip(ro, rd)]), kat["scene/march"])


def test_oracle_kat_camera(oracle, kat):
    _, opts, mcs = _kat_scene()
    out = np.stack([oracle.camera_ray(opts[0], mcs[0], int(i)) for i in kat["camera/id"]])
    assert _same(out, kat["camera/out"])


# ---- against the reference itself (only where oracle/_ref has been built) ----

LIVE = [
    dict(vres=64, width=96, height=64, iters=2, mat="ao"),
    dict(vres=128, width=80, height=45, iters=3, mat="metal"),
    dict(vres=48, width=40, height=40, iters=1, mat="metal2", volume="terrain"),
    dict(vres=32, width=32, height=32, iters=1, mat="metal", volume="empty"),
    dict(vres=32, width=32, height=32, iters=1, mat="orange-stripes", volume="full"),
    dict(vres=96, width=64, height=36, iters=2, mat="metal", volume="blob"),
]


@pytest.mark.parametrize("kw", LIVE, ids=lambda k: f"{k.get('volume', 'gyroid')}{k['vres']}_{k['mat']}")
def test_oracle_bit_identical_to_reference(oracle, ref_strict, kw):
    vol, opts, mcs = build_scene(**kw)
    pr, cr = ref_strict.render_frame(vol, mcs, opts, kw["width"], kw["height"])
    po, co = oracle.render_frame(vol, mcs, opts, kw["width"], kw["height"])
    assert np.array_equal(pr.view(np.uint32), po.view(np.uint32))
    assert np.array_equal(cr, co)
    assert np.array_equal(ref_strict.tonemap(pr, opts[0]), oracle.tonemap(po, opts[0]))


def test_reference_struct_layout(ref_strict, kat):
    from raymarchcl_b200.options import OPTS_BYTES, OPTS_FIELDS
    assert ref_strict.sizeof_opts() == OPTS_BYTES == 544
    assert [int(x) for x in ref_strict.opts_offsets()] == [off for _, off, _ in OPTS_FIELDS]
    assert [int(x) for x in kat["opts_offsets"]] == [off for _, off, _ in OPTS_FIELDS]
