"""meshvoxel.clj (mesh-scale, voxelize, voxelize-ks, load-mesh): host mirror vs the C restatement
(CPU), and the CUDA voxeliser through the C ABI vs the C restatement (GPU), byte for byte.
The reference has no fixture for this path (oracle/meshvoxel_oracle.c: parity unpinned)."""
import numpy as np
import pytest

from oracle import build_oracle, refso
from raymarchcl_b200 import load_mesh, mesh_scale, voxelize, voxelize_ks
from raymarchcl_b200.meshvoxel import save_stl


def _clouds():
    rng = np.random.default_rng(11)
    sphere = rng.normal(size=(4000, 3))
    sphere /= np.linalg.norm(sphere, axis=1, keepdims=True)
    yield "sphere", (sphere * [1.0, 0.6, 0.3] + [5.0, -2.0, 0.25]).astype(np.float32)
    yield "flat", np.concatenate([rng.uniform(-3, 3, size=(500, 2)), np.full((500, 1), 7.0)], axis=1).astype(np.float32)
    yield "tiny", (rng.uniform(0, 1e-3, size=(64, 3)) - 40.0).astype(np.float32)
    yield "two_points", np.array([[0, 0, 0], [1, 2, 3]], dtype=np.float32)
    yield "single_point", np.array([[0.5, 0.25, -1.0]] * 3, dtype=np.float32)  # md = 0: every coordinate is NaN -> voxel 0


CLOUDS = dict(_clouds())


@pytest.fixture(scope="module")
def orc():
    build_oracle.build(verbose=False)
    return refso.load("oracle")


@pytest.mark.parametrize("name", list(CLOUDS))
@pytest.mark.parametrize("res,ks", [(32, -1), (32, 0), (48, 1), (40, 3)])
def test_host_mirror_equals_the_restatement(orc, name, res, ks):
    pts = CLOUDS[name]
    ref = orc.voxelize_points(pts, res, ks)
    got = voxelize(pts, res) if ks < 0 else voxelize_ks(pts, res, ks)
    assert np.array_equal(got, ref)
    if name == "sphere":
        assert 0 < int((ref == 255).sum()) < ref.size
    if ks < 0 and name == "two_points":
        assert int((ref == 255).sum()) == 1  # the max corner maps to index res on the longest axis and is dropped


def test_mesh_scale_centres_the_shorter_axes():
    p, off, s = mesh_scale(np.array([[0, 0, 0], [4, 2, 1]], dtype=np.float32), 64)
    assert np.allclose(p, 0) and s == 16.0 and np.allclose(off, [0.0, 16.0, 24.0])


def test_stl_round_trip(tmp_path):
    tri = np.array([[[0, 0, 0], [1, 0, 0], [0, 1, 0]], [[1, 0, 0], [0, 1, 0], [0, 0, 1]]], dtype=np.float32)
    path = str(tmp_path / "t.stl")
    save_stl(path, tri)
    v = load_mesh(path)
    assert v.shape == (4, 3) and {tuple(x) for x in v.tolist()} == {(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1)}
    with open(path, "r+b") as f:
        f.truncate(100)
    with pytest.raises(ValueError):
        load_mesh(path)


def test_non_finite_points_are_rejected(orc):
    with pytest.raises(ValueError):
        orc.voxelize_points(np.array([[0, 0, np.nan], [1, 1, 1]], dtype=np.float32), 16, 0)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CLOUDS))
@pytest.mark.parametrize("res,ks", [(32, -1), (48, 1), (96, 3)])
def test_cuda_voxeliser_equals_the_restatement(gpu_renderer, orc, name, res, ks):
    from raymarchcl_b200.renderer import Renderer
    pts = CLOUDS[name]
    ref = orc.voxelize_points(pts, res, ks)
    with Renderer(0) as r:
        r.voxelize_points(pts, res, ks)
        got = r.read_volume()
    assert got.shape == ref.shape and np.array_equal(got, ref)


@pytest.mark.gpu
def test_cuda_voxeliser_large_cloud_and_render(gpu_renderer, orc):
    """A 2M-vertex cloud at 256^3, then rendered straight from the device-resident result."""
    from raymarchcl_b200.renderer import Renderer
    from tests.scenes import build_scene
    rng = np.random.default_rng(5)
    t = rng.uniform(0, 2 * np.pi, size=2_000_000)
    u = rng.uniform(0, 2 * np.pi, size=t.size)
    pts = np.stack([(1 + 0.35 * np.cos(u)) * np.cos(t), 0.35 * np.sin(u), (1 + 0.35 * np.cos(u)) * np.sin(t)], axis=1).astype(np.float32)
    ref = orc.voxelize_points(pts, 256, 1)
    _, opts, mcs = build_scene(vres=256, width=96, height=64, iters=1, mat="metal")
    ref_px, ref_cnt = orc.render_frame(ref, mcs, opts, 96, 64)
    with Renderer(0) as r:
        r.voxelize_points(pts, 256, 1)
        assert np.array_equal(r.read_volume(), ref)
        r.clear_accum(96, 64)
        r.count_work(True)
        r.render_frame(opts, mcs)
        px = r.read_accum()
        st = r.stats()
        with pytest.raises(Exception):
            r.voxelize_points(np.array([[0, np.inf, 0]], dtype=np.float32), 16, 0)
    assert [st["steps"], st["taps"], st["outer_iters"]] == [int(x) for x in ref_cnt]
    assert (np.abs(px - ref_px) <= 2e-5 * np.maximum(1.0, np.abs(ref_px))).all()
