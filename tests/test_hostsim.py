"""The PRODUCTION per-pixel-sample routine (raymarchcl_b200/csrc/rm_scene_plain.cuh: fetch elision,
irrelevance culling, cut marches, recurrence jumps), compiled for the host by tests/hostsim, against
the oracle -- BIT FOR BIT on the fp32 accumulator (same libm on both sides), exact work counters.
No GPU needed: this is the check that every "exact" optimisation of the kernel really is exact."""
import numpy as np
import pytest

from oracle import build_oracle, refso
from tests.hostsim.sim import HostSim
from tests.scenes import GOLDEN_SCENES, build_scene

EXTRA_SCENES = {
    # ground plane inside the volume, camera inside the box, ragged grid, coarse cells
    "ground_inside": dict(vres=64, width=40, height=24, iters=2, mat="metal", groundY=0.3),
    "eye_inside": dict(vres=64, width=40, height=24, iters=1, mat="metal2", dist=0.6),
    "blob_96": dict(vres=96, width=40, height=24, iters=2, mat="metal", volume="blob"),
    "empty_32": dict(vres=32, width=24, height=16, iters=1, mat="ao", volume="empty"),
    "full_32": dict(vres=32, width=24, height=16, iters=1, mat="metal", volume="full"),
    # mid-size frames: every shortcut of the production routine fires thousands of times
    "gyroid256_metal_240x136": dict(vres=256, width=240, height=136, iters=2, mat="metal"),
    "gyroid256_ao_close": dict(vres=256, width=160, height=90, iters=2, mat="ao", theta=200.0, dist=1.2),
    "blob192_ground_in_box": dict(vres=192, width=160, height=90, iters=2, mat="metal", volume="blob", groundY=0.5),
    "stripes128_dof": dict(vres=128, width=160, height=90, iters=2, mat="orange-stripes", theta=30.0, dof=0.025),
}


@pytest.fixture(scope="module")
def sim():
    return HostSim()


@pytest.fixture(scope="module")
def orc():
    build_oracle.build(verbose=False)
    return refso.load("oracle")


@pytest.mark.parametrize("name", list(GOLDEN_SCENES) + list(EXTRA_SCENES))
@pytest.mark.parametrize("mode,cell_shift", [("production", 2), ("production", 3), ("counting", 2), ("bytes", 2),
                                             ("wave", 2), ("wave_counting", 2),  # wave*: the wavefront stages of rm_wave.cuh
                                             # fused*: the default kernel's routine (rm_scene_fused.cuh) over the 4-bit
                                             # distance map it reads from shared memory / over the byte map
                                             ("fused", 2), ("fused", 3), ("fused_counting", 2), ("fused_bytemap", 2),
                                             ("fused_bytemap_counting", 3)])
def test_host_build_of_the_kernel_routine_is_bit_identical_to_the_oracle(sim, orc, name, mode, cell_shift):
    kw = GOLDEN_SCENES.get(name) or EXTRA_SCENES[name]
    vol, opts, mcs = build_scene(**kw)
    w, h = kw["width"], kw["height"]
    ref, ref_cnt = orc.render_frame(vol, mcs, opts, w, h)
    px, cnt = sim.render_frame(vol, mcs, opts, w, h, mode=mode, cell_shift=cell_shift)
    assert np.array_equal(px.view(np.uint32), ref.view(np.uint32)), (
        f"{name}/{mode}: {(px.view(np.uint32) != ref.view(np.uint32)).any(axis=-1).sum()} pixels differ")
    if mode not in ("production", "wave", "fused", "fused_bytemap"):
        assert np.array_equal(cnt, ref_cnt)


def test_production_march_elides_most_fetches(sim):
    kw = GOLDEN_SCENES["metal_64"]
    vol, opts, mcs = build_scene(**kw)
    sim.stats(reset=True)
    _, cnt = sim.render_frame(vol, mcs, opts, kw["width"], kw["height"], mode="counting")
    sim.stats(reset=True)
    sim.render_frame(vol, mcs, opts, kw["width"], kw["height"], mode="production")
    st = sim.stats()
    assert 0 < st["lookups"] < int(cnt[0]) // 2  # the production march looks at far fewer samples than the reference fetches


@pytest.mark.parametrize("vres", [(96, 40, 130), (33, 70, 45)], ids=str)
@pytest.mark.parametrize("mode", ["production", "counting", "wave", "fused", "fused_counting"])
def test_ragged_grids(sim, orc, vres, mode):
    """Extents that are neither equal nor multiples of the brick / macro-cell edge."""
    from raymarchcl_b200 import compute_eyepos, generate_scatter_offsets, make_gyroid_volume, make_render_option_buffers
    vol = make_gyroid_volume(vres)
    w, h = 48, 32
    opts = make_render_option_buffers(2, dict(width=w, height=h, vres=list(vres), iter=2, mat="metal",
                                              eyepos=compute_eyepos(120.0, 2.0, 0.5), targetpos=[0, -0.3, 0]))
    mcs = [generate_scatter_offsets(0x4000, 5 + i) for i in range(2)]
    ref, ref_cnt = orc.render_frame(vol, mcs, opts, w, h)
    px, cnt = sim.render_frame(vol, mcs, opts, w, h, mode=mode, cell_shift=3 if mode == "wave" else 2)
    assert np.array_equal(px.view(np.uint32), ref.view(np.uint32))
    if mode in ("counting", "fused_counting"):
        assert np.array_equal(cnt, ref_cnt)


def test_large_volume_with_brick_sized_cells(sim, orc):
    """A 512^3 volume with the distance map at 4-voxel cells (128^3 cells: what rm_api.cu:auto_cell_shift picks up to
    1024^3 since the end of round 2; 8-voxel cells until then): default kernel's routine over the byte map and over the
    4-bit map, production and counting, bit for bit against the oracle."""
    kw = dict(vres=512, width=64, height=36, iters=1, mat="metal", volume="blob")
    vol, opts, mcs = build_scene(**kw)
    ref, ref_cnt = orc.render_frame(vol, mcs, opts, kw["width"], kw["height"])
    for mode, shift in (("fused_bytemap", 2), ("fused", 2), ("fused_counting", 2), ("fused_bytemap_counting", 2),
                        ("production", 2), ("fused_bytemap", 3)):
        px, cnt = sim.render_frame(vol, mcs, opts, kw["width"], kw["height"], mode=mode, cell_shift=shift)
        assert np.array_equal(px.view(np.uint32), ref.view(np.uint32)), (mode, shift)
        if mode.endswith("counting"):
            assert np.array_equal(cnt, ref_cnt), (mode, shift)
