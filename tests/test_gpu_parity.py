"""CUDA path (through the C ABI) vs the oracle on identical seeded inputs, and vs the committed
golden fixtures generated from the reference's own kernel text.

Bars:
* work counters (inner steps, occupancy taps, outer iterations): EXACT -- geometry is bit-exact
  by construction (no FMA contraction, IEEE div/sqrt, pinned casts).
* fp32 RGBA accumulator: colour goes through exp/exp2/pow, whose CUDA and glibc implementations
  differ in the last ulp. Stated tolerance: >= 99 % of pixels within
  1e-4*max(1,|ref|) per channel and mean abs error <= 2e-3. Achieved and asserted here: EVERY
  pixel within 2e-5*max(1,|ref|).
* ARGB words: <= 1 LSB per channel on every pixel, identical on >= 99.5 % of pixels.
"""
import os

import numpy as np
import pytest

from tests.scenes import GOLDEN_SCENES, build_scene

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")

TIGHT = 2e-5


def check_frame(px_gpu, px_ref, argb_gpu, argb_ref):
    assert px_gpu.shape == px_ref.shape
    assert not np.isnan(px_gpu).any()
    tol = TIGHT * np.maximum(1.0, np.abs(px_ref))
    err = np.abs(px_gpu.astype(np.float64) - px_ref.astype(np.float64))
    bad = (err > tol).any(axis=-1)
    assert bad.mean() == 0.0, f"{bad.sum()} of {bad.size} pixels outside {TIGHT} rel; max err {err.max()}"
    assert err.mean() <= 2e-3
    d = np.zeros(argb_ref.shape, dtype=np.int64)
    for sh in (16, 8, 0):
        d = np.maximum(d, np.abs(((argb_gpu >> sh) & 255).astype(np.int64) - ((argb_ref >> sh) & 255).astype(np.int64)))
    assert (argb_gpu >> 24 == 0xFF).all()
    assert d.max() <= 1
    assert (d == 0).mean() >= 0.995


def render_gpu(r, vol, opts, mcs, w, h, fused=True, count=True):
    r.set_tile_shard(0, 1, 32, 32)
    r.set_volume(vol)
    r.clear_accum(w, h)
    r.reset_stats()
    r.count_work(count)
    if fused:
        r.render_frame(opts, mcs)
    else:
        for o, m in zip(opts, mcs):
            r.render_pass(o, m)
    st = r.stats()
    return r.read_accum(), r.tonemap(opts[0]), np.array([st["steps"], st["taps"], st["outer_iters"]], np.uint64)


@pytest.mark.parametrize("kernel", [0, 1, 3, 4], ids=["fast", "plain", "wave", "bricks"])
@pytest.mark.parametrize("name", sorted(GOLDEN_SCENES))
def test_gpu_matches_reference_golden(gpu_renderer, name, kernel):
    gold = np.load(os.path.join(GOLD, "frames.npz"))
    kw = GOLDEN_SCENES[name]
    vol, opts, mcs = build_scene(**kw)
    gpu_renderer.set_option(2, kernel)
    px, argb, cnt = render_gpu(gpu_renderer, vol, opts, mcs, kw["width"], kw["height"])
    assert np.array_equal(cnt, gold[name + "/counters"])
    check_frame(px, gold[name + "/accum"], argb, gold[name + "/argb"])


SCENES = [
    dict(vres=64, width=256, height=256, iters=1, mat="ao"),                       # BASELINE config 1
    dict(vres=256, width=320, height=180, iters=2, mat="metal"),                   # config 2, reduced frame
    dict(vres=128, width=200, height=120, iters=2, mat="metal2", dof=0.025),       # DoF + 3 bounces
    dict(vres=96, width=131, height=77, iters=1, mat="orange-stripes", theta=-45), # ragged frame
    dict(vres=64, width=64, height=64, iters=1, mat="metal", volume="empty"),
    dict(vres=32, width=64, height=64, iters=1, mat="metal", volume="full"),
    dict(vres=64, width=96, height=64, iters=1, mat="metal", volume="terrain"),
    dict(vres=160, width=160, height=90, iters=1, mat="metal", volume="blob"),     # bunny stand-in, non-pow2 res
]


@pytest.mark.parametrize("kernel", [0, 1, 3, 4], ids=["fast", "plain", "wave", "bricks"])
@pytest.mark.parametrize("kw", SCENES, ids=lambda k: f"{k.get('volume', 'gyroid')}{k['vres']}_{k['mat']}_{k['width']}x{k['height']}")
def test_gpu_matches_oracle(gpu_renderer, oracle, kw, kernel):
    vol, opts, mcs = build_scene(**kw)
    w, h = kw["width"], kw["height"]
    ref_px, ref_cnt = oracle.render_frame(vol, mcs, opts, w, h)
    gpu_renderer.set_option(2, kernel)
    px, argb, cnt = render_gpu(gpu_renderer, vol, opts, mcs, w, h)
    assert np.array_equal(cnt, ref_cnt), (cnt, ref_cnt)
    check_frame(px, ref_px, argb, oracle.tonemap(ref_px, opts[0]))


def test_pass_by_pass_equals_fused_frame(gpu_renderer):
    kw = dict(vres=64, width=96, height=64, iters=3, mat="metal")
    vol, opts, mcs = build_scene(**kw)
    a, argb_a, ca = render_gpu(gpu_renderer, vol, opts, mcs, 96, 64, fused=True)
    b, argb_b, cb = render_gpu(gpu_renderer, vol, opts, mcs, 96, 64, fused=False)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert np.array_equal(argb_a, argb_b) and np.array_equal(ca, cb)


def test_resident_path_equals_host_path(gpu_renderer):
    kw = dict(vres=64, width=96, height=64, iters=2, mat="metal")
    vol, opts, mcs = build_scene(**kw)
    a, _, _ = render_gpu(gpu_renderer, vol, opts, mcs, 96, 64)
    gpu_renderer.clear_accum(96, 64)
    gpu_renderer.upload_passes(opts, mcs)
    gpu_renderer.render_resident(0, 2)
    assert np.array_equal(a.view(np.uint32), gpu_renderer.read_accum().view(np.uint32))


def test_blend_weights_property(gpu_renderer):
    """N passes with frameBlend 1/N are NOT an average: they leave total weight 1-(1-1/N)^N
    (renderer.cl:492). Submitting the SAME pass (same opts, same table) N times
    makes every pass colour identical, so the accumulator must be colour * (1-(1-1/N)^N)."""
    kw = dict(vres=64, width=64, height=32, iters=16, mat="metal")
    vol, opts, mcs = build_scene(**kw)
    one, _, _ = render_gpu(gpu_renderer, vol, opts[:1], mcs[:1], 64, 32, count=False)
    px, _, _ = render_gpu(gpu_renderer, vol, [opts[0]] * 16, [mcs[0]] * 16, 64, 32, count=False)
    colour = one[..., :3].astype(np.float64) * 16.0   # one pass blended into zero with weight 1/16
    wsum = 1.0 - (1.0 - 1.0 / 16) ** 16
    assert np.allclose(px[..., :3], colour * wsum, rtol=1e-5, atol=1e-6)
    assert (px[..., 3] == 1.0).all()


@pytest.mark.parametrize("world", [2, 3, 8])
def test_tile_shards_partition_the_frame(gpu_renderer, world):
    kw = dict(vres=64, width=100, height=70, iters=1, mat="metal")
    vol, opts, mcs = build_scene(**kw)
    full, _, _ = render_gpu(gpu_renderer, vol, opts, mcs, 100, 70, count=False)
    acc = np.zeros_like(full)
    owned = 0
    for rank in range(world):
        gpu_renderer.set_tile_shard(rank, world, 16, 8)
        gpu_renderer.clear_accum(100, 70)
        gpu_renderer.render_frame(opts, mcs)
        part = gpu_renderer.read_accum()
        assert not ((part[..., 3] != 0) & (acc[..., 3] != 0)).any(), "tile rendered by two ranks"
        acc += part
        owned += gpu_renderer.shard_pixels()
    gpu_renderer.set_tile_shard(0, 1, 32, 32)
    assert owned == 100 * 70
    assert np.array_equal(acc.view(np.uint32), full.view(np.uint32))


def test_error_behaviour(gpu_renderer):
    from raymarchcl_b200._lib import RaymarchError
    from raymarchcl_b200.renderer import Renderer
    vol, opts, mcs = build_scene(vres=32, width=32, height=32, iters=1, mat="ao")
    r = Renderer(0)
    with pytest.raises(RaymarchError) as e:
        r.clear_accum(32, 32)
        r.render_pass(opts[0], mcs[0])
    assert e.value.code == -3  # no volume
    r.set_volume(vol)
    with pytest.raises(RaymarchError) as e:
        r.render_pass(opts[0][:-1] , mcs[0])
    assert e.value.code == -1  # blob size
    r.clear_accum(16, 16)
    with pytest.raises(RaymarchError) as e:
        r.render_pass(opts[0], mcs[0])
    assert e.value.code == -2 and "resolution" in e.value.message
    r.close()


def test_full_size_property_idempotent_and_deterministic(gpu_renderer):
    """At BASELINE's full frame size (1920x1080, one pass; the oracle is too slow here) check the
    size-independent properties: two runs are bit-identical, and the tile-sharded union equals
    the unsharded frame on a checksum of checksums."""
    kw = dict(vres=256, width=1920, height=1080, iters=1, mat="metal")
    vol, opts, mcs = build_scene(**kw)
    a, argb_a, ca = render_gpu(gpu_renderer, vol, opts, mcs, 1920, 1080)
    b, argb_b, cb = render_gpu(gpu_renderer, vol, opts, mcs, 1920, 1080)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32)) and np.array_equal(ca, cb)
    assert np.array_equal(argb_a, argb_b)
    per_ps = ca.astype(np.float64) / (1920 * 1080)
    assert 600 < per_ps[0] < 800 and 150 < per_ps[1] < 300, per_ps   # ~712 steps, ~210 taps
    rowsum = a.astype(np.float64).sum(axis=(1, 2))
    acc = np.zeros_like(rowsum)
    for rank in range(2):
        gpu_renderer.set_tile_shard(rank, 2, 32, 32)
        gpu_renderer.clear_accum(1920, 1080)
        gpu_renderer.render_frame(opts, mcs)
        acc += gpu_renderer.read_accum().astype(np.float64).sum(axis=(1, 2))
    gpu_renderer.set_tile_shard(0, 1, 32, 32)
    assert np.array_equal(acc, rowsum)


@pytest.mark.parametrize("knobs", [
    {3: 2}, {3: 3}, {3: 4}, {3: 5}, {6: 1}, {6: 2},
], ids=lambda k: "-".join(f"{a}={b}" for a, b in k.items()))
def test_fast_kernel_tuning_knobs_do_not_change_results(gpu_renderer, knobs):
    """Macro-cell size of the distance map and the pass-fusion limit are performance choices
    only: accumulator bits and work counters must not move."""
    kw = dict(vres=96, width=120, height=72, iters=3, mat="metal")
    vol, opts, mcs = build_scene(**kw)
    gpu_renderer.set_option(2, 1)
    ref, _, cref = render_gpu(gpu_renderer, vol, opts, mcs, 120, 72)
    gpu_renderer.set_option(2, 0)
    try:
        for k, v in knobs.items():
            gpu_renderer.set_option(k, v)
        px, _, cnt = render_gpu(gpu_renderer, vol, opts, mcs, 120, 72)
        px_nc, _, _ = render_gpu(gpu_renderer, vol, opts, mcs, 120, 72, count=False)
    finally:
        for k, v in {3: 0, 6: 32}.items():
            gpu_renderer.set_option(k, v)
    assert np.array_equal(cnt, cref)
    assert np.array_equal(px.view(np.uint32), ref.view(np.uint32))
    assert np.array_equal(px_nc.view(np.uint32), ref.view(np.uint32))


def test_fast_kernel_iso_boundary_and_non_uniform_passes(gpu_renderer, oracle):
    """Voxels exactly at isoVal are NOT solid for the march (v > iso) but DO count as occupied for
    normals (v >= iso); passes whose opts differ in more than `time` cannot share a launch."""
    kw = dict(vres=64, width=96, height=64, iters=3, mat="metal")
    vol, opts, mcs = build_scene(**kw)
    vol = vol.copy()
    vol[vol == 64] = 32  # one band sits exactly on the iso value
    from raymarchcl_b200.options import decode_render_opts, encode_render_opts
    f = decode_render_opts(opts[1])
    f["eyePos"] = [1.2, 0.5, -1.9]
    opts = [opts[0], encode_render_opts(f), opts[2]]
    ref_px, ref_cnt = oracle.render_frame(vol, mcs, opts, 96, 64)
    gpu_renderer.set_option(2, 0)
    px, argb, cnt = render_gpu(gpu_renderer, vol, opts, mcs, 96, 64)
    assert np.array_equal(cnt, ref_cnt)
    check_frame(px, ref_px, argb, oracle.tonemap(ref_px, opts[0]))


@pytest.mark.parametrize("world", [2, 5])
def test_packed_shards_unpack_to_the_full_frame(gpu_renderer, world):
    """rm_tonemap_device / rm_copy_accum_device (packed) per shard + rm_unpack_shards = the frame an
    unsharded context renders: the multi-GPU assembly step, exercised on one GPU."""
    import torch
    w, h = 150, 90
    kw = dict(vres=64, width=w, height=h, iters=2, mat="metal")
    vol, opts, mcs = build_scene(**kw)
    gpu_renderer.set_option(2, 0)
    full, argb_full, _ = render_gpu(gpu_renderer, vol, opts, mcs, w, h, count=False)
    r = gpu_renderer
    r.set_tile_shard(0, world, 32, 32)
    r.clear_accum(w, h)
    stride = r.shard_slots(0, world)
    parts = torch.zeros((world, stride), dtype=torch.int32, device="cuda:0")
    parts_acc = torch.zeros((world, stride, 4), dtype=torch.float32, device="cuda:0")
    frame = torch.zeros(w * h, dtype=torch.int32, device="cuda:0")
    frame_acc = torch.zeros((w * h, 4), dtype=torch.float32, device="cuda:0")
    torch.cuda.synchronize()
    for rank in range(world):
        r.set_tile_shard(rank, world, 32, 32)
        assert r.shard_slots(rank, world) <= stride
        r.clear_accum(w, h)
        r.render_frame(opts, mcs)
        r.tonemap_device(opts[0], parts[rank].data_ptr(), packed=True)
        r.copy_accum_device(parts_acc[rank].data_ptr(), packed=True)
    r.unpack_shards(parts.data_ptr(), world, stride, 4, frame.data_ptr())
    r.unpack_shards(parts_acc.data_ptr(), world, stride, 16, frame_acc.data_ptr())
    r.sync()
    r.set_tile_shard(0, 1, 32, 32)
    assert np.array_equal(frame.cpu().numpy().view(np.uint32).reshape(h, w), argb_full)
    assert np.array_equal(frame_acc.cpu().numpy().reshape(h, w, 4).view(np.uint32), full.view(np.uint32))


def test_warp_scheduled_kernel_matches_oracle(gpu_renderer, oracle):
    """RM_OPT_KERNEL = 2 (persistent warp-scheduled state machine): same results, exact counters."""
    kw = dict(vres=128, width=200, height=120, iters=2, mat="metal2", dof=0.025)
    vol, opts, mcs = build_scene(**kw)
    ref_px, ref_cnt = oracle.render_frame(vol, mcs, opts, 200, 120)
    gpu_renderer.set_option(2, 2)
    try:
        px, argb, cnt = render_gpu(gpu_renderer, vol, opts, mcs, 200, 120)
    finally:
        gpu_renderer.set_option(2, 0)
    assert np.array_equal(cnt, ref_cnt)
    check_frame(px, ref_px, argb, oracle.tonemap(ref_px, opts[0]))


@pytest.mark.parametrize("kw", SCENES + [
    dict(vres=192, width=96, height=64, iters=2, mat="metal", theta=20.0, dist=1.4),    # camera close to the box
    dict(vres=256, width=64, height=64, iters=1, mat="metal2", theta=200.0, dist=0.6),  # camera INSIDE the box
    dict(vres=100, width=80, height=50, iters=1, mat="ao", volume="terrain", theta=90.0),
], ids=lambda k: f"{k.get('volume', 'gyroid')}{k['vres']}_{k['mat']}_{k['width']}x{k['height']}")
def test_production_mode_equals_counting_mode_and_oracle(gpu_renderer, oracle, kw):
    """The non-counting kernels skip empty samples without looking at them (march_fast); the counting
    kernels visit every sample. Both must give the oracle's accumulator: bit-identical to each other."""
    vol, opts, mcs = build_scene(**kw)
    w, h = kw["width"], kw["height"]
    ref_px, ref_cnt = oracle.render_frame(vol, mcs, opts, w, h)
    gpu_renderer.set_option(2, 0)
    a, argb_a, cnt = render_gpu(gpu_renderer, vol, opts, mcs, w, h, count=True)
    b, argb_b, _ = render_gpu(gpu_renderer, vol, opts, mcs, w, h, count=False)
    assert np.array_equal(cnt, ref_cnt)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert np.array_equal(argb_a, argb_b)
    check_frame(b, ref_px, argb_b, oracle.tonemap(ref_px, opts[0]))


def test_two_contexts_on_one_device_do_not_disturb_each_other(gpu_renderer):
    """The kernels read per-launch constants from __constant__ memory; launches of different
    contexts (different streams) on one device are ordered by the library."""
    from raymarchcl_b200.renderer import Renderer
    kw_a = dict(vres=64, width=128, height=72, iters=3, mat="metal")
    kw_b = dict(vres=96, width=100, height=60, iters=2, mat="orange-stripes", theta=-45.0)
    va, oa, ma = build_scene(**kw_a)
    vb, ob, mb = build_scene(**kw_b)
    gpu_renderer.set_option(2, 0)
    ref_a, _, _ = render_gpu(gpu_renderer, va, oa, ma, 128, 72, count=False)
    ref_b, _, _ = render_gpu(gpu_renderer, vb, ob, mb, 100, 60, count=False)
    ra, rb = Renderer(0), Renderer(0)
    try:
        ra.set_volume(va); ra.clear_accum(128, 72); ra.upload_passes(oa, ma)
        rb.set_volume(vb); rb.clear_accum(100, 60); rb.upload_passes(ob, mb)
        for i in range(3):           # asynchronous launches, interleaved between the two contexts
            ra.render_resident(i, 1)
            if i < 2:
                rb.render_resident(i, 1)
        a, b = ra.read_accum(), rb.read_accum()
    finally:
        ra.close(); rb.close()
    assert np.array_equal(a.view(np.uint32), ref_a.view(np.uint32))
    assert np.array_equal(b.view(np.uint32), ref_b.view(np.uint32))


@pytest.mark.parametrize("kw", [
    dict(vres=512, width=160, height=90, iters=1, mat="metal", volume="blob"),      # BASELINE config 3 stand-in
    dict(vres=1024, width=96, height=54, iters=1, mat="metal2", volume="dragon"),   # BASELINE config 5 stand-in
], ids=["c3_blob512", "c5_thin1024"])
def test_large_volumes_match_oracle(gpu_renderer, oracle, kw):
    """512^3 (128 MiB) and 1024^3 (1 GiB, macro-cell 16 voxels, march step 5.3 voxels) volumes."""
    vol, opts, mcs = build_scene(**kw)
    w, h = kw["width"], kw["height"]
    ref_px, ref_cnt = oracle.render_frame(vol, mcs, opts, w, h)
    gpu_renderer.set_option(2, 0)
    px, argb, cnt = render_gpu(gpu_renderer, vol, opts, mcs, w, h)
    assert np.array_equal(cnt, ref_cnt)
    check_frame(px, ref_px, argb, oracle.tonemap(ref_px, opts[0]))
    b, _, _ = render_gpu(gpu_renderer, vol, opts, mcs, w, h, count=False)
    assert np.array_equal(px.view(np.uint32), b.view(np.uint32))


def test_vox_file_upload_and_error_codes(gpu_renderer, tmp_path):
    """rm_load_volume_file reads the reference's .vox format (io.clj:9-33) natively."""
    from raymarchcl_b200 import save_volume
    from raymarchcl_b200._lib import RaymarchError
    kw = dict(vres=48, width=64, height=40, iters=1, mat="metal")
    vol, opts, mcs = build_scene(**kw)
    gpu_renderer.set_option(2, 0)
    ref, _, _ = render_gpu(gpu_renderer, vol, opts, mcs, 64, 40, count=False)
    path = str(tmp_path / "v.vox")
    save_volume(path, vol)
    assert gpu_renderer.load_volume_file(path) == (48, 48, 48)
    gpu_renderer.clear_accum(64, 40)
    gpu_renderer.render_frame(opts, mcs)
    assert np.array_equal(gpu_renderer.read_accum().view(np.uint32), ref.view(np.uint32))
    bad = tmp_path / "bad.vox"
    bad.write_bytes(b"VOXEL" + b"\x00\x00\x00\x08" * 3 + b"\x01" + b"\x00" * 100)  # 8^3 declared, 100 bytes present
    for p in (str(bad), str(tmp_path / "missing.vox")):
        with pytest.raises(RaymarchError) as e:
            gpu_renderer.load_volume_file(p)
        assert e.value.code == -8
    gpu_renderer.set_volume(vol)


def test_anim_frames_equal_independent_renders(gpu_renderer):
    """test-anim (core.clj:181-213): the volume stays resident across frames, only the opts change."""
    import raymarchcl_b200.renderer as R
    from raymarchcl_b200 import compute_eyepos, generate_scatter_offsets, make_gyroid_volume, make_render_option_buffers
    vol = make_gyroid_volume(64)
    frames = R.test_anim(96, 54, 2, 64, "metal", frames=3, volume=vol)
    assert len(frames) == 3 and not np.array_equal(frames[0], frames[1])
    for f, got in enumerate(frames):
        t = f / 3.0
        args = dict(width=96, height=54, vres=[64, 64, 64], iter=2, mat="metal", fov=115.0, targetpos=[0, -0.15, 0],
                    eyepos=compute_eyepos(350.0 * t, 2.25, 0.44 + 0.01 * t))
        opts = make_render_option_buffers(2, args, t_step=0.3333)
        mcs = [generate_scatter_offsets(0x4000, 1000 + i) for i in range(2)]
        _, argb, _ = render_gpu(gpu_renderer, vol, opts, mcs, 96, 54, count=False)
        assert np.array_equal(argb, got)


@pytest.mark.parametrize("vres", [64, 256, (96, 40, 130), 512, 1024], ids=str)
def test_device_gyroid_generator_is_byte_identical_to_host_generator(gpu_renderer, vres):
    # (512^3 and 1024^3: the device's fp64 cos / sin are not correctly rounded, and the shell thresholds
    #  |0.2 - g| < 0.05, g > 0.35 could flip on a last-bit difference -- they do not, at any BASELINE size)
    from raymarchcl_b200 import make_gyroid_volume
    gpu_renderer.generate_gyroid_volume(vres)
    got = gpu_renderer.read_volume()
    want = make_gyroid_volume(vres)
    assert got.shape == want.shape
    assert np.array_equal(got, want), f"{(got != want).sum()} voxels differ"


@pytest.mark.parametrize("vres", [64, 256, (96, 40, 130), 512], ids=str)
def test_device_terrain_generator_is_byte_identical_to_host_generator(gpu_renderer, vres):
    from raymarchcl_b200 import make_terrain
    gpu_renderer.generate_terrain_volume(vres)
    got = gpu_renderer.read_volume()
    want = make_terrain(vres)
    assert got.shape == want.shape
    assert np.array_equal(got, want), f"{(got != want).sum()} voxels differ"
    with pytest.raises(Exception):
        gpu_renderer.generate_terrain_volume((64, 64, 32))  # the reference's second wall needs rz >= rx


def test_device_scatter_tables_reproduce_java_random(gpu_renderer):
    """Tables generated on the device (java.util.Random LCG, seeds 1000+i) drive a render that is
    bit-identical to the render from the host-generated tables."""
    kw = dict(vres=64, width=96, height=64, iters=3, mat="metal")
    vol, opts, mcs = build_scene(**kw)
    gpu_renderer.set_option(2, 0)
    ref, _, _ = render_gpu(gpu_renderer, vol, opts, mcs, 96, 64, count=False)
    r = gpu_renderer
    r.generate_gyroid_volume(64)
    r.clear_accum(96, 64)
    r.generate_scatter_tables(1000, 3)
    r.upload_passes(opts, None)
    r.render_resident(0, 3)
    assert np.array_equal(r.read_accum().view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize("kw", [
    dict(vres=64, width=96, height=64, iters=2, mat="metal", groundY=0.2),                  # ground plane cuts the volume
    dict(vres=128, width=80, height=60, iters=1, mat="metal2", groundY=-0.3, theta=60.0),   # plane at y=+0.3, camera below it
    dict(vres=64, width=64, height=48, iters=1, mat="ao", groundY=0.0, volume="full"),
], ids=["groundY0.2", "groundY-0.3", "full_groundY0"])
def test_ground_plane_inside_the_volume(gpu_renderer, oracle, kw):
    """The production kernels march only the samples that can change distanceToScene's result and
    repeat the last call in full when the reference's 'last voxel normal wins even if the ground is
    closer' quirk (renderer.cl:224-228) could apply. A ground plane that cuts through the voxels
    makes that quirk the common case."""
    vol, opts, mcs = build_scene(**kw)
    w, h = kw["width"], kw["height"]
    ref_px, ref_cnt = oracle.render_frame(vol, mcs, opts, w, h)
    gpu_renderer.set_option(2, 0)
    a, argb_a, cnt = render_gpu(gpu_renderer, vol, opts, mcs, w, h, count=True)
    b, argb_b, _ = render_gpu(gpu_renderer, vol, opts, mcs, w, h, count=False)
    assert np.array_equal(cnt, ref_cnt)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    check_frame(b, ref_px, argb_b, oracle.tonemap(ref_px, opts[0]))
