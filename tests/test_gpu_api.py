"""C-ABI behaviour added in round 2: the animation loop with resident state and double-buffered
asynchronous read-back (test-anim, core.clj:181-213), the ARGB frame written by the render launch
itself, device-resident volume input, and the state-machine corner cases the advisor found."""
import numpy as np
import pytest

from tests.scenes import build_scene
from tests.test_gpu_parity import render_gpu

pytestmark = pytest.mark.gpu


def anim_args(f, frames, w, h, vres, iters):
    from raymarchcl_b200 import compute_eyepos
    t = f / float(frames)
    return dict(width=w, height=h, vres=[vres] * 3, iter=iters, mat="metal", fov=115.0, targetpos=[0, -0.15, 0],
                eyepos=compute_eyepos(350.0 * t, 2.25, 0.44 + 0.01 * t))


def test_resident_animation_with_async_readback_equals_independent_renders(gpu_renderer):
    """35 frames of the reference's orbit (core.clj:194-212): volume + tables uploaded ONCE, per frame only
    the 544-byte opts blobs (rm_update_opts), rm_tonemap_async into alternating pinned buffers while the
    next frame renders. Every frame must equal an independent render of the same opts."""
    import torch
    from raymarchcl_b200 import generate_scatter_offsets, make_gyroid_volume, make_render_option_buffers
    w, h, vres, iters, frames = 96, 54, 64, 2, 35
    vol = make_gyroid_volume(vres)
    mcs = [generate_scatter_offsets(0x4000, 1000 + i) for i in range(iters)]
    all_opts = [make_render_option_buffers(iters, anim_args(f, frames, w, h, vres, iters), t_step=0.3333) for f in range(frames)]
    r = gpu_renderer
    r.set_option(2, 0)
    r.set_tile_shard(0, 1, 32, 32)
    r.set_volume(vol)
    r.clear_accum(w, h)
    r.upload_passes(all_opts[0], mcs)
    r.reset_stats()
    host = [torch.empty(w * h, dtype=torch.int32).pin_memory() for _ in range(2)]
    got = []
    for f in range(frames):
        slot = f & 1
        if f >= 2:
            r.wait(slot)
            got.append(host[slot].numpy().view(np.uint32).reshape(h, w).copy())
        if f > 0:
            r.update_opts(all_opts[f])
        r.clear_accum(w, h)
        r.render_resident(0, iters)
        r.tonemap_async(all_opts[f][0], host[slot].numpy().view(np.uint32), slot)
    for f in (frames - 2, frames - 1):
        r.wait(f & 1)
        got.append(host[f & 1].numpy().view(np.uint32).reshape(h, w).copy())
    st = r.stats()
    # per frame: iters opts blobs, nothing else crosses the bus on the way in
    assert st["h2d_bytes"] == 544 * iters * (frames - 1)
    assert len(got) == frames and not np.array_equal(got[0], got[1])
    for f in range(frames):
        _, argb, _ = render_gpu(r, vol, all_opts[f], mcs, w, h, count=False)
        assert np.array_equal(argb, got[f]), f


def test_test_anim_mirror_uses_the_resident_path(gpu_renderer):
    import raymarchcl_b200.renderer as R
    from raymarchcl_b200 import generate_scatter_offsets, make_gyroid_volume, make_render_option_buffers
    vol = make_gyroid_volume(64)
    frames = R.test_anim(96, 54, 2, 64, "metal", frames=4, volume=vol)
    assert len(frames) == 4
    mcs = [generate_scatter_offsets(0x4000, 1000 + i) for i in range(2)]
    for f, got in enumerate(frames):
        opts = make_render_option_buffers(2, anim_args(f, 4, 96, 54, 64, 2), t_step=0.3333)
        _, argb, _ = render_gpu(gpu_renderer, vol, opts, mcs, 96, 54, count=False)
        assert np.array_equal(argb, got)


def test_folded_tonemap_equals_the_tonemap_kernel(gpu_renderer):
    """rm_tonemap after a default-kernel launch returns the ARGB words that launch wrote; a different gamma,
    a second rm_tonemap, or a pass rendered in between must all still give TonemapImage of the accumulator."""
    from raymarchcl_b200.options import decode_render_opts, encode_render_opts
    kw = dict(vres=64, width=100, height=70, iters=4, mat="metal")
    vol, opts, mcs = build_scene(**kw)
    r = gpu_renderer
    r.set_option(2, 4)
    _, ref, _ = render_gpu(r, vol, opts, mcs, 100, 70, count=False)
    f = decode_render_opts(opts[0])
    f["gamma"] = 0.7
    other = encode_render_opts(f)
    ref_other = r.tonemap(other)
    r.set_option(2, 0)
    r.reset_stats()
    _, got, _ = render_gpu(r, vol, opts, mcs, 100, 70, count=False)
    launches = r.stats()["kernel_launches"]
    assert np.array_equal(got, ref)
    assert np.array_equal(r.tonemap(opts[0]), ref)
    assert r.stats()["kernel_launches"] == launches      # both served from the frame the render launch wrote
    assert np.array_equal(r.tonemap(other), ref_other)   # other gamma: the tonemap kernel runs
    assert r.stats()["kernel_launches"] == launches + 1
    assert np.array_equal(r.tonemap(opts[0]), ref)
    # a further pass changes the accumulator: the frame must follow
    r.render_pass(opts[1], mcs[1])
    a = r.tonemap(opts[0])
    r.set_option(2, 4)
    b = r.tonemap(opts[0])
    r.set_option(2, 0)
    assert np.array_equal(a, b) and not np.array_equal(a, ref)


@pytest.mark.parametrize("world", [2, 5])
def test_argb_target_packed_shards(gpu_renderer, world):
    """rm_set_argb_target(packed): the render launch leaves each shard's ARGB words in the caller's gather
    buffer; rm_tonemap_device on that buffer is then free; unpacking the shards gives the full frame."""
    import torch
    w, h = 150, 90
    kw = dict(vres=64, width=w, height=h, iters=2, mat="metal")
    vol, opts, mcs = build_scene(**kw)
    r = gpu_renderer
    r.set_option(2, 0)
    _, argb_full, _ = render_gpu(r, vol, opts, mcs, w, h, count=False)
    r.set_tile_shard(0, world, 32, 32)
    r.clear_accum(w, h)
    stride = r.shard_slots(0, world)
    parts = torch.full((world, stride), 0x55, dtype=torch.int32, device="cuda:0")
    frame = torch.zeros(w * h, dtype=torch.int32, device="cuda:0")
    torch.cuda.synchronize()
    try:
        for rank in range(world):
            r.set_tile_shard(rank, world, 32, 32)
            r.set_argb_target(parts[rank].data_ptr(), packed=True)
            r.clear_accum(w, h)
            r.reset_stats()
            r.render_frame(opts, mcs)
            n = r.stats()["kernel_launches"]
            r.tonemap_device(opts[0], parts[rank].data_ptr(), packed=True)
            assert r.stats()["kernel_launches"] == n  # nothing left to do
    finally:
        r.set_argb_target(None)
    r.unpack_shards(parts.data_ptr(), world, stride, 4, frame.data_ptr())
    r.sync()
    r.set_tile_shard(0, 1, 32, 32)
    assert np.array_equal(frame.cpu().numpy().view(np.uint32).reshape(h, w), argb_full)


def test_volume_from_device_memory(gpu_renderer):
    import torch
    kw = dict(vres=64, width=96, height=64, iters=2, mat="metal")
    vol, opts, mcs = build_scene(**kw)
    r = gpu_renderer
    r.set_option(2, 0)
    ref, _, _ = render_gpu(r, vol, opts, mcs, 96, 64, count=False)
    dvol = torch.from_numpy(np.ascontiguousarray(vol)).to("cuda:0")
    torch.cuda.synchronize()
    r.set_volume(np.zeros((8, 8, 8), np.uint8))
    r.set_volume_device(dvol.data_ptr(), 64, 64, 64)
    r.clear_accum(96, 64)
    r.render_frame(opts, mcs)
    assert np.array_equal(r.read_accum().view(np.uint32), ref.view(np.uint32))


def test_generated_tables_do_not_survive_a_host_frame(gpu_renderer):
    """generate -> rm_render_frame (overwrites the table slots with host tables) -> rm_upload_passes(NULL)
    must fail with RM_ERR_INVALID_ARG instead of rendering with the wrong tables."""
    from raymarchcl_b200._lib import RaymarchError
    kw = dict(vres=32, width=32, height=32, iters=2, mat="ao")
    vol, opts, mcs = build_scene(**kw)
    r = gpu_renderer
    r.set_volume(vol)
    r.clear_accum(32, 32)
    r.generate_scatter_tables(1000, 2)
    r.upload_passes(opts, None)          # fine: generated tables are in place
    r.render_frame(opts, mcs)            # host tables overwrite slots 0..1
    with pytest.raises(RaymarchError) as e:
        r.upload_passes(opts, None)
    assert e.value.code == -1
    r.generate_scatter_tables(1000, 2)
    r.upload_passes(opts, None)
    r.generate_scatter_tables(1000, 40)  # reallocation (more slots than before) keeps the new tables valid
    r.upload_passes(opts, None)


def test_failed_volume_call_leaves_no_stale_volume(gpu_renderer):
    from raymarchcl_b200._lib import RaymarchError
    from raymarchcl_b200.renderer import Renderer
    kw = dict(vres=32, width=32, height=32, iters=1, mat="ao")
    vol, opts, mcs = build_scene(**kw)
    with Renderer(0) as r:
        r.set_volume(vol)
        r.clear_accum(32, 32)
        r.render_frame(opts, mcs)
        with pytest.raises(RaymarchError):
            r.voxelize_points(np.array([[0, np.inf, 0], [1, 1, 1]], dtype=np.float32), 16, 0)
        with pytest.raises(RaymarchError) as e:  # the old volume is gone, the new one never arrived
            r.render_frame(opts, mcs)
        assert e.value.code == -3
        r.set_volume(vol)
        r.render_frame(opts, mcs)
