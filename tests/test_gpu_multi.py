"""rm_create_multi: ONE context over several GPUs of a box, behind the same C ABI (the reference is
single-device, core.clj:121-123; a host changes that one call site). Inputs are uploaded once and
broadcast device-to-device, tiles are dealt in diagonal stripes, and every GPU's render kernel stores the
ARGB words of its tiles straight into device 0's frame over NVLink -- no gather, no unpack.

The 1-member group runs everywhere (it exercises the whole group layer on a single GPU); the 2- and
N-member cases need that many devices and are skipped otherwise (they ran on a 2- and an 8-GPU box:
profiles/r02_multi_*.log)."""
import numpy as np
import pytest

from tests.scenes import build_scene
from tests.test_gpu_parity import render_gpu

pytestmark = pytest.mark.gpu


def device_count():
    from raymarchcl_b200 import _lib
    return _lib.load().rm_device_count()


def group_sizes():
    return [1, 2, 4, 8]


@pytest.mark.parametrize("n", group_sizes())
def test_group_frame_equals_single_gpu_frame(gpu_renderer, oracle, n):
    from raymarchcl_b200.renderer import Renderer
    if device_count() < n:
        pytest.skip(f"needs {n} GPUs")
    kw = dict(vres=96, width=200, height=120, iters=5, mat="metal2", dof=0.025)
    vol, opts, mcs = build_scene(**kw)
    w, h = kw["width"], kw["height"]
    gpu_renderer.set_option(2, 0)
    ref_px, ref_argb, ref_cnt = render_gpu(gpu_renderer, vol, opts, mcs, w, h, count=True)
    with Renderer(list(range(n))) as g:
        assert g.member_count() == n
        g.set_volume(vol)
        g.clear_accum(w, h)
        assert g.shard_pixels() == w * h
        g.reset_stats()
        g.count_work(True)
        g.render_frame(opts, mcs)
        px = g.read_accum()
        argb = g.tonemap(opts[0])
        st = g.stats()
        # production kernels, resident inputs, folded tonemap
        g.count_work(False)
        g.clear_accum(w, h)
        g.upload_passes(opts, mcs)
        g.reset_stats()
        g.render_resident(0, len(opts))
        argb2 = g.tonemap(opts[0])
        st2 = g.stats()
        per_member = [g.member_stats(i) for i in range(n)]
    assert np.array_equal(px.view(np.uint32), ref_px.view(np.uint32))
    assert np.array_equal(argb, ref_argb) and np.array_equal(argb2, ref_argb)
    assert [st["steps"], st["taps"], st["outer_iters"]] == [int(x) for x in ref_cnt]
    assert st2["kernel_launches"] == st2["render_launches"]  # nothing but the render launches: the tonemap is folded in
    assert sum(m["pixel_samples"] for m in per_member) == w * h * len(opts)
    if n > 1:
        assert all(m["pixel_samples"] > 0 for m in per_member)


@pytest.mark.parametrize("n", [1, 2, 8])
def test_group_animation_with_async_readback(gpu_renderer, n):
    """The resident animation loop (rm_update_opts + rm_tonemap_async / rm_wait, two ARGB frames on device 0)
    on a group: every frame equals the single-GPU render of the same opts."""
    from raymarchcl_b200 import generate_scatter_offsets, make_gyroid_volume, make_render_option_buffers
    from raymarchcl_b200.renderer import Renderer
    from tests.test_gpu_api import anim_args
    if device_count() < n:
        pytest.skip(f"needs {n} GPUs")
    w, h, vres, iters, frames = 96, 54, 64, 2, 7
    vol = make_gyroid_volume(vres)
    mcs = [generate_scatter_offsets(0x4000, 1000 + i) for i in range(iters)]
    all_opts = [make_render_option_buffers(iters, anim_args(f, frames, w, h, vres, iters), t_step=0.3333) for f in range(frames)]
    got = []
    with Renderer(list(range(n))) as g:
        g.set_volume(vol)
        g.clear_accum(w, h)
        g.upload_passes(all_opts[0], mcs)
        host = [g.alloc_pinned_argb() for _ in range(2)]
        for f in range(frames):
            slot = f & 1
            if f >= 2:
                g.wait(slot)
                got.append(host[slot].reshape(h, w).copy())
            if f > 0:
                g.update_opts(all_opts[f])
            g.clear_accum(w, h)
            g.render_resident(0, iters)
            g.tonemap_async(all_opts[f][0], host[slot], slot)
        for f in (frames - 2, frames - 1):
            g.wait(f & 1)
            got.append(host[f & 1].reshape(h, w).copy())
        g.free_pinned(host)
    gpu_renderer.set_option(2, 0)
    for f in range(frames):
        _, argb, _ = render_gpu(gpu_renderer, vol, all_opts[f], mcs, w, h, count=False)
        assert np.array_equal(argb, got[f]), f


def test_group_rejects_single_gpu_only_calls(gpu_renderer):
    from raymarchcl_b200._lib import RaymarchError
    from raymarchcl_b200.renderer import Renderer
    with Renderer([0]) as g:
        with pytest.raises(RaymarchError) as e:
            g.set_stream(None)
        assert e.value.code == -7
        with pytest.raises(RaymarchError) as e:
            g.set_tile_shard(1, 2, 32, 32)
        assert e.value.code == -1
        g.set_tile_shard(0, 1, 32, 16)  # tile size only
    with pytest.raises(RaymarchError):
        Renderer([0, 0])
