"""The N > 1 path on CPU: interleaved tile ownership and the one framebuffer gather
(raymarchcl_b200/dist.py), world_size 2 over gloo. The per-rank "render" is the oracle restricted
to the pixels that rank owns -- test infrastructure standing in for the GPU, exactly the pixels
rm_set_tile_shard would give the rank."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from raymarchcl_b200.dist import FrameGatherer, ShardLayout, assemble_frame
from tests.scenes import build_scene


@pytest.mark.parametrize("w,h,world,tw,th", [(100, 70, 2, 16, 8), (64, 64, 3, 32, 32), (1920, 1080, 8, 32, 32), (33, 9, 4, 8, 4)])
def test_shards_partition_every_pixel_exactly_once(w, h, world, tw, th):
    lay = ShardLayout(w, h, world, tw, th)
    seen = np.zeros(w * h, dtype=np.int32)
    for r in range(world):
        idx = lay.slot_pixel_index(r)
        assert idx.size == lay.slots(r) and lay.slots(r) <= lay.max_slots
        np.add.at(seen, idx[idx >= 0], 1)
    assert (seen == 1).all()
    assert lay.slots(0) == lay.max_slots  # rank 0 owns the most tiles (rm_shard_slots relies on it)


@pytest.mark.parametrize("w,h,world,tw,th", [(1920, 1080, 8, 16, 8), (1920, 1080, 6, 16, 8), (3840, 2160, 8, 16, 8), (100, 70, 5, 16, 8),
                                             (131, 77, 7, 8, 4), (64, 64, 9, 32, 32), (256, 256, 1, 32, 32)])
def test_python_shard_layout_is_the_kernels_layout(w, h, world, tw, th):
    """raymarchcl_b200/dist.py:ShardLayout restates rm_shard_layout / rm_slot_to_pixel (csrc/rm_types.h, rm_kernels.h);
    tests/hostsim compiles those very headers for the host: slot for slot the same pixel ids, for every rank."""
    import ctypes as C
    from tests.hostsim import build_hostsim
    lib = C.CDLL(build_hostsim.build())
    lib.sim_shard_slots.restype = C.c_longlong
    lib.sim_shard_slots.argtypes = [C.c_int] * 6 + [C.c_void_p, C.c_longlong]
    lay = ShardLayout(w, h, world, tw, th)
    loads = []
    for r in range(world):
        want = lay.slot_pixel_index(r)
        got = np.full(want.size + 8, -7, dtype=np.int32)
        n = lib.sim_shard_slots(w, h, r, world, tw, th, got.ctypes.data, got.size)
        assert n == want.size == lay.slots(r)
        assert np.array_equal(got[:n].astype(np.int64), want)
        loads.append(int((want >= 0).sum()))
    assert sum(loads) == w * h
    if world > 1 and w * h > 100000:
        assert max(loads) / (sum(loads) / world) < 1.02  # diagonal stripes deal the pixels evenly


def test_warp_bundles_are_8x4_pixel_blocks():
    lay = ShardLayout(64, 64, 2, 32, 32)
    idx = lay.slot_pixel_index(1)[:32]
    ys, xs = idx // 64, idx % 64
    assert xs.max() - xs.min() == 7 and ys.max() - ys.min() == 3


def test_assemble_numpy_roundtrip():
    lay = ShardLayout(50, 30, 3, 16, 8)
    full = np.arange(50 * 30, dtype=np.int64)
    parts = []
    for r in range(3):
        idx = lay.slot_pixel_index(r)
        p = np.zeros(lay.max_slots, dtype=np.int64)
        p[: idx.size][idx >= 0] = full[idx[idx >= 0]]
        parts.append(p)
    assert np.array_equal(assemble_frame(parts, lay), full)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, w, h, kw, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import refso
        orc = refso.load("oracle")
        orc.set_num_threads(2)
        vol, opts, mcs = build_scene(**kw)
        lay = ShardLayout(w, h, world, 16, 8)
        idx = lay.slot_pixel_index(rank)
        own = idx[idx >= 0].astype(np.int32)
        px, cnt = orc.render_frame(vol, mcs, opts, w, h, ids=own)       # this rank's pixels only
        argb = orc.tonemap(px, opts[0]).reshape(-1).view(np.int32)
        g = FrameGatherer(lay, rank, "cpu", torch.int32)
        packed = np.zeros(lay.max_slots, dtype=np.int32)
        packed[: idx.size][idx >= 0] = argb[own]
        g.local.copy_(torch.from_numpy(packed))
        ga = FrameGatherer(lay, rank, "cpu", torch.float32, elem_shape=(4,))
        pa = np.zeros((lay.max_slots, 4), dtype=np.float32)
        pa[: idx.size][idx >= 0] = px.reshape(-1, 4)[own]
        ga.local.copy_(torch.from_numpy(pa))
        frame = g.gather()
        accum = ga.gather()
        work = torch.from_numpy(cnt.astype(np.int64))
        dist.all_reduce(work)
        if rank == 0:
            np.savez(out_path, argb=frame.numpy().view(np.uint32), accum=accum.numpy(), work=work.numpy())
        else:
            assert frame is None and accum is None
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_frame_equals_single_rank(tmp_path, oracle):
    w, h = 72, 40
    kw = dict(vres=64, width=w, height=h, iters=2, mat="metal")
    out = str(tmp_path / "frame.npz")
    mp.spawn(_worker, args=(2, _free_port(), w, h, kw, out), nprocs=2, join=True)
    got = np.load(out)
    vol, opts, mcs = build_scene(**kw)
    px, cnt = oracle.render_frame(vol, mcs, opts, w, h)
    assert np.array_equal(got["accum"].reshape(h, w, 4).view(np.uint32), px.view(np.uint32))
    assert np.array_equal(got["argb"].reshape(h, w), oracle.tonemap(px, opts[0]))
    assert np.array_equal(got["work"], cnt.astype(np.int64))  # the ranks' work counters add up
