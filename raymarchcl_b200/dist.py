"""Multi-GPU plumbing of the render op: interleaved tile ownership + ONE framebuffer gather.

The reference is single-device (core.clj:121-123). A work-item reads only read-only buffers and
its own ``pixels[id]`` (renderer.cl:483-492), so the frame shards by pixels with no exchange until
the end: every rank holds the whole volume and renders the tiles ``(i + skew*j) % world == rank``
(diagonal stripes, because cost varies ~10x across the image and is far from uniform across columns), then the packed ARGB tiles are gathered
to rank 0 in one collective (NCCL over NVLink under torchrun; gloo in the CPU tests) and
de-interleaved into the frame. One process per GPU; ``torch.distributed`` is plumbing only.

``slot_pixel_index`` is the host mirror of ``rm_slot_to_pixel`` (csrc/rm_kernels.h).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np


def _skew(world: int) -> int:
    """Smallest of (3, 5, 7, 2, 1) coprime to ``world`` (csrc/rm_types.h:rm_shard_layout)."""
    from math import gcd
    for s in (3, 5, 7, 2, 1):
        if gcd(s, world) == 1:
            return s
    return 1


@dataclass(frozen=True)
class ShardLayout:
    """Host mirror of ``RmShard`` / ``rm_shard_layout`` / ``rm_slot_to_pixel``: tile (i, j) belongs to rank
    ``(i + skew*j) % world`` (diagonal stripes); every rank owns ``tiles_per_rank_row`` tile columns per
    tile row, columns beyond the frame being padding."""
    width: int
    height: int
    world: int
    tile_w: int = 32
    tile_h: int = 32

    @property
    def tiles_x(self) -> int:
        return (self.width + self.tile_w - 1) // self.tile_w

    @property
    def tiles_y(self) -> int:
        return (self.height + self.tile_h - 1) // self.tile_h

    @property
    def tiles(self) -> int:
        return self.tiles_x * self.tiles_y

    @property
    def tiles_per_rank_row(self) -> int:
        return (self.tiles_x + self.world - 1) // self.world

    @property
    def skew(self) -> int:
        return _skew(self.world)

    def owned_tiles(self, rank: int) -> int:
        return self.tiles_per_rank_row * self.tiles_y

    def slots(self, rank: int) -> int:
        return self.owned_tiles(rank) * self.tile_w * self.tile_h

    @property
    def max_slots(self) -> int:
        return self.slots(0)

    def slot_pixel_index(self, rank: int) -> np.ndarray:
        """int64[slots(rank)]: pixel id (y*W+x) of every work slot of ``rank``; -1 = padding."""
        n = self.slots(rank)
        slot = np.arange(n, dtype=np.int64)
        tile_px = self.tile_w * self.tile_h
        lt, r = slot // tile_px, slot % tile_px
        ty, k = lt // self.tiles_per_rank_row, lt % self.tiles_per_rank_row
        first = (rank - (self.skew * ty) % self.world) % self.world
        tx = first + k * self.world
        sb, l = r >> 5, r & 31
        sbw = self.tile_w >> 3
        sby, sbx = sb // sbw, sb % sbw
        x = tx * self.tile_w + sbx * 8 + (l & 7)
        y = ty * self.tile_h + sby * 4 + (l >> 3)
        pid = y * self.width + x
        pid[(tx >= self.tiles_x) | (x >= self.width) | (y >= self.height)] = -1
        return pid


def assemble_frame(parts, layout: ShardLayout, out=None):
    """De-interleave per-rank packed buffers (``parts[r][:slots(r)]``) into one flat frame of
    ``width*height`` elements. Works on numpy arrays and torch tensors (any device)."""
    import torch
    first = parts[0]
    is_torch = isinstance(first, torch.Tensor)
    n = layout.width * layout.height
    if out is None:
        out = (torch.zeros((n,) + tuple(first.shape[1:]), dtype=first.dtype, device=first.device)
               if is_torch else np.zeros((n,) + first.shape[1:], dtype=first.dtype))
    for r in range(layout.world):
        idx = layout.slot_pixel_index(r)
        valid = idx >= 0
        if is_torch:
            ti = torch.from_numpy(idx[valid]).to(first.device)
            tv = torch.from_numpy(np.nonzero(valid)[0]).to(first.device)
            out[ti] = parts[r][tv]
        else:
            out[idx[valid]] = parts[r][: idx.size][valid]
    return out


class FrameGatherer:
    """Buffers + index maps for the ONE collective of a frame: every rank contributes its packed
    shard (``local``, ``max_slots`` elements, tile-major), ``dst`` receives them side by side and
    de-interleaves them into the frame -- with the library's own unpack kernel
    (``rm_unpack_shards``) when a :class:`Renderer` is given, with a torch index copy otherwise
    (CPU / gloo tests)."""

    def __init__(self, layout: ShardLayout, rank: int, device, dtype, dst: int = 0, elem_shape=(), renderer=None):
        import torch
        self.layout, self.rank, self.dst, self.renderer = layout, rank, dst, renderer
        self.max_slots = layout.max_slots
        self.elem_shape = tuple(elem_shape)
        self.local = torch.zeros((self.max_slots,) + self.elem_shape, dtype=dtype, device=device)
        self.parts = None
        self.frame = None
        if rank == dst:
            self.parts = torch.zeros((layout.world, self.max_slots) + self.elem_shape, dtype=dtype, device=device)
            self.frame = torch.zeros((layout.width * layout.height,) + self.elem_shape, dtype=dtype, device=device)
            if renderer is None:
                src, dstpix = [], []
                for r in range(layout.world):
                    idx = layout.slot_pixel_index(r)
                    v = np.nonzero(idx >= 0)[0]
                    src.append(v + r * self.max_slots)
                    dstpix.append(idx[v])
                self._src = torch.from_numpy(np.concatenate(src)).to(device)
                self._dst = torch.from_numpy(np.concatenate(dstpix)).to(device)

    def gather(self):
        """Collective: returns the flat frame on ``dst`` (None elsewhere)."""
        import torch.distributed as dist
        if self.layout.world == 1:
            self.parts[0].copy_(self.local)
        else:
            dist.gather(self.local, list(self.parts.unbind(0)) if self.rank == self.dst else None, dst=self.dst)
            if self.rank != self.dst:
                return None
        if self.renderer is not None:
            elem_bytes = self.local.element_size() * int(np.prod(self.elem_shape, dtype=np.int64))
            self.renderer.unpack_shards(self.parts.data_ptr(), self.layout.world, self.max_slots, elem_bytes,
                                        self.frame.data_ptr())
        else:
            flat = self.parts.reshape((-1,) + self.elem_shape)
            self.frame[self._dst] = flat[self._src]
        return self.frame
