"""ctypes access to the two CPU checkers -- TEST INFRASTRUCTURE ONLY (oracle/).

* ``load("oracle")``      -> oracle/librm_oracle.so      the C restatement (rm_oracle.c), prefix ``orc_``
* ``load("ref_strict")``  -> oracle/_ref/libref_strict.so the reference's own kernel text, strict fp32, counters
* ``load("ref_fast")``    -> oracle/_ref/libref_fast.so   same text, -O3 -ffast-math (CPU timing baseline)

All three expose the same entry points, wrapped by :class:`CpuRenderer`. Only tests/,
``__graft_entry__.smoke()`` and bench.py's cpu_baseline / ``--impl reference`` legs may import this.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_PATHS = {
    "oracle": (os.path.join(HERE, "librm_oracle.so"), "orc_"),
    "ref_strict": (os.path.join(HERE, "_ref", "libref_strict.so"), "ref_"),
    "ref_fast": (os.path.join(HERE, "_ref", "libref_fast.so"), "ref_"),
}

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")


def available(kind: str) -> bool:
    return os.path.exists(_PATHS[kind][0])


class CpuRenderer:
    """Uniform wrapper over one CPU checker library."""

    def __init__(self, kind: str):
        path, pre = _PATHS[kind]
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing -- run oracle/build_oracle.py / oracle/build_ref.py")
        self.kind = kind
        self.lib = lib = C.CDLL(path)
        self._pre = pre
        f = lambda name: getattr(lib, pre + name)
        f("sizeof_opts").restype = C.c_int
        f("has_counters").restype = C.c_int
        f("num_threads").restype = C.c_int
        f("set_num_threads").argtypes = [C.c_int]
        f("render_pixels").argtypes = [_u8p, _f32p, C.c_char_p, _f32p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        f("render_pixels").restype = None
        f("tonemap").argtypes = [_f32p, C.c_char_p, _u32p, C.c_int]
        f("tonemap").restype = None
        f("intersects_box").argtypes = [_f32p] * 4
        f("intersects_box").restype = C.c_float
        f("voxel_lookup").argtypes = [_u8p, C.c_char_p, _f32p]
        f("voxel_lookup").restype = C.c_int
        f("voxel_normal").argtypes = [_u8p, C.c_char_p, _i32p, C.c_int, _f32p]
        f("voxel_normal").restype = None
        f("distance_to_scene").argtypes = [_u8p, C.c_char_p, _f32p, _f32p, C.c_int, C.c_int, _f32p]
        f("distance_to_scene").restype = None
        f("raymarch").argtypes = [_u8p, C.c_char_p, _f32p, _f32p, C.c_float, C.c_int, C.c_int, _f32p]
        f("raymarch").restype = None
        f("camera_ray").argtypes = [C.c_char_p, _f32p, C.c_int, _f32p]
        f("camera_ray").restype = None
        self._f = f

    # -- info --
    def sizeof_opts(self) -> int:
        return self._f("sizeof_opts")()

    def has_counters(self) -> bool:
        return bool(self._f("has_counters")())

    def num_threads(self) -> int:
        return self._f("num_threads")()

    def set_num_threads(self, n: int) -> None:
        self._f("set_num_threads")(int(n))

    def opts_offsets(self) -> Optional[np.ndarray]:
        if self._pre != "ref_":
            return None
        out = np.zeros(64, dtype=np.int32)
        fn = self.lib.ref_opts_offsets
        fn.argtypes = [_i32p, C.c_int]
        fn.restype = C.c_int
        k = fn(out, 64)
        return out[:k]

    # -- kernels --
    def render_pass(self, voxels: np.ndarray, mc: np.ndarray, opts: bytes, pixels: np.ndarray,
                    ids: Optional[np.ndarray] = None, counters: Optional[np.ndarray] = None) -> None:
        """One RenderImage pass over all pixels (or the listed ``ids``), in place on ``pixels``."""
        vox = np.ascontiguousarray(voxels, dtype=np.uint8).reshape(-1)
        n = pixels.size // 4
        assert pixels.dtype == np.float32 and pixels.flags.c_contiguous
        assert mc.dtype == np.float32 and mc.size == 4 * 16384
        idp, nid = None, 0
        if ids is not None:
            ids = np.ascontiguousarray(ids, dtype=np.int32)
            idp, nid = ids.ctypes.data_as(C.c_void_p), int(ids.size)
        cp = None
        if counters is not None:
            assert counters.dtype == np.uint64 and counters.size >= 3
            cp = counters.ctypes.data_as(C.c_void_p)
        self._f("render_pixels")(vox, mc.reshape(-1), opts, pixels.reshape(-1), n, idp, nid, cp)

    def render_frame(self, voxels, mcs: Sequence[np.ndarray], opts: Sequence[bytes], width: int, height: int,
                     ids: Optional[np.ndarray] = None):
        """All passes from a zero accumulator (core.clj:81-90). Returns (pixels[H,W,4], counters[3])."""
        pixels = np.zeros((height, width, 4), dtype=np.float32)
        counters = np.zeros(3, dtype=np.uint64)
        for o, mc in zip(opts, mcs):
            self.render_pass(voxels, mc, o, pixels, ids=ids, counters=counters)
        return pixels, counters

    def tonemap(self, pixels: np.ndarray, opts: bytes) -> np.ndarray:
        n = pixels.size // 4
        out = np.zeros(n, dtype=np.uint32)
        self._f("tonemap")(np.ascontiguousarray(pixels, dtype=np.float32).reshape(-1), opts, out, n)
        return out.reshape(pixels.shape[:-1])

    # -- per-function hooks --
    def intersects_box(self, bmin, bmax, p, d) -> float:
        a = [np.ascontiguousarray(v, dtype=np.float32) for v in (bmin, bmax, p, d)]
        return float(self._f("intersects_box")(*a))

    def voxel_lookup(self, voxels, opts: bytes, p) -> int:
        return int(self._f("voxel_lookup")(np.ascontiguousarray(voxels, dtype=np.uint8).reshape(-1), opts,
                                           np.ascontiguousarray(p, dtype=np.float32)))

    def voxel_normal(self, voxels, opts: bytes, q, smooth: bool) -> np.ndarray:
        out = np.zeros(3, dtype=np.float32)
        self._f("voxel_normal")(np.ascontiguousarray(voxels, dtype=np.uint8).reshape(-1), opts,
                                np.ascontiguousarray(q, dtype=np.int32), int(smooth), out)
        return out

    def distance_to_scene(self, voxels, opts: bytes, rpos, d, steps: int, smooth: bool) -> np.ndarray:
        out = np.zeros(5, dtype=np.float32)
        self._f("distance_to_scene")(np.ascontiguousarray(voxels, dtype=np.uint8).reshape(-1), opts,
                                     np.ascontiguousarray(rpos, dtype=np.float32),
                                     np.ascontiguousarray(d, dtype=np.float32), int(steps), int(smooth), out)
        return out

    def raymarch(self, voxels, opts: bytes, pos, d, max_dist: float, max_steps: int, smooth: bool) -> np.ndarray:
        out = np.zeros(8, dtype=np.float32)
        self._f("raymarch")(np.ascontiguousarray(voxels, dtype=np.uint8).reshape(-1), opts,
                            np.ascontiguousarray(pos, dtype=np.float32), np.ascontiguousarray(d, dtype=np.float32),
                            float(max_dist), int(max_steps), int(smooth), out)
        return out

    def voxelize_points(self, vertices, res: int, ks: int = -1) -> np.ndarray:
        """meshvoxel.clj voxelize (ks < 0) / voxelize-ks restated in oracle/meshvoxel_oracle.c (C restatement only)."""
        fn = self.lib.orc_voxelize_points
        fn.argtypes = [_f32p, C.c_longlong, C.c_int, C.c_int, _u8p]
        fn.restype = C.c_int
        v = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 3)
        out = np.zeros(res * res * res, dtype=np.uint8)
        if fn(v.reshape(-1), v.shape[0], int(res), int(ks), out) != 0:
            raise ValueError("NaN or infinite coordinate")
        return out.reshape(res, res, res)

    def camera_ray(self, opts: bytes, mc: np.ndarray, pid: int) -> np.ndarray:
        out = np.zeros(8, dtype=np.float32)
        self._f("camera_ray")(opts, mc.reshape(-1), int(pid), out)
        return out


_cache = {}


def load(kind: str = "oracle") -> CpuRenderer:
    if kind not in _cache:
        _cache[kind] = CpuRenderer(kind)
    return _cache[kind]
