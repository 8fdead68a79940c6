"""Host-side mirror of the reference's render pipeline over the C ABI.

Reference interface (Clojure, /root/reference/src/thi/ng/raymarchcl/core.clj) -> here:

* ``init-renderer`` (:119-148)   -> :func:`init_renderer`  (state map with the same keys)
* ``make-pipeline`` (:76-97)     -> :func:`make_pipeline`  (the same step list, as Python dicts)
* ``ops/execute-pipeline`` (:171)-> :func:`execute_pipeline` (runs the steps through libraymarch_b200.so)
* ``update-render-option-buffer`` (:108-117) -> :func:`update_render_option_buffer`
* ``test-render`` (:154-179)     -> :func:`test_render`

:class:`Renderer` is the thin object wrapper of one ``rm_ctx``. No CPU fallback exists: without the
CUDA library and a B200 every call raises :class:`RaymarchError` / ``ImportError``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Any, Dict, List, Mapping, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import RaymarchError, RmStats
from .generators import generate_scatter_offsets, make_gyroid_volume
from .options import (OPTS_BYTES, compute_eyepos, encode_render_opts, make_render_option_buffers,
                      render_options)
from .volio import load_volume


def _as_table(mc) -> np.ndarray:
    t = np.ascontiguousarray(mc, dtype=np.float32).reshape(-1)
    if t.size != _lib.TABLE_FLOATS:
        raise ValueError(f"scatter table must hold {_lib.TABLE_FLOATS} floats, got {t.size}")
    return t


class Renderer:
    """One render context on one GPU (``rm_create`` .. ``rm_destroy``)."""

    def __init__(self, device=0):
        """``device``: one CUDA device id (``rm_create``) or a sequence of ids (``rm_create_multi``: one
        context over several GPUs of a box, the frame assembled on the first one over NVLink)."""
        self._lib = _lib.load()
        h = C.c_void_p()
        if isinstance(device, (list, tuple)):
            ids = (C.c_int * len(device))(*[int(d) for d in device])
            rc = self._lib.rm_create_multi(ids, len(device), C.byref(h))
            self.device = int(device[0])
        else:
            rc = self._lib.rm_create(int(device), C.byref(h))
            self.device = int(device)
        if rc != _lib.RM_OK:
            raise RaymarchError(rc, self._lib.rm_last_error(None).decode())
        self._h = h
        self.width = self.height = 0
        self.vres = None
        self._keep: List[Any] = []

    # -- plumbing --
    def _check(self, rc: int) -> None:
        if rc != _lib.RM_OK:
            raise RaymarchError(rc, self._lib.rm_last_error(self._h).decode())

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.rm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- inputs --
    def set_volume(self, voxels: np.ndarray) -> None:
        v = np.ascontiguousarray(voxels, dtype=np.uint8)
        if v.ndim != 3:
            raise ValueError("volume must be uint8[rz, ry, rx]")
        rz, ry, rx = v.shape
        self._check(self._lib.rm_set_volume(self._h, v.ctypes.data, rx, ry, rz))
        self.vres = (rx, ry, rz)

    def load_volume_file(self, path: str):
        """Upload a .vox file (io.clj:19-33) straight from disk; returns (rx, ry, rz)."""
        rx, ry, rz = C.c_int(), C.c_int(), C.c_int()
        self._check(self._lib.rm_load_volume_file(self._h, os.fsencode(path), C.byref(rx), C.byref(ry), C.byref(rz)))
        self.vres = (rx.value, ry.value, rz.value)
        return self.vres

    def generate_gyroid_volume(self, vres) -> None:
        """``make-gyroid-volume`` (generators.clj:27-42) on the device instead of host + upload."""
        rx, ry, rz = [int(vres)] * 3 if isinstance(vres, (int, np.integer)) else [int(v) for v in vres]
        self._check(self._lib.rm_generate_gyroid_volume(self._h, rx, ry, rz))
        self.vres = (rx, ry, rz)

    def generate_terrain_volume(self, vres) -> None:
        """``make-terrain`` (generators.clj:44-60) on the device instead of host + upload."""
        rx, ry, rz = [int(vres)] * 3 if isinstance(vres, (int, np.integer)) else [int(v) for v in vres]
        self._check(self._lib.rm_generate_terrain_volume(self._h, rx, ry, rz))
        self.vres = (rx, ry, rz)

    def voxelize_points(self, vertices, res: int, ks: int = -1) -> None:
        """``meshvoxel/voxelize`` (``ks < 0``, meshvoxel.clj:60-69) or ``meshvoxel/voxelize-ks``
        (meshvoxel.clj:45-58) of mesh vertices on the device; the result becomes the volume."""
        v = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 3)
        self._check(self._lib.rm_voxelize_points(self._h, v.ctypes.data, v.shape[0], int(res), int(ks)))
        self.vres = (int(res),) * 3

    def generate_scatter_tables(self, seed0: int, count: int) -> None:
        """``generate-scatter-offsets`` for seeds seed0..seed0+count-1 into the resident table slots."""
        self._check(self._lib.rm_generate_scatter_tables(self._h, int(seed0), int(count)))

    def read_volume(self) -> np.ndarray:
        rx, ry, rz = self.vres
        out = np.empty((rz, ry, rx), dtype=np.uint8)
        self._check(self._lib.rm_read_volume(self._h, out.ctypes.data))
        return out

    def clear_accum(self, width: int, height: int) -> None:
        self._check(self._lib.rm_clear_accum(self._h, int(width), int(height)))
        self.width, self.height = int(width), int(height)

    # -- hot path --
    def render_pass(self, opts: bytes, mc) -> None:
        t = _as_table(mc)
        self._check(self._lib.rm_render_pass(self._h, C.c_char_p(opts), len(opts), t.ctypes.data, t.size))

    def _ptr_arrays(self, opts: Sequence[bytes], mcs: Sequence[np.ndarray]):
        n = len(opts)
        if n != len(mcs):
            raise ValueError("opts and mc lists differ in length")
        for o in opts:
            if len(o) != OPTS_BYTES:
                raise ValueError(f"TRenderOpts blob must be {OPTS_BYTES} bytes")
        tabs = [_as_table(m) for m in mcs]
        obufs = [C.create_string_buffer(o, OPTS_BYTES) for o in opts]
        oarr = (C.c_void_p * n)(*[C.cast(b, C.c_void_p) for b in obufs])
        marr = (C.c_void_p * n)(*[C.c_void_p(t.ctypes.data) for t in tabs])
        return n, oarr, marr, (tabs, obufs)

    def render_frame(self, opts: Sequence[bytes], mcs: Sequence[np.ndarray]) -> None:
        n, oarr, marr, keep = self._ptr_arrays(opts, mcs)
        self._check(self._lib.rm_render_frame(self._h, oarr, marr, n))

    def upload_passes(self, opts: Sequence[bytes], mcs: Optional[Sequence[np.ndarray]]) -> None:
        """``mcs=None``: use the tables made in place by :meth:`generate_scatter_tables`."""
        if mcs is None:
            n = len(opts)
            obufs = [C.create_string_buffer(o, OPTS_BYTES) for o in opts]
            oarr = (C.c_void_p * n)(*[C.cast(b, C.c_void_p) for b in obufs])
            self._check(self._lib.rm_upload_passes(self._h, oarr, None, n))
            return
        n, oarr, marr, keep = self._ptr_arrays(opts, mcs)
        self._check(self._lib.rm_upload_passes(self._h, oarr, marr, n))

    def update_opts(self, opts: Sequence[bytes]) -> None:
        """``update-render-option-buffer`` (core.clj:108-117) for resident passes: new opts, same tables."""
        n = len(opts)
        obufs = [C.create_string_buffer(o, OPTS_BYTES) for o in opts]
        oarr = (C.c_void_p * n)(*[C.cast(b, C.c_void_p) for b in obufs])
        self._check(self._lib.rm_update_opts(self._h, oarr, n))

    def set_volume_device(self, dptr: int, rx: int, ry: int, rz: int) -> None:
        """Volume already in device memory (``uint8[rz][ry][rx]``): one device-to-device copy."""
        self._check(self._lib.rm_set_volume_device(self._h, C.c_void_p(dptr), int(rx), int(ry), int(rz)))
        self.vres = (int(rx), int(ry), int(rz))

    def tonemap_async(self, opts: bytes, out: np.ndarray, slot: int) -> None:
        """Queue TonemapImage + the read-back into ``out`` (pinned uint32[H*W]); :meth:`wait` completes it."""
        if out.dtype != np.uint32 or out.size != self.width * self.height or not out.flags.c_contiguous:
            raise ValueError("out must be a contiguous uint32 array of width*height words")
        self._check(self._lib.rm_tonemap_async(self._h, C.c_char_p(opts), len(opts), out.ctypes.data, int(slot)))

    def wait(self, slot: int) -> None:
        self._check(self._lib.rm_wait(self._h, int(slot)))

    def alloc_pinned_argb(self) -> np.ndarray:
        """A page-locked uint32[width*height] host buffer (``rm_host_alloc``) for :meth:`tonemap_async`."""
        n = self.width * self.height
        p = C.c_void_p()
        self._check(self._lib.rm_host_alloc(self._h, n * 4, C.byref(p)))
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint32)), shape=(n,))

    def free_pinned(self, buffers) -> None:
        for b in buffers:
            if b is not None and self._h:
                self._lib.rm_host_free(self._h, C.c_void_p(b.ctypes.data))

    def set_argb_target(self, dptr: Optional[int], packed: bool = False) -> None:
        """Device buffer the default kernel fills with ARGB words while it renders (None: own frame)."""
        self._check(self._lib.rm_set_argb_target(self._h, C.c_void_p(dptr or 0), int(packed)))

    def render_resident(self, first: int, count: int) -> None:
        self._check(self._lib.rm_render_resident(self._h, int(first), int(count)))

    def tonemap(self, opts: bytes, out: Optional[np.ndarray] = None) -> np.ndarray:
        if out is None:
            out = np.empty((self.height, self.width), dtype=np.uint32)
        elif out.dtype != np.uint32 or out.size != self.width * self.height or not out.flags.c_contiguous:
            raise ValueError("out must be a contiguous uint32 array of width*height words")
        self._check(self._lib.rm_tonemap(self._h, C.c_char_p(opts), len(opts), out.ctypes.data))
        return out

    def tonemap_device(self, opts: bytes, dptr: int, packed: bool) -> None:
        self._check(self._lib.rm_tonemap_device(self._h, C.c_char_p(opts), len(opts), C.c_void_p(dptr), int(packed)))

    def copy_accum_device(self, dptr: int, packed: bool) -> None:
        self._check(self._lib.rm_copy_accum_device(self._h, C.c_void_p(dptr), int(packed)))

    def read_accum(self) -> np.ndarray:
        out = np.empty((self.height, self.width, 4), dtype=np.float32)
        self._check(self._lib.rm_read_accum(self._h, out.ctypes.data))
        return out

    def sync(self) -> None:
        self._check(self._lib.rm_sync(self._h))

    def set_stream(self, cuda_stream: Optional[int]) -> None:
        self._check(self._lib.rm_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    # -- sharding, options, stats --
    def set_tile_shard(self, rank: int, world: int, tile_w: int = 32, tile_h: int = 32) -> None:
        self._check(self._lib.rm_set_tile_shard(self._h, rank, world, tile_w, tile_h))

    def shard_pixels(self) -> int:
        return int(self._lib.rm_shard_pixels(self._h))

    def shard_slots(self, rank: int, world: int) -> int:
        return int(self._lib.rm_shard_slots(self._h, int(rank), int(world)))

    def unpack_shards(self, parts_dptr: int, world: int, stride_slots: int, elem_bytes: int, frame_dptr: int) -> None:
        """De-interleave gathered per-rank packed buffers into a frame (all device memory)."""
        self._check(self._lib.rm_unpack_shards(self._h, C.c_void_p(parts_dptr), int(world), int(stride_slots),
                                               int(elem_bytes), C.c_void_p(frame_dptr)))

    def set_option(self, option: int, value: int) -> None:
        self._check(self._lib.rm_set_option(self._h, int(option), int(value)))

    def count_work(self, on: bool = True) -> None:
        self.set_option(_lib.RM_OPT_COUNT_WORK, int(on))

    def stats(self) -> Dict[str, Any]:
        st = RmStats()
        self._check(self._lib.rm_get_stats(self._h, C.byref(st)))
        return st.as_dict()

    def member_count(self) -> int:
        return int(self._lib.rm_member_count(self._h))

    def member_stats(self, member: int) -> Dict[str, Any]:
        st = RmStats()
        self._check(self._lib.rm_get_member_stats(self._h, int(member), C.byref(st)))
        return st.as_dict()

    def reset_stats(self) -> None:
        self._check(self._lib.rm_reset_stats(self._h))


# ----------------------------------------------------------------------------------------------
# The reference's pipeline vocabulary
# ----------------------------------------------------------------------------------------------

def init_renderer(args: Mapping[str, Any], device: int = 0, seed0: int = 1000) -> Dict[str, Any]:
    """State map of ``init-renderer`` (core.clj:119-148): context, per-pass opts blobs, per-pass
    scatter tables (seed ``seed0 + i`` instead of nanoTime), the volume, pixel count, pipeline.

    ``args`` keys as in the reference: width height vres iter vname (+ the render-options keys).
    ``vname`` is a .vox path; when absent a gyroid volume of ``vres`` is generated
    (``make-gyroid-volume`` is the commented-out alternative at core.clj:142).
    """
    width, height, it = int(args["width"]), int(args["height"]), int(args.get("iter", 1))
    vol = args.get("volume")
    if vol is None:
        vol = load_volume(args["vname"]) if args.get("vname") else make_gyroid_volume(args["vres"])
    r = Renderer(device)
    state: Dict[str, Any] = {
        "cl-state": r,
        "opts-buffers": make_render_option_buffers(it, args),
        "mc-buffers": [generate_scatter_offsets(0x4000, seed0 + i) for i in range(it)],
        "num": width * height,
        "p-buf": (width, height),
        "q-buf": (width, height),
        "v-buf": vol,
        "args": dict(args),
    }
    state["pipeline"] = make_pipeline(state)
    return state


def make_pipeline(state: Mapping[str, Any]) -> List[Dict[str, Any]]:
    """The step list of ``make-pipeline`` (core.clj:76-97)."""
    steps: List[Dict[str, Any]] = [{"write": ["p-buf", "v-buf"]}]
    for i in range(len(state["opts-buffers"])):
        steps.append({"write": [("o-buf", i), ("mc-buf", i)]})
        steps.append({"name": "RenderImage", "in": ["v-buf", ("mc-buf", i), ("o-buf", i)], "out": "p-buf",
                      "n": state["num"], "args": [[state["num"], "int"]]})
    steps.append({"write": "q-buf"})
    steps.append({"name": "TonemapImage", "in": ["p-buf", ("o-buf", 0)], "out": "q-buf", "n": state["num"],
                  "read": ["out"], "args": [[state["num"], "int"]]})
    return steps


def execute_pipeline(state: Mapping[str, Any], pipeline: Optional[List[Dict[str, Any]]] = None,
                     fused: bool = True) -> np.ndarray:
    """``ops/execute-pipeline`` (core.clj:171): run the step list, return the ARGB words [H, W].

    ``fused=True`` submits all RenderImage steps with one ``rm_render_frame`` call; ``False`` issues
    one ``rm_render_pass`` per step, exactly like the reference's queue.
    """
    r: Renderer = state["cl-state"]
    steps = pipeline if pipeline is not None else state["pipeline"]
    width, height = state["p-buf"]
    passes: List[int] = []
    argb = None
    for st in steps:
        if "write" in st and st.get("name") is None:
            w = st["write"]
            if isinstance(w, list) and "p-buf" in w:
                if r.vres is None or state.get("_volume_dirty", True):
                    r.set_volume(state["v-buf"])
                r.clear_accum(width, height)
            continue
        if st["name"] == "RenderImage":
            i = st["in"][2][1]
            if fused:
                passes.append(i)
            else:
                r.render_pass(state["opts-buffers"][i], state["mc-buffers"][i])
        elif st["name"] == "TonemapImage":
            if passes:
                r.render_frame([state["opts-buffers"][i] for i in passes],
                               [state["mc-buffers"][i] for i in passes])
                passes = []
            argb = r.tonemap(state["opts-buffers"][st["in"][1][1]])
    return argb


def update_render_option_buffer(state: Dict[str, Any], args: Mapping[str, Any]) -> None:
    """``update-render-option-buffer`` (core.clj:108-117): re-encode every pass (t = i*0.3333)."""
    n = len(state["opts-buffers"])
    state["opts-buffers"] = make_render_option_buffers(n, {**state["args"], **args}, t_step=0.3333)
    state["_volume_dirty"] = False


def test_render(width: int = 640, height: int = 360, iter: int = 1, vres: int = 256, mat: str = "metal",
                vname: Optional[str] = None, out_path: Optional[str] = None, theta: float = 135,
                dist: float = 2.25, device: int = 0, **opts) -> np.ndarray:
    """``test-render`` (core.clj:154-179). Returns the ARGB words; writes a PNG when ``out_path``."""
    args = {"width": width, "height": height, "vres": vres, "iter": iter,
            "eyepos": compute_eyepos(theta, dist, 0.35), "targetpos": [0, -0.4, 0], "mat": mat,
            "vname": vname, **opts}
    state = init_renderer(args, device=device)
    try:
        argb = execute_pipeline(state)
    finally:
        state["cl-state"].close()
    if out_path:
        from PIL import Image
        rgb = np.stack([(argb >> 16) & 255, (argb >> 8) & 255, argb & 255], axis=-1).astype(np.uint8)
        Image.fromarray(rgb, "RGB").save(out_path)
    return argb


def test_anim(width: int, height: int, iter: int, res: int, mat: str, vname: Optional[str] = None,
              frames: int = 35, out_dir: Optional[str] = None, device: int = 0, volume=None) -> List[np.ndarray]:
    """``test-anim`` (core.clj:181-213): one ``init-renderer``, then per frame an orbiting camera
    (theta 0..350 deg over ``frames`` frames, r 2.25, eye y 0.44..0.45, target y -0.15, fov 115) and
    ``update-render-option-buffer``. The reference's step list re-uploads the volume and every table
    on every frame (core.clj:81,84); here they are uploaded ONCE (``rm_set_volume``,
    ``rm_upload_passes``), a frame costs ``iter`` x 544 bytes of opts (``rm_update_opts``), and the
    ARGB read-back of frame k (``rm_tonemap_async`` into one of two pinned host buffers) overlaps the
    render of frame k+1. Returns the ARGB frames; writes ``frame-%04d.png`` into ``out_dir`` when given."""
    args = {"width": width, "height": height, "vres": [res, res, res], "iter": iter, "mat": mat,
            "vname": vname, "volume": volume}
    state = init_renderer(args, device=device)
    r: Renderer = state["cl-state"]
    out: List[np.ndarray] = []

    def collect(frame: int, slot: int) -> None:
        r.wait(slot)
        argb = host[slot].reshape(height, width).copy()
        out.append(argb)
        if out_dir:
            from PIL import Image
            rgb = np.stack([(argb >> 16) & 255, (argb >> 8) & 255, argb & 255], axis=-1).astype(np.uint8)
            Image.fromarray(rgb, "RGB").save(os.path.join(out_dir, "frame-%04d.png" % frame))

    host = []
    try:
        r.set_volume(state["v-buf"])
        r.clear_accum(width, height)
        host = [r.alloc_pinned_argb() for _ in range(2)]
        r.upload_passes(state["opts-buffers"], state["mc-buffers"])
        for frame in range(frames):
            slot = frame & 1
            if frame >= 2:
                collect(frame - 2, slot)
            t = frame / float(frames)                       # m/map-interval frame 0 35 0.0 1.0
            frame_args = {"fov": 115.0, "targetpos": [0, -0.15, 0],
                          "eyepos": compute_eyepos(350.0 * t, 2.25, 0.44 + 0.01 * t)}
            update_render_option_buffer(state, frame_args)
            r.update_opts(state["opts-buffers"])
            r.clear_accum(width, height)
            r.render_resident(0, len(state["opts-buffers"]))
            r.tonemap_async(state["opts-buffers"][0], host[slot], slot)
        for frame in range(max(0, frames - 2), frames):
            collect(frame, frame & 1)
    finally:
        r.free_pinned(host)
        r.close()
    return out
