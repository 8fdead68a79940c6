"""ctypes binding of libraymarch_b200.so (include/raymarch_b200.h).

This is the same C ABI a JVM host binds through JNA. There is NO CPU fallback:
if the shared library is missing or no sm_100 device is usable, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RAYMARCH_B200_LIB") or os.path.join(PKG_DIR, "libraymarch_b200.so")

OPTS_BYTES = 544
TABLE_FLOATS = 65536

RM_OK = 0
STATUS_NAMES = {0: "RM_OK", -1: "RM_ERR_INVALID_ARG", -2: "RM_ERR_BAD_OPTS", -3: "RM_ERR_NO_VOLUME",
                -4: "RM_ERR_NO_FRAMEBUFFER", -5: "RM_ERR_CUDA", -6: "RM_ERR_NO_DEVICE", -7: "RM_ERR_UNSUPPORTED", -8: "RM_ERR_IO"}
RM_OPT_COUNT_WORK = 1
RM_OPT_KERNEL = 2
RM_OPT_CELL_SHIFT = 3
RM_OPT_FUSE_LIMIT = 6
RM_OPT_TRIP_LIMIT = 7
RM_OPT_WAVE_CHUNK = 8
RM_OPT_WAVE_REFILL = 9
RM_OPT_PERSIST_BLOCK = 10
RM_OPT_PERSIST_GROUP = 11
RM_OPT_PERSIST_SMEM = 12
RM_OPT_PERSIST_ORDER = 13

# every symbol include/raymarch_b200.h declares (tests check the .so exports all of them)
EXPORTS = [
    "rm_abi_version", "rm_device_count", "rm_create", "rm_destroy", "rm_last_error", "rm_set_volume", "rm_load_volume_file", "rm_generate_gyroid_volume", "rm_generate_terrain_volume", "rm_voxelize_points", "rm_generate_scatter_tables", "rm_read_volume",
    "rm_clear_accum", "rm_render_pass", "rm_render_frame", "rm_tonemap", "rm_read_accum",
    "rm_upload_passes", "rm_render_resident", "rm_tonemap_device", "rm_copy_accum_device", "rm_sync",
    "rm_set_stream", "rm_set_tile_shard", "rm_shard_pixels", "rm_shard_slots", "rm_unpack_shards", "rm_set_option", "rm_get_stats",
    "rm_reset_stats", "rm_set_volume_device", "rm_tonemap_async", "rm_wait", "rm_update_opts", "rm_set_argb_target", "rm_host_alloc", "rm_host_free", "rm_create_multi", "rm_member_count", "rm_get_member_stats",
]


class RmStats(C.Structure):
    _fields_ = [("steps", C.c_uint64), ("taps", C.c_uint64), ("outer_iters", C.c_uint64),
                ("pixel_samples", C.c_uint64), ("kernel_launches", C.c_uint64), ("render_launches", C.c_uint64),
                ("render_ms", C.c_double), ("tonemap_ms", C.c_double),
                ("h2d_ms", C.c_double), ("d2h_ms", C.c_double),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64)]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


class RaymarchError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"{STATUS_NAMES.get(code, code)}: {message}")
        self.code = code
        self.message = message


_lib = None


def load() -> C.CDLL:
    """Load libraymarch_b200.so (built in-tree by raymarchcl_b200/csrc/build.py). Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python raymarchcl_b200/csrc/build.py` "
            "(or __graft_entry__.build()). There is no CPU fallback for the render op.")
    lib = C.CDLL(LIB_PATH)
    vp, ip, sz = C.c_void_p, C.c_int, C.c_size_t
    lib.rm_abi_version.restype = ip
    lib.rm_device_count.restype = ip
    lib.rm_create.argtypes = [ip, C.POINTER(vp)]
    lib.rm_destroy.argtypes = [vp]
    lib.rm_destroy.restype = None
    lib.rm_last_error.argtypes = [vp]
    lib.rm_last_error.restype = C.c_char_p
    lib.rm_set_volume.argtypes = [vp, vp, ip, ip, ip]
    lib.rm_load_volume_file.argtypes = [vp, C.c_char_p, C.POINTER(ip), C.POINTER(ip), C.POINTER(ip)]
    lib.rm_generate_gyroid_volume.argtypes = [vp, ip, ip, ip]
    lib.rm_generate_terrain_volume.argtypes = [vp, ip, ip, ip]
    lib.rm_voxelize_points.argtypes = [vp, vp, C.c_int64, ip, ip]
    lib.rm_generate_scatter_tables.argtypes = [vp, C.c_int64, ip]
    lib.rm_read_volume.argtypes = [vp, vp]
    lib.rm_clear_accum.argtypes = [vp, ip, ip]
    lib.rm_render_pass.argtypes = [vp, vp, sz, vp, sz]
    lib.rm_render_frame.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), ip]
    lib.rm_tonemap.argtypes = [vp, vp, sz, vp]
    lib.rm_read_accum.argtypes = [vp, vp]
    lib.rm_upload_passes.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), ip]
    lib.rm_render_resident.argtypes = [vp, ip, ip]
    lib.rm_tonemap_device.argtypes = [vp, vp, sz, vp, ip]
    lib.rm_copy_accum_device.argtypes = [vp, vp, ip]
    lib.rm_sync.argtypes = [vp]
    lib.rm_set_stream.argtypes = [vp, vp]
    lib.rm_set_tile_shard.argtypes = [vp, ip, ip, ip, ip]
    lib.rm_shard_pixels.argtypes = [vp]
    lib.rm_shard_pixels.restype = C.c_int64
    lib.rm_shard_slots.argtypes = [vp, ip, ip]
    lib.rm_shard_slots.restype = C.c_int64
    lib.rm_unpack_shards.argtypes = [vp, vp, ip, C.c_int64, ip, vp]
    lib.rm_set_option.argtypes = [vp, ip, C.c_int64]
    lib.rm_get_stats.argtypes = [vp, C.POINTER(RmStats)]
    lib.rm_reset_stats.argtypes = [vp]
    lib.rm_set_volume_device.argtypes = [vp, vp, ip, ip, ip]
    lib.rm_tonemap_async.argtypes = [vp, vp, sz, vp, ip]
    lib.rm_wait.argtypes = [vp, ip]
    lib.rm_update_opts.argtypes = [vp, C.POINTER(vp), ip]
    lib.rm_set_argb_target.argtypes = [vp, vp, ip]
    lib.rm_host_alloc.argtypes = [vp, sz, C.POINTER(vp)]
    lib.rm_host_free.argtypes = [vp, vp]
    lib.rm_create_multi.argtypes = [C.POINTER(ip), ip, C.POINTER(vp)]
    lib.rm_member_count.argtypes = [vp]
    lib.rm_get_member_stats.argtypes = [vp, ip, C.POINTER(RmStats)]
    _lib = lib
    return lib
