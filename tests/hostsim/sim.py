"""ctypes access to tests/hostsim/libhostsim.so -- TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import build_hostsim

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")

STAT_NAMES = ["lookups", "skips", "skipped_samples", "jumps", "jump_samples", "seq_adds", "marches", "traces"]
MODES = {"production": 0, "counting": 1, "bytes": 2, "wave": 0, "wave_counting": 1,
         "fused": 3, "fused_counting": 4, "fused_bytemap": 5, "fused_bytemap_counting": 6}


class HostSim:
    def __init__(self):
        self.lib = lib = C.CDLL(build_hostsim.build())
        lib.sim_render_pixels.argtypes = [_u8p, _f32p, C.c_char_p, _f32p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                          C.c_int, C.c_int]
        lib.sim_render_pixels.restype = None
        lib.sim_render_pixels_wave.argtypes = lib.sim_render_pixels.argtypes
        lib.sim_render_pixels_wave.restype = None
        lib.sim_stats_words.restype = C.c_int
        lib.sim_get_stats.argtypes = [_u64p, C.c_int]

    def render_frame(self, voxels, mcs: Sequence[np.ndarray], opts: Sequence[bytes], width: int, height: int,
                     ids: Optional[np.ndarray] = None, mode: str = "production", cell_shift: int = 2):
        """All passes from a zero accumulator. Returns (pixels[H,W,4] float32, counters[3] uint64)."""
        vox = np.ascontiguousarray(voxels, dtype=np.uint8).reshape(-1)
        self._keep = vox  # the accel tables are keyed on this buffer's address
        pixels = np.zeros((height, width, 4), dtype=np.float32)
        counters = np.zeros(3, dtype=np.uint64)
        idp, nid = None, 0
        if ids is not None:
            ids = np.ascontiguousarray(ids, dtype=np.int32)
            idp, nid = ids.ctypes.data_as(C.c_void_p), int(ids.size)
        fn = self.lib.sim_render_pixels_wave if mode.startswith("wave") else self.lib.sim_render_pixels
        for o, mc in zip(opts, mcs):
            fn(vox, np.ascontiguousarray(mc, dtype=np.float32).reshape(-1), o,
                                       pixels.reshape(-1), width * height, idp, nid,
                                       counters.ctypes.data_as(C.c_void_p), MODES[mode], cell_shift)
        return pixels, counters

    def stats(self, reset: bool = True) -> dict:
        n = self.lib.sim_stats_words()
        out = np.zeros(n, dtype=np.uint64)
        self.lib.sim_get_stats(out, int(reset))
        d = {k: int(out[i]) for i, k in enumerate(STAT_NAMES)}
        k = len(STAT_NAMES)
        d["skip_hist_log2"] = [int(x) for x in out[k:k + 16]]
        d["events"] = [int(x) for x in out[k + 16:k + 48]]
        return d
