#!/usr/bin/env python3
"""Where a launch of the default kernel spends its fixed cost: per-warp start / first-bundle / end timestamps (%globaltimer)
of a full C2 frame and of one rank's 1/8 shard. Needs a library built with the instrumentation,
    python raymarchcl_b200/csrc/build.py -o build_ab/tl.so -DRM_PERSIST_TIMELINE
    RAYMARCH_B200_LIB=$PWD/build_ab/tl.so python tools/timeline_probe.py
Result (profiles/r02_timeline_probe.jsonl): the ramp is 2 us; the warps run out of bundles between T - 260 us and T, and the
mean idle time at the end is ~0.2 ms whatever the launch size -- one bundle's duration. That IS the per-launch fixed cost."""
import os, sys, ctypes as C, json
sys.path.insert(0, os.getcwd())
import numpy as np
from raymarchcl_b200.renderer import Renderer
from raymarchcl_b200 import _lib
from tests.scenes import build_scene
import bench
sc = bench.WORKLOADS["c2"]["scene"]
w, h, iters = sc["width"], sc["height"], sc["iters"]
vol, opts, mcs = build_scene(**sc)
lib = _lib.load()
lib.rm_debug_timeline.argtypes = [C.c_void_p]
N = 148 * 40
with Renderer(0) as r:
    r.set_volume(vol); r.clear_accum(w, h); r.upload_passes(opts, mcs)
    for world in (1, 8):
        r.set_tile_shard(0, world, 16, 8)
        for _ in range(3):
            r.clear_accum(w, h); r.render_resident(0, iters)
        r.sync(); r.reset_stats()
        r.clear_accum(w, h); r.render_resident(0, iters); r.sync()
        ms = r.stats()["render_ms"]
        buf = np.zeros(3 * N, np.uint64)
        lib.rm_debug_timeline(buf.ctypes.data)
        st, fi, en = buf[:N], buf[N:2*N], buf[2*N:]
        used = en > 0
        t0 = st[used].min()
        st_, fi_, en_ = (st[used] - t0) / 1e3, (fi[used] - t0) / 1e3, (en[used] - t0) / 1e3
        T = en_.max()
        # per-SM-block finish: idle warp-time at the end
        idle = (T - en_).sum() / used.sum()
        print(json.dumps({"world": world, "render_ms": round(ms, 3), "warps": int(used.sum()), "kernel_span_us": round(float(T), 1),
                          "block_start_us_max": round(float(st_.max()), 1), "first_bundle_us_mean": round(float(fi_.mean()), 1), "first_bundle_us_max": round(float(fi_.max()), 1),
                          "warp_end_us_min": round(float(en_.min()), 1), "warp_end_us_p10": round(float(np.percentile(en_, 10)), 1),
                          "warp_end_us_median": round(float(np.median(en_)), 1), "mean_idle_tail_us": round(float(idle), 1)}))
