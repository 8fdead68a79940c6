#!/usr/bin/env python3
"""Render BASELINE config 1 (64^3 gyroid, 256x256, :ao) -- plus a 3-pass :metal frame that exercises the
fused blend, the bounce loop and the shadow rays -- once with every render kernel, counting on and off.
Run under `compute-sanitizer --tool memcheck|racecheck|initcheck|synccheck` (tools/gpu_sanitize.sh);
exits non-zero when a call fails. No torch: numpy + the C ABI only."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from raymarchcl_b200.renderer import Renderer
from tests.scenes import build_scene

kernels = [int(a) for a in sys.argv[1:]] or [0, 1, 3, 4]
SCENES = [dict(vres=64, width=256, height=256, iters=1, mat="ao"),
          dict(vres=64, width=96, height=64, iters=3, mat="metal")]
with Renderer(0) as r:
    for kw in SCENES:
        vol, opts, mcs = build_scene(**kw)
        ref = None
        for k in kernels:
            r.set_option(2, k)
            # kernel 0: (block layout, scheduler, map location) -- the automatic choice, both layouts with the
            # TMA-staged shared-memory map (mbarrier + cp.async.bulk), block-synchronous rounds
            knobs = [(0, -1, 2), (1024, 0, 1), (256, 0, 1), (256, 1, 0)] if k == 0 else [(0, -1, 2)]
            for block, group, smem in knobs:
                r.set_option(10, block); r.set_option(11, group); r.set_option(12, smem)
                for count in (True, False):
                    r.set_volume(vol)
                    r.clear_accum(kw["width"], kw["height"])
                    r.count_work(count)
                    r.render_frame(opts, mcs)
                    px = r.read_accum()
                    argb = r.tonemap(opts[0])
                    if ref is None:
                        ref = px
                    assert np.array_equal(px.view(np.uint32), ref.view(np.uint32)), (k, count, block, group, smem)
                print(f"kernel {k} (block {block}, group {group}, smem {smem}): {kw['width']}x{kw['height']}x{kw['iters']} ok", flush=True)
            r.set_option(10, 0); r.set_option(11, -1); r.set_option(12, 2)
    # the asynchronous read-back path and the accel rebuild kernels
    out = [r.alloc_pinned_argb() for _ in range(2)]
    r.set_option(2, 0)
    for f in range(4):
        r.clear_accum(kw["width"], kw["height"])
        r.render_frame(opts, mcs)
        r.tonemap_async(opts[0], out[f & 1], f & 1)
    r.wait(0); r.wait(1)
    r.free_pinned(out)
print("sanitize_c1 done")
