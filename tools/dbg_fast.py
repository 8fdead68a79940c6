"""Debug helper: tiny renders through the fast kernel with a low watchdog limit."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from raymarchcl_b200.renderer import Renderer
from tests.scenes import build_scene
from oracle import refso, build_oracle
build_oracle.build(verbose=False)
orc = refso.load("oracle")
r = Renderer(0)
r.set_option(7, 200000)
for kw in [dict(vres=32, width=32, height=32, iters=1, mat="ao", volume="empty"),
           dict(vres=64, width=64, height=48, iters=1, mat="ao"),
           dict(vres=64, width=64, height=48, iters=2, mat="metal"),
           dict(vres=256, width=320, height=180, iters=2, mat="metal")]:
    for count in (True, False):
        vol, opts, mcs = build_scene(**kw)
        w, h = kw["width"], kw["height"]
        try:
            t = time.time()
            r.set_volume(vol); r.clear_accum(w, h); r.reset_stats(); r.count_work(count)
            r.render_frame(opts, mcs)
            st = r.stats(); px = r.read_accum()
            ref, cnt = orc.render_frame(vol, mcs, opts, w, h)
            err = np.abs(px - ref).max()
            print(kw, "count", count, "ok %.3fs" % (time.time() - t), "maxerr", err, "steps", st["steps"], int(cnt[0]),
                  "taps", st["taps"], int(cnt[1]), "outer", st["outer_iters"], int(cnt[2]), "ms", st["render_ms"], flush=True)
        except Exception as e:
            print(kw, "count", count, "FAILED", e, flush=True)
