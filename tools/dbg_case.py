"""Debug helper: render one small scene through a chosen kernel (KERNEL=0|1|2, TRIPS = watchdog
limit of kernel 2) and compare with the oracle. usage: dbg_case.py VRES WIDTH HEIGHT PASSES"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from raymarchcl_b200.renderer import Renderer
from tests.scenes import build_scene
from oracle import refso, build_oracle
build_oracle.build(verbose=False)
orc = refso.load("oracle")
r = Renderer(0)
r.set_option(2, int(os.environ.get("KERNEL", "2")))
r.set_option(7, int(os.environ.get("TRIPS", "2000000")))
vres, w, h, it = [int(a) for a in sys.argv[1:5]]
kw = dict(vres=vres, width=w, height=h, iters=it, mat="metal")
vol, opts, mcs = build_scene(**kw)
for count in (True, False):
    try:
        r.set_volume(vol); r.clear_accum(w, h); r.reset_stats(); r.count_work(count)
        r.render_frame(opts, mcs)
        st = r.stats(); px = r.read_accum()
        ref, cnt = orc.render_frame(vol, mcs, opts, w, h)
        print(os.environ.get("RAYMARCH_B200_LIB"), kw, "count", count, "maxerr", np.abs(px - ref).max(), "steps", st["steps"], int(cnt[0]), "ms", st["render_ms"], flush=True)
    except Exception as e:
        print(os.environ.get("RAYMARCH_B200_LIB"), "count", count, "FAILED", e, flush=True)
        break
