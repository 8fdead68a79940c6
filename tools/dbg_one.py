import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from raymarchcl_b200.renderer import Renderer
from tests.scenes import build_scene
r = Renderer(0)
r.set_option(7, 200000)
vres = int(sys.argv[1]) if len(sys.argv) > 1 else 256
kw = dict(vres=vres, width=96, height=64, iters=2, mat="metal")
vol, opts, mcs = build_scene(**kw)
r.set_volume(vol); r.clear_accum(96, 64); r.count_work(False)
try:
    r.render_frame(opts, mcs)
    print("ok", r.stats())
except Exception as e:
    print("FAILED", e)
