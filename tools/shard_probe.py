#!/usr/bin/env python3
"""One GPU renders ONE shard of an N-way sharded C2 frame (what each rank of an N-GPU run does): kernel time
(CUDA events inside the library) against 1/N of the full-frame time, for the scheduling knobs. Finds the
within-rank loss of the multi-GPU runs without needing N GPUs."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from raymarchcl_b200.renderer import Renderer
from tests.scenes import build_scene
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--world", type=int, default=8)
ap.add_argument("--steps", type=int, default=8)
ap.add_argument("--tiles", default="16x8,32x32")
ap.add_argument("--knobs", default="256:1:1,256:1:0,256:0:1")
args = ap.parse_args()
sc = bench.WORKLOADS["c2"]["scene"]
w, h, iters = sc["width"], sc["height"], sc["iters"]
vol, opts, mcs = build_scene(**sc)
with Renderer(0) as r:
    r.set_volume(vol)
    r.clear_accum(w, h)
    r.upload_passes(opts, mcs)

    def timed(rank, world, tw, th):
        r.set_tile_shard(rank, world, tw, th)
        for _ in range(2):
            r.clear_accum(w, h); r.render_resident(0, iters)
        r.sync(); r.reset_stats()
        for _ in range(args.steps):
            r.clear_accum(w, h); r.render_resident(0, iters)
        return r.stats()["render_ms"] / args.steps

    for knob in args.knobs.split(","):
        b, k, up, sm = ([int(x) for x in knob.split(":")] + [2])[:4]  # block : rounds : bottom-up [: shared-memory map 0/1/2]
        r.set_option(10, b); r.set_option(11, k); r.set_option(13, up); r.set_option(12, sm)
        full = timed(0, 1, 32, 32)
        for tile in args.tiles.split(","):
            tw, th = [int(x) for x in tile.split("x")]
            ts = [timed(rank, args.world, tw, th) for rank in range(args.world)]
            print(json.dumps({"block": b, "round": k, "bottom_up": up, "smem": sm, "tile": tile, "world": args.world, "full_ms": round(full, 3),
                              "ideal_ms": round(full / args.world, 3), "shard_ms_min": round(min(ts), 3),
                              "shard_ms_mean": round(sum(ts) / len(ts), 3), "shard_ms_max": round(max(ts), 3),
                              "efficiency_max": round(full / args.world / max(ts), 4)}), flush=True)
