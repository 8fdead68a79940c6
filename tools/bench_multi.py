#!/usr/bin/env python3
"""The multi-GPU path behind the C ABI (rm_create_multi: ONE process, one context over N GPUs, the frame
assembled on GPU 0 by the render kernels' own peer stores) on BASELINE config 2 -- the same frame bench.py
times under torchrun with one process per GPU and an NCCL gather. Prints one JSON line per N.

  python tools/bench_multi.py --gpus 1,2,4,8 --steps 20

resident: volume / tables / opts on the devices; per frame rm_clear_accum + rm_render_resident +
          rm_tonemap_async into alternating pinned buffers (the read-back of frame k overlaps frame k+1);
          wall clock over `steps` frames after warm-up.
e2e:      per frame rm_set_volume (one PCIe upload + NVLink broadcast) + rm_clear_accum + rm_render_frame
          (tables: one upload + NVLink broadcast) + rm_tonemap (blocking read-back), wall clock.
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from raymarchcl_b200.renderer import Renderer
from tests.scenes import build_scene

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", default="1,2")
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--workload", default="c2")
ap.add_argument("--tile", default="16,8")
args = ap.parse_args()
import bench
sc = bench.WORKLOADS[args.workload]["scene"]
w, h, iters = sc["width"], sc["height"], sc["iters"]
vol, opts, mcs = build_scene(**sc)
tw, th = [int(x) for x in args.tile.split(",")]
for n in [int(x) for x in args.gpus.split(",")]:
    with Renderer(list(range(n))) as g:
        g.set_tile_shard(0, 1, tw, th)
        g.set_volume(vol)
        g.clear_accum(w, h)
        g.upload_passes(opts, mcs)
        g.count_work(True)
        g.reset_stats()
        g.render_resident(0, iters)
        steps_frame = g.stats()["steps"]
        g.count_work(False)
        host = [g.alloc_pinned_argb() for _ in range(2)]

        def frame(f):
            g.clear_accum(w, h)
            g.render_resident(0, iters)
            g.tonemap_async(opts[0], host[f & 1], f & 1)

        for f in range(args.warmup):
            frame(f)
        g.wait(0); g.wait(1)
        g.reset_stats()
        t0 = time.perf_counter()
        for f in range(args.steps):
            if f >= 2:
                g.wait(f & 1)
            frame(f)
        g.wait(0); g.wait(1)
        dt = time.perf_counter() - t0
        members = [g.member_stats(i)["render_ms"] / args.steps for i in range(n)]
        ref = host[(args.steps - 1) & 1].copy()
        # end to end from host buffers
        for _ in range(2):
            g.set_volume(vol); g.clear_accum(w, h); g.render_frame(opts, mcs); out = g.tonemap(opts[0])
        t0 = time.perf_counter()
        for _ in range(args.steps):
            g.set_volume(vol); g.clear_accum(w, h); g.render_frame(opts, mcs); out = g.tonemap(opts[0])
        dte = time.perf_counter() - t0
        assert np.array_equal(out.reshape(-1), ref), "resident and host-buffer frames differ"
        g.free_pinned(host)
    print(json.dumps({"path": "rm_create_multi (single process, peer stores into GPU 0's frame)", "n_gpus": n, "tile": [tw, th],
                      "workload": bench.WORKLOADS[args.workload]["name"], "steps": args.steps,
                      "ms_per_frame_resident": 1e3 * dt / args.steps, "Mray_steps_per_s": steps_frame * args.steps / dt / 1e6,
                      "render_ms_per_member": {"min": min(members), "mean": sum(members) / n, "max": max(members)},
                      "ms_per_frame_e2e": 1e3 * dte / args.steps, "timing": "host wall clock around the frame loop"}), flush=True)
