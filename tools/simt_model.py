#!/usr/bin/env python3
"""SIMT lane-utilisation model of the render op, from the host simulation (tests/hostsim) -- the
numbers behind the table in DESIGN.md section 6. No GPU needed; ~1 minute on 8 cores.

The production routine is run on the CPU for a reduced C2 frame (256^3 gyroid, 480x272, 16 passes,
:metal); its statistics hooks log, per pixel-sample and per call site (primary trace, AO probe k /
shadow ray of light i / bounce trace of level L), a cost estimate in instructions: 15 per
ground-only evaluation, 150 per full distanceToScene call, 50 per table lookup, 3 per skipped
sample. Pixel-samples are then grouped into warps exactly like the fused kernel's items
(slot-major / pass-minor over 8x4 pixel blocks) and different execution schemes are costed:

  lock step per call site      each warp pays, per call site, the maximum over its 32 lanes
  lane-private batches         a lane runs several of its own jobs back to back; the warp pays the
                               maximum over lanes of the sums
  compacted jobs, no refill    wavefront with one job per thread, warps of 32 consecutive jobs
  warp job pool                the jobs of a warp's 32 items per wave, list-scheduled over 32 lanes
  refill, ideal                every lane always busy: 32

"lanes" = total thread cost / total lock-step cost = useful lanes per issued instruction. That model
ignores divergence INSIDE one distanceToScene evaluation; march_model() adds it: the number of table
lookups per evaluation is heavy-tailed (mean 3.6, tail beyond 15), neighbouring pixel-samples have
correlated march lengths (9 of 32 lanes busy in the fused kernel's march loop -- what ncu measures),
unrelated rays do not (5 of 32 -- what the refilling wavefront kernel measured).
"""
from __future__ import annotations

import argparse
import ctypes as C
import heapq
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tests.hostsim.sim import HostSim  # noqa: E402
from tests.scenes import build_scene  # noqa: E402


def site_costs(width: int, height: int, passes: int, vres: int, mat: str) -> np.ndarray:
    """cost[pass, pixel, 64 call sites] of the production routine."""
    sim = HostSim()
    vol, opts, mcs = build_scene(vres=vres, width=width, height=height, iters=passes, mat=mat)
    n = width * height
    cost = np.zeros((passes, n, 64), np.float32)
    sim.lib.sim_set_cost_buffer.argtypes = [C.c_void_p]
    vox = np.ascontiguousarray(vol).reshape(-1)
    px = np.zeros((height, width, 4), np.float32)
    for p in range(passes):
        sim.lib.sim_set_cost_buffer(cost[p].ctypes.data)
        sim.lib.sim_render_pixels(vox, np.ascontiguousarray(mcs[p]).reshape(-1), opts[p], px.reshape(-1), n,
                                  None, 0, None, 0, 2)
    sim.lib.sim_set_cost_buffer(None)
    return cost


def eval_logs(width: int, height: int, passes: int, vres: int, mat: str) -> np.ndarray:
    """log[pass, pixel, 256]: the full distanceToScene evaluations of each pixel-sample in order,
    (call site << 8 | table lookups of its march), 0xffff = end."""
    sim = HostSim()
    vol, opts, mcs = build_scene(vres=vres, width=width, height=height, iters=passes, mat=mat)
    n = width * height
    log = np.zeros((passes, n, 256), np.uint16)
    sim.lib.sim_set_eval_buffer.argtypes = [C.c_void_p]
    vox = np.ascontiguousarray(vol).reshape(-1)
    px = np.zeros((height, width, 4), np.float32)
    for p in range(passes):
        sim.lib.sim_set_eval_buffer(log[p].ctypes.data)
        sim.lib.sim_render_pixels(vox, np.ascontiguousarray(mcs[p]).reshape(-1), opts[p], px.reshape(-1), n,
                                  None, 0, None, 0, 2)
    sim.lib.sim_set_eval_buffer(None)
    return log


def march_model(width: int, height: int, passes: int, vres: int, mat: str, sample: int = 3000) -> None:
    """Divergence INSIDE the full evaluations: how many lanes are busy in the march loop when the
    evaluations a warp runs together are (a) the fused kernel's -- the k-th evaluation of the same
    call site of neighbouring pixel-samples -- or (b) unrelated, as after per-lane refill."""
    log = eval_logs(width, height, passes, vres, mat)
    warps = as_warps(log, width, height)
    valid = warps != 0xFFFF
    site, cnt = (warps >> 8).astype(np.int32), (warps & 255).astype(np.int32)
    n_ps = log.shape[0] * log.shape[1]
    print(f"\nfull evaluations per pixel-sample {valid.sum() / n_ps:.1f}, lookups per evaluation {cnt[valid].sum() / valid.sum():.2f}")
    rng = np.random.default_rng(0)
    ev_slots = lk_slots = ev = lk = 0
    for wi in rng.choice(len(warps), size=min(sample, len(warps)), replace=False):
        s_, c_, v_ = site[wi], cnt[wi], valid[wi]
        for st in np.unique(s_[v_]):
            seqs = [c_[lane][(s_[lane] == st) & v_[lane]] for lane in range(32)]
            m = max(len(q) for q in seqs)
            a = np.zeros((32, m), np.int32)
            for lane, q in enumerate(seqs):
                a[lane, :len(q)] = q
                ev += len(q)
            ev_slots += m
            lk_slots += int(a.max(axis=0).sum())
            lk += int(a.sum())
    print(f"fused kernel (lock step per call site and iteration): {ev / ev_slots:4.1f} of 32 lanes in an evaluation, "
          f"{lk / lk_slots:4.1f} of 32 in its march loop")
    allc = cnt[valid]
    for k in (32, 16):
        d = rng.choice(allc, size=(200000, k))
        print(f"{k} unrelated evaluations together (per-lane refill):  {k * d.mean() / d.max(axis=1).mean():4.1f} of {k} lanes in the march loop")


def as_warps(cost: np.ndarray, width: int, height: int) -> np.ndarray:
    """[warps, 32 lanes, 64 sites] in the fused kernel's item order (rm_kernels.h:rm_slot_to_pixel)."""
    passes, n = cost.shape[0], cost.shape[1]
    ys, xs = np.meshgrid(np.arange(height), np.arange(width), indexing="ij")
    slot = (((ys // 4) * (width // 8) + xs // 8) * 32 + (ys % 4) * 8 + xs % 8).reshape(-1)
    order = np.argsort(slot)
    k = cost.shape[2]
    items = cost[:, order, :].transpose(1, 0, 2).reshape(n * passes, k)
    return items.reshape(-1, 32, k)


def ao_sites(level):
    return list(range(level * 16 + 8, level * 16 + 16))


def shadow_sites(level):
    return [level * 16 + 1, level * 16 + 2, level * 16 + 3, level * 16 + 4]


def lockstep(warps, groups):
    """groups: lists of call sites run back to back by each lane; cost = sum over groups of max over lanes."""
    return float(sum(warps[:, :, g].sum(axis=2).max(axis=1).sum() for g in groups))


def makespan(jobs, workers=32, setup=0.0):
    if len(jobs) == 0:
        return 0.0
    h = [0.0] * workers
    heapq.heapify(h)
    for c in jobs:
        heapq.heappush(h, heapq.heappop(h) + c + setup)
    return max(h)


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--width", type=int, default=480)
    ap.add_argument("--height", type=int, default=272)
    ap.add_argument("--passes", type=int, default=16)
    ap.add_argument("--vres", type=int, default=256)
    ap.add_argument("--mat", default="metal")
    ap.add_argument("--pool-sample", type=int, default=4000, help="warps sampled for the job-pool model")
    args = ap.parse_args()
    assert args.width % 8 == 0 and args.height % 4 == 0 and 32 % args.passes == 0

    cost = site_costs(args.width, args.height, args.passes, args.vres, args.mat)
    warps = as_warps(cost, args.width, args.height)
    total = float(warps.sum())
    print(f"thread cost per pixel-sample (model instructions): {total / (cost.shape[0] * cost.shape[1]):.0f}")

    per_site = [[0]]
    for L in range(4):
        per_site += [[c] for c in ao_sites(L)] + [[c] for c in shadow_sites(L)] + ([[(L + 1) * 16]] if L < 3 else [])
    base = lockstep(warps, per_site)

    def row(name, t):
        print(f"{name:52s} lanes {total / t:5.2f}   time vs lock step {t / base:5.3f}")

    row("fused kernel: lock step per call site", base)
    for name, with_ao, with_bounce in (("lane-private batch {shadows}", False, False),
                                       ("lane-private batch {shadows, next bounce}", False, True),
                                       ("lane-private batch {AO, shadows, next bounce}", True, True)):
        g = [[0]]
        for L in range(4):
            batch = shadow_sites(L) + ([(L + 1) * 16] if with_bounce and L < 3 else [])
            if with_ao:
                g.append(ao_sites(L) + batch)
            else:
                g += [[c] for c in ao_sites(L)] + [batch]
                if not with_bounce and L < 3:
                    g.append([(L + 1) * 16])
        row(name, lockstep(warps, g))
    row("everything after the primary hit as one batch", lockstep(warps, [[0], list(range(1, 64))]))

    # wavefront, one compacted job per thread: jobs in item order, warps of 32 consecutive jobs
    items = warps.reshape(-1, 64)
    t = float(items[:, 0].reshape(-1, 32).max(axis=1).sum())
    for L in range(4):
        for cols, per_job in ((ao_sites(L), False), (shadow_sites(L) + ([(L + 1) * 16] if L < 3 else []), True)):
            jobs = items[:, cols].reshape(-1) if per_job else items[:, cols].sum(axis=1)
            jobs = jobs[jobs > 0]
            jobs = np.concatenate([jobs, np.zeros((-len(jobs)) % 32, np.float32)])
            t += float(jobs.reshape(-1, 32).max(axis=1).sum())
    row("wavefront, one compacted job per thread, no refill", t)

    # warp-level job pool per wave (AO of a surface = one job), list scheduling, 150 instructions per job switch
    rng = np.random.default_rng(0)
    idx = rng.choice(len(warps), size=min(args.pool_sample, len(warps)), replace=False)
    s_tot, t_tot = 0.0, 0.0
    for wi in idx:
        w = warps[wi]
        s_tot += float(w.sum())
        t_tot += float(w[:, 0].max())
        for L in range(4):
            jobs = np.concatenate([w[:, ao_sites(L)].sum(axis=1),
                                   w[:, shadow_sites(L) + ([(L + 1) * 16] if L < 3 else [])].reshape(-1)])
            t_tot += makespan(jobs[jobs > 0], 32, 150.0)
    print(f"{'warp job pool per wave (150 per job switch)':52s} lanes {s_tot / t_tot:5.2f}")
    print(f"{'refilling trace kernel, ideal':52s} lanes 32.00")

    # multi-GPU: interleaved tile ownership
    pix = cost.sum(axis=(0, 2)).reshape(args.height, args.width)
    print("\ntile balance, max / mean load over ranks (tile edge in pixels of this reduced frame):")
    for tile in (4, 8, 16):
        out = {}
        for world in (2, 4, 8):
            ty, tx = (args.height + tile - 1) // tile, (args.width + tile - 1) // tile
            loads = np.zeros(world)
            for j in range(ty):
                for i in range(tx):
                    loads[(j * tx + i) % world] += pix[j * tile:(j + 1) * tile, i * tile:(i + 1) * tile].sum()
            out[world] = round(float(loads.max() / loads.mean()), 4)
        print(f"  tile {tile:2d}: {out}")

    march_model(args.width // 2, args.height // 2, args.passes, args.vres, args.mat)


if __name__ == "__main__":
    main()
