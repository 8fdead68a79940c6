"""Randomised TRenderOpts: the production routine (host build, tests/hostsim) against the oracle,
bit for bit. The reference only ever sets a handful of these fields (core.clj:28-74), but the
544-byte blob is the ABI, so every shortcut of the kernel has to hold for ANY values a caller can
put there: far-away eyes and huge maxDist (march windows untrusted), start distances, boxes that are
not centred, negative aoAmp (AO culling off), zero / four lights, minLightAtt cut-offs, coarse eps,
tiny iteration budgets, the ground plane above the volume, eye positions inside solid voxels."""
import struct

import numpy as np
import pytest

from oracle import build_oracle, refso
from raymarchcl_b200 import (compute_eyepos, decode_render_opts, encode_render_opts, generate_scatter_offsets,
                             make_blob_volume, make_gyroid_volume, make_terrain, render_options)
from tests.hostsim.sim import HostSim

W, H = 40, 24


def _random_fields(rng: np.random.Generator, vres: int):
    mat = rng.choice(["metal", "metal2", "ao", "orange-stripes"])
    theta, dist = rng.uniform(0, 360), rng.choice([0.3, 0.9, 1.6, 2.25, 4.0, 40.0, 500.0])
    f = render_options(dict(width=W, height=H, vres=vres, iter=1, mat=mat, dof=float(rng.choice([0.0, 0.001, 0.05])),
                            eyepos=compute_eyepos(theta, dist, float(rng.uniform(-0.5, 1.5))),
                            targetpos=[float(x) for x in rng.uniform(-0.5, 0.5, 3)], t=float(rng.uniform(0, 12)),
                            groundY=float(rng.choice([1.05, 0.3, -0.5, 0.0, 1.5]))))
    pick = lambda *v: v[int(rng.integers(len(v)))]
    f["maxDist"] = pick(30, 30, 5.0, 100.0, 1e4)
    f["startDist"] = pick(0.0, 0.0, 0.25, -0.1)
    f["eps"] = pick(0.005, 0.005, 0.05, 1e-4)
    f["maxIter"] = pick(128, 128, 16, 1)  # (0 makes the reference read an uninitialised TIsec, renderer.cl:241-256)
    f["shadowIter"] = pick(128, 128, 8, 0)
    f["maxVoxelIter"] = pick(192, 192, 64, 500, 7)
    f["aoIter"] = pick(5, 5, 0, 9)
    f["aoAmp"] = pick(0.25, 0.25, -0.25, 2.0)
    f["aoStepDist"] = pick(0.05, 0.05, 0.3)
    f["shadowBias"] = pick(0.1, 0.1, 0.0, 0.01)
    f["minLightAtt"] = pick(0.0, 0.0, 0.2)
    f["lightScatter"] = pick(0.2, 0.2, 0.0, 1.5)
    f["reflectIter"] = pick(f["reflectIter"], 0, 3, 5)
    f["isoVal"] = pick(32, 32, 0, 64, 128, 254)
    f["voxelSize"] = pick(1.0 / vres, 1.0 / vres, 0.02, 0.0)
    nl = pick(f["numLights"], 0, 1, 3, 4)
    f["numLights"] = nl
    f["lightPos"] = [[float(x) for x in rng.uniform(-3, 3, 3)] + [0.0] for _ in range(4)]
    f["lightColor"] = [[float(x) for x in rng.uniform(0, 60, 3)] for _ in range(4)]
    if rng.random() < 0.3:  # a box that is neither centred nor the unit cube
        lo = rng.uniform(-1.0, -0.5, 3)
        hi = rng.uniform(0.5, 1.0, 3)
        f["voxelBoundsMin"] = [float(x) for x in lo]
        f["voxelBoundsMax"] = [float(x) for x in hi]
    if rng.random() < 0.2:
        f["up"] = [0.1, 1.0, -0.2]
    return f


def _noise(r, solid):
    """Random bytes: the worst case for the bit-bricks / distance map (solid voxels everywhere or almost nowhere)."""
    rng = np.random.default_rng(r)
    v = rng.integers(0, 256, size=(r, r, r), dtype=np.uint8)
    return np.where(rng.random((r, r, r)) < solid, v, 0).astype(np.uint8)


VOLUMES = {"gyroid": lambda r: make_gyroid_volume(r), "terrain": lambda r: make_terrain(r),
           "blob": lambda r: make_blob_volume(r, ks=1),
           "noise_dense": lambda r: _noise(r, 0.3), "noise_sparse": lambda r: _noise(r, 0.004)}


@pytest.fixture(scope="module")
def checkers():
    build_oracle.build(verbose=False)
    return refso.load("oracle"), HostSim()


@pytest.mark.parametrize("seed", range(160))
def test_random_options_production_routine_is_bit_identical(checkers, seed):
    orc, sim = checkers
    rng = np.random.default_rng(1000 + seed)
    vres = int(rng.choice([32, 48, 64, 96]))
    vol = VOLUMES[str(rng.choice(list(VOLUMES)))](vres)
    fields = _random_fields(rng, vres)
    opts = [encode_render_opts({**fields, "frameBlend": 0.5, "time": fields["time"] + 0.333 * i}) for i in range(2)]
    assert decode_render_opts(opts[0])["numLights"] == fields["numLights"]
    mcs = [generate_scatter_offsets(0x4000, 77 + seed + i) for i in range(2)]
    ref, ref_cnt = orc.render_frame(vol, mcs, opts, W, H)
    for mode, shift in (("production", 2), ("production", 3), ("counting", 2), ("wave", 2), ("wave_counting", 2),
                        ("fused", 2), ("fused", 3), ("fused_counting", 2), ("fused_bytemap", 2)):
        px, cnt = sim.render_frame(vol, mcs, opts, W, H, mode=mode, cell_shift=shift)
        same = px.view(np.uint32) == ref.view(np.uint32)
        both_nan = np.isnan(px) & np.isnan(ref)  # NaN payloads may differ; NaN-ness may not
        assert (same | both_nan).all(), f"seed {seed} {mode}: {(~(same | both_nan)).any(axis=-1).sum()} pixels differ; {fields}"
        if mode in ("counting", "wave_counting", "fused_counting"):
            assert np.array_equal(cnt, ref_cnt)


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", [0, 3], ids=["fast", "wave"])
@pytest.mark.parametrize("seed", range(0, 160, 5))
def test_random_options_on_the_gpu(checkers, gpu_renderer, seed, kernel):
    """The same randomised blobs through the C ABI on the GPU: production and counting kernels
    agree bit for bit with each other, the work counters equal the oracle's, and the accumulator is
    within 2e-5 relative of it wherever it is finite (exp / exp2 / pow differ in the last ulp
    between CUDA and glibc; NaN / infinite pixels -- zero lights, degenerate cameras -- must be
    non-finite on both sides)."""
    orc, _ = checkers
    rng = np.random.default_rng(1000 + seed)
    vres = int(rng.choice([32, 48, 64, 96]))
    vol = VOLUMES[str(rng.choice(list(VOLUMES)))](vres)
    fields = _random_fields(rng, vres)
    opts = [encode_render_opts({**fields, "frameBlend": 0.5, "time": fields["time"] + 0.333 * i}) for i in range(2)]
    mcs = [generate_scatter_offsets(0x4000, 77 + seed + i) for i in range(2)]
    ref, ref_cnt = orc.render_frame(vol, mcs, opts, W, H)
    r = gpu_renderer
    r.set_option(2, kernel)
    r.set_option(8, 1024 if seed % 2 else 1 << 21)  # wavefront path: several chunks per frame, or one
    r.set_tile_shard(0, 1, 32, 32)
    out = {}
    for count in (True, False):
        r.set_volume(vol)
        r.clear_accum(W, H)
        r.reset_stats()
        r.count_work(count)
        r.render_frame(opts, mcs)
        st = r.stats()
        out[count] = r.read_accum()
        if count:
            assert [st["steps"], st["taps"], st["outer_iters"]] == [int(x) for x in ref_cnt]
    r.count_work(False)
    if kernel == 0:  # the layout of long launches (1024 x 1, distance map staged into shared memory by TMA): same bits
        try:
            r.set_option(10, 1024)
            r.set_option(12, 1)
            r.set_volume(vol)
            r.clear_accum(W, H)
            r.render_frame(opts, mcs)
            big = r.read_accum()
        finally:
            r.set_option(10, 0)
            r.set_option(12, 2)
        assert ((big.view(np.uint32) == out[False].view(np.uint32)) | (np.isnan(big) & np.isnan(out[False]))).all(), "layouts differ"
    r.set_option(2, 0)
    a, b = out[True].view(np.uint32), out[False].view(np.uint32)
    assert ((a == b) | (np.isnan(out[True]) & np.isnan(out[False]))).all(), "production and counting kernels differ"
    px = out[False].astype(np.float64)
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(px), fin)
    assert (np.abs(px - ref)[fin] <= 2e-5 * np.maximum(1.0, np.abs(ref[fin]))).all()


def _reference_safe(fields):
    """Keep the randomised options inside the domain where the reference's own text has defined behaviour: a
    trace that runs out of iterations returns (int)ground-distance as its material id (renderer.cl:211) and
    the reference then indexes opts->materials[] out of bounds (:392, :417) -- a wild read, a crash for far
    misses (maxDist 1e4) or a plane above the eye. The restatement and the kernels clamp the index."""
    f = dict(fields)
    f["maxIter"] = 128
    f["maxDist"] = min(float(f["maxDist"]), 30.0)
    f["groundY"] = max(float(f["groundY"]), 1.05)
    f["eyePos"] = [f["eyePos"][0], max(float(f["eyePos"][1]), -0.5), f["eyePos"][2]]
    return f


def _pin_worker(seeds):
    """Runs in a child process (a wild read of the reference must not take the test session down)."""
    build_oracle.build(verbose=False)
    orc, ref = refso.load("oracle"), refso.load("ref_strict")
    bad = []
    for seed in seeds:
        rng = np.random.default_rng(1000 + seed)
        vres = int(rng.choice([32, 48, 64, 96]))
        vol = VOLUMES[str(rng.choice(list(VOLUMES)))](vres)
        fields = _reference_safe(_random_fields(rng, vres))
        opts = [encode_render_opts({**fields, "frameBlend": 0.5, "time": fields["time"] + 0.333 * i}) for i in range(2)]
        mcs = [generate_scatter_offsets(0x4000, 77 + seed + i) for i in range(2)]
        pr, cr = ref.render_frame(vol, mcs, opts, W, H)
        po, co = orc.render_frame(vol, mcs, opts, W, H)
        same = (pr.view(np.uint32) == po.view(np.uint32)) | (np.isnan(pr) & np.isnan(po))
        if not same.all() or not np.array_equal(cr, co):
            bad.append((seed, int((~same).any(axis=-1).sum()), cr.tolist(), co.tolist()))
    return bad


def test_random_options_oracle_is_bit_identical_to_the_reference_text(ref_strict):
    """The pin of the C restatement itself on randomised blobs: oracle/rm_oracle.c against the reference's own
    kernel text (oracle/_ref/libref_strict.so, built from /root/reference/resources/renderer.cl), bit for bit on
    accumulators and work counters, 80 seeds. Skipped where oracle/_ref has not been built."""
    import multiprocessing as mp
    with mp.get_context("spawn").Pool(1) as pool:  # (not fork: the parent already runs OpenMP threads)
        res = pool.apply_async(_pin_worker, (list(range(0, 160, 2)),))
        try:
            bad = res.get(timeout=180)
        except Exception as e:  # the child died: the reference read out of bounds
            pytest.fail(f"the reference text crashed or timed out on a randomised blob: {e!r}")
    assert not bad, bad
