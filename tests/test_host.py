"""Host-side mirror of the reference's Clojure interface: options, presets, generators, .vox IO,
pipeline description, and the C-ABI library's exported surface (no GPU needed)."""
import ctypes
import math
import os
import re

import numpy as np
import pytest

import raymarchcl_b200 as rm
from raymarchcl_b200 import _lib
from raymarchcl_b200.generators import java_random_next_doubles
from raymarchcl_b200.options import OPTS_FIELDS, OPTS_OFFSETS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_opts_layout_is_544_bytes_and_non_overlapping():
    size = {"f3": 12, "i4": 16, "i2": 8, "f": 4, "i": 4, "u8": 1, "f4x4": 64, "mat4": 128}
    end = 0
    for name, off, kind in OPTS_FIELDS:
        assert off >= end, name
        end = off + size[kind]
    assert end == rm.OPTS_BYTES == 544


def test_render_options_defaults_match_reference_table():
    f = rm.render_options(dict(width=640, height=360, vres=256, iter=4, t=0.333, mat="metal"))
    assert f["maxIter"] == 128 and f["maxVoxelIter"] == 192 and f["shadowIter"] == 128
    assert f["isoVal"] == 32 and f["aoIter"] == 5 and f["maxDist"] == 30
    assert f["voxelRes"] == [256, 256, 256, 65536]
    assert f["frameBlend"] == 0.25 and f["voxelSize"] == 1.0 / 256
    assert f["fov"] == math.radians(90)
    assert f["aoAmp"] == 0.25 and f["reflectIter"] == 3 and f["numLights"] == 2   # preset overrides
    assert f["lightPos"][0] == [0, 2, 0, 0]
    # unknown material falls back to :ao (core.clj:74)
    g = rm.render_options(dict(width=8, height=8, vres=16, iter=1, mat="nope"))
    assert g["numLights"] == 1 and g["reflectIter"] == 0


def test_encode_decode_roundtrip():
    f = rm.render_options(dict(width=1920, height=1080, vres=[256, 128, 64], iter=16, t=1.332, mat="metal2",
                               eyepos=[1.5, 0.35, -1.5], dof=0.025))
    blob = rm.encode_render_opts(f)
    assert len(blob) == 544
    d = rm.decode_render_opts(blob)
    assert d["voxelRes"] == [256, 128, 64, 256 * 128]
    assert d["resolution"] == [1920, 1080]
    assert d["isoVal"] == 32 and d["numLights"] == 2
    assert d["time"] == np.float32(1.332)
    assert d["materials"][3]["r0"] == np.float32(0.75)
    assert d["lightColor"][1][:3] == [8.0, 18.0, 28.0]
    assert blob[OPTS_OFFSETS["mcTableLength"]:OPTS_OFFSETS["mcTableLength"] + 4] == b"\0\0\0\0"


def test_per_pass_buffers():
    bufs = rm.make_render_option_buffers(3, dict(width=16, height=8, vres=32, mat="ao"))
    times = [rm.decode_render_opts(b)["time"] for b in bufs]
    assert times == [np.float32(0.0), np.float32(0.333), np.float32(0.666)]
    assert rm.decode_render_opts(bufs[0])["frameBlend"] == np.float32(1 / 3)


def test_java_random_known_answers():
    # java.util.Random(42).nextDouble() x2 and Random(0).nextDouble(): published JDK behaviour
    assert java_random_next_doubles(42, 2).tolist() == [0.7275636800328681, 0.6832234717598454]
    assert java_random_next_doubles(0, 1)[0] == 0.730967787376657


def test_scatter_table():
    t = rm.generate_scatter_offsets(0x4000, 1000)
    assert t.dtype == np.float32 and t.size == 65536
    n = np.linalg.norm(t.reshape(-1, 4).astype(np.float64), axis=1)
    assert np.allclose(n, 1.0, atol=1e-6)
    assert not np.array_equal(t, rm.generate_scatter_offsets(0x4000, 1001))


def test_gyroid_volume_values_and_slabs():
    v = rm.make_gyroid_volume(64)
    assert v.shape == (64, 64, 64) and set(np.unique(v)) <= {0, 64, 128, 255}
    assert not v[:32].any() and v[32:].any()          # only slabs with (z & 63) >= 32
    assert not (v[:, :, :32] == 128).any() and not (v[:, :, 32:] == 64).any()
    assert 0.02 < (v > 32).mean() < 0.15


def test_vox_roundtrip(tmp_path):
    v = rm.make_gyroid_volume([16, 24, 32])
    p = str(tmp_path / "g.vox")
    rm.save_volume(p, v)
    raw = open(p, "rb").read()
    assert raw[:5] == b"VOXEL" and raw[5:17] == b"\0\0\0\x10\0\0\0\x18\0\0\0\x20" and raw[17] == 1
    assert len(raw) == 18 + v.size
    assert np.array_equal(rm.load_volume(p), v)
    with open(p, "wb") as f:
        f.write(raw[:100])
    with pytest.raises(ValueError):
        rm.load_volume(p)


def test_pipeline_description_matches_reference_step_list():
    from raymarchcl_b200.renderer import make_pipeline
    state = {"opts-buffers": [b""] * 3, "num": 12}
    steps = make_pipeline(state)
    names = [s.get("name", "write") for s in steps]
    assert names == ["write", "write", "RenderImage", "write", "RenderImage", "write", "RenderImage", "write", "TonemapImage"]
    assert steps[-1]["read"] == ["out"] and steps[-1]["in"][1] == ("o-buf", 0)


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "raymarch_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(rm_[a-z_]+)\s*\(", header)))
    assert declared == sorted(_lib.EXPORTS)
    if not os.path.exists(_lib.LIB_PATH):
        pytest.skip("libraymarch_b200.so not built")
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    lib.rm_abi_version.restype = ctypes.c_int
    assert lib.rm_abi_version() == 2


def test_no_cpu_fallback_without_gpu():
    """The product path must fail loudly when no device is usable."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    if not os.path.exists(_lib.LIB_PATH):
        pytest.skip("libraymarch_b200.so not built")
    from raymarchcl_b200.renderer import Renderer
    with pytest.raises(_lib.RaymarchError) as e:
        Renderer(0)
    assert e.value.code == -6


def test_product_code_never_touches_the_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "raymarchcl_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"(from|import)\s+oracle|oracle/|librm_oracle|libref_", txt):
                    bad.append(f)
    assert not bad, bad


def test_pass_chunking_of_the_default_kernel():
    """rm_persist_pick_passes (csrc/rm_kernels.h, compiled for the host by tests/hostsim): a frame's passes are cut
    into launches of m <= 32 whose bundles (32 // m pixels x m passes) fill >= 80 % of a warp, else a power of two."""
    import ctypes as C
    from tests.hostsim import build_hostsim
    lib = C.CDLL(build_hostsim.build())
    lib.sim_pick_passes.argtypes = [C.c_int]
    for n in range(1, 201):
        left, chunks = n, []
        while left:
            m = lib.sim_pick_passes(left)
            assert 1 <= m <= min(left, 32)
            assert (32 // m) * m >= 26 or (m & (m - 1)) == 0, (n, m)
            chunks.append(m)
            left -= m
        assert sum(chunks) == n
    assert [lib.sim_pick_passes(k) for k in (16, 100, 11, 17, 3, 33)] == [16, 32, 8, 16, 3, 32]


def test_packed_fp32_is_used_and_never_contracted():
    """The default kernel adds the (x, y) lanes of its float3 values with Blackwell's packed FADD2 (csrc/rm_math.cuh).
    ptxas contracts a packed multiply feeding a packed add into FFMA2 even with --fmad false and explicit .rn
    modifiers -- one rounding where the reference has two -- so the library must hold packed ADDS only: no FFMA2 and no
    FMUL2 anywhere in its SASS."""
    import shutil
    import subprocess
    if not os.path.exists(_lib.LIB_PATH):
        pytest.skip("libraymarch_b200.so not built")
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    assert len(re.findall(r"\bFADD2\b", sass)) > 100
    assert not re.findall(r"\b(FFMA2|FMUL2)\b", sass)


def test_layout_choice_of_the_default_kernel():
    """rm_persist_pick_layout (csrc/rm_kernels.h, compiled for the host by tests/hostsim): one 1024-thread block per SM with
    the TMA-staged shared-memory distance map for long launches whose map fits, five 256-thread blocks with the byte map in
    global memory otherwise; RM_OPT_PERSIST_BLOCK / _SMEM override either half."""
    import ctypes as C
    from tests.hostsim import build_hostsim
    lib = C.CDLL(build_hostsim.build())
    lib.sim_pick_layout.argtypes = [C.c_longlong, C.c_int, C.c_uint, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]

    def pick(bundles, nib_bytes, block=0, smem=2, counting=0, sms=148):
        out = (C.c_int * 3)()
        lib.sim_pick_layout(bundles, sms, nib_bytes, block, smem, counting, out)
        return tuple(out)

    c2 = 1920 * 1080 * 16 // 32
    nib256 = 64 ** 3 // 2                                   # 128 KiB: C2's map
    assert pick(c2, nib256) == (1024, 1, 1)                 # C2, one GPU
    assert pick(c2 // 8, nib256) == (1024, 1, 1)            # one rank's shard of eight
    assert pick(256 * 256 // 32, 16 ** 3 // 2) == (256, 5, 0)   # C1: a short launch does not pay for staging the map
    assert pick(4 * 148 * 32 - 1, nib256) == (256, 5, 0) and pick(4 * 148 * 32, nib256) == (1024, 1, 1)  # the threshold
    assert pick(960 * 540 * 4 // 32, nib256) == (1024, 1, 1)  # 13.7 bundles per warp slot: measured 2.45 vs 2.55 ms
    assert pick(c2, 160 * 160 * 33 // 2) == (256, 5, 0)     # 422 KB of nibbles: does not fit an SM
    assert pick(c2, 0) == (256, 5, 0)                       # no 4-bit map built
    assert pick(c2, nib256, smem=0) == (256, 5, 0)          # shared-memory map switched off
    assert pick(c2, nib256, block=256) == (256, 5, 0)       # 5 x 128 KiB do not fit
    assert pick(c2, 16 * 1024, block=256) == (256, 5, 0)    # ... and where they would, automatic mode still says no
    assert pick(c2, 16 * 1024, block=256, smem=1) == (256, 5, 1)
    assert pick(c2, nib256, block=1024, smem=0) == (1024, 1, 0)
    assert pick(100, nib256, block=1024) == (1024, 1, 1)    # forced layout: the map comes along whenever it fits
    assert pick(100, nib256, counting=1) == (1024, 1, 1)    # counting kernels: always the big layout
    assert pick(100, nib256, counting=1, smem=0) == (1024, 1, 0)
