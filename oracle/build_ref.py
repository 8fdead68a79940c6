#!/usr/bin/env python3
"""Build the reference's own kernel text into oracle/_ref/ -- TEST INFRASTRUCTURE ONLY.

Reads /root/reference/resources/renderer.cl where it lies, applies three mechanical rewrites
into a TEMPORARY directory (nothing of the reference's text is written into the repository):

  1. OpenCL vector literals  `(float3)(a, b, c)`  ->  constructor calls `float3(a, b, c)`
  2. swizzles                `v.xyz`              ->  member calls `v.xyz()`
  3. (strict build only) work counters after the inner-march loop head (renderer.cl:219), the
     occupancy-tap function head (renderer.cl:172) and the sphere-trace loop head (renderer.cl:243)

and compiles oracle/ref_driver.cpp + oracle/clshim.h around it with g++:

  oracle/_ref/libref_strict.so  -O2 -ffp-contract=off, counters on   = parity anchor
  oracle/_ref/libref_fast.so    -O3 -ffast-math -march=x86-64-v3     = CPU timing baseline
                                (mirrors -cl-fast-relaxed-math -cl-mad-enable, core.clj:128)

oracle/_ref/ is git-ignored but not gpurun-ignored: the .so files travel to the GPU box, where
/root/reference does not exist. If the reference tree is absent this script leaves existing
builds alone and reports so.
"""
from __future__ import annotations

import os
import re
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF_CL = os.environ.get("RM_REFERENCE_CL", "/root/reference/resources/renderer.cl")

VARIANTS = {
    "strict": dict(counters=1, flags=["-O2", "-ffp-contract=off"]),
    "fast": dict(counters=0, flags=["-O3", "-ffast-math", "-march=x86-64-v3"]),
}


def transform(text: str, counters: bool) -> str:
    text = re.sub(r"\((float2|float3|float4|int3)\)\(", r"\1(", text)
    text = re.sub(r"\.(xyz|xyy|yxy|yyx|zxy|zyx|zw)\b", r".\1()", text)
    if counters:
        n_total = 0
        text, n = re.subn(r"(while\s*\(\s*--steps\s*>=\s*0\s*\)\s*\{)", r"\1 RM_CNT(step);", text)
        n_total += n
        text, n = re.subn(r"(while\s*\(\s*--maxSteps\s*>=\s*0\s*\)\s*\{)", r"\1 RM_CNT(outer);", text)
        n_total += n
        text, n = re.subn(r"(float\s+voxelLookupI\s*\([^;{]*\)\s*\{)", r"\1 RM_CNT(tap);", text)
        n_total += n
        if n_total != 3:
            raise RuntimeError(f"counter injection matched {n_total} sites, expected 3")
    return text


def build(verbose: bool = True) -> bool:
    if not os.path.exists(REF_CL):
        have = all(os.path.exists(os.path.join(OUT, f"libref_{v}.so")) for v in VARIANTS)
        if verbose:
            print(f"[build_ref] {REF_CL} not present; "
                  f"{'keeping prebuilt oracle/_ref' if have else 'oracle/_ref NOT available'}")
        return have
    os.makedirs(OUT, exist_ok=True)
    src = open(REF_CL, "r").read()
    with tempfile.TemporaryDirectory(prefix="rm_ref_") as tmp:
        for name, cfg in VARIANTS.items():
            xf = os.path.join(tmp, f"renderer_{name}.inc")
            with open(xf, "w") as f:
                f.write(transform(src, bool(cfg["counters"])))
            so = os.path.join(OUT, f"libref_{name}.so")
            cmd = ["g++", "-std=c++17", "-shared", "-fPIC", "-fopenmp", "-w", *cfg["flags"],
                   f"-DRM_COUNTERS={cfg['counters']}", f'-DRM_REF_SOURCE="{xf}"',
                   "-I", HERE, os.path.join(HERE, "ref_driver.cpp"), "-o", so]
            if verbose:
                print("[build_ref]", " ".join(cmd))
            subprocess.check_call(cmd)
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
