// rm_math.cuh -- fp32 vector helpers with the evaluation order the reference's OpenCL built-ins
// imply. This translation unit is compiled with -fmad=false so every a*b+c below is
// two roundings, like mad() without -cl-mad-enable; division and sqrt are IEEE (nvcc defaults).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#define RM_DEV __device__ __forceinline__

RM_DEV float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
RM_DEV float3 f3s(float s) { return make_float3(s, s, s); }
#if defined(__CUDA_ARCH__) && defined(RM_PACKED_F3)
// float3 arithmetic with the x and y lanes as one packed instruction (FADD2 / FMUL2, see add2 below)
#define RM_F3_PACKED_OP(name, ptx)                                                                            \
  RM_DEV float3 name(float3 a, float3 b) {                                                                    \
    float3 r;                                                                                                 \
    asm("{\n .reg .b64 a, b;\n mov.b64 a, {%2, %3};\n mov.b64 b, {%4, %5};\n " ptx ".rn.f32x2 a, a, b;\n mov.b64 {%0, %1}, a;\n}" \
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));                                     \
    return r;                                                                                                 \
  }
RM_F3_PACKED_OP(f3_add_xy, "add")
RM_F3_PACKED_OP(f3_sub_xy, "sub")
RM_DEV float3 operator+(float3 a, float3 b) { float3 r = f3_add_xy(a, b); r.z = a.z + b.z; return r; }
RM_DEV float3 operator-(float3 a, float3 b) { float3 r = f3_sub_xy(a, b); r.z = a.z - b.z; return r; }
// (multiplies stay scalar: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even with --fmad false and explicit
//  rounding modifiers -- one rounding instead of the reference's two)
RM_DEV float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
RM_DEV float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
#else
RM_DEV float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
RM_DEV float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
RM_DEV float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
RM_DEV float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
#endif
RM_DEV float3 operator/(float3 a, float s) { return f3(a.x / s, a.y / s, a.z / s); }
RM_DEV float3 operator/(float3 a, float3 b) { return f3(a.x / b.x, a.y / b.y, a.z / b.z); }
RM_DEV float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }

// OpenCL 6.12.4 common functions: min(x,y) = y<x?y:x, max(x,y) = x<y?y:x (NaN-order sensitive)
RM_DEV float cl_min(float x, float y) { return y < x ? y : x; }
RM_DEV float cl_max(float x, float y) { return x < y ? y : x; }
RM_DEV float cl_clamp(float x, float lo, float hi) { return cl_min(cl_max(x, lo), hi); }
RM_DEV float dot3(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
RM_DEV float3 cross3(float3 a, float3 b) {
  return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
RM_DEV float len3(float3 a) { return sqrtf(dot3(a, a)); }
// normalize(0) = 0 (pinned)
// Shared, NOT inlined (like light_pos and atmosphere in rm_scene_plain.cuh): three IEEE divisions
// and a square root per call site add up; the render kernel's code (~60 KB) is larger than the
// 32 KB L1.5 instruction cache, instruction-fetch stalls are its top stall reason, and every KB
// taken out of it shows (B200, C2: unit3 41.0 -> 39.5, + atmosphere 39.1, + light_pos 38.4 ms).
#define RM_SHARED_FN static __device__ __noinline__
#ifndef RM_UNIT3_ATTR
#define RM_UNIT3_ATTR RM_SHARED_FN
#endif
RM_UNIT3_ATTR float3 unit3(float3 a) {
  const float l = len3(a);
  return l == 0.0f ? a : a / l;
}
// 1 / x for a NORMAL x (every caller checks |x| >= 1e-5 first): the bare MUFU.RCP. __fdividef(1, x) is the same
// instruction wrapped in six more that pre-scale a denormal divisor (this build keeps denormals). Only used for
// conservative bounds (march windows, skip lengths), never for a value of the result.
RM_DEV float rcp_fast(float x) {
#ifdef __CUDA_ARCH__
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return 1.0f / x;
#endif
}
// (x, y) += (dx, dy) as ONE instruction. Blackwell (sm_100) has packed fp32 arithmetic -- SASS FADD2 / FMUL2 / FFMA2,
// PTX add / mul / fma .f32x2 on 64-bit register pairs -- that performs two IEEE round-to-nearest operations per issued
// instruction: the same bits as two scalar adds (no flush to zero), in half the issue slots. The operands stay
// ordinary floats in the source (ptxas allocates x, y and dx, dy to aligned pairs and the packing moves vanish), so
// that a conditional scalar add on one of them next to it costs nothing extra. The host build (tests/hostsim) uses
// two scalar adds.
RM_DEV void add2(float& x, float& y, float dx, float dy) {
#ifdef __CUDA_ARCH__
  asm("{\n .reg .b64 a, b;\n mov.b64 a, {%0, %1};\n mov.b64 b, {%2, %3};\n add.rn.f32x2 a, a, b;\n mov.b64 {%0, %1}, a;\n}"
      : "+f"(x), "+f"(y) : "f"(dx), "f"(dy));
#else
  x += dx;
  y += dy;
#endif
}
RM_DEV float3 lerp3(float3 a, float3 b, float t) { return a + (b - a) * t; }
// (uint)float with two's-complement wrap of the truncated value, also for negatives (8c-1)
RM_DEV uint32_t f2u_wrap(float f) { return (uint32_t)(long long)f; }
// convert_int_sat: truncate toward zero, saturate, NaN -> 0 (8c-3) == cvt.rzi.s32.f32
RM_DEV int f2i_sat(float f) { return __float2int_rz(f); }
