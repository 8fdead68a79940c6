// TEST INFRASTRUCTURE ONLY. A stand-in for <cuda_runtime.h> that lets g++ compile the device
// routine of raymarchcl_b200/csrc/rm_scene_plain.cuh for the host (tests/hostsim/hostsim.cpp), so
// that the production algorithm -- fetch elision, culling, closed-form recurrence jumps -- can be
// checked BIT FOR BIT against the oracle on the CPU. Never part of the product library.
#pragma once
#include <math.h>
#include <cmath>
#include <algorithm>
#include <cstdint>
#include <cstddef>
#include <cstring>

#define __device__
#define __host__
#define __constant__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))

struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct float4 { float x, y, z, w; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float3 make_float3(float x, float y, float z) { return float3{x, y, z}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

template <class T> static inline T __ldg(const T* p) { return *p; }
// cvt.rzi.s32.f32: truncate toward zero, saturate, NaN -> 0
static inline int __float2int_rz(float f) {
  if (f != f) return 0;
  if (f >= 2147483648.0f) return 2147483647;
  if (f <= -2147483648.0f) return -2147483647 - 1;
  return (int)f;
}
static inline float __fdividef(float a, float b) { return a / b; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline unsigned __float_as_uint(float f) { unsigned i; std::memcpy(&i, &f, 4); return i; }
static inline float __uint_as_float(unsigned i) { float f; std::memcpy(&f, &i, 4); return f; }
typedef int cudaError_t;
typedef void* cudaStream_t;
