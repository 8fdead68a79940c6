// clshim.h -- TEST INFRASTRUCTURE ONLY (oracle/). Not part of the product path.
//
// A minimal OpenCL-C 1.x compatibility layer for g++ so that the reference's own kernel text
// (/root/reference/resources/renderer.cl) can be compiled as host C++ WITHOUT copying it into
// this repository: oracle/build_ref.py reads the .cl file where it lies, applies two mechanical
// rewrites (vector literals -> constructors, swizzles -> member calls) into a temp dir and
// #includes the result inside namespace refcl (see oracle/ref_driver.cpp).
//
// Semantics pinned here (the reference leaves them to the OpenCL implementation):
//   * float3/int3 occupy 16 bytes (OpenCL 6.1.5) -> sizeof(TRenderOpts) == 544.
//   * min(x,y) = y<x?y:x ; max(x,y) = x<y?y:x ; clamp = min(max(x,lo),hi)   (OpenCL 6.12.4)
//   * step(e,x) = x<e?0:1 ; mix(a,b,t) = a+(b-a)*t ; mad(a,b,c) = a*b+c as TWO roundings
//   * dot = x*x' + y*y' + z*z' left to right ; length = sqrt(dot) ; normalize(v) = v/length(v),
//     normalize(0) = 0
//   * convert_int3_sat: truncate toward zero, saturate to int range, NaN -> 0
//   * (uint)float : two's-complement wrap of the truncated value ((uint32)(int64)x), also for
//     negatives (x86 / CPU-OpenCL behaviour)
//   * transcendental functions are the float versions of the host libm.
#pragma once
#include <cmath>
#include <cstdint>

namespace refcl {

typedef unsigned char uchar;

// `uint` as a class so that the C-style casts `(uint)(float expr)` in the kernel text get the
// pinned wrap-around conversion instead of C++ undefined behaviour.
struct uint {
  uint32_t v;
  uint() : v(0) {}
  uint(float f) : v((uint32_t)(int64_t)f) {}
  uint(int i) : v((uint32_t)i) {}
  uint(unsigned i) : v(i) {}
  operator uint32_t() const { return v; }
  uint& operator+=(int k) { v += (uint32_t)k; return *this; }
};
inline uint operator+(uint a, uint b) { return uint(a.v + b.v); }
inline uint32_t operator&(uint a, int m) { return a.v & (uint32_t)m; }

struct float2;
struct float3;
struct float4;
struct int3;

struct alignas(8) float2 {
  float x, y;
  float2() : x(0), y(0) {}
  float2(float s) : x(s), y(s) {}
  float2(float a, float b) : x(a), y(b) {}
};
struct alignas(16) float3 {
  float x, y, z, _pad;
  float3() : x(0), y(0), z(0), _pad(0) {}
  float3(float s) : x(s), y(s), z(s), _pad(0) {}
  float3(float a, float b, float c) : x(a), y(b), z(c), _pad(0) {}
  float3 zxy() const { return float3(z, x, y); }
  float3 zyx() const { return float3(z, y, x); }
  float3 xyz() const { return *this; }
};
struct alignas(16) float4 {
  float x, y, z, w;
  float4() : x(0), y(0), z(0), w(0) {}
  float4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
  float4(const float3& v, float d) : x(v.x), y(v.y), z(v.z), w(d) {}
  float3 xyz() const { return float3(x, y, z); }
  float2 zw() const { return float2(z, w); }
};
struct alignas(8) int2 { int x, y; };
struct alignas(16) int3 {
  int x, y, z, _pad;
  int3() : x(0), y(0), z(0), _pad(0) {}
  int3(int a, int b, int c) : x(a), y(b), z(c), _pad(0) {}
  int3 xyy() const { return int3(x, y, y); }
  int3 yxy() const { return int3(y, x, y); }
  int3 yyx() const { return int3(y, y, x); }
};
struct alignas(16) int4 {
  int x, y, z, w;
  int3 xyz() const { return int3(x, y, z); }
};

// ---- float2 ----
inline float2 operator+(float2 a, float2 b) { return float2(a.x + b.x, a.y + b.y); }
inline float2 operator-(float2 a, float s) { return float2(a.x - s, a.y - s); }
inline float2 operator*(float2 a, float s) { return float2(a.x * s, a.y * s); }
inline float2 operator/(float2 a, float2 b) { return float2(a.x / b.x, a.y / b.y); }

// ---- float3 ----
inline float3 operator+(float3 a, float3 b) { return float3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline float3 operator-(float3 a, float3 b) { return float3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline float3 operator*(float3 a, float3 b) { return float3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline float3 operator/(float3 a, float3 b) { return float3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline float3 operator+(float3 a, float s) { return float3(a.x + s, a.y + s, a.z + s); }
inline float3 operator+(float s, float3 a) { return float3(s + a.x, s + a.y, s + a.z); }
inline float3 operator-(float3 a, float s) { return float3(a.x - s, a.y - s, a.z - s); }
inline float3 operator*(float3 a, float s) { return float3(a.x * s, a.y * s, a.z * s); }
inline float3 operator*(float s, float3 a) { return float3(s * a.x, s * a.y, s * a.z); }
inline float3 operator/(float3 a, float s) { return float3(a.x / s, a.y / s, a.z / s); }
inline float3 operator-(float3 a) { return float3(-a.x, -a.y, -a.z); }
inline float3& operator+=(float3& a, float3 b) { a = a + b; return a; }
inline float3& operator*=(float3& a, float3 b) { a = a * b; return a; }

// ---- int3 ----
inline int3 operator+(int3 a, int3 b) { return int3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline int3 operator-(int3 a, int3 b) { return int3(a.x - b.x, a.y - b.y, a.z - b.z); }

// ---- scalar builtins ----
inline float min(float x, float y) { return y < x ? y : x; }
inline float max(float x, float y) { return x < y ? y : x; }
inline float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
inline float mad(float a, float b, float c) { return a * b + c; }
inline float step(float e, float x) { return x < e ? 0.0f : 1.0f; }
inline float exp(float x) { return ::expf(x); }
inline float exp2(float x) { return ::exp2f(x); }
inline float pow(float x, float y) { return ::powf(x, y); }
inline float sqrt(float x) { return ::sqrtf(x); }
inline float fabs(float x) { return ::fabsf(x); }

// ---- vector builtins ----
inline float3 min(float3 a, float3 b) { return float3(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)); }
inline float3 max(float3 a, float3 b) { return float3(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }
inline float3 mad(float3 a, float3 b, float3 c) { return a * b + c; }
inline float3 mad(float3 a, float b, float3 c) { return a * b + c; }
inline float3 mix(float3 a, float3 b, float3 t) { return a + (b - a) * t; }
inline float3 mix(float3 a, float3 b, float t) { return a + (b - a) * t; }
inline float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float3 cross(float3 a, float3 b) {
  return float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline float length(float3 a) { return sqrt(dot(a, a)); }
inline float3 normalize(float3 a) {
  const float l = length(a);
  return l == 0.0f ? a : a / l;
}
inline int sat_int(float f) {
  if (f != f) return 0;
  if (f >= 2147483648.0f) return 2147483647;
  if (f <= -2147483648.0f) return (-2147483647 - 1);
  return (int)f;
}
inline int3 convert_int3_sat(float3 a) { return int3(sat_int(a.x), sat_int(a.y), sat_int(a.z)); }
inline float3 convert_float3(int3 a) { return float3((float)a.x, (float)a.y, (float)a.z); }

// ---- execution model ----
extern thread_local int g_global_id;
inline int get_global_id(int) { return g_global_id; }

}  // namespace refcl

#define __kernel
#define __global
#define __private
#define __constant static const
