// rm_scene_plain.cuh -- the render op as one per-pixel-sample routine: the CUDA form of
// RenderImage (renderer.cl:478-494 and its call tree), templated on how the volume is read:
//
//   ByteVolume   the raw uint8 grid, one byte per reference fetch. This is the in-library
//                comparison kernel (RM_OPT_KERNEL = 1): no derived data at all.
//   BrickVolume  the occupancy data of rm_accel.cu: the march predicate comes from 4x4x4
//                bit-bricks, and a macro-cell Chebyshev distance map says how many of the next
//                samples of the fixed-step march cannot be solid, so that those samples cost
//                three adds (the fp32 recurrence p += delta is part of the result) and no fetch.
//
// Both evaluate the surface normal once per sphere-trace, from the hit of the LAST
// distanceToScene call (the reference recomputes it at every outer iteration and overwrites it,
// renderer.cl:224-228, so only the last one is ever visible), and both skip the slab test's six
// IEEE divisions when the ray origin is strictly inside the voxel box (the test then returns
// exactly +0 whatever the quotients are). The production instantiations (kCount == false) also
// leave out work that provably cannot change the result -- shadow rays of lights that contribute
// exactly zero, march samples beyond the distance at which a hit still matters, distanceToScene
// calls outside the ray's march window (march_window), the ground-only tail of rays that can only
// miss, AO probes far from the box -- while the counting instantiations (kCount == true) do the
// reference's full work and report its reference-equivalent counters. Results are identical to the
// reference's order of operations: tests/hostsim compiles this very header for the host and
// compares it with the oracle bit for bit; on the GPU the work counters are compared exactly.
#pragma once
#include "rm_math.cuh"
#include "rm_types.h"

// Statistics hooks of the host simulation (tests/hostsim): empty in the CUDA build.
#ifndef RM_STAT_LOOKUP
#define RM_STAT_LOOKUP() ((void)0)
#define RM_STAT_SKIP(n) ((void)0)
#define RM_STAT_JUMP(n) ((void)0)
#define RM_STAT_SEQ(n) ((void)0)
#define RM_STAT_MARCH() ((void)0)
#define RM_STAT_TRACE() ((void)0)
#define RM_STAT_EVENT(id) ((void)0)
#endif
#ifndef RM_STAT_SITE
#define RM_STAT_SITE(id) ((void)0)      // which call site of the pixel-sample the following work belongs to
#define RM_STAT_LEVEL(l) ((void)0)     // 0 = primary surface, k = k-th bounce
#define RM_STAT_LEVEL_GET() 0
#endif

namespace plain {

// Per-launch constants. They live in __constant__ memory (one copy per translation unit that
// includes this header; each launcher uploads them on its stream right before the launch) rather
// than in kernel parameters, because the shared, NON-inlined distanceToScene below can then read
// them as constant-bank operands; through a reference to a kernel parameter it had to re-load
// them with generic loads on every march iteration (the register budget leaves no room to keep
// them). Launches that come from different streams on the same device are serialised by
// rm_api.cu (DeviceGuard), so a later upload can never overtake a running kernel.
static __constant__ RmOpts g_opts;
static __constant__ RmAccel g_accel;

struct Work {  // reference-equivalent work counters of this thread
  unsigned steps, taps, outer;
};

struct Scene {
  const uint8_t* __restrict__ vox;
  const float4* __restrict__ table;
  float time;  // TRenderOpts.time of the pass this thread renders (per lane in the fused kernel)
  Work w;
  RM_DEV Scene(const uint8_t* v, const float4* t) : vox(v), table(t), time(g_opts.time) {
    w.steps = w.taps = w.outer = 0;
  }
};

struct PixelState {  // TRenderState, renderer.cl:27-33
  float3 eye, mcNormal;
  float px, py;
};

struct Isec {  // TIsec, renderer.cl:6-12
  float3 pos, normal;
  float distance;
  int objectID;
};

RM_DEV float4 table_at(const Scene& s, uint32_t seed) { return __ldg(s.table + (seed & RM_TABLE_MASK)); }
RM_DEV float3 table_xyz(const Scene& s, uint32_t seed) {
  const float4 t = table_at(s, seed);
  return f3(t.x, t.y, t.z);
}

// renderer.cl:153-161
RM_DEV float box_entry(float3 bmin, float3 bmax, float3 p, float3 d) {
  const float3 t0 = (bmin - p) / d;
  const float3 t1 = (bmax - p) / d;
  const float a = cl_max(cl_max(cl_min(t1.x, t0.x), 0.0f), cl_max(cl_min(t1.y, t0.y), cl_min(t1.z, t0.z)));
  const float b = cl_min(cl_max(t1.x, t0.x), cl_min(cl_max(t1.y, t0.y), cl_max(t1.z, t0.z)));
  return b > a ? a : -1.0f;
}

RM_DEV bool in_grid(const RmOpts& o, int x, int y, int z) {
  return (unsigned)x < (unsigned)o.rx && (unsigned)y < (unsigned)o.ry && (unsigned)z < (unsigned)o.rz;
}

// ---- volume access policies ----------------------------------------------------------------

struct ByteVolume {
  const uint8_t* __restrict__ vox;
  RM_DEV int value(const RmOpts& o, int x, int y, int z) const {
    return __ldg(vox + ((size_t)z * o.rxy + (size_t)y * o.rx + x));
  }
  // voxelLookupI (renderer.cl:172-178): v >= isoVal, 0 outside the grid
  RM_DEV int occ(const RmOpts& o, int x, int y, int z) const {
    if (!in_grid(o, x, y, z)) return 0;
    return value(o, x, y, z) >= o.isoVal ? 1 : 0;
  }
};

struct BrickVolume {  // stateless: everything comes from g_accel
  // brick / cell indices fit 32 bits (rm_set_volume bounds the grid), so index math stays 32-bit
  RM_DEV uint64_t word(const uint64_t* __restrict__ bricks, int x, int y, int z) const {
    return __ldg(bricks + (unsigned)(((z >> 2) * g_accel.by + (y >> 2)) * g_accel.bx + (x >> 2)));
  }
  RM_DEV int cell_dist(int x, int y, int z) const {
    const int cs = g_accel.cell_shift;
    return __ldg(g_accel.dist + (unsigned)(((z >> cs) * g_accel.my + (y >> cs)) * g_accel.mx + (x >> cs)));
  }
  static RM_DEV unsigned bit(int x, int y, int z) { return (x & 3) | ((y & 3) << 2) | ((z & 3) << 4); }
  RM_DEV int value(const RmOpts& o, int x, int y, int z) const {
    return __ldg(g_accel.vox + ((size_t)z * o.rxy + (size_t)y * o.rx + x));
  }
  RM_DEV int occ(const RmOpts& o, int x, int y, int z) const {
    if (!in_grid(o, x, y, z)) return 0;
    return (int)((word(g_accel.occ, x, y, z) >> bit(x, y, z)) & 1ull);
  }
};

// voxelNormal (renderer.cl:180-188) as integers: -(occ(+1) - occ(-1)) per axis
template <class Vol>
RM_DEV void gradient6_i(const Vol& V, const RmOpts& o, int x, int y, int z, int& gx, int& gy, int& gz) {
  gx = V.occ(o, x - 1, y, z) - V.occ(o, x + 1, y, z);
  gy = V.occ(o, x, y - 1, z) - V.occ(o, x, y + 1, z);
  gz = V.occ(o, x, y, z - 1) - V.occ(o, x, y, z + 1);
}

// unit3(voxelNormal): the reference negates a float difference, so equal taps give -0.0f
template <class Vol>
RM_DEV float3 normal_6tap(const Vol& V, const RmOpts& o, int x, int y, int z) {
  int gx, gy, gz;
  gradient6_i(V, o, x, y, z, gx, gy, gz);
  return unit3(f3(-(float)(-gx), -(float)(-gy), -(float)(-gz)));
}

// voxelNormalSmooth (renderer.cl:190-203): the reference's float sums are sums of small
// integers, hence exact (and never -0): the integer sum converted once is the same value.
template <class Vol>
RM_DEV float3 normal_smooth(const Vol& V, const RmOpts& o, int x, int y, int z) {
  int sx = 0, sy = 0, sz = 0;
  for (int dz = -1; dz <= 1; ++dz)
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx)
        if (V.occ(o, x + dx, y + dy, z + dz)) {
          int gx, gy, gz;
          gradient6_i(V, o, x + dx, y + dy, z + dz, gx, gy, gz);
          sx += gx; sy += gy; sz += gz;
        }
  return unit3(f3((float)sx, (float)sy, (float)sz));
}

// reference-equivalent occupancy taps of one hit (counting mode only)
template <class Vol>
RM_DEV unsigned taps_of_hit(const Vol& V, const RmOpts& o, int x, int y, int z, bool smooth) {
  if (!smooth) return 6u;
  unsigned n = 0;
  for (int dz = -1; dz <= 1; ++dz)
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx) n += V.occ(o, x + dx, y + dy, z + dz);
  return 27u + 6u * n;
}

// One distanceToScene call (renderer.cl:209-237) without its normal: what the march found.
struct JobResult {
  float dist;     // .x of the returned pair
  float g;        // ground-plane distance of this call (the ground's "id" is (int)g, :211)
  bool hit;       // the march stopped on a solid voxel at sample position p
  bool closer;    // ... and the voxel distance won against the ground plane
  float3 p;
};

// The fixed-step march (renderer.cl:219-234). Returns true on a solid voxel; p is then the
// sample position of the hit.
template <bool kCount>
RM_DEV bool march(Scene& s, const ByteVolume& V, float3& p, float3 delta, int steps, float) {
  const RmOpts& o = g_opts;
  while (--steps >= 0) {
    const int x = f2i_sat(p.x * (float)o.rx), y = f2i_sat(p.y * (float)o.ry), z = f2i_sat(p.z * (float)o.rz);
    if (kCount) s.w.steps++;
    if (!in_grid(o, x, y, z)) return false;  // voxelLookup < 0 -> break
    if (V.value(o, x, y, z) > o.isoVal) return true;
    p = p + delta;
  }
  return false;
}

// Counting form: visits every sample the reference fetches (exact step counter).
RM_DEV bool march_counting(Scene& s, const BrickVolume& V, float3& p, float3 delta, int rem, float invS) {
  const RmOpts& o = g_opts;
  const float rxf = (float)o.rx, ryf = (float)o.ry, rzf = (float)o.rz;
  while (rem > 0) {
    const int x = f2i_sat(p.x * rxf), y = f2i_sat(p.y * ryf), z = f2i_sat(p.z * rzf);
    s.w.steps++;
    if (!in_grid(o, x, y, z)) return false;
    const int d = V.cell_dist(x, y, z);
    if (d != 0) {
      // This sample and the next n-1 lie in cells known to hold no solid voxel: advance the
      // recurrence without fetching. n-1 further steps of at most 1/invS voxels each stay within
      // (d-1) cells; 0.25 voxel of slack covers the rounding drift of the recurrence.
      const float reach = (float)(d - 1) * g_accel.cellf - 0.25f;
      int n = reach > 0.0f ? 1 + f2i_sat(fminf(reach * invS, 1e6f)) : 1;
      n = n < rem ? n : rem;
      rem -= n;
      for (int j = 1; j <= n; ++j) {
        p = p + delta;
        if (j < n) {  // the reference fetches (and counts) every one of these samples
          s.w.steps++;
          if (!in_grid(o, f2i_sat(p.x * rxf), f2i_sat(p.y * ryf), f2i_sat(p.z * rzf))) return false;
        }
      }
    } else {
      if ((V.word(g_accel.solid, x, y, z) >> BrickVolume::bit(x, y, z)) & 1ull) return true;
      p = p + delta;
      rem -= 1;
    }
  }
  return false;
}

// Production form: the same walk without the per-sample bookkeeping of the counting form. Samples
// known to be empty cost three adds each (the fp32 recurrence p += delta is part of the result:
// the position of a hit feeds the next sphere-trace step) and no fetch.
// (Tried and measured slower, 153 vs 104 ms per C2 frame: deferring the adds until a hit needs
// them, locating samples approximately at p + k*delta meanwhile -- the extra per-lookup work and
// the long catch-up loops at low lane counts cost more than the adds saved on missing rays.
// Also without effect, 80.3 vs 79.8 ms: returning early when a long skip lands outside the grid.)
RM_DEV bool march_fast(const RmOpts& o, const BrickVolume& V, float3& p, float3 delta, int rem, float invS) {
  RM_STAT_MARCH();
  // Conversions run on the quarter-rate XU pipe, which this loop saturates (ncu: pipe_xu the busiest
  // pipe): only the three truncations that ARE the reference's convert_int3_sat stay on it. The
  // grid extents come as floats from the constant bank, and the skip length is computed with
  // magic-number adds: (float)(d-1) = as_float(0x4b000000 | (d-1)) - 2^23, and
  // round-to-nearest(x - 0.5) <= floor(x) = as_int((x - 0.5) + 1.5*2^23) - as_int(1.5*2^23) for
  // 0 <= x < 2^22 (a skip may be one sample shorter than the bound allows, never longer).
  const float A = g_accel.cellf * invS, B = 0.25f * invS + 0.5f;  // invS <= 2000 (march_delta)
  while (rem > 0) {
    const int x = f2i_sat(p.x * g_accel.rxf), y = f2i_sat(p.y * g_accel.ryf), z = f2i_sat(p.z * g_accel.rzf);
    if (!in_grid(o, x, y, z)) return false;
    const int d = V.cell_dist(x, y, z);
    RM_STAT_LOOKUP();
    RM_STAT_EVENT(d == 0 ? 13 : (d == 1 ? 14 : (d == 2 ? 15 : 16)));
    int n = 1;  // samples consumed by this iteration: this one plus the ones known to be empty
    if (d > 1) {
      // reach = (d-1)*cell - 0.25 voxels (the slack covers the rounding drift of the recurrence);
      // n = 1 + floor(reach / largest step), conservatively
      const float dm1 = __int_as_float(0x4b000000 | (d - 1)) - 8388608.0f;
      const float k = (dm1 * A - B) + 12582912.0f;
      n = 1 + (__float_as_int(k) - 0x4b400000);
    } else if (d == 0 && ((V.word(g_accel.solid, x, y, z) >> BrickVolume::bit(x, y, z)) & 1ull)) {
      return true;
    }
    if (n >= rem) return false;  // the march runs out inside space known to be empty: a miss
    rem -= n;
    RM_STAT_SKIP(n);
    // n sequential adds, binary-decomposed so that short skips (the common case) take no loop
    for (; n >= 8; n -= 8) {
      p = p + delta; p = p + delta; p = p + delta; p = p + delta;
      p = p + delta; p = p + delta; p = p + delta; p = p + delta;
    }
    if (n & 4) { p = p + delta; p = p + delta; p = p + delta; p = p + delta; }
    if (n & 2) { p = p + delta; p = p + delta; }
    if (n & 1) p = p + delta;
  }
  return false;
}

template <bool kCount>
RM_DEV bool march(Scene& s, const BrickVolume& V, float3& p, float3 delta, int rem, float invS) {
  if (kCount) return march_counting(s, V, p, delta, rem, invS);
  return march_fast(g_opts, V, p, delta, rem, invS);
}

// renderer.cl:209-237 without the normal. Deliberately NOT inlined: the four call sites (primary,
// bounce and shadow sphere-traces, AO probes) then share one copy of the march loop, which cuts
// the kernel's code size ~3x; measured on B200 (C2): 220 -> 178 ms per frame from this alone
// (instruction-fetch stalls were the top stall reason), and it frees registers for occupancy.
#ifndef RM_SD_INLINE
#define RM_SD_INLINE __device__ __noinline__
#endif
template <bool kCount, class Vol>
RM_SD_INLINE JobResult scene_distance(Scene& s, const Vol& V, float3 rpos, float3 dir, float3 delta, int steps, float invS,
                                bool smooth) {
  const RmOpts& o = g_opts;
  JobResult r;
  r.g = rpos.y + o.groundY;
  r.dist = r.g < 1e5f ? r.g : 1e5f;
  r.hit = false;
  r.closer = false;
  r.p = f3s(0.0f);
  const bool inside = rpos.x > o.boundsMin.x && rpos.x < o.boundsMax.x && rpos.y > o.boundsMin.y &&
                      rpos.y < o.boundsMax.y && rpos.z > o.boundsMin.z && rpos.z < o.boundsMax.z;
  // strictly inside: every entry quotient is < 0 and every exit quotient > 0, so the slab test
  // returns max(..., 0) = +0 without evaluating the divisions
  // strictly beyond one slab and moving away from it: both quotients of that axis are negative,
  // so the exit distance is negative and the test returns -1 (also when another axis yields
  // NaN: OpenCL's min/max as written in rm_math.cuh then keep the negative operand)
  const bool away = (rpos.x > o.boundsMax.x && dir.x > 0.0f) || (rpos.x < o.boundsMin.x && dir.x < 0.0f) ||
                    (rpos.y > o.boundsMax.y && dir.y > 0.0f) || (rpos.y < o.boundsMin.y && dir.y < 0.0f) ||
                    (rpos.z > o.boundsMax.z && dir.z > 0.0f) || (rpos.z < o.boundsMin.z && dir.z < 0.0f);
  const float idist = inside ? 0.0f : (away ? -1.0f : box_entry(o.boundsMin, o.boundsMax, rpos, dir));
  RM_STAT_EVENT(0);
  RM_STAT_EVENT(inside ? 1 : (away ? 2 : 3));
  if (idist >= 0.0f && idist < r.dist) {
    RM_STAT_EVENT(4);
    float3 p = rpos + o.voxelBounds;
    if (idist > 0.0f) p = dir * idist + p;
    p = p * o.invVoxelScale;
    if (march<kCount>(s, V, p, delta, steps, invS)) {
      RM_STAT_EVENT(5);
      r.hit = true;
      r.p = p;
      if (kCount) {
        const int x = f2i_sat(p.x * (float)o.rx), y = f2i_sat(p.y * (float)o.ry), z = f2i_sat(p.z * (float)o.rz);
        s.w.taps += taps_of_hit(V, o, x, y, z, smooth);
      }
      const float3 hp = p * o.voxelBounds2 + (-o.voxelBounds);
      const float d = len3(rpos - hp) - o.voxelSize;
      if (d < r.dist) { r.dist = d; r.closer = true; }
    }
  }
  return r;
}

// per-direction constants of the march: delta (renderer.cl:215) and 1 / (largest step in voxels)
#ifndef RM_MDELTA_ATTR
#define RM_MDELTA_ATTR RM_DEV  // (sharing this one too measured slower: 40.0 vs 39.1 ms)
#endif
RM_MDELTA_ATTR float3 march_delta(const RmOpts& o, float3 dir, int steps, float& invS) {
  const float3 delta = (dir / ((float)steps * 0.5f)) * o.invVoxelScale;
  const float sm = fmaxf(fmaxf(fabsf(delta.x) * (float)o.rx, fabsf(delta.y) * (float)o.ry), fabsf(delta.z) * (float)o.rz);
  invS = sm > 5e-4f ? __fdividef(1.0f, sm) : 2000.0f;  // capped: (d-1)*cell*invS <= 31*64*2000 < 2^22 in march_fast
  return delta;
}

// A conservative window [tin, tout] of ray parameters outside of which a distanceToScene call at
// ro + rd*d cannot reach the voxel march (production kernels only). The line is clipped against the
// voxel box GROWN by kGrow on every side, in ordinary (not pinned) arithmetic:
//   d > tout            the evaluation point lies beyond a slab of the true box by >= kGrow and moves
//                       away from it: this is exactly the `away` case of scene_distance (-1, no march);
//   tin - d > g + slack the fp32 entry distance the reference computes from that point differs from
//                       the true one by ~1e-5 at these magnitudes (|coordinates| <= 128, see `ok`), the
//                       true one is >= tin - d, so it is >= g: `idist < res.x` (renderer.cl:213) fails;
//   tin = +inf          the line misses the grown box altogether: the slab test returns -1 everywhere.
// Whenever the window cannot be trusted (huge coordinates, NaNs) it is (-inf, +inf): every call is
// then evaluated in full. The window only selects WHICH calls are evaluated in full; a call that is
// skipped returns the ground-plane pair, which is what the full evaluation returns without a march.
RM_DEV void march_window(const RmOpts& o, float3 ro, float3 rd, float maxDist, float& tin, float& tout) {
  const float kGrow = 0.01f, kTiny = 1e-5f, kBig = 64.0f, kInf = 3.0e38f;
  tin = -kInf;
  tout = kInf;
  const float ro_[3] = {ro.x, ro.y, ro.z}, rd_[3] = {rd.x, rd.y, rd.z};
  const float lo_[3] = {o.boundsMin.x, o.boundsMin.y, o.boundsMin.z}, hi_[3] = {o.boundsMax.x, o.boundsMax.y, o.boundsMax.z};
  bool ok = maxDist <= kBig, miss = false;
#pragma unroll
  for (int i = 0; i < 3; ++i)
    ok = ok && fabsf(ro_[i]) <= kBig && fabsf(lo_[i]) <= kBig && fabsf(hi_[i]) <= kBig && fabsf(rd_[i]) <= 2.0f;
  if (!ok) return;
  float a = -kInf, b = kInf;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float l = lo_[i] - kGrow - ro_[i], h = hi_[i] + kGrow - ro_[i];
    if (fabsf(rd_[i]) < kTiny) {
      // the coordinate moves by < kTiny * (kBig + a ground step) over the whole trace: outside the
      // grown slab now means outside the true slab for good, inside means no constraint
      if (l > 0.0f || h < 0.0f) miss = true;
    } else {
      const float inv = __fdividef(1.0f, rd_[i]);
      const float t0 = l * inv, t1 = h * inv;
      a = fmaxf(a, fminf(t0, t1));
      b = fminf(b, fmaxf(t0, t1));
    }
  }
  // (a box wholly behind the origin, b < 0, needs no special case: every d > tout is "beyond" -- and a
  // trace with a negative startDist may well begin inside it)
  if (miss || b < a) { tin = kInf; tout = -kInf; return; }
  tin = a;
  tout = b;
}

// renderer.cl:239-257
#ifndef RM_ST_INLINE
#define RM_ST_INLINE __device__ __noinline__  // one copy for primary / bounce / shadow traces (code size)
#endif
template <bool kCount, class Vol>
RM_ST_INLINE void sphere_trace(Scene& s, const Vol& V, float3 ro, float3 rd, Isec& r, float maxDist, int maxSteps,
                         bool smooth, bool wantSurface) {
  const RmOpts& o = g_opts;
  RM_STAT_TRACE();
  RM_STAT_EVENT(wantSurface ? 6 : 7);
  float invS;
  const float3 delta = march_delta(o, rd, o.maxVoxelIter, invS);
  JobResult j;
  j.g = 0.0f; j.dist = 0.0f; j.hit = false; j.closer = false; j.p = f3s(0.0f);
  // the trace state lives in registers; `r` (the caller's memory) is written once at the end
  float dist = o.startDist;
  float3 pos = ro;
  // distanceToScene returns min(voxel distance, ground distance g): a voxel hit whose distance
  // len - voxelSize is >= g returns the same pair as no hit, so only the samples within
  // g + voxelSize of the ray position can change the returned distance; for a shadow ray
  // (distance-only consumer, wantSurface == false) the same holds for hits beyond the light
  // (they and "no hit" both end the trace as lit). The production kernels march only those samples
  // (plus two for rounding). The one thing a far hit still decides is the reference's quirk that
  // the LAST call's voxel normal wins even when the ground is closer (renderer.cl:224-228): when a
  // surface is wanted and the last call was cut short without a hit, that call is repeated in full.
  const float inv_step = kCount ? 0.0f : 1.01f / len3(delta * o.voxelBounds2);
  bool cut = false;
  // Most distanceToScene calls of a trace cannot reach the voxel march at all (C2: 65 of 81 per
  // pixel-sample -- the ray has left the box, has not come within the ground distance of it, or
  // never meets it): they return the ground-plane pair, and the production kernels evaluate just
  // that (march_window says which calls those are). The counting kernels evaluate every call in full.
  float tin = -3.0e38f, tout = 3.0e38f;
  if (!kCount) march_window(o, ro, rd, maxDist, tin, tout);
  while (--maxSteps >= 0) {
    if (kCount) s.w.outer++;
    RM_STAT_EVENT(wantSurface ? 8 : 9);
    pos = ro + rd * dist;
    const float g = pos.y + o.groundY;
    if (!kCount && (dist > tout || tin - dist > g * 1.0001f + 1e-3f || g <= 0.0f)) {
      RM_STAT_EVENT(12);
      j.g = g;
      j.dist = g < 1e5f ? g : 1e5f;
      j.hit = false;
      j.closer = false;
      cut = false;
      // A ray that is past the box (or never meets it) and does not descend only sees the ground
      // from here on, farther away at every step: every remaining evaluation advances it by at
      // least g (fp32 rounding is monotone), none can converge, so if the remaining evaluations
      // must carry it beyond maxDist the trace ends as a miss ("lit" for a shadow ray) -- and of a
      // miss its consumers read nothing but that fact (renderer.cl:252-255, 292-301, 389-392,
      // 413-416: distance 1000, objectID -1). 1% covers the rounding of <= 2^16 additions.
      if (dist > tout && rd.y >= 0.0f && g > o.eps && g < 1e5f && maxSteps < 65536 &&
          (float)(maxSteps + 1) * g * 0.99f >= maxDist - dist) {
        RM_STAT_EVENT(17);
        dist = maxDist;
        break;
      }
    } else {
      int limit = o.maxVoxelIter;
      if (!kCount) {
        float reach = g;
        // (a hit just beyond the light must also be farther than eps, or it could end the trace as
        // "converged", i.e. shadowed, before the light is reached)
        if (!wantSurface) reach = fminf(reach, fmaxf(maxDist - dist, o.eps));
        const float k = (reach + o.voxelSize) * inv_step;
        cut = k < (float)(limit - 2);
        if (cut) limit = f2i_sat(k) + 2;
      }
      j = scene_distance<kCount>(s, V, pos, rd, delta, limit, invS, smooth);
    }
    if (fabsf(j.dist) <= o.eps || dist >= maxDist) break;
    dist += j.dist;
  }
  if (!kCount && wantSurface && cut && !j.hit) RM_STAT_EVENT(10);
  if (!kCount && wantSurface && cut && !j.hit)
    j = scene_distance<kCount>(s, V, pos, rd, delta, o.maxVoxelIter, invS, smooth);
  const bool miss = dist >= maxDist;
  if (miss) {
    pos = ro + rd * dist;
    dist = 1000.0f;
  }
  r.distance = dist;
  r.pos = pos;
  r.objectID = -1;
  r.normal = f3s(0.0f);
  if (!wantSurface) return;  // shadow rays use the distance only
  // object id and normal of the LAST distanceToScene call
  const int x = f2i_sat(j.p.x * (float)o.rx), y = f2i_sat(j.p.y * (float)o.ry), z = f2i_sat(j.p.z * (float)o.rz);
  if (!miss) {
    if (j.closer) {
      const int v = V.value(o, x, y, z);
      r.objectID = v < 168 ? (v < 84 ? 1 : 2) : 3;  // voxelMaterial, renderer.cl:205-207
    } else {
      r.objectID = f2i_sat(j.g < 1e5f ? j.g : -1.0f);
    }
  }
  if (j.hit) r.normal = smooth ? normal_smooth(V, o, x, y, z) : normal_6tap(V, o, x, y, z);
  else r.normal = j.g < 1e5f ? f3(0.0f, 1.0f, 0.0f) : -rd;
}

// ---- the same trace, one distanceToScene evaluation at a time -------------------------------------
// For the persistent trace kernel of the wavefront path (rm_render_wave.cu): the lanes of a warp
// run DIFFERENT rays there and must meet once per evaluation, so the loop of sphere_trace is lifted
// into a state + a step function. trace_begin / trace_step / trace_finish / trace_surface are
// sphere_trace cut at its loop; tests/hostsim runs them (via rm_wave.cuh) against the oracle.
struct TraceSetup {  // per-ray constants, computed where the ray is created (coherently)
  float3 delta;
  float invS, inv_step, tin, tout;
};

template <bool kCount>
RM_DEV TraceSetup trace_setup(float3 ro, float3 rd, float maxDist) {
  const RmOpts& o = g_opts;
  TraceSetup t;
  t.delta = march_delta(o, rd, o.maxVoxelIter, t.invS);
  t.inv_step = kCount ? 0.0f : 1.01f / len3(t.delta * o.voxelBounds2);
  t.tin = -3.0e38f;
  t.tout = 3.0e38f;
  if (!kCount) march_window(o, ro, rd, maxDist, t.tin, t.tout);
  return t;
}

struct TraceState {
  float3 ro, rd, pos;
  TraceSetup c;
  float maxDist, dist;
  int maxSteps;
  bool wantSurface, cut;
  JobResult j;
};

RM_DEV void trace_begin(TraceState& t, float3 ro, float3 rd, const TraceSetup& c, float maxDist, int maxSteps, bool wantSurface) {
  t.ro = ro; t.rd = rd; t.pos = ro; t.c = c;
  t.maxDist = maxDist; t.dist = g_opts.startDist; t.maxSteps = maxSteps;
  t.wantSurface = wantSurface; t.cut = false;
  t.j.g = 0.0f; t.j.dist = 0.0f; t.j.hit = false; t.j.closer = false; t.j.p = f3s(0.0f);
}

// The loop of sphere_trace in two halves, so that the lanes of the persistent trace kernel meet at
// the expensive one: trace_run_cheap runs this ray's ground-only evaluations (a few instructions
// each) until the ray ends (returns true) or its next evaluation has to be a full distanceToScene
// call (returns false, t.pos set); trace_full makes that call and returns true when the ray ends.
// Together they are exactly one or more iterations of the loop of sphere_trace.
enum { kTraceNeedsFull = 0, kTraceDone = 1, kTraceMoreCheap = 2 };
template <bool kCount>
RM_DEV int trace_run_cheap(Scene& s, TraceState& t, int budget) {
  const RmOpts& o = g_opts;
  for (;; --budget) {
    if (budget <= 0) return kTraceMoreCheap;  // the lanes of a warp wait for the longest run: keep runs short
    if (--t.maxSteps < 0) return kTraceDone;
    if (kCount) s.w.outer++;
    RM_STAT_EVENT(t.wantSurface ? 8 : 9);
    t.pos = t.ro + t.rd * t.dist;
    const float g = t.pos.y + o.groundY;
    if (kCount || !(t.dist > t.c.tout || t.c.tin - t.dist > g * 1.0001f + 1e-3f || g <= 0.0f)) return kTraceNeedsFull;
    RM_STAT_EVENT(12);
    t.j.g = g;
    t.j.dist = g < 1e5f ? g : 1e5f;
    t.j.hit = false;
    t.j.closer = false;
    t.cut = false;
    if (t.dist > t.c.tout && t.rd.y >= 0.0f && g > o.eps && g < 1e5f && t.maxSteps < 65536 &&
        (float)(t.maxSteps + 1) * g * 0.99f >= t.maxDist - t.dist) {
      RM_STAT_EVENT(17);
      t.dist = t.maxDist;
      return kTraceDone;
    }
    if (fabsf(t.j.dist) <= o.eps || t.dist >= t.maxDist) return kTraceDone;
    t.dist += t.j.dist;
  }
}

template <bool kCount, class Vol>
RM_DEV bool trace_full(Scene& s, const Vol& V, TraceState& t) {
  const RmOpts& o = g_opts;
  int limit = o.maxVoxelIter;
  if (!kCount) {
    float reach = t.pos.y + o.groundY;
    if (!t.wantSurface) reach = fminf(reach, fmaxf(t.maxDist - t.dist, o.eps));
    const float k = (reach + o.voxelSize) * t.c.inv_step;
    t.cut = k < (float)(limit - 2);
    if (t.cut) limit = f2i_sat(k) + 2;
  }
  t.j = scene_distance<kCount>(s, V, t.pos, t.rd, t.c.delta, limit, t.c.invS, false);
  if (fabsf(t.j.dist) <= o.eps || t.dist >= t.maxDist) return true;
  t.dist += t.j.dist;
  return false;
}

// After the loop: the repeated full call of the voxel-normal quirk, and the miss bookkeeping.
// Returns true on a miss; t.dist / t.pos are then TIsec.distance (1000) / TIsec.pos.
template <bool kCount, class Vol>
RM_DEV bool trace_finish(Scene& s, const Vol& V, TraceState& t) {
  const RmOpts& o = g_opts;
  if (!kCount && t.wantSurface && t.cut && !t.j.hit) {
    RM_STAT_EVENT(10);
    t.j = scene_distance<kCount>(s, V, t.pos, t.rd, t.c.delta, o.maxVoxelIter, t.c.invS, false);
  }
  const bool miss = t.dist >= t.maxDist;
  if (miss) {
    t.pos = t.ro + t.rd * t.dist;
    t.dist = 1000.0f;
  }
  return miss;
}

// objectID and normal of a finished (non-smooth) surface trace, from what its last evaluation found
template <class Vol>
RM_DEV void trace_surface(const Vol& V, float3 rd, bool miss, bool hit, bool closer, float3 jp, float jg, int& objectID, float3& normal) {
  const RmOpts& o = g_opts;
  objectID = -1;
  const int x = f2i_sat(jp.x * (float)o.rx), y = f2i_sat(jp.y * (float)o.ry), z = f2i_sat(jp.z * (float)o.rz);
  if (!miss) {
    if (closer) {
      const int v = V.value(o, x, y, z);
      objectID = v < 168 ? (v < 84 ? 1 : 2) : 3;
    } else {
      objectID = f2i_sat(jg < 1e5f ? jg : -1.0f);
    }
  }
  if (hit) normal = normal_6tap(V, o, x, y, z);
  else normal = jg < 1e5f ? f3(0.0f, 1.0f, 0.0f) : -rd;
}

RM_DEV float3 sky(const RmOpts& o, float3 d) { return lerp3(o.sky1, o.sky2, d.y * 0.5f + 0.5f); }  // :259-261

// renderer.cl:263-269
#ifndef RM_LPOS_ATTR
#define RM_LPOS_ATTR RM_SHARED_FN  // see unit3 in rm_math.cuh
#endif
RM_LPOS_ATTR float3 light_pos(const Scene& s, const PixelState& st, int i) {
  const uint32_t seed = f2u_wrap(st.px * 1957.0f + st.py * 2173.0f + s.time * 4763.742f);
  return table_xyz(s, seed) * g_opts.lightScatter + g_opts.lightPos[i];
}

RM_DEV float3 reflect3(float3 v, float3 n) { return v - n * (2.0f * dot3(v, n)); }  // :271-273

// renderer.cl:275-290
#ifndef RM_ATMO_ATTR
#define RM_ATMO_ATTR RM_SHARED_FN  // two call sites (primary, bounce); see unit3 in rm_math.cuh
#endif
RM_ATMO_ATTR float3 atmosphere(const Scene& s, const PixelState& st, float3 ro, float3 rd, float distance, float3 col) {
  const RmOpts& o = g_opts;
  const float fa = 1.0f - expf(distance * distance * -o.fogPow);
  col = (sky(o, rd) - col) * fa + col;
  for (int i = 0; i < o.numLights; ++i) {
    float3 lp = light_pos(s, st, i);
    const float d = cl_clamp(dot3(lp - ro, rd), 0.0f, distance);
    lp = rd * d + (ro - lp);
    col = o.lightColor[i] * (o.flareAmp / dot3(lp, lp)) + col;
  }
  return col;
}

// renderer.cl:304-311
RM_DEV float schlick(float r0, float smooth, float3 n, float3 view) {
  const float d = cl_clamp(1.0f - dot3(n, -view), 0.0f, 1.0f);
  if (d > 0.0f) {
    const float d2 = d * d;
    return (1.0f - r0) * (smooth * d2 * d2 * d) + r0;
  }
  return 0.0f;
}

// renderer.cl:317-325
RM_DEV float blinn_phong(float smooth, float3 rd, float3 ldir, float3 n) {
  const float nh = dot3(unit3(ldir - rd), n);
  if (nh > 0.0f) {
    const float sp = exp2f(6.0f * smooth + 4.0f);
    return powf(nh, sp) * (sp + 2.0f) * 0.125f;
  }
  return 0.0f;
}

// renderer.cl:327-346
template <bool kCount, class Vol>
RM_DEV float ambient_occlusion(Scene& s, const Vol& V, float3 pos, float3 n0) {
  const RmOpts& o = g_opts;
  float ao = 1.0f, d = 0.0f;
  uint32_t seed = f2u_wrap(pos.x * 3183.75f + pos.y * 1831.42f + pos.z * 2945.87f + s.time * 2671.918f);
  for (int i = 0; i <= o.aoIter && ao > 0.01f; ++i) {
    d += o.aoStepDist;
    seed += 37u;
    const float3 n = unit3(table_xyz(s, seed) * 0.2f + n0);
    const float3 q = n * d + pos;
    float hdist;
    // Probe origins that lie farther from the voxel box than d + voxelSize along some axis (ground
    // far from the object): whatever the march might hit is at least that far away, i.e. beyond
    // the distance d up to which a hit matters, and the call returns the ground pair or an
    // equivalent one. 1e-3 dwarfs the fp32 error of the slab test at these magnitudes.
    const float out = fmaxf(fmaxf(fmaxf(o.boundsMin.x - q.x, q.x - o.boundsMax.x), fmaxf(o.boundsMin.y - q.y, q.y - o.boundsMax.y)),
                            fmaxf(o.boundsMin.z - q.z, q.z - o.boundsMax.z));
    if (!kCount && o.aoAmp >= 0.0f && d > 0.0f && out > d + o.voxelSize + 1e-3f && out < 1e3f) {
      RM_STAT_EVENT(18);
      const float g = q.y + o.groundY;
      hdist = g < 1e5f ? g : 1e5f;
    } else {
      float invS;
      const int msteps = o.maxVoxelIter / 2;
      const float3 delta = march_delta(o, n, msteps, invS);
      // A probe changes ao only through max((d - h)*aoAmp/d, 0), which is 0 (factor exactly 1) for
      // every h >= d when aoAmp >= 0: a voxel hit whose distance len - voxelSize is >= d, and the
      // ground plane beyond d, are the same as no hit at all. Sample k of the march lies k world
      // steps from the probe origin, so only the first (d + voxelSize)/step samples can matter;
      // the production kernels march those (plus two for rounding), the counting kernels all of
      // them like the reference (renderer.cl:342 marches maxVoxelIter/2 = 96 samples per probe).
      int limit = msteps;
      if (!kCount && o.aoAmp >= 0.0f && d > 0.0f) {
        const float k = (d + o.voxelSize) * 1.01f / len3(delta * o.voxelBounds2);
        if (k < (float)msteps) limit = f2i_sat(k) + 2 < msteps ? f2i_sat(k) + 2 : msteps;
      }
      RM_STAT_EVENT(11);
      RM_STAT_SITE(RM_STAT_LEVEL_GET() * 16 + 8 + i);
      hdist = scene_distance<kCount>(s, V, q, n, delta, limit, invS, false).dist;
    }
    ao *= 1.0f - cl_max((d - hdist) * o.aoAmp / d, 0.0f);
  }
  return ao;
}

// renderer.cl:348-381 (shadow :292-301 inlined)
#ifndef RM_OL_INLINE
#define RM_OL_INLINE __device__ __noinline__  // one copy for the primary and the bounce surfaces
#endif
template <bool kCount, class Vol>
RM_OL_INLINE float3 object_lighting(Scene& s, const Vol& V, const PixelState& st, float3 rd, float3 ipos, const RmMaterial& m,
                              float3 n, float3 reflectCol) {
  const RmOpts& o = g_opts;
  const float ao = ambient_occlusion<kCount>(s, V, ipos, n);
  float3 diff = sky(o, n) * ao;
  float3 spec = reflectCol * ao;
  float3 fin = f3s(0.0f);
  for (int i = 0; i < o.numLights; ++i) {
    const float3 dl = light_pos(s, st, i) - ipos;
    const float ld2 = dot3(dl, dl);
    const float att = 1.0f / ld2;
    if (att > o.minLightAtt) {
      const float3 ldir = unit3(dl);
      const float lmax = cl_min(sqrtf(ld2) - o.shadowBias, o.maxDist);
      const float kd = cl_max(0.0f, dot3(ldir, n));
      const float ks = blinn_phong(m.smoothness, rd, ldir, n);
      const float3 inc = (o.lightColor[i] * 1.0f) * att;  // (lightColor * shadow factor 1) * att, :369
      // A light that faces neither the surface (kd = 0) nor its highlight (ks = 0) adds inc*0 to
      // both sums whether or not it is shadowed, i.e. nothing as long as inc is finite: its shadow
      // ray cannot change the result and is not traced. (The counting kernels trace it: the
      // reference does, and its work is part of the reference-equivalent counters.)
      const float3 zero = inc * 0.0f;
      const bool irrelevant = !kCount && kd == 0.0f && ks == 0.0f && zero.x == 0.0f && zero.y == 0.0f && zero.z == 0.0f;
      if (!irrelevant) {
        Isec sh;
        RM_STAT_SITE(RM_STAT_LEVEL_GET() * 16 + 1 + i);
        sphere_trace<kCount>(s, V, ipos + ldir * o.shadowBias, ldir, sh, lmax, o.shadowIter, false, false);
        const float sf = sh.distance < lmax ? 0.0f : 1.0f;
        if (sf > 0.0f) {
          diff = diff + inc * kd;
          spec = spec + inc * ks;
        }
      }
    }
    diff = diff * m.albedo;
    fin = fin + lerp3(diff, spec, schlick(m.r0, m.smoothness, n, rd));
  }
  return fin / (float)o.numLights;
}

RM_DEV int mat_index(int id) { return id < 0 ? 0 : (id > 3 ? 3 : id); }

// renderer.cl:383-405
template <bool kCount, class Vol>
RM_DEV float3 bounce_color(Scene& s, const Vol& V, const PixelState& st, float3 ro, float3 rd, Isec& isec) {
  const RmOpts& o = g_opts;
  RM_STAT_SITE(RM_STAT_LEVEL_GET() * 16);
  sphere_trace<kCount>(s, V, ro, rd, isec, o.maxDist, o.maxIter, false, true);
  float3 col;
  if (isec.objectID < 0) {
    col = sky(o, rd);
  } else {
    col = object_lighting<kCount>(s, V, st, rd, isec.pos, o.mat[mat_index(isec.objectID)], isec.normal,
                                  sky(o, reflect3(rd, isec.normal)));
  }
  return atmosphere(s, st, ro, rd, isec.distance, col);
}

// renderer.cl:407-446
template <bool kCount, class Vol>
RM_DEV float3 scene_color(Scene& s, const Vol& V, const PixelState& st, float3 ro, float3 rd) {
  const RmOpts& o = g_opts;
  Isec isec;
  RM_STAT_LEVEL(0);
  RM_STAT_SITE(0);
  sphere_trace<kCount>(s, V, ro, rd, isec, o.maxDist, o.maxIter, true, true);
  float3 col;
  if (isec.distance >= o.maxDist) {
    col = sky(o, rd);
  } else {
    const RmMaterial& m = o.mat[mat_index(isec.objectID)];
    const float3 n = st.mcNormal * (1.0f / (m.smoothness * 200.0f + 5.0f)) + isec.normal;
    float3 reflectCol = f3s(0.0f);
    if (m.r0 > 0.0f && o.reflectIter > 0) {
      Isec ri;
      ri.pos = isec.pos;
      ri.normal = n;
      float3 bd = rd;
      for (int i = 0; i < o.reflectIter; ++i) {
        bd = reflect3(bd, ri.normal);
        const float3 bo = ri.pos + bd * 0.0075f;
        RM_STAT_LEVEL(i + 1);
        reflectCol = reflectCol + bounce_color<kCount>(s, V, st, bo, bd, ri);
        if (ri.objectID < 0) break;
        if (o.mat[mat_index(ri.objectID)].r0 < 0.001f) break;
      }
    } else {
      reflectCol = sky(o, reflect3(rd, n));
    }
    RM_STAT_LEVEL(0);
    col = object_lighting<kCount>(s, V, st, rd, isec.pos, m, n, reflectCol);
  }
  return atmosphere(s, st, ro, rd, isec.distance, col);
}

// renderer.cl:467-476 + :456-465
RM_DEV float3 setup_pixel(const Scene& s, int id, PixelState& st) {
  const RmOpts& o = g_opts;
  const float4 a = table_at(s, (uint32_t)(id * 17) + f2u_wrap(s.time * 3141.3862f));
  st.mcNormal = unit3(table_xyz(s, (uint32_t)(id * 37) + f2u_wrap(s.time * 1859.1467f)));
  st.px = (float)(id % o.width) + a.z;
  st.py = (float)(id / o.width) + a.w;
  st.eye = f3(st.mcNormal.z, st.mcNormal.x, st.mcNormal.y) * o.dof + o.eyePos;
  const float3 fwd = unit3(o.targetPos - st.eye);
  const float3 right = unit3(cross3(fwd, o.up));
  const float vx = st.px / (float)o.width * o.fov - o.fov * 0.5f;
  float vy = st.py / (float)o.height * o.fov - o.fov * 0.5f;
  vy = vy * -o.invAspect;
  return unit3(right * vx + cross3(right, fwd) * vy + fwd);
}

// One work-item of RenderImage (renderer.cl:478-494): returns sceneColor * exposure.
template <bool kCount, class Vol>
RM_DEV float3 render_pixel_sample(Scene& s, const Vol& V, int id) {
  PixelState st;
  const float3 rd = setup_pixel(s, id, st);
  return scene_color<kCount>(s, V, st, st.eye, rd) * g_opts.exposure;
}

}  // namespace plain
