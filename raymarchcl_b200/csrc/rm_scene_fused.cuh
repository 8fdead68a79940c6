// rm_scene_fused.cuh -- the render op (RenderImage, renderer.cl:478-494 and its call tree) as the
// per-pixel-sample routine of the DEFAULT kernel (rm_render_persist.cu, RM_OPT_KERNEL = 0).
//
// Same arithmetic, same order of operations and the same exact shortcuts as rm_scene_plain.cuh
// (fetch elision over bit-bricks + a macro-cell Chebyshev distance map, lazy normals, slab-test
// shortcuts, irrelevance culling, march windows, early misses; the proofs are written out there and
// are not repeated here) -- what differs is how the routine sits on the machine:
//
//   * every shared (non-inlined) function takes and returns VALUES, which nvcc passes in registers:
//     nothing is forced into local memory by having its address taken (the by-reference Isec / Scene /
//     JobResult / PixelState of rm_scene_plain.cuh were). What remains in local memory is register
//     spill of the 48-register build;
//   * the march loop is rewritten for few paths and few instructions (march_fast below);
//   * the distance map can be read from SHARED MEMORY (kMapNib): 4 bits per macro-cell (Chebyshev
//     distance saturated at 15), staged once per resident block with a bulk TMA copy (cp.async.bulk +
//     mbarrier, rm_render_persist.cu): the default of long launches, in the 1024-thread layout;
//     the byte map in global memory / L1 otherwise;
//   * float3 adds and the march recurrence use Blackwell's packed FADD2 on the (x, y) lanes (rm_math.cuh).
//
// tests/hostsim compiles this header for the host (RM_NIB_BASE is then a plain pointer) and compares
// the routine with the oracle bit for bit.
#pragma once
#include "rm_math.cuh"
#include "rm_types.h"

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
#define RM_STAT_SITE(id) ((void)0)
#define RM_STAT_LEVEL(l) ((void)0)
#define RM_STAT_LEVEL_GET() 0
#endif

namespace fused {

// Per-launch constants in __constant__ memory (one copy per translation unit; see the note in
// rm_scene_plain.cuh: the shared functions read them as constant-bank operands).
static __constant__ RmOpts g_opts;
static __constant__ RmAccel g_accel;

#if defined(__CUDACC__)
extern __shared__ __align__(16) uint8_t rm_smem_nib[];  // the 4-bit distance map of this block
#define RM_NIB_BASE rm_smem_nib
// The map's address in the shared window, computed once per march and kept in a register (the empty asm
// makes it opaque: otherwise the compiler re-derives it from SR_CgaCtaId inside the loop, 4 instructions
// per lookup), and a plain ld.shared through it.
RM_DEV unsigned nib_base() {
  unsigned a = (unsigned)__cvta_generic_to_shared(rm_smem_nib);
  asm("" : "+r"(a));
  return a;
}
RM_DEV unsigned nib_load(unsigned base, unsigned byte) {
  unsigned v;
  asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(base + byte));
  return v;
}
#else
static const uint8_t* rm_host_nib = nullptr;  // host simulation: the same packed map in host memory
#define RM_NIB_BASE rm_host_nib
inline unsigned nib_base() { return 0u; }
inline unsigned nib_load(unsigned, unsigned byte) { return rm_host_nib[byte]; }
#endif

// Reference-equivalent work counters: a real object (by reference) in the counting kernels, an
// empty value in the production kernels, so that those carry no address-taken local at all.
template <bool kCount> struct Cnt {};
template <> struct Cnt<true> { unsigned steps = 0, taps = 0, outer = 0; };
template <bool kCount> struct CntArg { using type = Cnt<false>; };
template <> struct CntArg<true> { using type = Cnt<true>&; };
#define RM_CNT typename CntArg<kCount>::type

struct Lane {  // what differs between the passes of one fused launch: TRenderOpts.time and the table
  const float4* table;
  float time;
};

struct Isec {  // TIsec, renderer.cl:6-12
  float3 pos, normal;
  float distance;
  int objectID;
};

struct Hit {  // one distanceToScene call (renderer.cl:209-237) without its normal
  float3 p;      // sample position of the solid voxel the march stopped on (when kHit)
  float dist;    // .x of the returned pair
  int flags;     // kHit: the march stopped on a solid voxel; kCloser: ... and the voxel distance won against the
                 // ground plane. (One word: two bools in a struct cost a dozen byte-permutes per sphere-trace trip.)
};
enum { kHit = 1, kCloser = 2 };

RM_DEV float4 table_at(const Lane& s, uint32_t seed) { return __ldg(s.table + (seed & RM_TABLE_MASK)); }
RM_DEV float3 table_xyz(const Lane& s, uint32_t seed) {
  const float4 t = table_at(s, seed);
  return f3(t.x, t.y, t.z);
}

// renderer.cl:153-161. Shared and out of line: six IEEE divisions (~100 instructions) that the inside / away
// shortcuts of scene_distance make rare -- kept out of that function's hot body.
RM_SHARED_FN float box_entry(float3 bmin, float3 bmax, float3 p, float3 d) {
  const float3 t0 = (bmin - p) / d;
  const float3 t1 = (bmax - p) / d;
  const float a = cl_max(cl_max(cl_min(t1.x, t0.x), 0.0f), cl_max(cl_min(t1.y, t0.y), cl_min(t1.z, t0.z)));
  const float b = cl_min(cl_max(t1.x, t0.x), cl_min(cl_max(t1.y, t0.y), cl_max(t1.z, t0.z)));
  return b > a ? a : -1.0f;
}

RM_DEV bool in_grid(int x, int y, int z) {
  return (unsigned)x < (unsigned)g_opts.rx && (unsigned)y < (unsigned)g_opts.ry && (unsigned)z < (unsigned)g_opts.rz;
}

// ---- occupancy data (rm_accel.cu) ----------------------------------------------------------------
RM_DEV uint64_t brick_word(const uint64_t* __restrict__ bricks, int x, int y, int z) {
  return __ldg(bricks + (unsigned)(((z >> 2) * g_accel.by + (y >> 2)) * g_accel.bx + (x >> 2)));
}
RM_DEV unsigned brick_bit(int x, int y, int z) { return (x & 3) | ((y & 3) << 2) | ((z & 3) << 4); }
RM_DEV bool solid_at(int x, int y, int z) { return (brick_word(g_accel.solid, x, y, z) >> brick_bit(x, y, z)) & 1ull; }
// voxelLookupI (renderer.cl:172-178): v >= isoVal, 0 outside the grid
RM_DEV int occ_at(int x, int y, int z) {
  if (!in_grid(x, y, z)) return 0;
  return (int)((brick_word(g_accel.occ, x, y, z) >> brick_bit(x, y, z)) & 1ull);
}
RM_DEV int voxel_value(int x, int y, int z) {
  return __ldg(g_accel.vox + ((size_t)z * g_opts.rxy + (size_t)y * g_opts.rx + x));
}
// How the distance map is read (template parameter kMap of everything below, a bit set):
//   kMapNib    the 4-bit copy in shared memory (else the byte map in global memory)
//   kMapCell4  the macro-cell is one 4x4x4 brick (cell_shift == 2: every volume up to 256^3), so the cell
//              index is the brick index and the shift is an immediate
//   kMapPow2   every grid extent is a power of two: the march runs its recurrence in voxel units (exact, see
//              march_fast) and saves the three multiplies per sample
enum { kMapNib = 1, kMapCell4 = 2, kMapPow2 = 4 };
// Chebyshev distance (in macro-cells, saturated) from the cell of voxel (x, y, z) to the nearest cell
// that holds a solid voxel.
template <int kMap>
RM_DEV int cell_dist(int x, int y, int z) {
  const int cs = g_accel.cell_shift;
  const unsigned c = (unsigned)(((z >> cs) * g_accel.my + (y >> cs)) * g_accel.mx + (x >> cs));
  if (kMap & kMapNib) return (RM_NIB_BASE[c >> 1] >> ((c & 1u) << 2)) & 15;
  return __ldg(g_accel.dist + c);
}

// voxelNormal (renderer.cl:180-188) as integers
RM_DEV void gradient6_i(int x, int y, int z, int& gx, int& gy, int& gz) {
  gx = occ_at(x - 1, y, z) - occ_at(x + 1, y, z);
  gy = occ_at(x, y - 1, z) - occ_at(x, y + 1, z);
  gz = occ_at(x, y, z - 1) - occ_at(x, y, z + 1);
}
// unit3(voxelNormal): the reference negates a float difference, so equal taps give -0.0f
RM_DEV float3 normal_6tap(int x, int y, int z) {
  int gx, gy, gz;
  gradient6_i(x, y, z, gx, gy, gz);
  return unit3(f3(-(float)(-gx), -(float)(-gy), -(float)(-gz)));
}
// Five occupancy bits of the voxels (x0 .. x0+4, Y, Z), bit i = voxel x0 + i; 0 outside the grid. The five
// voxels always straddle exactly two bricks of the row.
RM_DEV unsigned occ_row5(int x0, int Y, int Z) {
  if ((unsigned)Y >= (unsigned)g_opts.ry || (unsigned)Z >= (unsigned)g_opts.rz) return 0u;
  const int bxi = x0 >> 2;  // (x0 >= -2: an arithmetic shift gives brick -1 for x0 < 0)
  const unsigned row = (unsigned)(((Z >> 2) * g_accel.by + (Y >> 2)) * g_accel.bx);
  const unsigned nyb = (unsigned)(((Y & 3) << 2) | ((Z & 3) << 4));  // where the x-row sits inside a brick word
  unsigned lo = 0u, hi = 0u;
  if ((unsigned)bxi < (unsigned)g_accel.bx) lo = (unsigned)(__ldg(g_accel.occ + row + (unsigned)bxi) >> nyb) & 15u;
  if ((unsigned)(bxi + 1) < (unsigned)g_accel.bx) hi = (unsigned)(__ldg(g_accel.occ + row + (unsigned)(bxi + 1)) >> nyb) & 15u;
  return ((lo | (hi << 4)) >> (x0 & 3)) & 31u;
}

// voxelNormalSmooth (renderer.cl:190-203). The reference sums, over the occupied voxels q + d of the 3x3x3
// neighbourhood, their 6-tap gradients (occ(q+d-e) - occ(q+d+e) per axis e): <= 27 + 162 taps. The sums are sums
// of 0 / +-1, hence exact integers in any order, and along one axis they telescope: with a_k = occ(q + k*e + rest),
//   sum_{k=-1..1} a_k (a_{k-1} - a_{k+1}) = a_{-2} a_{-1} - a_1 a_2,
// so each component is a difference of two population counts over 9 rows: 21 five-voxel rows (42 brick words)
// instead of up to 189 single-voxel lookups, and the same integers.
RM_DEV float3 normal_smooth(int x, int y, int z) {
  int sx = 0, sy = 0, sz = 0;
  unsigned zprev = 0u;  // rows (dy = -1, 0, 1) of the previous z-slice, 5 bits each
#pragma unroll 1
  for (int dz = -2; dz <= 2; ++dz) {
    const bool zin = dz >= -1 && dz <= 1;
    unsigned zcur = 0u, yprev = 0u;
#pragma unroll 1
    for (int dy = zin ? -2 : -1; dy <= (zin ? 2 : 1); ++dy) {
      const unsigned r = occ_row5(x - 2, y + dy, z + dz);
      if (zin) {
        if (dy == -1) sy += __popc(yprev & r & 14u);  // rows y-2, y-1
        if (dy == 2) sy -= __popc(yprev & r & 14u);   // rows y+1, y+2
        if (dy >= -1 && dy <= 1) sx += (int)((r & (r >> 1)) & 1u) - (int)(((r >> 3) & (r >> 4)) & 1u);
      }
      if (dy >= -1 && dy <= 1) zcur |= r << (5 * (dy + 1));
      yprev = r;
    }
    if (dz == -1) sz += __popc(zprev & zcur & 0x39ceu);  // slices z-2, z-1 (mask: x offsets -1..1 of the three rows)
    if (dz == 2) sz -= __popc(zprev & zcur & 0x39ceu);   // slices z+1, z+2
    zprev = zcur;
  }
  return unit3(f3((float)sx, (float)sy, (float)sz));
}
// reference-equivalent occupancy taps of one hit (counting kernels only)
RM_DEV unsigned taps_of_hit(int x, int y, int z, bool smooth) {
  if (!smooth) return 6u;
  unsigned n = 0;
  for (int dz = -1; dz <= 1; ++dz)
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx) n += occ_at(x + dx, y + dy, z + dz);
  return 27u + 6u * n;
}

// ---- the fixed-step march (renderer.cl:219-234) ----------------------------------------------------
// Counting form: visits every sample the reference fetches (exact step counter).
template <int kMap>
RM_DEV bool march_counting(Cnt<true>& c, float3& p, float3 delta, int rem, float invS) {
  const float rxf = (float)g_opts.rx, ryf = (float)g_opts.ry, rzf = (float)g_opts.rz;
  while (rem > 0) {
    const int x = f2i_sat(p.x * rxf), y = f2i_sat(p.y * ryf), z = f2i_sat(p.z * rzf);
    c.steps++;
    if (!in_grid(x, y, z)) return false;
    const int d = cell_dist<kMap>(x, y, z);
    if (d != 0) {
      const float reach = (float)(d - 1) * g_accel.cellf - 0.25f;
      int n = reach > 0.0f ? 1 + f2i_sat(fminf(reach * invS, 1e6f)) : 1;
      n = n < rem ? n : rem;
      rem -= n;
      for (int j = 1; j <= n; ++j) {
        p = p + delta;
        if (j < n) {  // the reference fetches (and counts) every one of these samples
          c.steps++;
          if (!in_grid(f2i_sat(p.x * rxf), f2i_sat(p.y * ryf), f2i_sat(p.z * rzf))) return false;
        }
      }
    } else {
      if (solid_at(x, y, z)) return true;
      p = p + delta;
      rem -= 1;
    }
  }
  return false;
}

// Production form (see march_fast in rm_scene_plain.cuh for the derivation of the skip length and
// of the XU-free conversions): samples known to be empty cost three adds each and no fetch.
//
// This loop is more than half of the kernel's executed instructions, at ~10 active lanes: every lane
// of a warp pays for the union of the paths any lane takes, so it is written to have few of them.
//   * the skip length is computed without a branch for every d: (float)(d-1) = as_float(2^23 + d) -
//     (2^23 + 1), and k = rint((d-1)*A - B) clamps to 0 for d <= 1;
//   * the skipped adds run in a 2x-unrolled loop (+1 predicated add): 14 instructions of code instead
//     of the 60 of an 8/4/2/1 ladder, and 6 executed instructions instead of ~27 for the most common
//     skip (none: 53 % of all lookups sit in or next to an occupied cell);
//   * with 4-voxel cells the cell index IS the brick index.
// Which samples are looked at only decides how many fetches are saved, never the result: a hit is
// always decided by the solid bit of the exact voxel of a sample of the fp32 recurrence.
template <int kMap>
RM_DEV bool march_fast(float3& p, float3 delta, int rem, float invS) {
  RM_STAT_MARCH();
  const float A = g_accel.cellf * invS, B = 0.25f * invS + 0.5f;  // invS <= 2000 (march_delta)
  // Power-of-two grids: q = p * res and dq = delta * res are exact (a change of exponent), and so is every
  // q += dq against p += delta -- fl((a + b) * 2^k) = fl(a + b) * 2^k as long as nothing is subnormal, and a sum
  // of two such floats is either 0 or no smaller than an ulp of the larger one. The sample's voxel is then
  // trunc(q) with no multiply, and p = q / res (exact again) is handed back at the end.
  constexpr bool kPow2 = (kMap & kMapPow2) != 0;
  float px = p.x, py = p.y, pz = p.z;
  if (kPow2) {
    px *= g_accel.rxf; py *= g_accel.ryf; pz *= g_accel.rzf;
    delta = f3(delta.x * g_accel.rxf, delta.y * g_accel.ryf, delta.z * g_accel.rzf);
  }
  bool found = false;
  const unsigned nib = (kMap & kMapNib) ? nib_base() : 0u;
  const unsigned rx = g_opts.rx, ry = g_opts.ry, rz = g_opts.rz;
  const int cs = (kMap & kMapCell4) ? 2 : g_accel.cell_shift, my = g_accel.my, mx = g_accel.mx;
  const float rxf = kPow2 ? 1.0f : g_accel.rxf, ryf = kPow2 ? 1.0f : g_accel.ryf, rzf = kPow2 ? 1.0f : g_accel.rzf;
  while (rem > 0) {
    const int x = kPow2 ? f2i_sat(px) : f2i_sat(px * rxf);
    const int y = kPow2 ? f2i_sat(py) : f2i_sat(py * ryf);
    const int z = kPow2 ? f2i_sat(pz) : f2i_sat(pz * rzf);
    if ((unsigned)x >= rx || (unsigned)y >= ry || (unsigned)z >= rz) break;
    const unsigned c = (unsigned)(((z >> cs) * my + (y >> cs)) * mx + (x >> cs));
    int d;
    if (kMap & kMapNib) d = (nib_load(nib, c >> 1) >> ((c & 1u) << 2)) & 15;
    else d = __ldg(g_accel.dist + c);
    RM_STAT_LOOKUP();
    RM_STAT_EVENT(d == 0 ? 13 : (d == 1 ? 14 : (d == 2 ? 15 : 16)));
    if (d == 0) {
      const unsigned bi = (kMap & kMapCell4) ? c : (unsigned)(((z >> 2) * g_accel.by + (y >> 2)) * g_accel.bx + (x >> 2));
      if ((unsigned)(__ldg(g_accel.solid + bi) >> brick_bit(x, y, z)) & 1u) { found = true; break; }
    }
    // samples consumed by this iteration: this one plus the ones known to be empty
    const float dm1 = __int_as_float(0x4b000000 + d) - 8388609.0f;
    const int k = __float_as_int(fmaf(dm1, A, -B) + 12582912.0f) - 0x4b400000;  // (own bound, not reference arithmetic: one rounding is fine)
    const int n = 1 + (k > 0 ? k : 0);
    if (n >= rem) break;  // the march runs out inside space known to be empty: a miss
    rem -= n;
    RM_STAT_SKIP(n);
#pragma unroll 1
    for (int h = n >> 1; h > 0; --h) {  // (x, y) as one FADD2 (rm_math.cuh:add2): 7 instructions per two samples
      add2(px, py, delta.x, delta.y); pz += delta.z;
      add2(px, py, delta.x, delta.y); pz += delta.z;
    }
    if (n & 1) { px += delta.x; py += delta.y; pz += delta.z; }
  }
  p = kPow2 ? f3(px * g_accel.inv_rxf, py * g_accel.inv_ryf, pz * g_accel.inv_rzf) : f3(px, py, pz);
  return found;
}

// renderer.cl:209-237 without the normal; g = rpos.y + groundY is passed in by the caller (it has it).
// Shared by the sphere traces and the AO probes: one copy of the march loop in the kernel.
#ifndef RM_FUSED_SD_ATTR
#define RM_FUSED_SD_ATTR __device__ __noinline__
#endif
template <bool kCount, int kMap>
RM_FUSED_SD_ATTR Hit scene_distance(RM_CNT c, float3 rpos, float3 dir, float3 delta, int steps, float invS, float g,
                                    bool smooth) {
  const RmOpts& o = g_opts;
  Hit r;
  r.dist = g < 1e5f ? g : 1e5f;
  r.flags = 0;
  r.p = rpos;  // (only read after a hit)
  const bool inside = rpos.x > o.boundsMin.x && rpos.x < o.boundsMax.x && rpos.y > o.boundsMin.y &&
                      rpos.y < o.boundsMax.y && rpos.z > o.boundsMin.z && rpos.z < o.boundsMax.z;
  float idist = 0.0f;
  RM_STAT_EVENT(0);
  if (!inside) {  // (a real branch: the call sites that reach this function are mostly inside the box already)
    const bool away = (rpos.x > o.boundsMax.x && dir.x > 0.0f) || (rpos.x < o.boundsMin.x && dir.x < 0.0f) ||
                      (rpos.y > o.boundsMax.y && dir.y > 0.0f) || (rpos.y < o.boundsMin.y && dir.y < 0.0f) ||
                      (rpos.z > o.boundsMax.z && dir.z > 0.0f) || (rpos.z < o.boundsMin.z && dir.z < 0.0f);
    idist = away ? -1.0f : box_entry(o.boundsMin, o.boundsMax, rpos, dir);
    RM_STAT_EVENT(away ? 2 : 3);
  } else {
    RM_STAT_EVENT(1);
  }
  if (idist >= 0.0f && idist < r.dist) {
    RM_STAT_EVENT(4);
    float3 p = rpos + o.voxelBounds;
    if (idist > 0.0f) p = dir * idist + p;
    p = p * o.invVoxelScale;
    bool found;
    if constexpr (kCount) found = march_counting<kMap>(c, p, delta, steps, invS);
    else found = march_fast<kMap>(p, delta, steps, invS);
    r.p = p;
    if (found) {
      RM_STAT_EVENT(5);
      r.flags = kHit;
      if constexpr (kCount) {
        const int x = f2i_sat(p.x * (float)o.rx), y = f2i_sat(p.y * (float)o.ry), z = f2i_sat(p.z * (float)o.rz);
        c.taps += taps_of_hit(x, y, z, smooth);
      }
      const float3 hp = p * o.voxelBounds2 + (-o.voxelBounds);
      const float d = len3(rpos - hp) - o.voxelSize;
      if (d < r.dist) { r.dist = d; r.flags = kHit | kCloser; }
    }
  }
  return r;
}

// per-direction constants of the march: delta (renderer.cl:215) and 1 / (largest step in voxels)
#ifndef RM_FUSED_MD_ATTR
#define RM_FUSED_MD_ATTR RM_DEV
#endif
struct Delta { float3 d; float invS; };
RM_FUSED_MD_ATTR Delta march_delta(float3 dir, int steps) {
  const RmOpts& o = g_opts;
  Delta r;
  r.d = (dir / ((float)steps * 0.5f)) * o.invVoxelScale;
  const float sm = fmaxf(fmaxf(fabsf(r.d.x) * (float)o.rx, fabsf(r.d.y) * (float)o.ry), fabsf(r.d.z) * (float)o.rz);
  r.invS = sm > 5e-4f ? rcp_fast(sm) : 2000.0f;
  return r;
}

// The march window of a trace (see rm_scene_plain.cuh:march_window for the argument).
RM_DEV void march_window(float3 ro, float3 rd, float maxDist, float& tin, float& tout) {
  const RmOpts& o = g_opts;
  const float kGrow = 0.01f, kTiny = 1e-5f, kBig = 64.0f, kInf = 3.0e38f;
  tin = -kInf;
  tout = kInf;
  const float ro_[3] = {ro.x, ro.y, ro.z}, rd_[3] = {rd.x, rd.y, rd.z};
  const float lo_[3] = {o.boundsMin.x, o.boundsMin.y, o.boundsMin.z}, hi_[3] = {o.boundsMax.x, o.boundsMax.y, o.boundsMax.z};
  // (the box itself is checked once per launch: o.window_ok, rm_derive_opts)
  bool ok = o.window_ok && maxDist <= kBig, miss = false;
#pragma unroll
  for (int i = 0; i < 3; ++i) ok = ok && fabsf(ro_[i]) <= kBig && fabsf(rd_[i]) <= 2.0f;
  if (!ok) return;
  float a = -kInf, b = kInf;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float l = lo_[i] - kGrow - ro_[i], h = hi_[i] + kGrow - ro_[i];
    if (fabsf(rd_[i]) < kTiny) {
      if (l > 0.0f || h < 0.0f) miss = true;
    } else {
      const float inv = rcp_fast(rd_[i]);
      const float t0 = l * inv, t1 = h * inv;
      a = fmaxf(a, fminf(t0, t1));
      b = fminf(b, fmaxf(t0, t1));
    }
  }
  if (miss || b < a) { tin = kInf; tout = -kInf; return; }
  tin = a;
  tout = b;
}

// renderer.cl:239-257. One copy for the primary / bounce / shadow traces.
#ifndef RM_FUSED_ST_ATTR
#define RM_FUSED_ST_ATTR __device__ __noinline__
#endif
template <bool kCount, int kMap>
RM_FUSED_ST_ATTR Isec sphere_trace(RM_CNT c, float3 ro, float3 rd, float maxDist, int maxSteps, bool smooth,
                                   bool wantSurface, bool unitDir) {
  const RmOpts& o = g_opts;
  RM_STAT_TRACE();
  RM_STAT_EVENT(wantSurface ? 6 : 7);
  const Delta md = march_delta(rd, o.maxVoxelIter);
  const float3 delta = md.d;
  const float invS = md.invS;
  Hit j;
  j.dist = 0.0f; j.flags = 0; j.p = f3s(0.0f);
  float jg = 0.0f;  // ground distance of the last evaluation (the ground's "id" is (int)g, renderer.cl:211)
  float dist = o.startDist;
  float3 pos = ro;
  // samples per world unit along the ray, a conservative upper bound (only used to cut marches short): for a unit
  // direction the per-launch constant of rm_derive_opts, else from the step's length
  const float inv_step = kCount ? 0.0f : (unitDir ? o.st_k : 1.01f / len3(delta * o.voxelBounds2));
  int cut = 0;
  float tin = -3.0e38f, tout = 3.0e38f;
  if (!kCount) march_window(ro, rd, maxDist, tin, tout);
  while (--maxSteps >= 0) {
    if constexpr (kCount) c.outer++;
    RM_STAT_EVENT(wantSurface ? 8 : 9);
    pos = ro + rd * dist;
    const float g = pos.y + o.groundY;
    jg = g;
    if (!kCount && (dist > tout || tin - dist > g * 1.0001f + 1e-3f || g <= 0.0f)) {
      RM_STAT_EVENT(12);
      j.dist = g < 1e5f ? g : 1e5f;
      j.flags = 0;
      cut = 0;
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
        if (!wantSurface) reach = fminf(reach, fmaxf(maxDist - dist, o.eps));
        const float k = (reach + o.voxelSize) * inv_step;
        cut = k < (float)(limit - 2) ? 1 : 0;
        if (cut) limit = f2i_sat(k) + 2;
      }
      j = scene_distance<kCount, kMap>(c, pos, rd, delta, limit, invS, g, smooth);
    }
    if (fabsf(j.dist) <= o.eps || dist >= maxDist) break;
    dist += j.dist;
  }
  if (!kCount && wantSurface && cut && !(j.flags & kHit)) {
    RM_STAT_EVENT(10);
    j = scene_distance<kCount, kMap>(c, pos, rd, delta, o.maxVoxelIter, invS, jg, smooth);
  }
  const bool miss = dist >= maxDist;
  if (miss) {
    pos = ro + rd * dist;
    dist = 1000.0f;
  }
  Isec r;
  r.distance = dist;
  r.pos = pos;
  r.objectID = -1;
  r.normal = f3s(0.0f);
  if (!wantSurface) return r;  // shadow rays use the distance only
  const int x = f2i_sat(j.p.x * (float)o.rx), y = f2i_sat(j.p.y * (float)o.ry), z = f2i_sat(j.p.z * (float)o.rz);
  if (!miss) {
    if (j.flags & kCloser) {
      const int v = voxel_value(x, y, z);
      r.objectID = v < 168 ? (v < 84 ? 1 : 2) : 3;  // voxelMaterial, renderer.cl:205-207
    } else {
      r.objectID = f2i_sat(jg < 1e5f ? jg : -1.0f);
    }
  }
  if (j.flags & kHit) r.normal = smooth ? normal_smooth(x, y, z) : normal_6tap(x, y, z);
  else r.normal = jg < 1e5f ? f3(0.0f, 1.0f, 0.0f) : -rd;
  return r;
}

RM_DEV float3 sky(float3 d) { return lerp3(g_opts.sky1, g_opts.sky2, d.y * 0.5f + 0.5f); }  // :259-261

// renderer.cl:263-269. The jitter of the light positions, table[(uint)(px*1957 + py*2173 + time*4763.742)] *
// lightScatter, is the same for every light and every surface of a pixel-sample: it is fetched and scaled once
// (render_pixel_sample) and handed down instead of (px, py); lightPos(i) is then one vector add -- the same
// operands and roundings as the reference's expression.
RM_DEV float3 light_jitter(Lane s, float px, float py) {
  const uint32_t seed = f2u_wrap(px * 1957.0f + py * 2173.0f + s.time * 4763.742f);
  return table_xyz(s, seed) * g_opts.lightScatter;
}
RM_DEV float3 light_pos(float3 ljit, int i) { return ljit + g_opts.lightPos[i]; }

RM_DEV float3 reflect3(float3 v, float3 n) { return v - n * (2.0f * dot3(v, n)); }  // :271-273

// renderer.cl:275-290
RM_SHARED_FN float3 atmosphere(float3 ljit, float3 ro, float3 rd, float distance, float3 col) {
  const RmOpts& o = g_opts;
  const float fa = 1.0f - expf(distance * distance * -o.fogPow);
  col = (sky(rd) - col) * fa + col;
#pragma unroll 1  // (ptxas unrolls this four times otherwise: 1.6 KB of code for a loop that runs 1-4 times per surface)
  for (int i = 0; i < o.numLights; ++i) {
    float3 lp = light_pos(ljit, i);
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
template <bool kCount, int kMap>
RM_DEV float ambient_occlusion(RM_CNT c, Lane s, float3 pos, float3 n0) {
  const RmOpts& o = g_opts;
  float ao = 1.0f, d = 0.0f;
  uint32_t seed = f2u_wrap(pos.x * 3183.75f + pos.y * 1831.42f + pos.z * 2945.87f + s.time * 2671.918f);
  for (int i = 0; i <= o.aoIter && ao > 0.01f; ++i) {
    d += o.aoStepDist;
    seed += 37u;
    const float3 n = unit3(table_xyz(s, seed) * 0.2f + n0);
    const float3 q = n * d + pos;
    const float g = q.y + o.groundY;
    float hdist;
    const float out = fmaxf(fmaxf(fmaxf(o.boundsMin.x - q.x, q.x - o.boundsMax.x), fmaxf(o.boundsMin.y - q.y, q.y - o.boundsMax.y)),
                            fmaxf(o.boundsMin.z - q.z, q.z - o.boundsMax.z));
    if (!kCount && o.aoAmp >= 0.0f && d > 0.0f && out > d + o.voxelSize + 1e-3f && out < 1e3f) {
      RM_STAT_EVENT(18);
      hdist = g < 1e5f ? g : 1e5f;
    } else {
      const int msteps = o.maxVoxelIter / 2;
      const Delta md = march_delta(n, msteps);
      const float3 delta = md.d;
      const float invS = md.invS;
      int limit = msteps;
      if (!kCount && o.aoAmp >= 0.0f && d > 0.0f) {
        // (n is a unit vector -- or 0, and then the march does not move and its length is immaterial --, so the
        //  samples within the reach are bounded by a per-launch constant: no sqrt, no division per probe)
        const float k = (d + o.voxelSize) * o.ao_k;
        if (k < (float)msteps) limit = f2i_sat(k) + 2 < msteps ? f2i_sat(k) + 2 : msteps;
      }
      RM_STAT_EVENT(11);
      RM_STAT_SITE(RM_STAT_LEVEL_GET() * 16 + 8 + i);
      hdist = scene_distance<kCount, kMap>(c, q, n, delta, limit, invS, g, false).dist;
    }
    ao *= 1.0f - cl_max((d - hdist) * o.aoAmp / d, 0.0f);
  }
  return ao;
}

// renderer.cl:348-381 (shadow :292-301 inlined). One copy for the primary and the bounce surfaces.
#ifndef RM_FUSED_OL_ATTR
#define RM_FUSED_OL_ATTR __device__ __noinline__
#endif
template <bool kCount, int kMap>
RM_FUSED_OL_ATTR float3 object_lighting(RM_CNT c, Lane s, float3 ljit, float3 rd, float3 ipos, int mat, float3 n,
                                        float3 reflectCol) {
  const RmOpts& o = g_opts;
  const RmMaterial& m = o.mat[mat];
  const float ao = ambient_occlusion<kCount, kMap>(c, s, ipos, n);
  float3 diff = sky(n) * ao;
  float3 spec = reflectCol * ao;
  float3 fin = f3s(0.0f);
  for (int i = 0; i < o.numLights; ++i) {
    const float3 dl = light_pos(ljit, i) - ipos;
    const float ld2 = dot3(dl, dl);
    const float att = 1.0f / ld2;
    if (att > o.minLightAtt) {
      const float3 ldir = unit3(dl);
      const float lmax = cl_min(sqrtf(ld2) - o.shadowBias, o.maxDist);
      const float kd = cl_max(0.0f, dot3(ldir, n));
      const float ks = blinn_phong(m.smoothness, rd, ldir, n);
      const float3 inc = (o.lightColor[i] * 1.0f) * att;  // (lightColor * shadow factor 1) * att, :369
      const float3 zero = inc * 0.0f;
      const bool irrelevant = !kCount && kd == 0.0f && ks == 0.0f && zero.x == 0.0f && zero.y == 0.0f && zero.z == 0.0f;
      if (!irrelevant) {
        RM_STAT_SITE(RM_STAT_LEVEL_GET() * 16 + 1 + i);
        const Isec sh = sphere_trace<kCount, kMap>(c, ipos + ldir * o.shadowBias, ldir, lmax, o.shadowIter, false, false, true);
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

// renderer.cl:407-446 with basicSceneColor (:383-405) in its bounce loop
template <bool kCount, int kMap>
RM_DEV float3 scene_color(RM_CNT c, Lane s, float3 ljit, float3 mcNormal, float3 ro, float3 rd) {
  const RmOpts& o = g_opts;
  RM_STAT_LEVEL(0);
  RM_STAT_SITE(0);
  const Isec isec = sphere_trace<kCount, kMap>(c, ro, rd, o.maxDist, o.maxIter, true, true, true);
  float3 col;
  if (isec.distance >= o.maxDist) {
    col = sky(rd);
  } else {
    const int mi = mat_index(isec.objectID);
    const RmMaterial& m = o.mat[mi];
    const float3 n = mcNormal * (1.0f / (m.smoothness * 200.0f + 5.0f)) + isec.normal;
    float3 reflectCol = f3s(0.0f);
    if (m.r0 > 0.0f && o.reflectIter > 0) {
      float3 bpos = isec.pos, bn = n, bd = rd;
      for (int i = 0; i < o.reflectIter; ++i) {
        bd = reflect3(bd, bn);
        const float3 bo = bpos + bd * 0.0075f;
        RM_STAT_LEVEL(i + 1);
        RM_STAT_SITE(RM_STAT_LEVEL_GET() * 16);
        const Isec ri = sphere_trace<kCount, kMap>(c, bo, bd, o.maxDist, o.maxIter, false, true, false);  // (a reflection about an un-normalised normal is not a unit vector)
        float3 bc;
        if (ri.objectID < 0) bc = sky(bd);
        else bc = object_lighting<kCount, kMap>(c, s, ljit, bd, ri.pos, mat_index(ri.objectID), ri.normal,
                                                sky(reflect3(bd, ri.normal)));
        reflectCol = reflectCol + atmosphere(ljit, bo, bd, ri.distance, bc);
        if (ri.objectID < 0) break;
        if (o.mat[mat_index(ri.objectID)].r0 < 0.001f) break;
        bpos = ri.pos;
        bn = ri.normal;
      }
    } else {
      reflectCol = sky(reflect3(rd, n));
    }
    RM_STAT_LEVEL(0);
    col = object_lighting<kCount, kMap>(c, s, ljit, rd, isec.pos, mi, n, reflectCol);
  }
  return atmosphere(ljit, ro, rd, isec.distance, col);
}

// One work-item of RenderImage (renderer.cl:478-494; initRenderState :467-476, cameraRayLookat
// :456-465): returns sceneColor * exposure.
template <bool kCount, int kMap>
RM_DEV float3 render_pixel_sample(RM_CNT c, Lane s, int id) {
  const RmOpts& o = g_opts;
  const float4 a = table_at(s, (uint32_t)(id * 17) + f2u_wrap(s.time * 3141.3862f));
  const float3 mcNormal = unit3(table_xyz(s, (uint32_t)(id * 37) + f2u_wrap(s.time * 1859.1467f)));
  const float px = (float)(id % o.width) + a.z;
  const float py = (float)(id / o.width) + a.w;
  const float3 eye = f3(mcNormal.z, mcNormal.x, mcNormal.y) * o.dof + o.eyePos;
  const float3 fwd = unit3(o.targetPos - eye);
  const float3 right = unit3(cross3(fwd, o.up));
  const float vx = px / (float)o.width * o.fov - o.fov * 0.5f;
  float vy = py / (float)o.height * o.fov - o.fov * 0.5f;
  vy = vy * -o.invAspect;
  const float3 rd = unit3(right * vx + cross3(right, fwd) * vy + fwd);
  return scene_color<kCount, kMap>(c, s, light_jitter(s, px, py), mcNormal, eye, rd) * o.exposure;
}

}  // namespace fused
