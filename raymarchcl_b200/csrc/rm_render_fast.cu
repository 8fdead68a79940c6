// rm_render_fast.cu -- the production form of RenderImage (renderer.cl:478-494) for sm_100a.
//
// Same arithmetic as the plain kernel (rm_scene_plain.cuh; both are checked against the oracle),
// organised for the machine instead of for the source text:
//
//  * persistent warps, lane refill: every lane runs a small state machine over work items
//    (pixel, pass); a lane that finishes an item takes the next one from a warp-local pool that
//    is refilled 32 items at a time from one global counter. All passes of a frame are items of
//    ONE launch (per-pass colours go to a colour buffer and are blended in pass order afterwards,
//    k_blend_passes), so the 10x spread in per-pixel cost is balanced over ~450 items per lane.
//  * one march loop for every ray kind: primary / bounce / shadow sphere-traces and the AO probes
//    are all "distanceToScene" jobs (renderer.cl:209-237); the lanes of a warp sit in the same
//    inner loop whatever kind of ray they are tracing.
//  * fetch elision: the reference's fixed-step march keeps its fp32 recurrence p += delta (the
//    sample positions are part of the result), but a byte-per-macro-cell Chebyshev distance map
//    says how many of the next samples cannot be solid, so those samples cost three adds and no
//    fetch; near the surface the predicate v > isoVal is read from 4x4x4 bit-bricks (8x less
//    data than the byte volume). Normals read the 27-neighbourhood from bit-bricks too, and are
//    evaluated once per trace (only the last distanceToScene call's normal is ever used).
//  * the slab test's six IEEE divisions are skipped when the ray origin is strictly inside the
//    voxel box (the test then returns exactly 0 whatever the quotients are).
//
// Compiled with -fmad=false like the rest of the library (pinned two-rounding evaluation order).
#include <cooperative_groups.h>

#include "rm_kernels.h"
#include "rm_scene_plain.cuh"

namespace fast {

using plain::PixelState;
using plain::Scene;

enum : int {
  S_IDLE = 0,      // needs a work item
  S_INIT,          // item assigned: set up the pixel, start the primary trace
  S_SLOW_FIRST,
  S_TRACE_END = S_SLOW_FIRST,  // sphere-trace finished: continuation by trace kind
  S_AO_NEXT,       // next ambient-occlusion probe, or start of the light loop
  S_LIGHT_NEXT,    // next light: shadow trace or skip
  S_LIGHT_POST,    // tail of the light loop body
  S_SURFACE_DONE,  // objectLighting finished for the current surface
  S_SLOW_LAST = S_SURFACE_DONE,
  S_TRACE_ITER,    // head of the sphere-trace loop: set up one distanceToScene job
  S_MARCH,         // inside the inner march
  S_MARCH_END,     // distanceToScene result -> consumer
  S_DONE
};

enum : int { T_PRIMARY = 0, T_BOUNCE = 1, T_SHADOW = 2 };
enum : int { C_TRACE = 0, C_AO = 1 };

struct Lane {
  int state;
  // work item
  int id;
  long long item;
  PixelState st;
  float3 rd0;
  // sphere-trace (raymarch, renderer.cl:239-257)
  int tkind, itersLeft;
  float3 ro;
  float tdist, tmax;
  // distanceToScene job (renderer.cl:209-237)
  int consumer, msteps, rem;
  float3 rpos, mdir, delta, p;
  float g, invS;
  bool hit;         // this job ended on a solid voxel (p is the sample position of the hit)
  bool closer;      // ... and the voxel distance won against the ground plane
  float hx;         // result distance
  // primary surface
  float3 ppos, pn, reflAcc;
  float pdist;
  int pmi, bi;
  // current bounce
  float3 bo, bd, rpos_b, rn_b;
  float rdist_b;
  int robj_b;
  // lighting of the current surface (objectLighting, renderer.cl:348-381)
  int lsurf, lmi, li, aoI;
  float3 ipos, ln, lvdir, lrefl, diff, spec, fin;
  float ao, aoD, att;
  uint32_t aoSeed;
};

struct Grid {
  const RmOpts& o;
  const RmAccel& a;
  float rxf, ryf, rzf;
};

RM_DEV bool in_grid(const RmOpts& o, int x, int y, int z) {
  return (unsigned)x < (unsigned)o.rx && (unsigned)y < (unsigned)o.ry && (unsigned)z < (unsigned)o.rz;
}

RM_DEV uint64_t brick_word(const uint64_t* __restrict__ bricks, const RmAccel& a, int x, int y, int z) {
  return __ldg(bricks + ((size_t)(z >> 2) * a.by + (y >> 2)) * a.bx + (x >> 2));
}
RM_DEV unsigned brick_bit(int x, int y, int z) { return (x & 3) | ((y & 3) << 2) | ((z & 3) << 4); }

// voxelLookupI (renderer.cl:172-178) from the (v >= isoVal) bit-bricks; 0 outside the grid
RM_DEV int occ_at(const Grid& G, int x, int y, int z) {
  if (!in_grid(G.o, x, y, z)) return 0;
  return (int)((brick_word(G.a.occ, G.a, x, y, z) >> brick_bit(x, y, z)) & 1ull);
}

// voxelNormal (renderer.cl:180-188) as integers
RM_DEV void gradient6_i(const Grid& G, int x, int y, int z, int& gx, int& gy, int& gz) {
  gx = -(occ_at(G, x + 1, y, z) - occ_at(G, x - 1, y, z));
  gy = -(occ_at(G, x, y + 1, z) - occ_at(G, x, y - 1, z));
  gz = -(occ_at(G, x, y, z + 1) - occ_at(G, x, y, z - 1));
}

// voxelNormalSmooth (renderer.cl:190-203): the float sums of the reference are sums of small
// integers, hence exact; the integer sum converted once is the same value.
RM_DEV float3 normal_smooth(const Grid& G, int x, int y, int z) {
  int sx = 0, sy = 0, sz = 0;
  for (int dz = -1; dz <= 1; ++dz)
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx)
        if (occ_at(G, x + dx, y + dy, z + dz)) {
          int gx, gy, gz;
          gradient6_i(G, x + dx, y + dy, z + dz, gx, gy, gz);
          sx += gx; sy += gy; sz += gz;
        }
  return unit3(f3((float)sx, (float)sy, (float)sz));
}

RM_DEV float3 normal_6tap(const Grid& G, int x, int y, int z) {
  int gx, gy, gz;
  gradient6_i(G, x, y, z, gx, gy, gz);
  // -(a - b) in floats gives -0.0f for a == b; unit3 then divides: keep the reference's signs
  return unit3(f3(-(float)(-gx), -(float)(-gy), -(float)(-gz)));
}

// reference-equivalent occupancy taps of one hit (counting mode only)
RM_DEV unsigned taps_of_hit(const Grid& G, int x, int y, int z, bool smooth) {
  if (!smooth) return 6u;
  unsigned n = 0;
  for (int dz = -1; dz <= 1; ++dz)
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx) n += occ_at(G, x + dx, y + dy, z + dz);
  return 27u + 6u * n;
}

// Set up one distanceToScene job (renderer.cl:209-218): ground plane, slab test, march start.
RM_DEV void job_begin(Lane& L, const RmOpts& o) {
  L.g = L.rpos.y + o.groundY;
  const float gx = L.g < 1e5f ? L.g : 1e5f;
  L.hit = false;
  L.closer = false;
  L.hx = gx;
  const bool inside = L.rpos.x > o.boundsMin.x && L.rpos.x < o.boundsMax.x && L.rpos.y > o.boundsMin.y &&
                      L.rpos.y < o.boundsMax.y && L.rpos.z > o.boundsMin.z && L.rpos.z < o.boundsMax.z;
  // strictly inside: every entry quotient is < 0 and every exit quotient > 0, so the slab test
  // returns max(...,0) = +0 without evaluating the divisions
  const float idist = inside ? 0.0f : plain::box_entry(o.boundsMin, o.boundsMax, L.rpos, L.mdir);
  if (idist >= 0.0f && idist < gx && L.msteps > 0) {
    float3 p = L.rpos + o.voxelBounds;
    if (idist > 0.0f) p = L.mdir * idist + p;
    L.p = p * o.invVoxelScale;
    L.rem = L.msteps;
    L.state = S_MARCH;
  } else {
    L.state = S_MARCH_END;
  }
}

// per-job constants of the march: delta (renderer.cl:215) and 1 / (largest step in voxels)
RM_DEV void job_direction(Lane& L, const Grid& G, float3 dir, int steps) {
  L.mdir = dir;
  L.msteps = steps;
  L.delta = (dir / ((float)steps * 0.5f)) * G.o.invVoxelScale;
  const float s = fmaxf(fmaxf(fabsf(L.delta.x) * G.rxf, fabsf(L.delta.y) * G.ryf), fabsf(L.delta.z) * G.rzf);
  L.invS = s > 1e-12f ? __fdividef(1.0f, s) : 1e12f;
}

RM_DEV void trace_begin(Lane& L, const Grid& G, int kind, float3 ro, float3 rd, float maxDist, int iters) {
  L.tkind = kind;
  L.ro = ro;
  L.tdist = G.o.startDist;
  L.tmax = maxDist;
  L.itersLeft = iters;
  L.consumer = C_TRACE;
  L.rpos = ro;
  L.g = 0.0f;
  L.hit = false;
  L.closer = false;
  job_direction(L, G, rd, G.o.maxVoxelIter);
  L.state = S_TRACE_ITER;
}

RM_DEV void lighting_begin(Lane& L, const Scene& s, int surf, float3 ipos, float3 n, int mi, float3 vdir, float3 refl) {
  L.lsurf = surf; L.ipos = ipos; L.ln = n; L.lmi = mi; L.lvdir = vdir; L.lrefl = refl;
  L.ao = 1.0f; L.aoD = 0.0f; L.aoI = 0;
  L.aoSeed = f2u_wrap(ipos.x * 3183.75f + ipos.y * 1831.42f + ipos.z * 2945.87f + s.time * 2671.918f);
  L.state = S_AO_NEXT;
}

RM_DEV void bounce_begin(Lane& L, const Grid& G) {
  L.bd = plain::reflect3(L.bd, L.rn_b);
  L.bo = L.rpos_b + L.bd * 0.0075f;
  trace_begin(L, G, T_BOUNCE, L.bo, L.bd, G.o.maxDist, G.o.maxIter);
}

}  // namespace fast

namespace {

using namespace fast;
namespace cg = cooperative_groups;

constexpr int kFastBlock = 128;
constexpr int kMinMarchIters = 4;

struct FastParams {
  const float4* tables;    // passes x 16384 float4
  float times[RM_MAX_FUSED_PASSES];  // TRenderOpts.time per pass
  float4* colour;          // passes x slots (null when passes == 1: blend straight into accum)
  float4* accum;
  unsigned long long* queue;
  RmCounters* counters;
  int passes;
  unsigned* watchdog;      // [0] = tripped flag, [1..15] = state of the first lane that tripped
  unsigned trip_limit;     // trips round the state machine a lane may take before giving up
  int march_quota;         // max march iterations per trip round the state machine
  int min_marchers;        // leave the march loop when fewer lanes are marching together
};

// Every lane runs its own state machine; nothing below depends on the lanes of a warp being
// converged (no *_sync intrinsic names a fixed mask): the hardware's SIMT reconvergence only
// decides how many lanes execute a section together, never what they compute.
template <bool kCount>
__global__ void __launch_bounds__(kFastBlock)
k_render_fast(const __grid_constant__ RmOpts o, const __grid_constant__ RmShard sh,
              const __grid_constant__ RmAccel acc, const __grid_constant__ FastParams P) {
  const long long total = (long long)P.passes * sh.slots;
  const Grid G{o, acc, (float)o.rx, (float)o.ry, (float)o.rz};
  Scene s(acc.vox, P.tables, o);
  Lane L = {};
  L.state = S_IDLE;
  const float cellf = (float)(1 << acc.cell_shift);
  unsigned trips = 0;

  for (;;) {
    // ---- 1. a lane without work takes the next item; lanes that are idle at the same time
    //         share one atomic (coalesced group) and get consecutive items, i.e. adjacent pixels
    if (L.state == S_IDLE) {
      cg::coalesced_group grp = cg::coalesced_threads();
      unsigned long long base = 0;
      if (grp.thread_rank() == 0) base = atomicAdd(P.queue, (unsigned long long)grp.size());
      base = grp.shfl(base, 0);
      L.item = (long long)base + grp.thread_rank();
      if (L.item >= total) break;  // queue exhausted: this lane is finished

      // ---- 2. item set-up (initRenderState + cameraRayLookat, renderer.cl:456-476) ----
      const int pass = (int)(L.item / sh.slots);
      const long long slot = L.item - (long long)pass * sh.slots;
      L.id = rm_slot_to_pixel(sh, slot, o.width, o.height);
      if (L.id < 0) continue;  // padding slot of an edge tile
      s.time = P.times[pass];
      s.table = P.tables + (size_t)pass * (RM_TABLE_MASK + 1);
      L.rd0 = plain::setup_pixel(s, L.id, L.st);
      trace_begin(L, G, T_PRIMARY, L.st.eye, L.rd0, o.maxDist, o.maxIter);
    }

    // watchdog: a lane that does not finish within trip_limit trips records why and gives up, so
    // that a logic error can never hang the device (the host turns the flag into an error)
    if (++trips > P.trip_limit) {
      if (atomicCAS(P.watchdog, 0u, 1u) == 0u) {
        P.watchdog[1] = (unsigned)L.state; P.watchdog[2] = (unsigned)L.tkind; P.watchdog[3] = (unsigned)L.consumer;
        P.watchdog[4] = (unsigned)L.rem; P.watchdog[5] = (unsigned)L.itersLeft; P.watchdog[6] = (unsigned)L.id;
        P.watchdog[7] = (unsigned)L.item; P.watchdog[8] = (unsigned)L.li; P.watchdog[9] = (unsigned)L.aoI;
        P.watchdog[10] = (unsigned)L.bi; P.watchdog[14] = blockIdx.x; P.watchdog[15] = threadIdx.x;
      }
      break;
    }

    // ---- 3. shading transitions (rare, divergent) ----
#pragma unroll 1
    for (int round = 0; round < 4 && L.state >= S_SLOW_FIRST && L.state <= S_SLOW_LAST; ++round) {
      switch (L.state) {
        case S_TRACE_END: {
          // tail of raymarch (renderer.cl:252-256)
          float distance = L.tdist;
          const bool miss = L.tdist >= L.tmax;
          if (miss) {
            L.rpos = L.ro + L.mdir * L.tdist;
            distance = 1000.0f;
          }
          if (L.tkind == T_SHADOW) {
            // shadow (renderer.cl:292-301) and the lit branch of objectLighting (:366-374)
            const float sf = distance < L.tmax ? 0.0f : 1.0f;
            if (sf > 0.0f) {
              const RmMaterial& m = o.mat[L.lmi];
              const float3 inc = (o.lightColor[L.li] * sf) * L.att;
              L.diff = L.diff + inc * cl_max(0.0f, dot3(L.mdir, L.ln));
              L.spec = L.spec + inc * plain::blinn_phong(m.smoothness, L.lvdir, L.mdir, L.ln);
            }
            L.state = S_LIGHT_POST;
            break;
          }
          // result of the LAST distanceToScene call of the trace: object id and normal
          int objectID = -1;
          float3 normal;
          {
            const int x = f2i_sat(L.p.x * G.rxf), y = f2i_sat(L.p.y * G.ryf), z = f2i_sat(L.p.z * G.rzf);
            if (!miss) {
              if (L.closer) {
                const int v = __ldg(acc.vox + ((size_t)z * o.rxy + (size_t)y * o.rx + x));
                objectID = v < 168 ? (v < 84 ? 1 : 2) : 3;  // voxelMaterial, renderer.cl:205-207
              } else {
                objectID = f2i_sat(L.g < 1e5f ? L.g : -1.0f);  // the ground's "id" is its distance (:211)
              }
            }
            if (L.hit) normal = L.tkind == T_PRIMARY ? normal_smooth(G, x, y, z) : normal_6tap(G, x, y, z);
            else normal = L.g < 1e5f ? f3(0.0f, 1.0f, 0.0f) : -L.mdir;
          }
          if (L.tkind == T_PRIMARY) {
            // sceneColor (renderer.cl:407-446)
            L.pdist = distance;
            if (distance >= o.maxDist) {
              const float3 col = plain::atmosphere(s, L.st, L.st.eye, L.rd0, distance, plain::sky(o, L.rd0));
              L.fin = col;
              L.lsurf = 2;  // pixel finished
              L.state = S_SURFACE_DONE;
              break;
            }
            L.pmi = plain::mat_index(objectID);
            const RmMaterial& m = o.mat[L.pmi];
            L.pn = L.st.mcNormal * (1.0f / (m.smoothness * 200.0f + 5.0f)) + normal;
            L.ppos = L.rpos;
            if (m.r0 > 0.0f && o.reflectIter > 0) {
              L.reflAcc = f3s(0.0f);
              L.bi = 0;
              L.rpos_b = L.ppos;
              L.rn_b = L.pn;
              L.bd = L.rd0;
              bounce_begin(L, G);
            } else {
              lighting_begin(L, s, 0, L.ppos, L.pn, L.pmi, L.rd0, plain::sky(o, plain::reflect3(L.rd0, L.pn)));
            }
          } else {
            // basicSceneColor (renderer.cl:383-405)
            L.rpos_b = L.rpos;
            L.rn_b = normal;
            L.robj_b = objectID;
            L.rdist_b = distance;
            if (objectID < 0) {
              const float3 col = plain::atmosphere(s, L.st, L.bo, L.bd, distance, plain::sky(o, L.bd));
              L.reflAcc = L.reflAcc + col;
              lighting_begin(L, s, 0, L.ppos, L.pn, L.pmi, L.rd0, L.reflAcc);  // bounce loop breaks (:428)
            } else {
              lighting_begin(L, s, 1, L.rpos_b, L.rn_b, plain::mat_index(objectID), L.bd,
                             plain::sky(o, plain::reflect3(L.bd, L.rn_b)));
            }
          }
          break;
        }
        case S_AO_NEXT: {
          // ambientOcclusion loop head (renderer.cl:333-336)
          if (L.aoI <= o.aoIter && L.ao > 0.01f) {
            L.aoD += o.aoStepDist;
            L.aoSeed += 37u;
            const float3 n = unit3(plain::table_xyz(s, L.aoSeed) * 0.2f + L.ln);
            L.consumer = C_AO;
            job_direction(L, G, n, o.maxVoxelIter / 2);
            L.rpos = n * L.aoD + L.ipos;
            job_begin(L, o);
          } else {
            L.diff = plain::sky(o, L.ln) * L.ao;
            L.spec = L.lrefl * L.ao;
            L.fin = f3s(0.0f);
            L.li = 0;
            L.state = S_LIGHT_NEXT;
          }
          break;
        }
        case S_LIGHT_NEXT: {
          if (L.li >= o.numLights) { L.state = S_SURFACE_DONE; break; }
          const float3 dl = plain::light_pos(s, L.st, L.li) - L.ipos;
          const float ld2 = dot3(dl, dl);
          L.att = 1.0f / ld2;
          if (L.att > o.minLightAtt) {
            const float3 ldir = unit3(dl);
            const float lmax = cl_min(sqrtf(ld2) - o.shadowBias, o.maxDist);
            trace_begin(L, G, T_SHADOW, L.ipos + ldir * o.shadowBias, ldir, lmax, o.shadowIter);
          } else {
            L.state = S_LIGHT_POST;
          }
          break;
        }
        case S_LIGHT_POST: {
          const RmMaterial& m = o.mat[L.lmi];
          L.diff = L.diff * m.albedo;  // compounding per light, renderer.cl:376
          L.fin = L.fin + lerp3(L.diff, L.spec, plain::schlick(m.r0, m.smoothness, L.ln, L.lvdir));
          L.li += 1;
          L.state = S_LIGHT_NEXT;
          break;
        }
        case S_SURFACE_DONE: {
          if (L.lsurf == 1) {
            // bounce surface lit: back in the reflection loop of sceneColor (renderer.cl:424-431)
            float3 col = L.fin / (float)o.numLights;
            col = plain::atmosphere(s, L.st, L.bo, L.bd, L.rdist_b, col);
            L.reflAcc = L.reflAcc + col;
            L.bi += 1;
            if (o.mat[plain::mat_index(L.robj_b)].r0 < 0.001f || L.bi >= o.reflectIter)
              lighting_begin(L, s, 0, L.ppos, L.pn, L.pmi, L.rd0, L.reflAcc);
            else
              bounce_begin(L, G);
          } else {
            float3 col = L.fin;
            if (L.lsurf == 0) {
              col = L.fin / (float)o.numLights;
              col = plain::atmosphere(s, L.st, L.st.eye, L.rd0, L.pdist, col);
            }
            col = col * o.exposure;
            if (P.colour) {
              P.colour[L.item] = make_float4(col.x, col.y, col.z, 1.0f);
            } else {
              const float4 old = P.accum[L.id];
              const float3 m = lerp3(f3(old.x, old.y, old.z), col, o.frameBlend);  // mix(), renderer.cl:492
              P.accum[L.id] = make_float4(m.x, m.y, m.z, 1.0f);
            }
            L.state = S_IDLE;
          }
          break;
        }
        default: break;
      }
    }

    // ---- 4. head of the sphere-trace loop (renderer.cl:243-245) ----
    if (L.state == S_TRACE_ITER) {
      if (--L.itersLeft < 0) {
        L.state = S_TRACE_END;
      } else {
        if (kCount) s.w.outer++;
        L.rpos = L.ro + L.mdir * L.tdist;
        job_begin(L, o);
      }
    }

    // ---- 5. the march (renderer.cl:219-234) ----
#pragma unroll 1
    for (int it = 0; it < P.march_quota && L.state == S_MARCH; ++it) {
      // leave early when only a few lanes are still marching together (the others of the warp
      // are waiting to shade / start their next job); a heuristic, never a correctness matter
      if (it >= kMinMarchIters && (int)__popc(__activemask()) < P.min_marchers) break;
      const int x = f2i_sat(L.p.x * G.rxf), y = f2i_sat(L.p.y * G.ryf), z = f2i_sat(L.p.z * G.rzf);
      if (kCount) s.w.steps++;
      if (!in_grid(o, x, y, z)) { L.state = S_MARCH_END; break; }  // voxelLookup < 0 -> break
      const int cs = acc.cell_shift;
      const int d = __ldg(acc.dist + ((size_t)(z >> cs) * acc.my + (y >> cs)) * acc.mx + (x >> cs));
      if (d != 0) {
        // This sample and the next n-1 lie in cells known to hold no solid voxel: advance the
        // recurrence without fetching. Displacement bound: n-1 further steps of at most 1/invS
        // voxels each stay within (d-1) cells; 0.25 voxel of slack covers the rounding drift.
        const float reach = (float)(d - 1) * cellf - 0.25f;
        int n = reach > 0.0f ? 1 + f2i_sat(fminf(reach * L.invS, 1e6f)) : 1;
        n = n < L.rem ? n : L.rem;
        L.rem -= n;
        if (kCount) {
          for (int j = 1; j <= n; ++j) {
            L.p = L.p + L.delta;
            if (j < n) {
              s.w.steps++;
              const int xx = f2i_sat(L.p.x * G.rxf), yy = f2i_sat(L.p.y * G.ryf), zz = f2i_sat(L.p.z * G.rzf);
              if (!in_grid(o, xx, yy, zz)) { L.state = S_MARCH_END; break; }
            }
          }
        } else {
          for (int j = 0; j < n; ++j) L.p = L.p + L.delta;
        }
      } else {
        const uint64_t w = brick_word(acc.solid, acc, x, y, z);
        if ((w >> brick_bit(x, y, z)) & 1ull) {
          L.hit = true;
          L.state = S_MARCH_END;
          break;
        }
        L.p = L.p + L.delta;
        L.rem -= 1;
      }
      if (L.rem <= 0 && L.state == S_MARCH) L.state = S_MARCH_END;
    }

    // ---- 6. distanceToScene result (renderer.cl:223-236) -> consumer ----
    if (L.state == S_MARCH_END) {
      if (L.hit) {
        if (kCount) {
          const int x = f2i_sat(L.p.x * G.rxf), y = f2i_sat(L.p.y * G.ryf), z = f2i_sat(L.p.z * G.rzf);
          s.w.taps += taps_of_hit(G, x, y, z, L.consumer == C_TRACE && L.tkind == T_PRIMARY);
        }
        const float3 hp = L.p * o.voxelBounds2 + (-o.voxelBounds);
        const float dv = len3(L.rpos - hp) - o.voxelSize;
        if (dv < L.hx) { L.hx = dv; L.closer = true; }
      }
      if (L.consumer == C_AO) {
        L.ao *= 1.0f - cl_max((L.aoD - L.hx) * o.aoAmp / L.aoD, 0.0f);  // renderer.cl:343
        L.aoI += 1;
        L.state = S_AO_NEXT;
      } else if (fabsf(L.hx) <= o.eps || L.tdist >= L.tmax) {
        L.state = S_TRACE_END;
      } else {
        L.tdist += L.hx;
        L.state = S_TRACE_ITER;
      }
    }
  }

  if (kCount) {
    atomicAdd(&P.counters->steps, (unsigned long long)s.w.steps);
    atomicAdd(&P.counters->taps, (unsigned long long)s.w.taps);
    atomicAdd(&P.counters->outer, (unsigned long long)s.w.outer);
  }
}

struct BlendWeights { float w[RM_MAX_FUSED_PASSES]; };

// Blend the per-pass colours of one launch into the accumulator in pass order:
// pixels = mix(pixels, colour, frameBlend) per pass (renderer.cl:492).
__global__ void __launch_bounds__(256)
k_blend_passes(const float4* __restrict__ colour, const __grid_constant__ BlendWeights bw, int passes,
               const __grid_constant__ RmShard sh, int W, int H, float4* __restrict__ accum) {
  const long long slot = (long long)blockIdx.x * 256 + threadIdx.x;
  if (slot >= sh.slots) return;
  const int id = rm_slot_to_pixel(sh, slot, W, H);
  if (id < 0) return;
  const float4 old = accum[id];
  float3 p = f3(old.x, old.y, old.z);
  for (int k = 0; k < passes; ++k) {
    const float4 c = __ldcs(colour + (size_t)k * sh.slots + slot);
    p = lerp3(p, f3(c.x, c.y, c.z), bw.w[k]);
  }
  accum[id] = make_float4(p.x, p.y, p.z, 1.0f);
}

}  // namespace

int rm_fast_blocks_per_sm(int count) {
  int n = 0;
  cudaError_t e = count ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_render_fast<true>, kFastBlock, 0)
                        : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_render_fast<false>, kFastBlock, 0);
  return e == cudaSuccess ? n : 0;
}

cudaError_t rm_launch_render_fast(const RmOpts& opts, const RmShard& shard, const RmAccel& accel,
                                  const float4* d_tables, const float* times, const float* blend,
                                  int passes, float4* d_colour, float4* d_accum,
                                  unsigned long long* d_queue, RmCounters* d_counters, int grid_blocks,
                                  int march_quota, int min_marchers, unsigned* d_watchdog, unsigned trip_limit,
                                  cudaStream_t stream) {
  if (shard.slots <= 0 || passes <= 0) return cudaSuccess;
  if (passes > RM_MAX_FUSED_PASSES) return cudaErrorInvalidValue;
  cudaError_t e = cudaMemsetAsync(d_queue, 0, sizeof(unsigned long long), stream);
  if (e != cudaSuccess) return e;
  FastParams P;
  P.tables = d_tables;
  BlendWeights bw;
  for (int i = 0; i < RM_MAX_FUSED_PASSES; ++i) {
    P.times[i] = i < passes ? times[i] : 0.0f;
    bw.w[i] = i < passes ? blend[i] : 0.0f;
  }
  P.colour = passes > 1 ? d_colour : nullptr; P.accum = d_accum;
  P.queue = d_queue; P.counters = d_counters; P.passes = passes;
  P.march_quota = march_quota; P.min_marchers = min_marchers;
  P.watchdog = d_watchdog; P.trip_limit = trip_limit;
  const long long total = (long long)passes * shard.slots;
  long long need = (total + kFastBlock - 1) / kFastBlock;
  const unsigned blocks = (unsigned)(need < grid_blocks ? need : grid_blocks);
  if (d_counters)
    k_render_fast<true><<<blocks, kFastBlock, 0, stream>>>(opts, shard, accel, P);
  else
    k_render_fast<false><<<blocks, kFastBlock, 0, stream>>>(opts, shard, accel, P);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  if (passes > 1) {
    k_blend_passes<<<(unsigned)((shard.slots + 255) / 256), 256, 0, stream>>>(d_colour, bw, passes, shard,
                                                                            opts.width, opts.height, d_accum);
    e = cudaGetLastError();
  }
  return e;
}
