// rm_render_warp.cu -- RenderImage (renderer.cl:478-494) as a warp-scheduled state machine.
//
// Same arithmetic as the per-pixel kernels (rm_scene_plain.cuh; all are checked against the
// oracle), organised so that the lanes of a warp spend their time in the SAME loop:
//
//  * every lane owns one work item (pixel, pass) at a time and walks it through four states:
//      MARCH  inside the fixed-step march of one distanceToScene call (renderer.cl:219-234);
//             primary / bounce / shadow sphere-traces and AO probes all march here
//      JOB    between two marches: finish the distanceToScene call, feed its consumer (sphere
//             trace or AO loop) and set up the next call
//      SHADE  end of a sphere-trace or of the AO loop: normals, materials, lights, atmosphere
//      IDLE   item finished; take the next one from the global queue (persistent lanes)
//  * each trip round the loop the warp votes and runs ONE phase, the one most lanes wait for;
//    lanes in other states wait (they keep their registers), so a phase runs with many lanes
//    instead of every lane dragging the warp through its own private call tree;
//  * the march itself is the fetch-eliding BrickVolume march of rm_scene_plain.cuh, one sample
//    per loop iteration.
//
// Every control decision is derived from full-warp votes, all 32 lanes stay in the loop until
// the warp has no work left, and a trip counter (watchdog) bounds the loop whatever happens.
// Compiled with -fmad=false like the rest of the library (pinned two-rounding evaluation order).
#include "rm_kernels.h"
#include "rm_scene_plain.cuh"

namespace {

using plain::BrickVolume;
using plain::PixelState;
using plain::Scene;

constexpr unsigned kFull = 0xffffffffu;
constexpr int kWarpBlock = 128;

enum : int { S_IDLE = 0, S_MARCH = 1, S_JOB = 2, S_SHADE = 3, S_DONE = 4 };
enum : int { JP_RESULT = 0, JP_ITER = 1 };                                       // JOB sub-state
enum : int { SH_TRACE_END = 0, SH_AO_DONE, SH_LIGHT_NEXT, SH_LIGHT_POST, SH_SURFACE_DONE };  // SHADE sub-state
enum : int { T_PRIMARY = 0, T_BOUNCE = 1, T_SHADOW = 2 };
enum : int { C_TRACE = 0, C_AO = 1 };

struct Lane {
  int state, sub;
  // work item
  int id;
  long long item;
  PixelState st;
  float3 rd0;
  // sphere-trace (raymarch, renderer.cl:239-257)
  int tkind, itersLeft;
  float3 ro;
  float tdist, tmax;
  // distanceToScene call (renderer.cl:209-237)
  int consumer, msteps, rem;
  float3 rpos, mdir, delta, p;
  float g, invS, hx;
  bool hit;     // the call's march stopped on a solid voxel (p = sample position of the hit)
  bool closer;  // ... and the voxel distance won against the ground plane
  // primary surface
  float3 ppos, pn, reflAcc;
  float pdist;
  int pmi, bi;
  // current bounce
  float3 bo, bd, rpos_b, rn_b;
  float rdist_b;
  int robj_b;
  // objectLighting of the current surface (renderer.cl:348-381)
  int lsurf, lmi, li, aoI;
  float3 ipos, ln, lvdir, lrefl, diff, spec, fin;
  float ao, aoD, att;
  uint32_t aoSeed;
};

struct Ctx {
  const RmOpts& o;
  const BrickVolume& V;
  float rxf, ryf, rzf;
};

// distanceToScene head (renderer.cl:209-218): ground plane, slab test, march start
RM_DEV void job_begin(Lane& L, const RmOpts& o) {
  L.g = L.rpos.y + o.groundY;
  const float gx = L.g < 1e5f ? L.g : 1e5f;
  L.hit = false;
  L.closer = false;
  L.hx = gx;
  const bool inside = L.rpos.x > o.boundsMin.x && L.rpos.x < o.boundsMax.x && L.rpos.y > o.boundsMin.y &&
                      L.rpos.y < o.boundsMax.y && L.rpos.z > o.boundsMin.z && L.rpos.z < o.boundsMax.z;
  const float idist = inside ? 0.0f : plain::box_entry(o.boundsMin, o.boundsMax, L.rpos, L.mdir);
  if (idist >= 0.0f && idist < gx && L.msteps > 0) {
    float3 p = L.rpos + o.voxelBounds;
    if (idist > 0.0f) p = L.mdir * idist + p;
    L.p = p * o.invVoxelScale;
    L.rem = L.msteps;
    L.state = S_MARCH;
  } else {
    L.state = S_JOB;
    L.sub = JP_RESULT;
  }
}

RM_DEV void job_direction(Lane& L, const RmOpts& o, float3 dir, int steps) {
  L.mdir = dir;
  L.msteps = steps;
  L.delta = plain::march_delta(o, dir, steps, L.invS);
}

RM_DEV void trace_begin(Lane& L, const RmOpts& o, int kind, float3 ro, float3 rd, float maxDist, int iters) {
  L.tkind = kind;
  L.ro = ro;
  L.tdist = o.startDist;
  L.tmax = maxDist;
  L.itersLeft = iters;
  L.consumer = C_TRACE;
  L.rpos = ro;
  L.g = 0.0f;
  L.hit = false;
  L.closer = false;
  job_direction(L, o, rd, o.maxVoxelIter);
  L.state = S_JOB;
  L.sub = JP_ITER;
}

// ambientOcclusion loop head (renderer.cl:333-336): next probe, or on to the lights
RM_DEV void ao_next(Lane& L, const Scene& s) {
  const RmOpts& o = plain::g_opts;
  if (L.aoI <= o.aoIter && L.ao > 0.01f) {
    L.aoD += o.aoStepDist;
    L.aoSeed += 37u;
    const float3 n = unit3(plain::table_xyz(s, L.aoSeed) * 0.2f + L.ln);
    L.consumer = C_AO;
    job_direction(L, o, n, o.maxVoxelIter / 2);
    L.rpos = n * L.aoD + L.ipos;
    job_begin(L, o);
  } else {
    L.state = S_SHADE;
    L.sub = SH_AO_DONE;
  }
}

RM_DEV void lighting_begin(Lane& L, const Scene& s, int surf, float3 ipos, float3 n, int mi, float3 vdir, float3 refl) {
  L.lsurf = surf; L.ipos = ipos; L.ln = n; L.lmi = mi; L.lvdir = vdir; L.lrefl = refl;
  L.ao = 1.0f; L.aoD = 0.0f; L.aoI = 0;
  L.aoSeed = f2u_wrap(ipos.x * 3183.75f + ipos.y * 1831.42f + ipos.z * 2945.87f + s.time * 2671.918f);
  ao_next(L, s);
}

RM_DEV void bounce_begin(Lane& L, const RmOpts& o) {
  L.bd = plain::reflect3(L.bd, L.rn_b);
  L.bo = L.rpos_b + L.bd * 0.0075f;
  trace_begin(L, o, T_BOUNCE, L.bo, L.bd, o.maxDist, o.maxIter);
}

struct WarpParams {
  const float4* tables;              // passes x 16384 float4
  float times[RM_MAX_FUSED_PASSES];  // TRenderOpts.time per pass
  float4* colour;                    // passes x slots (null when passes == 1: blend straight into accum)
  float4* accum;
  unsigned long long* queue;
  RmCounters* counters;
  unsigned* watchdog;                // 16 words: [0] tripped flag, [1..] state of a lane of the warp that tripped
  unsigned trip_limit;
  int passes;
};

// ---- SHADE phase: one transition of the shading continuation of a lane --------------------
template <bool kCount>
RM_DEV void shade_step(Lane& L, Scene& s, const Ctx& C, const WarpParams& P) {
  const RmOpts& o = C.o;
  switch (L.sub) {
    case SH_TRACE_END: {
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
        L.sub = SH_LIGHT_POST;
        break;
      }
      // object id and normal of the LAST distanceToScene call of the trace
      int objectID = -1;
      float3 normal;
      {
        const int x = f2i_sat(L.p.x * C.rxf), y = f2i_sat(L.p.y * C.ryf), z = f2i_sat(L.p.z * C.rzf);
        if (!miss) {
          if (L.closer) {
            const int v = C.V.value(o, x, y, z);
            objectID = v < 168 ? (v < 84 ? 1 : 2) : 3;  // voxelMaterial, renderer.cl:205-207
          } else {
            objectID = f2i_sat(L.g < 1e5f ? L.g : -1.0f);  // the ground's "id" is its distance (:211)
          }
        }
        if (L.hit) normal = L.tkind == T_PRIMARY ? plain::normal_smooth(C.V, o, x, y, z) : plain::normal_6tap(C.V, o, x, y, z);
        else normal = L.g < 1e5f ? f3(0.0f, 1.0f, 0.0f) : -L.mdir;
      }
      if (L.tkind == T_PRIMARY) {
        // sceneColor (renderer.cl:407-446)
        L.pdist = distance;
        if (distance >= o.maxDist) {
          L.fin = plain::atmosphere(s, L.st, L.st.eye, L.rd0, distance, plain::sky(o, L.rd0));
          L.lsurf = 2;  // no surface: the pixel-sample is finished
          L.sub = SH_SURFACE_DONE;
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
          bounce_begin(L, o);
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
          lighting_begin(L, s, 0, L.ppos, L.pn, L.pmi, L.rd0, L.reflAcc);  // the bounce loop breaks (:428)
        } else {
          lighting_begin(L, s, 1, L.rpos_b, L.rn_b, plain::mat_index(objectID), L.bd,
                         plain::sky(o, plain::reflect3(L.bd, L.rn_b)));
        }
      }
      break;
    }
    case SH_AO_DONE: {
      L.diff = plain::sky(o, L.ln) * L.ao;
      L.spec = L.lrefl * L.ao;
      L.fin = f3s(0.0f);
      L.li = 0;
      L.sub = SH_LIGHT_NEXT;
      break;
    }
    case SH_LIGHT_NEXT: {
      if (L.li >= o.numLights) { L.sub = SH_SURFACE_DONE; break; }
      const float3 dl = plain::light_pos(s, L.st, L.li) - L.ipos;
      const float ld2 = dot3(dl, dl);
      L.att = 1.0f / ld2;
      if (L.att > o.minLightAtt) {
        const float3 ldir = unit3(dl);
        const float lmax = cl_min(sqrtf(ld2) - o.shadowBias, o.maxDist);
        trace_begin(L, o, T_SHADOW, L.ipos + ldir * o.shadowBias, ldir, lmax, o.shadowIter);
      } else {
        L.sub = SH_LIGHT_POST;
      }
      break;
    }
    case SH_LIGHT_POST: {
      const RmMaterial& m = o.mat[L.lmi];
      L.diff = L.diff * m.albedo;  // compounding per light, renderer.cl:376
      L.fin = L.fin + lerp3(L.diff, L.spec, plain::schlick(m.r0, m.smoothness, L.ln, L.lvdir));
      L.li += 1;
      L.sub = SH_LIGHT_NEXT;
      break;
    }
    default: {  // SH_SURFACE_DONE
      if (L.lsurf == 1) {
        // bounce surface lit: back in the reflection loop of sceneColor (renderer.cl:424-431)
        float3 col = L.fin / (float)o.numLights;
        col = plain::atmosphere(s, L.st, L.bo, L.bd, L.rdist_b, col);
        L.reflAcc = L.reflAcc + col;
        L.bi += 1;
        if (o.mat[plain::mat_index(L.robj_b)].r0 < 0.001f || L.bi >= o.reflectIter)
          lighting_begin(L, s, 0, L.ppos, L.pn, L.pmi, L.rd0, L.reflAcc);
        else
          bounce_begin(L, o);
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
  }
}

// ---- JOB phase: finish one distanceToScene call, feed its consumer, set up the next call ----
template <bool kCount>
RM_DEV void job_step(Lane& L, Scene& s, const Ctx& C) {
  const RmOpts& o = C.o;
  if (L.sub == JP_RESULT) {
    // renderer.cl:223-236
    if (L.hit) {
      if (kCount) {
        const int x = f2i_sat(L.p.x * C.rxf), y = f2i_sat(L.p.y * C.ryf), z = f2i_sat(L.p.z * C.rzf);
        s.w.taps += plain::taps_of_hit(C.V, o, x, y, z, L.consumer == C_TRACE && L.tkind == T_PRIMARY);
      }
      const float3 hp = L.p * o.voxelBounds2 + (-o.voxelBounds);
      const float dv = len3(L.rpos - hp) - o.voxelSize;
      if (dv < L.hx) { L.hx = dv; L.closer = true; }
    }
    if (L.consumer == C_AO) {
      L.ao *= 1.0f - cl_max((L.aoD - L.hx) * o.aoAmp / L.aoD, 0.0f);  // renderer.cl:343
      L.aoI += 1;
      ao_next(L, s);
      return;
    }
    if (fabsf(L.hx) <= o.eps || L.tdist >= L.tmax) {  // renderer.cl:249
      L.state = S_SHADE;
      L.sub = SH_TRACE_END;
      return;
    }
    L.tdist += L.hx;
    L.sub = JP_ITER;
  }
  // head of the sphere-trace loop (renderer.cl:243-245)
  if (--L.itersLeft < 0) {
    L.state = S_SHADE;
    L.sub = SH_TRACE_END;
  } else {
    if (kCount) s.w.outer++;
    L.rpos = L.ro + L.mdir * L.tdist;
    job_begin(L, o);
  }
}

template <bool kCount>
__global__ void __launch_bounds__(kWarpBlock)
k_render_warp(const __grid_constant__ RmShard sh, const __grid_constant__ WarpParams P) {
  const RmOpts& o = plain::g_opts;
  const RmAccel& acc = plain::g_accel;
  const unsigned lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  const long long total = (long long)P.passes * sh.slots;
  const BrickVolume V{};
  const Ctx C{o, V, (float)o.rx, (float)o.ry, (float)o.rz};
  Scene s(acc.vox, P.tables);
  Lane L = {};
  L.state = S_IDLE;
  bool exhausted = false;
  unsigned trips = 0;

  for (;;) {
    __syncwarp();
    const unsigned mI = __ballot_sync(kFull, L.state == S_IDLE);
    const unsigned mM = __ballot_sync(kFull, L.state == S_MARCH);
    const unsigned mJ = __ballot_sync(kFull, L.state == S_JOB);
    const unsigned mS = __ballot_sync(kFull, L.state == S_SHADE);
    if ((mI | mM | mJ | mS) == 0u) break;  // every lane is DONE
    if (++trips > P.trip_limit) {
      // watchdog: give up (uniformly) instead of hanging the device; the host reports it
      if (L.state != S_DONE && atomicCAS(P.watchdog, 0u, 1u) == 0u) {
        P.watchdog[1] = (unsigned)L.state; P.watchdog[2] = (unsigned)L.sub; P.watchdog[3] = (unsigned)L.tkind;
        P.watchdog[4] = (unsigned)L.consumer; P.watchdog[5] = (unsigned)L.rem; P.watchdog[6] = (unsigned)L.itersLeft;
        P.watchdog[7] = (unsigned)L.id; P.watchdog[8] = (unsigned)L.item; P.watchdog[9] = mI; P.watchdog[10] = mM;
        P.watchdog[11] = mJ; P.watchdog[12] = mS; P.watchdog[13] = (unsigned)exhausted;
        P.watchdog[14] = blockIdx.x; P.watchdog[15] = threadIdx.x;
      }
      break;
    }
    const int cI = __popc(mI), cM = __popc(mM), cJ = __popc(mJ), cS = __popc(mS);

    if (cM > 0 && cM >= cJ && cM >= cS && cM >= cI) {
      // ---- MARCH phase (renderer.cl:219-234): runs while at least as many lanes march as
      //      wait for a JOB phase
      const int others = cS > cI ? cS : cI;
      for (int it = 0;; ++it) {
        const bool m = L.state == S_MARCH;
        const int marchers = __popc(__ballot_sync(kFull, m));
        const int waiting = __popc(__ballot_sync(kFull, L.state == S_JOB));
        if (marchers == 0 || (it > 0 && (marchers < waiting || marchers < others))) break;
        if (m) {
          const int x = f2i_sat(L.p.x * C.rxf), y = f2i_sat(L.p.y * C.ryf), z = f2i_sat(L.p.z * C.rzf);
          if (kCount) s.w.steps++;
          if (!plain::in_grid(o, x, y, z)) {  // voxelLookup < 0 -> break
            L.state = S_JOB;
            L.sub = JP_RESULT;
          } else {
            const int d = V.cell_dist(x, y, z);
            if (d != 0) {
              // this sample and the next n-1 cannot be solid (see BrickVolume march)
              const float reach = (float)(d - 1) * acc.cellf - 0.25f;
              int n = reach > 0.0f ? 1 + f2i_sat(fminf(reach * L.invS, 1e6f)) : 1;
              n = n < L.rem ? n : L.rem;
              L.rem -= n;
              if (kCount) {
                for (int j = 1; j <= n; ++j) {
                  L.p = L.p + L.delta;
                  if (j < n) {
                    s.w.steps++;
                    if (!plain::in_grid(o, f2i_sat(L.p.x * C.rxf), f2i_sat(L.p.y * C.ryf), f2i_sat(L.p.z * C.rzf))) {
                      L.rem = 0;
                      break;
                    }
                  }
                }
              } else {
                int j = 0;
                for (; j + 4 <= n; j += 4) {
                  L.p = L.p + L.delta; L.p = L.p + L.delta; L.p = L.p + L.delta; L.p = L.p + L.delta;
                }
                for (; j < n; ++j) L.p = L.p + L.delta;
              }
            } else if ((V.word(acc.solid, x, y, z) >> BrickVolume::bit(x, y, z)) & 1ull) {
              L.hit = true;
              L.rem = 0;
            } else {
              L.p = L.p + L.delta;
              L.rem -= 1;
            }
            if (L.rem <= 0) {
              L.state = S_JOB;
              L.sub = JP_RESULT;
            }
          }
        }
      }
    } else if (cJ > 0 && cJ >= cS && cJ >= cI) {
      // ---- JOB phase: up to three call boundaries per lane (ground-only iterations chain here)
#pragma unroll 1
      for (int r = 0; r < 3; ++r) {
        const bool j = L.state == S_JOB;
        if (!__any_sync(kFull, j)) break;
        if (j) job_step<kCount>(L, s, C);
      }
    } else if (cS > 0 && cS >= cI) {
      // ---- SHADE phase: up to four chained transitions per lane
#pragma unroll 1
      for (int r = 0; r < 4; ++r) {
        const bool sh_ = L.state == S_SHADE;
        if (!__any_sync(kFull, sh_)) break;
        if (sh_) shade_step<kCount>(L, s, C, P);
      }
    } else {
      // ---- REFILL phase: the idle lanes take consecutive items with one atomic
      if (exhausted) {
        if (L.state == S_IDLE) L.state = S_DONE;
      } else {
        const int leader = __ffs(mI) - 1;
        unsigned long long base = 0;
        if ((int)lane == leader) base = atomicAdd(P.queue, (unsigned long long)cI);
        base = __shfl_sync(kFull, base, leader);
        if (base + (unsigned long long)cI >= (unsigned long long)total) exhausted = true;
        if (L.state == S_IDLE) {
          L.item = (long long)base + __popc(mI & lt_mask);
          if (L.item >= total) {
            L.state = S_DONE;
          } else {
            // initRenderState + cameraRayLookat (renderer.cl:456-476)
            const long long slot = L.item / P.passes;  // pass-minor order, see rm_render_fast.cu
            const int pass = (int)(L.item - slot * P.passes);
            L.id = rm_slot_to_pixel(sh, slot, o.width, o.height);
            if (L.id >= 0) {  // else: padding slot of an edge tile, stay idle
              s.time = P.times[pass];
              s.table = P.tables + (size_t)pass * (RM_TABLE_MASK + 1);
              L.rd0 = plain::setup_pixel(s, L.id, L.st);
              trace_begin(L, o, T_PRIMARY, L.st.eye, L.rd0, o.maxDist, o.maxIter);
            }
          }
        }
      }
    }
  }

  if (kCount) {
    unsigned long long a = s.w.steps, b = s.w.taps, c = s.w.outer;
    for (int off = 16; off > 0; off >>= 1) {
      a += __shfl_down_sync(kFull, a, off);
      b += __shfl_down_sync(kFull, b, off);
      c += __shfl_down_sync(kFull, c, off);
    }
    if (lane == 0) {
      atomicAdd(&P.counters->steps, a);
      atomicAdd(&P.counters->taps, b);
      atomicAdd(&P.counters->outer, c);
    }
  }
}

}  // namespace

int rm_warp_blocks_per_sm(int count) {
  int n = 0;
  cudaError_t e = count ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_render_warp<true>, kWarpBlock, 0)
                        : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_render_warp<false>, kWarpBlock, 0);
  return e == cudaSuccess ? n : 0;
}

cudaError_t rm_launch_render_warp(const RmOpts& opts, const RmShard& shard, const RmAccel& accel,
                                  const float4* d_tables, const float* times, int passes, float4* d_colour,
                                  float4* d_accum, unsigned long long* d_queue, RmCounters* d_counters,
                                  unsigned* d_watchdog, unsigned trip_limit, int grid_blocks, cudaStream_t stream) {
  if (shard.slots <= 0 || passes <= 0) return cudaSuccess;
  if (passes > RM_MAX_FUSED_PASSES) return cudaErrorInvalidValue;
  cudaError_t e = cudaMemsetAsync(d_queue, 0, sizeof(unsigned long long), stream);
  if (e != cudaSuccess) return e;
  WarpParams P;
  P.tables = d_tables;
  for (int i = 0; i < RM_MAX_FUSED_PASSES; ++i) P.times[i] = i < passes ? times[i] : 0.0f;
  P.colour = passes > 1 ? d_colour : nullptr;
  P.accum = d_accum;
  P.queue = d_queue;
  P.counters = d_counters;
  P.watchdog = d_watchdog;
  P.trip_limit = trip_limit;
  P.passes = passes;
  const long long total = (long long)passes * shard.slots;
  const long long need = (total + kWarpBlock - 1) / kWarpBlock;
  const unsigned blocks = (unsigned)(need < grid_blocks ? need : grid_blocks);
  if ((e = cudaMemcpyToSymbolAsync(plain::g_opts, &opts, sizeof(RmOpts), 0, cudaMemcpyHostToDevice, stream)) != cudaSuccess) return e;
  if ((e = cudaMemcpyToSymbolAsync(plain::g_accel, &accel, sizeof(RmAccel), 0, cudaMemcpyHostToDevice, stream)) != cudaSuccess) return e;
  if (d_counters)
    k_render_warp<true><<<blocks, kWarpBlock, 0, stream>>>(shard, P);
  else
    k_render_warp<false><<<blocks, kWarpBlock, 0, stream>>>(shard, P);
  return cudaGetLastError();
}
