// rm_wave.cuh -- RenderImage (renderer.cl:478-494) as a WAVEFRONT: the same per-pixel-sample
// arithmetic as rm_scene_plain.cuh, cut at its sphere traces into stages that communicate through
// per-item records and a queue of trace jobs, so that the traces -- whose length varies wildly
// between neighbouring shadow / bounce rays (10-14 of 32 lanes busy in the fused
// kernel) -- can be run by a kernel in which a lane that finishes its ray fetches the next one.
//
//   stage            per item (pixel, pass)                                     reference
//   wave_primary     ray setup, primary trace; hit -> record of level 0         sceneColor :407-416
//   wave_prepare(L)  AO of the level-L surface, one shadow job per light that   objectLighting :348-371,
//                    matters, the ray of bounce L+1 (record L+1 + bounce job)   sceneColor :421-432
//   wave_trace       one job: sphere trace; shadow -> lit bit of its record,    raymarch :239-257,
//                    bounce -> position / normal / id of record L+1             shadow :292-301
//   wave_finish(L)   L >= 1: colour of bounce surface L (or sky), atmosphere,   basicSceneColor :383-405
//                    added to the item's reflection colour in bounce order
//   wave_final       primary surface lit with the summed reflection colour,     sceneColor :433-446,
//                    atmosphere, exposure                                       RenderImage :491
//
// Every value is computed by the same expression, in the same order, as in the fused routine: the
// stages only move WHERE the work runs. tests/hostsim runs these stage bodies on the host and
// demands bit-identical accumulators against the oracle.
#pragma once
#include "rm_scene_plain.cuh"

namespace wave {

using namespace plain;

enum : unsigned {
  kLitMask = 0xfu, kValid = 1u << 8, kSurface = 1u << 9,
  // a bounce record as the trace kernel leaves it: nrm = sample position of the last hit, ao = ground
  // distance of the last evaluation, and these bits; resolve_record turns that into normal / objectID
  kRaw = 1u << 10, kRawHit = 1u << 11, kRawCloser = 1u << 12, kRawMiss = 1u << 13
};
constexpr int kMaxLevels = 8;       // primary + up to 7 bounces (TRenderOpts.reflectIter beyond that: fused kernel)
constexpr int kBounceKind = 4;      // job kinds 0..3 = shadow ray of light i

struct WaveRec {  // 64 bytes: the ray of one level of one item and what it found
  float3 pos;  float dist;     // TIsec.pos, TIsec.distance (1000 after a miss)
  float3 nrm;  int obj;        // shading normal; TIsec.objectID
  float3 dir;  float ao;       // the ray's direction; ambient occlusion of the surface
  float3 org;  unsigned flags; // the ray's origin; kValid: a ray exists at this level, kSurface: it was shaded,
                               // bit i: the shadow ray of light i reached the light
};

struct WaveJob {  // 64 bytes
  float3 org;  float maxDist;
  float3 dir;  unsigned info;  // item | level << 24 | kind << 27
  TraceSetup c;                // march step, skip scale, march window: computed where the ray is created
  float pad;
};

struct WaveBuf {
  WaveRec* rec;        // [level][cap]
  float4* refl;        // [cap] summed bounce colours (reflectCol of the primary surface)
  float2* pxy;         // [cap] jittered pixel position (lightPos seed, renderer.cl:267)
  WaveJob* jobs;       // [job_cap]
  unsigned* njobs;     // number of jobs appended since the last reset
  unsigned cap, job_cap;
  int levels;          // record levels allocated (<= kMaxLevels): reflectIter + 2
};

#if defined(__CUDA_ARCH__)
// one atomic per warp: the lanes that append right now (whatever subset of the warp that is) take
// consecutive slots. 80 M appends per C2 frame to ONE counter serialise in L2 otherwise.
RM_DEV unsigned wave_atomic_inc(unsigned* p) {
  const unsigned mask = __activemask();
  const unsigned lane = threadIdx.x & 31;
  const int leader = __ffs(mask) - 1;
  unsigned base = 0;
  if ((int)lane == leader) base = atomicAdd(p, (unsigned)__popc(mask));
  base = __shfl_sync(mask, base, leader);
  return base + __popc(mask & ((1u << lane) - 1u));
}
RM_DEV void wave_atomic_or(unsigned* p, unsigned v) { atomicOr(p, v); }
#else
inline unsigned wave_atomic_inc(unsigned* p) { return __atomic_fetch_add(p, 1u, __ATOMIC_RELAXED); }
inline void wave_atomic_or(unsigned* p, unsigned v) { __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
#endif

RM_DEV WaveRec& rec_at(const WaveBuf& B, int level, unsigned it) { return B.rec[(size_t)level * B.cap + it]; }

template <bool kCount>
RM_DEV void push_job(const WaveBuf& B, float3 org, float3 dir, float maxDist, unsigned it, int level, int kind) {
  const unsigned k = wave_atomic_inc(B.njobs);
  if (k >= B.job_cap) return;  // cannot happen: the launcher sizes the queue for (numLights + 1) jobs per item
  WaveJob j;
  j.org = org; j.maxDist = maxDist; j.dir = dir;
  j.c = trace_setup<kCount>(org, dir, maxDist);
  j.pad = 0.0f;
  j.info = it | ((unsigned)level << 24) | ((unsigned)kind << 27);
  B.jobs[k] = j;
}

// The per-light terms of objectLighting (renderer.cl:356-370), shared by the stage that decides
// which shadow rays to trace and the stage that consumes their results.
struct LightTerms {
  bool on;       // att > minLightAtt: the light is considered at all
  bool traced;   // its shadow ray can change the result (see object_lighting in rm_scene_plain.cuh)
  float3 ldir, inc;
  float lmax, kd, ks;
};

template <bool kCount>
RM_DEV LightTerms light_terms(const Scene& s, const PixelState& st, float3 rd, float3 ipos, const RmMaterial& m, float3 n, int i) {
  const RmOpts& o = g_opts;
  LightTerms t;
  t.on = false; t.traced = false; t.ldir = f3s(0.0f); t.inc = f3s(0.0f); t.lmax = 0.0f; t.kd = 0.0f; t.ks = 0.0f;
  const float3 dl = light_pos(s, st, i) - ipos;
  const float ld2 = dot3(dl, dl);
  const float att = 1.0f / ld2;
  if (att > o.minLightAtt) {
    t.on = true;
    t.ldir = unit3(dl);
    t.lmax = cl_min(sqrtf(ld2) - o.shadowBias, o.maxDist);
    t.kd = cl_max(0.0f, dot3(t.ldir, n));
    t.ks = blinn_phong(m.smoothness, rd, t.ldir, n);
    t.inc = (o.lightColor[i] * 1.0f) * att;
    const float3 zero = t.inc * 0.0f;
    const bool irrelevant = !kCount && t.kd == 0.0f && t.ks == 0.0f && zero.x == 0.0f && zero.y == 0.0f && zero.z == 0.0f;
    t.traced = !irrelevant;
  }
  return t;
}

// objectLighting (renderer.cl:348-381) with the shadow factors read from `flags` instead of traced
template <bool kCount>
RM_DEV float3 light_finish(const Scene& s, const PixelState& st, float3 rd, float3 ipos, const RmMaterial& m, float3 n,
                           float3 reflectCol, float ao, unsigned flags) {
  const RmOpts& o = g_opts;
  float3 diff = sky(o, n) * ao;
  float3 spec = reflectCol * ao;
  float3 fin = f3s(0.0f);
  for (int i = 0; i < o.numLights; ++i) {
    const LightTerms t = light_terms<kCount>(s, st, rd, ipos, m, n, i);
    if (t.on && t.traced && ((flags >> i) & 1u)) {
      diff = diff + t.inc * t.kd;
      spec = spec + t.inc * t.ks;
    }
    diff = diff * m.albedo;
    fin = fin + lerp3(diff, spec, schlick(m.r0, m.smoothness, n, rd));
  }
  return fin / (float)o.numLights;
}

// ---- stage bodies ------------------------------------------------------------------------------

// Stage 1. Returns nothing; record 0 of the item says what happened.
template <bool kCount, class Vol>
RM_DEV void wave_primary(const WaveBuf& B, unsigned it, Scene& s, const Vol& V, int id) {
  const RmOpts& o = g_opts;
  PixelState st;
  const float3 rd = setup_pixel(s, id, st);
  Isec isec;
  sphere_trace<kCount>(s, V, st.eye, rd, isec, o.maxDist, o.maxIter, true, true);
  WaveRec r;
  r.pos = isec.pos; r.dist = isec.distance; r.obj = isec.objectID;
  r.dir = rd; r.org = st.eye; r.ao = 0.0f; r.flags = kValid;
  r.nrm = isec.normal;
  if (isec.distance < o.maxDist) {  // a surface (sceneColor :417-420): the normal it is shaded and mirrored with
    const RmMaterial& m = o.mat[mat_index(isec.objectID)];
    r.nrm = st.mcNormal * (1.0f / (m.smoothness * 200.0f + 5.0f)) + isec.normal;
  }
  rec_at(B, 0, it) = r;
  B.pxy[it] = make_float2(st.px, st.py);
  B.refl[it] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
}

RM_DEV PixelState pixel_state_of(const WaveBuf& B, unsigned it) {
  PixelState st;
  const float2 p = B.pxy[it];
  st.px = p.x; st.py = p.y;
  st.eye = f3s(0.0f); st.mcNormal = f3s(0.0f);  // not read by the later stages
  return st;
}

// normal and objectID of a bounce record the trace kernel has filled (done once, by the first stage that reads it)
template <class Vol>
RM_DEV void resolve_record(WaveRec& r, const Vol& V) {
  if (!(r.flags & kRaw)) return;
  int obj;
  float3 n;
  trace_surface(V, r.dir, (r.flags & kRawMiss) != 0, (r.flags & kRawHit) != 0, (r.flags & kRawCloser) != 0, r.nrm, r.ao, obj, n);
  r.obj = obj;
  r.nrm = n;
  r.ao = 0.0f;
  r.flags &= ~(kRaw | kRawHit | kRawCloser | kRawMiss);
}

// does record L of this item hold a surface that is to be shaded?
RM_DEV bool is_surface(const WaveRec& r, int level) {
  if (!(r.flags & kValid)) return false;
  return level == 0 ? r.dist < g_opts.maxDist : r.obj >= 0;  // sceneColor :416 / basicSceneColor :390
}

// Stage 2 for level L.
template <bool kCount, class Vol>
RM_DEV void wave_prepare(const WaveBuf& B, unsigned it, Scene& s, const Vol& V, int level) {
  const RmOpts& o = g_opts;
  WaveRec& r = rec_at(B, level, it);
  if (level + 1 < B.levels) rec_at(B, level + 1, it).flags = 0;
  resolve_record(r, V);
  if (!is_surface(r, level)) return;
  const PixelState st = pixel_state_of(B, it);
  const RmMaterial& m = o.mat[mat_index(r.obj)];
  RM_STAT_LEVEL(level);
  r.ao = ambient_occlusion<kCount>(s, V, r.pos, r.nrm);
  unsigned flags = r.flags | kSurface;
  r.flags = flags;
  for (int i = 0; i < o.numLights; ++i) {
    const LightTerms t = light_terms<kCount>(s, st, r.dir, r.pos, m, r.nrm, i);
    if (t.on && t.traced) push_job<kCount>(B, r.pos + t.ldir * o.shadowBias, t.ldir, t.lmax, it, level, i);
  }
  // the next bounce (sceneColor :421-432): from the primary surface when it reflects at all, from a
  // bounce surface while the budget lasts and the surface is not (nearly) matt
  const bool more = level == 0 ? (m.r0 > 0.0f && o.reflectIter > 0) : (level < o.reflectIter && !(m.r0 < 0.001f));
  if (more && level + 1 < B.levels) {
    const float3 bd = reflect3(r.dir, r.nrm);
    const float3 bo = r.pos + bd * 0.0075f;
    WaveRec& nx = rec_at(B, level + 1, it);
    nx.dir = bd; nx.org = bo; nx.flags = kValid; nx.ao = 0.0f;
    push_job<kCount>(B, bo, bd, o.maxDist, it, level + 1, kBounceKind);
  }
}

// Stage 3: one job = trace_begin, trace_step until done, job_end. (The persistent kernel runs the
// three pieces itself, one step per trip of its loop.)
RM_DEV void job_begin(const WaveJob& j, TraceState& t) {
  const RmOpts& o = g_opts;
  const bool bounce = (j.info >> 27) == (unsigned)kBounceKind;
  trace_begin(t, j.org, j.dir, j.c, j.maxDist, bounce ? o.maxIter : o.shadowIter, bounce);
}

template <bool kCount, class Vol>
RM_DEV void job_end(const WaveBuf& B, unsigned info, Scene& s, const Vol& V, TraceState& t) {
  const unsigned it = info & 0xffffffu;
  const int level = (int)((info >> 24) & 7u), kind = (int)(info >> 27);
  const bool miss = trace_finish<kCount>(s, V, t);
  WaveRec& w = rec_at(B, level, it);
  if (kind == kBounceKind) {
    w.pos = t.pos; w.dist = t.dist; w.nrm = t.j.p; w.ao = t.j.g; w.obj = -1;
    w.flags = kValid | kRaw | (t.j.hit ? kRawHit : 0u) | (t.j.closer ? kRawCloser : 0u) | (miss ? kRawMiss : 0u);
  } else if (!(t.dist < t.maxDist)) {
    wave_atomic_or(&w.flags, 1u << kind);  // shadow() :292-301: the ray reached the light
  }
}

template <bool kCount, class Vol>
RM_DEV void wave_trace(const WaveBuf& B, const WaveJob& j, Scene& s, const Vol& V) {
  TraceState t;
  RM_STAT_SITE((int)((j.info >> 24) & 7u) * 16 + ((j.info >> 27) == (unsigned)kBounceKind ? 0 : 1 + (int)(j.info >> 27)));
  job_begin(j, t);
  for (;;) {
    const int st = trace_run_cheap<kCount>(s, t, 8);
    if (st == kTraceDone) break;
    if (st == kTraceNeedsFull && trace_full<kCount>(s, V, t)) break;
  }
  job_end<kCount>(B, j.info, s, V, t);
}

// Stage 4 for level L >= 1 (basicSceneColor :383-405 after its trace).
template <bool kCount, class Vol>
RM_DEV void wave_finish(const WaveBuf& B, unsigned it, const Scene& s, const Vol& V, int level) {
  const RmOpts& o = g_opts;
  WaveRec& r = rec_at(B, level, it);
  if (!(r.flags & kValid)) return;
  resolve_record(r, V);  // (a no-op when wave_prepare(level) has run, i.e. always but for a level beyond the last prepare)
  const PixelState st = pixel_state_of(B, it);
  float3 col;
  if (r.obj < 0) {
    col = sky(o, r.dir);
  } else {
    col = light_finish<kCount>(s, st, r.dir, r.pos, o.mat[mat_index(r.obj)], r.nrm, sky(o, reflect3(r.dir, r.nrm)), r.ao, r.flags);
  }
  col = atmosphere(s, st, r.org, r.dir, r.dist, col);
  const float4 a = B.refl[it];
  const float3 sum = f3(a.x, a.y, a.z) + col;  // reflectCol += basicSceneColor(...), in bounce order
  B.refl[it] = make_float4(sum.x, sum.y, sum.z, 0.0f);
}

// Stage 5: sceneColor :433-446 and the exposure of RenderImage :491.
template <bool kCount>
RM_DEV float3 wave_final(const WaveBuf& B, unsigned it, const Scene& s) {
  const RmOpts& o = g_opts;
  const WaveRec& r = rec_at(B, 0, it);
  const PixelState st = pixel_state_of(B, it);
  float3 col;
  if (r.dist >= o.maxDist) {
    col = sky(o, r.dir);
  } else {
    const RmMaterial& m = o.mat[mat_index(r.obj)];
    float3 reflectCol;
    if (m.r0 > 0.0f && o.reflectIter > 0) {
      const float4 a = B.refl[it];
      reflectCol = f3(a.x, a.y, a.z);
    } else {
      reflectCol = sky(o, reflect3(r.dir, r.nrm));
    }
    col = light_finish<kCount>(s, st, r.dir, r.pos, m, r.nrm, reflectCol, r.ao, r.flags);
  }
  return atmosphere(s, st, r.org, r.dir, r.dist, col) * o.exposure;
}

}  // namespace wave
