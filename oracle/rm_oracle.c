/* rm_oracle.c -- TEST INFRASTRUCTURE ONLY (oracle/). Never linked, imported or executed by the
 * product path (raymarchcl_b200/); only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may use it, and only as the checker.
 *
 * Plain-C restatement of the voxel ray-march render op of thi-ng/raymarchcl
 * (/root/reference/resources/renderer.cl, kernels RenderImage :478-494 and TonemapImage :496-508),
 * written from the algorithm, scalar, with the evaluation order of every fp32 expression pinned
 * to the source order of the reference and NO fused multiply-add (build with -ffp-contract=off).
 * Each function cites the reference lines it follows.
 *
 * Parity pinning: the reference ships no tests or golden vectors, so this file
 * is pinned against the reference's own kernel text compiled through oracle/clshim.h
 * (oracle/_ref/libref_strict.so, built by oracle/build_ref.py): tests/test_oracle_vs_ref.py
 * demands bit-identical accumulators, ARGB words and work counters, and tests/golden/ holds
 * vectors generated from that build (tests/golden/make_golden.py) for machines without
 * /root/reference.
 *
 * Semantics the reference leaves undefined, pinned here:
 *   (1) float->uint of a negative value wraps: (uint32)(int64)trunc(x)         (:267, :334)
 *   (2) normalize(0) = 0                                                       (:202, :228)
 *   (3) float->int conversion of voxel coordinates truncates toward zero, saturates, NaN->0 (:165)
 *   (4) the slab test uses IEEE division (+-inf for zero direction components)   (:154-155)
 *   (5) mad(a,b,c) is a*b then +c (two roundings); expressions associate left to right
 *   (6) min(x,y)=y<x?y:x, max(x,y)=x<y?y:x, step(e,x)=x<e?0:1, mix(a,b,t)=a+(b-a)t
 *   (7) material index clamped to 0..3 (the reference would read out of bounds)
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <omp.h>

typedef struct { float x, y, z; } v3;

typedef struct { v3 albedo; float r0, smoothness; } material_t;

/* TRenderOpts (renderer.cl:35-78) decoded from the 544-byte blob; offsets per OpenCL layout rules
 * (float3/int3 = 16 bytes), cross-checked against the compiled reference in the tests. */
typedef struct {
  v3 eyePos, targetPos, up, voxelBounds, voxelBounds2, voxelBoundsMin, voxelBoundsMax,
      invVoxelScale, skyColor1, skyColor2;
  int rx, ry, rz, rxy;
  int width, height;
  float invAspect, time, fov;
  int maxIter, maxVoxelIter;
  float maxDist, startDist, eps;
  int aoIter;
  float aoStepDist, aoAmp, voxelSize, groundY;
  int shadowIter, reflectIter;
  float shadowBias, lightScatter, minLightAtt, gamma, exposure, dof, frameBlend, fogPow, flareAmp;
  int isoVal, numLights;
  v3 lightPos[4], lightColor[4];
  material_t materials[4];
} opts_t;

#define RM_OPTS_BYTES 544
#define RM_TABLE_MASK 0x3fffu /* renderer.cl:143 */

typedef struct { uint64_t step, tap, outer; } counters_t;

typedef struct {
  const uint8_t* vox;
  const float* mc; /* 16384 x float4 */
  opts_t o;
  counters_t c;
} scene_t;

static float rd_f(const uint8_t* b, int off) { float f; memcpy(&f, b + off, 4); return f; }
static int rd_i(const uint8_t* b, int off) { int32_t i; memcpy(&i, b + off, 4); return (int)i; }
static v3 rd_v3(const uint8_t* b, int off) { v3 r = { rd_f(b, off), rd_f(b, off + 4), rd_f(b, off + 8) }; return r; }

static void decode_opts(const void* blob, opts_t* o) {
  const uint8_t* b = (const uint8_t*)blob;
  int i;
  o->eyePos = rd_v3(b, 0);          o->targetPos = rd_v3(b, 16);      o->up = rd_v3(b, 32);
  o->voxelBounds = rd_v3(b, 48);    o->voxelBounds2 = rd_v3(b, 64);   o->voxelBoundsMin = rd_v3(b, 80);
  o->voxelBoundsMax = rd_v3(b, 96); o->invVoxelScale = rd_v3(b, 112); o->skyColor1 = rd_v3(b, 128);
  o->skyColor2 = rd_v3(b, 144);
  o->rx = rd_i(b, 160); o->ry = rd_i(b, 164); o->rz = rd_i(b, 168); o->rxy = rd_i(b, 172);
  o->width = rd_i(b, 176); o->height = rd_i(b, 180);
  o->invAspect = rd_f(b, 184); o->time = rd_f(b, 188); o->fov = rd_f(b, 192);
  o->maxIter = rd_i(b, 196); o->maxVoxelIter = rd_i(b, 200);
  o->maxDist = rd_f(b, 204); o->startDist = rd_f(b, 208); o->eps = rd_f(b, 212);
  o->aoIter = rd_i(b, 216);
  o->aoStepDist = rd_f(b, 220); o->aoAmp = rd_f(b, 224); o->voxelSize = rd_f(b, 228);
  o->groundY = rd_f(b, 232);
  o->shadowIter = rd_i(b, 236); o->reflectIter = rd_i(b, 240);
  o->shadowBias = rd_f(b, 244); o->lightScatter = rd_f(b, 248); o->minLightAtt = rd_f(b, 252);
  o->gamma = rd_f(b, 256); o->exposure = rd_f(b, 260); o->dof = rd_f(b, 264);
  o->frameBlend = rd_f(b, 268); o->fogPow = rd_f(b, 272); o->flareAmp = rd_f(b, 276);
  /* 280: mcTableLength, never read by the kernel */
  o->isoVal = b[284]; o->numLights = b[285];
  for (i = 0; i < 4; ++i) {
    o->lightPos[i] = rd_v3(b, 288 + 16 * i);
    o->lightColor[i] = rd_v3(b, 352 + 16 * i);
    o->materials[i].albedo = rd_v3(b, 416 + 32 * i);
    o->materials[i].r0 = rd_f(b, 416 + 32 * i + 16);
    o->materials[i].smoothness = rd_f(b, 416 + 32 * i + 20);
  }
}

/* ---- fp32 helpers with pinned evaluation order ---- */
static v3 V(float x, float y, float z) { v3 r = { x, y, z }; return r; }
static v3 add(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static v3 sub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static v3 mul(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
static v3 scale(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
static v3 divs(v3 a, float s) { return V(a.x / s, a.y / s, a.z / s); }
static v3 neg(v3 a) { return V(-a.x, -a.y, -a.z); }
static float fmin_cl(float x, float y) { return y < x ? y : x; }
static float fmax_cl(float x, float y) { return x < y ? y : x; }
static float clamp_cl(float x, float lo, float hi) { return fmin_cl(fmax_cl(x, lo), hi); }
static float dot3(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static v3 cross3(v3 a, v3 b) {
  return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static float len3(v3 a) { return sqrtf(dot3(a, a)); }
static v3 unit3(v3 a) { const float l = len3(a); return l == 0.0f ? a : divs(a, l); }
static v3 lerp3(v3 a, v3 b, float t) { return add(a, scale(sub(b, a), t)); }
static int f2i_sat(float f) {
  if (f != f) return 0;
  if (f >= 2147483648.0f) return 2147483647;
  if (f <= -2147483648.0f) return -2147483647 - 1;
  return (int)f;
}
static uint32_t f2u_wrap(float f) { return (uint32_t)(int64_t)f; }
static int clamp_mat(int id) { return id < 0 ? 0 : (id > 3 ? 3 : id); }

/* renderer.cl:142-144 randFloat4 */
static const float* table_at(const scene_t* s, uint32_t seed) { return s->mc + 4u * (seed & RM_TABLE_MASK); }
static v3 table_xyz(const scene_t* s, uint32_t seed) { const float* t = table_at(s, seed); return V(t[0], t[1], t[2]); }

/* renderer.cl:153-161 intersectsBox: slab test, entry distance clamped to >= 0, or -1 */
static float box_entry(v3 bmin, v3 bmax, v3 p, v3 dir) {
  const v3 omin = V((bmin.x - p.x) / dir.x, (bmin.y - p.y) / dir.y, (bmin.z - p.z) / dir.z);
  const v3 omax = V((bmax.x - p.x) / dir.x, (bmax.y - p.y) / dir.y, (bmax.z - p.z) / dir.z);
  const v3 lo = V(fmin_cl(omax.x, omin.x), fmin_cl(omax.y, omin.y), fmin_cl(omax.z, omin.z));
  const v3 hi = V(fmax_cl(omax.x, omin.x), fmax_cl(omax.y, omin.y), fmax_cl(omax.z, omin.z));
  const float a = fmax_cl(fmax_cl(lo.x, 0.0f), fmax_cl(lo.y, lo.z));
  const float b = fmin_cl(hi.x, fmin_cl(hi.y, hi.z));
  return b > a ? a : -1.0f;
}

static int in_grid(const opts_t* o, int x, int y, int z) {
  return z >= 0 && z < o->rz && y >= 0 && y < o->ry && x >= 0 && x < o->rx;
}

/* renderer.cl:163-170 voxelLookup: nearest-voxel point sample of the unit-cube position p */
static int voxel_at(const scene_t* s, v3 p, int* qx, int* qy, int* qz) {
  const opts_t* o = &s->o;
  const int x = f2i_sat(p.x * (float)o->rx), y = f2i_sat(p.y * (float)o->ry), z = f2i_sat(p.z * (float)o->rz);
  *qx = x; *qy = y; *qz = z;
  if (in_grid(o, x, y, z)) return (int)s->vox[(int64_t)z * o->rxy + (int64_t)y * o->rx + x];
  return -1;
}

/* renderer.cl:172-178 voxelLookupI: occupancy with >= isoVal, 0 outside */
static float occ(scene_t* s, int x, int y, int z) {
  const opts_t* o = &s->o;
  s->c.tap++;
  if (in_grid(o, x, y, z)) {
    const float v = (float)s->vox[(int64_t)z * o->rxy + (int64_t)y * o->rx + x];
    return v < (float)o->isoVal ? 0.0f : 1.0f;
  }
  return 0.0f;
}

/* renderer.cl:180-188 voxelNormal: negated central differences of occupancy (un-normalised) */
static v3 grad6(scene_t* s, int x, int y, int z) {
  const float nx = occ(s, x + 1, y, z) - occ(s, x - 1, y, z);
  const float ny = occ(s, x, y + 1, z) - occ(s, x, y - 1, z);
  const float nz = occ(s, x, y, z + 1) - occ(s, x, y, z - 1);
  return V(-nx, -ny, -nz);
}

/* renderer.cl:190-203 voxelNormalSmooth: sum of grad6 over occupied 3x3x3 neighbours, normalised */
static v3 grad27(scene_t* s, int x, int y, int z) {
  v3 n = V(0.0f, 0.0f, 0.0f);
  int dx, dy, dz;
  for (dz = -1; dz <= 1; ++dz)
    for (dy = -1; dy <= 1; ++dy)
      for (dx = -1; dx <= 1; ++dx)
        if (occ(s, x + dx, y + dy, z + dz) > 0.0f) n = add(n, grad6(s, x + dx, y + dy, z + dz));
  return unit3(n);
}

/* renderer.cl:205-207 voxelMaterial */
static float band(int v) { return v < 168 ? (v < 84 ? 1.0f : 2.0f) : 3.0f; }

typedef struct { float dist, id; } hit_t;

/* renderer.cl:209-237 distanceToScene: ground plane, then a fixed-step point-sampled march
 * through the voxel box. *normal is overwritten exactly where the reference writes isec->normal. */
static hit_t scene_distance(scene_t* s, v3 rpos, v3 dir, int steps, int smooth, v3* normal) {
  const opts_t* o = &s->o;
  const float g = rpos.y + o->groundY;
  hit_t res;
  float idist;
  if (g < 1e5f) { res.dist = g; res.id = g; } else { res.dist = 1e5f; res.id = -1.0f; }
  *normal = (res.dist < 1e5f) ? V(0.0f, 1.0f, 0.0f) : neg(dir);
  idist = box_entry(o->voxelBoundsMin, o->voxelBoundsMax, rpos, dir);
  if (idist >= 0.0f && idist < res.dist) {
    const float sf = (float)steps * 0.5f;
    const v3 delta = mul(divs(dir, sf), o->invVoxelScale);
    v3 p = add(rpos, o->voxelBounds);
    if (idist > 0.0f) p = add(scale(dir, idist), p);
    p = mul(p, o->invVoxelScale);
    while (--steps >= 0) {
      int qx, qy, qz;
      const int v = voxel_at(s, p, &qx, &qy, &qz);
      s->c.step++;
      if (v < 0) break;
      if (v > o->isoVal) {
        hit_t h;
        v3 hp;
        *normal = smooth ? grad27(s, qx, qy, qz) : unit3(grad6(s, qx, qy, qz));
        hp = add(mul(p, o->voxelBounds2), neg(o->voxelBounds));
        h.dist = len3(sub(rpos, hp)) - o->voxelSize;
        h.id = band(v);
        return h.dist < res.dist ? h : res;
      }
      p = add(p, delta);
    }
  }
  return res;
}

typedef struct { v3 pos, normal; float distance; int objectID; } isec_t;

/* renderer.cl:239-257 raymarch: sphere trace driven by scene_distance */
static void sphere_trace(scene_t* s, v3 rpos, v3 rdir, isec_t* r, float maxDist, int maxSteps, int smooth) {
  const opts_t* o = &s->o;
  r->distance = o->startDist;
  while (--maxSteps >= 0) {
    hit_t h;
    s->c.outer++;
    r->pos = add(rpos, scale(rdir, r->distance));
    h = scene_distance(s, r->pos, rdir, o->maxVoxelIter, smooth, &r->normal);
    r->objectID = f2i_sat(h.id);
    if (fabsf(h.dist) <= o->eps || r->distance >= maxDist) break;
    r->distance += h.dist;
  }
  if (r->distance >= maxDist) {
    r->pos = add(rpos, scale(rdir, r->distance));
    r->objectID = -1;
    r->distance = 1000.0f;
  }
}

/* renderer.cl:259-261 skyGradient */
static v3 sky(const opts_t* o, v3 dir) { return lerp3(o->skyColor1, o->skyColor2, dir.y * 0.5f + 0.5f); }

typedef struct { v3 eyePos; float mcPos[4]; v3 mcNormal; float px, py; } pixstate_t;

/* renderer.cl:263-269 lightPos: jittered light position (same jitter for all lights of a pixel) */
static v3 light_pos(const scene_t* s, const pixstate_t* st, int i) {
  const opts_t* o = &s->o;
  const uint32_t seed = f2u_wrap(st->px * 1957.0f + st->py * 2173.0f + o->time * 4763.742f);
  return add(scale(table_xyz(s, seed), o->lightScatter), o->lightPos[i]);
}

/* renderer.cl:271-273 reflect */
static v3 reflect3(v3 v, v3 n) { return sub(v, scale(n, 2.0f * dot3(v, n))); }

/* renderer.cl:275-290 applyAtmosphere: distance fog toward the sky colour + light flares */
static v3 atmosphere(const scene_t* s, const pixstate_t* st, v3 rpos, v3 rdir, float distance, v3 col) {
  const opts_t* o = &s->o;
  const float fa = 1.0f - expf(distance * distance * -o->fogPow);
  int i;
  col = add(scale(sub(sky(o, rdir), col), fa), col);
  for (i = 0; i < o->numLights; ++i) {
    v3 lp = light_pos(s, st, i);
    const float d = clamp_cl(dot3(sub(lp, rpos), rdir), 0.0f, distance);
    lp = add(scale(rdir, d), sub(rpos, lp));
    col = add(scale(o->lightColor[i], o->flareAmp / dot3(lp, lp)), col);
  }
  return col;
}

/* renderer.cl:292-301 shadow */
static float shadow_factor(scene_t* s, v3 p, v3 ldir, float lightMaxDist) {
  isec_t si;
  sphere_trace(s, p, ldir, &si, lightMaxDist, s->o.shadowIter, 0);
  return si.distance < lightMaxDist ? 0.0f : 1.0f;
}

/* renderer.cl:304-311 schlick */
static float schlick(float r0, float smooth, v3 normal, v3 view) {
  const float d = clamp_cl(1.0f - dot3(normal, neg(view)), 0.0f, 1.0f);
  if (d > 0.0f) {
    const float d2 = d * d;
    return (1.0f - r0) * (smooth * d2 * d2 * d) + r0;
  }
  return 0.0f;
}

/* renderer.cl:317-325 blinnPhongIntensity */
static float blinn_phong(float smooth, v3 rdir, v3 lightDir, v3 normal) {
  const float nh = dot3(unit3(sub(lightDir, rdir)), normal);
  if (nh > 0.0f) {
    const float specPow = exp2f(6.0f * smooth + 4.0f);
    return powf(nh, specPow) * (specPow + 2.0f) * 0.125f;
  }
  return 0.0f;
}

/* renderer.cl:327-346 ambientOcclusion: aoIter+1 randomised half-length probes */
static float ambient_occlusion(scene_t* s, v3 pos, v3 normal) {
  const opts_t* o = &s->o;
  float ao = 1.0f, d = 0.0f;
  uint32_t seed = f2u_wrap(pos.x * 3183.75f + pos.y * 1831.42f + pos.z * 2945.87f + o->time * 2671.918f);
  int i;
  for (i = 0; i <= o->aoIter && ao > 0.01f; ++i) {
    v3 n, ignored;
    hit_t h;
    d += o->aoStepDist;
    seed += 37u;
    n = unit3(add(scale(table_xyz(s, seed), 0.2f), normal));
    h = scene_distance(s, add(scale(n, d), pos), n, o->maxVoxelIter / 2, 0, &ignored);
    ao *= 1.0f - fmax_cl((d - h.dist) * o->aoAmp / d, 0.0f);
  }
  return ao;
}

/* renderer.cl:348-381 objectLighting */
static v3 object_lighting(scene_t* s, const pixstate_t* st, v3 rdir, v3 ipos, const material_t* mat,
                          v3 normal, v3 reflectCol) {
  const opts_t* o = &s->o;
  const float ao = ambient_occlusion(s, ipos, normal);
  v3 diff = scale(sky(o, normal), ao);
  v3 spec = scale(reflectCol, ao);
  v3 fin = V(0.0f, 0.0f, 0.0f);
  int i;
  for (i = 0; i < o->numLights; ++i) {
    const v3 dl = sub(light_pos(s, st, i), ipos);
    const float lightDist = dot3(dl, dl);
    const float att = 1.0f / lightDist;
    if (att > o->minLightAtt) {
      const v3 ldir = unit3(dl);
      const float sf = shadow_factor(s, add(ipos, scale(ldir, o->shadowBias)), ldir,
                                     fmin_cl(sqrtf(lightDist) - o->shadowBias, o->maxDist));
      if (sf > 0.0f) {
        const v3 incident = scale(scale(o->lightColor[i], sf), att);
        diff = add(diff, scale(incident, fmax_cl(0.0f, dot3(ldir, normal)))); /* :313-315 */
        spec = add(spec, scale(incident, blinn_phong(mat->smoothness, rdir, ldir, normal)));
      }
    }
    diff = mul(diff, mat->albedo); /* compounding per light, as in the reference (:376) */
    fin = add(fin, lerp3(diff, spec, schlick(mat->r0, mat->smoothness, normal, rdir)));
  }
  return divs(fin, (float)o->numLights);
}

/* renderer.cl:383-405 basicSceneColor: one reflection bounce */
static v3 bounce_color(scene_t* s, const pixstate_t* st, v3 rpos, v3 rdir, isec_t* isec) {
  const opts_t* o = &s->o;
  v3 col;
  sphere_trace(s, rpos, rdir, isec, o->maxDist, o->maxIter, 0);
  if (isec->objectID < 0) {
    col = sky(o, rdir);
  } else {
    const material_t* mat = &o->materials[clamp_mat(isec->objectID)];
    col = object_lighting(s, st, rdir, isec->pos, mat, isec->normal, sky(o, reflect3(rdir, isec->normal)));
  }
  return atmosphere(s, st, rpos, rdir, isec->distance, col);
}

/* renderer.cl:407-446 sceneColor */
static v3 scene_color(scene_t* s, const pixstate_t* st, v3 rpos, v3 rdir) {
  const opts_t* o = &s->o;
  isec_t isec;
  v3 col;
  sphere_trace(s, rpos, rdir, &isec, o->maxDist, o->maxIter, 1);
  if (isec.distance >= o->maxDist) {
    col = sky(o, rdir);
  } else {
    const material_t* mat = &o->materials[clamp_mat(isec.objectID)];
    const v3 norm = add(scale(st->mcNormal, 1.0f / (mat->smoothness * 200.0f + 5.0f)), isec.normal);
    v3 reflectCol = V(0.0f, 0.0f, 0.0f);
    if (mat->r0 > 0.0f && o->reflectIter > 0) {
      isec_t ri;
      v3 bdir = rdir;
      int i;
      ri.pos = isec.pos;
      ri.normal = norm;
      for (i = 0; i < o->reflectIter; ++i) {
        v3 bpos;
        bdir = reflect3(bdir, ri.normal);
        bpos = add(ri.pos, scale(bdir, 0.0075f));
        reflectCol = add(reflectCol, bounce_color(s, st, bpos, bdir, &ri));
        if (ri.objectID < 0) break;
        if (o->materials[clamp_mat(ri.objectID)].r0 < 0.001f) break;
      }
    } else {
      reflectCol = sky(o, reflect3(rdir, norm));
    }
    col = object_lighting(s, st, rdir, isec.pos, mat, norm, reflectCol);
  }
  return atmosphere(s, st, rpos, rdir, isec.distance, col);
}

/* renderer.cl:467-476 initRenderState */
static void init_pixel(const scene_t* s, int id, pixstate_t* st) {
  const opts_t* o = &s->o;
  const float* a = table_at(s, (uint32_t)(id * 17) + f2u_wrap(o->time * 3141.3862f));
  const v3 nrm = unit3(table_xyz(s, (uint32_t)(id * 37) + f2u_wrap(o->time * 1859.1467f)));
  memcpy(st->mcPos, a, 16);
  st->mcNormal = nrm;
  st->px = (float)(id % o->width) + st->mcPos[2];
  st->py = (float)(id / o->width) + st->mcPos[3];
  st->eyePos = add(scale(V(nrm.z, nrm.x, nrm.y), o->dof), o->eyePos);
}

/* renderer.cl:456-465 cameraRayLookat */
static v3 camera_dir(const opts_t* o, const pixstate_t* st) {
  const v3 fwd = unit3(sub(o->targetPos, st->eyePos));
  const v3 right = unit3(cross3(fwd, o->up));
  const float vx = st->px / (float)o->width * o->fov - o->fov * 0.5f;
  float vy = st->py / (float)o->height * o->fov - o->fov * 0.5f;
  vy = vy * -o->invAspect;
  return unit3(add(add(scale(right, vx), scale(cross3(right, fwd), vy)), fwd));
}

/* renderer.cl:478-494 RenderImage, one work-item */
static void render_pixel(scene_t* s, float* pixels, int id) {
  const opts_t* o = &s->o;
  pixstate_t st;
  v3 rdir, col, old;
  float* px = pixels + 4 * (int64_t)id;
  init_pixel(s, id, &st);
  rdir = camera_dir(o, &st);
  col = scale(scene_color(s, &st, st.eyePos, rdir), o->exposure);
  old = V(px[0], px[1], px[2]);
  col = lerp3(old, col, o->frameBlend);
  px[0] = col.x; px[1] = col.y; px[2] = col.z; px[3] = 1.0f;
}

/* ================================ C entry points ================================ */

int orc_sizeof_opts(void) { return RM_OPTS_BYTES; }
int orc_has_counters(void) { return 1; }
int orc_num_threads(void) { return omp_get_max_threads(); }
void orc_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }

void orc_render_pixels(const uint8_t* voxels, const float* mc, const void* opts, float* pixels,
                       int n, const int* ids, int n_ids, uint64_t* counters) {
  const int count = ids ? n_ids : n;
  uint64_t cs = 0, ct = 0, co = 0;
#pragma omp parallel reduction(+ : cs, ct, co)
  {
    scene_t s;
    int k;
    s.vox = voxels; s.mc = mc;
    decode_opts(opts, &s.o);
    s.c.step = s.c.tap = s.c.outer = 0;
#pragma omp for schedule(dynamic, 64)
    for (k = 0; k < count; ++k) {
      const int id = ids ? ids[k] : k;
      if (id >= 0 && id < n) render_pixel(&s, pixels, id);
    }
    cs += s.c.step; ct += s.c.tap; co += s.c.outer;
  }
  if (counters) { counters[0] += cs; counters[1] += ct; counters[2] += co; }
}

/* renderer.cl:496-508 TonemapImage (+ :448-454 gamma/tonemap) */
void orc_tonemap(const float* pixels, const void* opts, uint32_t* argb, int n) {
  opts_t o;
  int k;
  decode_opts(opts, &o);
#pragma omp parallel for schedule(static)
  for (k = 0; k < n; ++k) {
    const float* px = pixels + 4 * (int64_t)k;
    uint32_t ch[3];
    int c;
    for (c = 0; c < 3; ++c) {
      float t = px[c] / (o.gamma + px[c]);
      t = t * t * 255.0f;
      t = clamp_cl(t, 0.0f, 255.0f);
      ch[c] = (uint32_t)f2i_sat(t);
    }
    argb[k] = 0xff000000u | (ch[0] << 16) | (ch[1] << 8) | ch[2];
  }
}

float orc_intersects_box(const float* bmin, const float* bmax, const float* p, const float* dir) {
  return box_entry(V(bmin[0], bmin[1], bmin[2]), V(bmax[0], bmax[1], bmax[2]), V(p[0], p[1], p[2]),
                   V(dir[0], dir[1], dir[2]));
}

int orc_voxel_lookup(const uint8_t* voxels, const void* opts, const float* p) {
  scene_t s;
  int x, y, z;
  s.vox = voxels; s.mc = 0;
  decode_opts(opts, &s.o);
  return voxel_at(&s, V(p[0], p[1], p[2]), &x, &y, &z);
}

void orc_voxel_normal(const uint8_t* voxels, const void* opts, const int* q, int smooth, float* out) {
  scene_t s;
  v3 n;
  s.vox = voxels; s.mc = 0;
  decode_opts(opts, &s.o);
  memset(&s.c, 0, sizeof(s.c));
  n = smooth ? grad27(&s, q[0], q[1], q[2]) : grad6(&s, q[0], q[1], q[2]);
  out[0] = n.x; out[1] = n.y; out[2] = n.z;
}

void orc_distance_to_scene(const uint8_t* voxels, const void* opts, const float* rpos,
                           const float* dir, int steps, int smooth, float* out) {
  scene_t s;
  v3 n = V(0.0f, 0.0f, 0.0f);
  hit_t h;
  s.vox = voxels; s.mc = 0;
  decode_opts(opts, &s.o);
  memset(&s.c, 0, sizeof(s.c));
  h = scene_distance(&s, V(rpos[0], rpos[1], rpos[2]), V(dir[0], dir[1], dir[2]), steps, smooth, &n);
  out[0] = h.dist; out[1] = h.id; out[2] = n.x; out[3] = n.y; out[4] = n.z;
}

void orc_raymarch(const uint8_t* voxels, const void* opts, const float* pos, const float* dir,
                  float max_dist, int max_steps, int smooth, float* out) {
  scene_t s;
  isec_t r;
  s.vox = voxels; s.mc = 0;
  decode_opts(opts, &s.o);
  memset(&s.c, 0, sizeof(s.c));
  memset(&r, 0, sizeof(r));
  sphere_trace(&s, V(pos[0], pos[1], pos[2]), V(dir[0], dir[1], dir[2]), &r, max_dist, max_steps, smooth);
  out[0] = r.pos.x; out[1] = r.pos.y; out[2] = r.pos.z;
  out[3] = r.normal.x; out[4] = r.normal.y; out[5] = r.normal.z;
  out[6] = r.distance; out[7] = (float)r.objectID;
}

void orc_camera_ray(const void* opts, const float* mc, int id, float* out) {
  scene_t s;
  pixstate_t st;
  v3 d;
  s.vox = 0; s.mc = mc;
  decode_opts(opts, &s.o);
  init_pixel(&s, id, &st);
  d = camera_dir(&s.o, &st);
  out[0] = st.eyePos.x; out[1] = st.eyePos.y; out[2] = st.eyePos.z;
  out[3] = d.x; out[4] = d.y; out[5] = d.z;
  out[6] = st.px; out[7] = st.py;
}
