// rm_scene_plain.cuh -- the render op as one self-contained per-pixel routine over the raw uint8
// volume: the straightforward CUDA form of RenderImage (renderer.cl:478-494 and its call tree).
// It is the in-library comparison kernel (RM_OPT_KERNEL = 1): same results as the fast kernel,
// no acceleration data, one thread per pixel-sample.
#pragma once
#include "rm_math.cuh"
#include "rm_types.h"

namespace plain {

struct Work {  // reference-equivalent work counters of this thread
  unsigned steps, taps, outer;
};

struct Scene {
  const uint8_t* __restrict__ vox;
  const float4* __restrict__ table;
  const RmOpts& o;
  float time;  // TRenderOpts.time of the pass this thread renders (per lane in the fused kernel)
  Work w;
  RM_DEV Scene(const uint8_t* v, const float4* t, const RmOpts& opts) : vox(v), table(t), o(opts), time(opts.time) {
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

// renderer.cl:172-178
RM_DEV float occupancy(Scene& s, int x, int y, int z) {
  s.w.taps++;
  if (!in_grid(s.o, x, y, z)) return 0.0f;
  const int v = __ldg(s.vox + ((size_t)z * s.o.rxy + (size_t)y * s.o.rx + x));
  return v < s.o.isoVal ? 0.0f : 1.0f;
}

// renderer.cl:180-188
RM_DEV float3 gradient6(Scene& s, int x, int y, int z) {
  const float gx = occupancy(s, x + 1, y, z) - occupancy(s, x - 1, y, z);
  const float gy = occupancy(s, x, y + 1, z) - occupancy(s, x, y - 1, z);
  const float gz = occupancy(s, x, y, z + 1) - occupancy(s, x, y, z - 1);
  return f3(-gx, -gy, -gz);
}

// renderer.cl:190-203
RM_DEV float3 gradient27(Scene& s, int x, int y, int z) {
  float3 n = f3s(0.0f);
  for (int dz = -1; dz <= 1; ++dz)
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx)
        if (occupancy(s, x + dx, y + dy, z + dz) > 0.0f) n = n + gradient6(s, x + dx, y + dy, z + dz);
  return unit3(n);
}

// renderer.cl:209-237. Returns (distance, id); writes *normal where the reference writes isec->normal.
RM_DEV float2 scene_distance(Scene& s, float3 rpos, float3 dir, int steps, bool smooth, float3* normal) {
  const RmOpts& o = s.o;
  const float g = rpos.y + o.groundY;
  float2 res = (g < 1e5f) ? make_float2(g, g) : make_float2(1e5f, -1.0f);
  *normal = (res.x < 1e5f) ? f3(0.0f, 1.0f, 0.0f) : -dir;
  const float idist = box_entry(o.boundsMin, o.boundsMax, rpos, dir);
  if (idist >= 0.0f && idist < res.x) {
    const float3 delta = (dir / ((float)steps * 0.5f)) * o.invVoxelScale;
    float3 p = rpos + o.voxelBounds;
    if (idist > 0.0f) p = dir * idist + p;
    p = p * o.invVoxelScale;
    while (--steps >= 0) {
      const int x = f2i_sat(p.x * (float)o.rx), y = f2i_sat(p.y * (float)o.ry), z = f2i_sat(p.z * (float)o.rz);
      s.w.steps++;
      if (!in_grid(o, x, y, z)) break;
      const int v = __ldg(s.vox + ((size_t)z * o.rxy + (size_t)y * o.rx + x));
      if (v > o.isoVal) {
        *normal = smooth ? gradient27(s, x, y, z) : unit3(gradient6(s, x, y, z));
        const float3 hp = p * o.voxelBounds2 + (-o.voxelBounds);
        const float d = len3(rpos - hp) - o.voxelSize;
        const float band = v < 168 ? (v < 84 ? 1.0f : 2.0f) : 3.0f;  // renderer.cl:205-207
        return d < res.x ? make_float2(d, band) : res;
      }
      p = p + delta;
    }
  }
  return res;
}

// renderer.cl:239-257
RM_DEV void sphere_trace(Scene& s, float3 ro, float3 rd, Isec& r, float maxDist, int maxSteps, bool smooth) {
  r.distance = s.o.startDist;
  while (--maxSteps >= 0) {
    s.w.outer++;
    r.pos = ro + rd * r.distance;
    const float2 h = scene_distance(s, r.pos, rd, s.o.maxVoxelIter, smooth, &r.normal);
    r.objectID = f2i_sat(h.y);
    if (fabsf(h.x) <= s.o.eps || r.distance >= maxDist) break;
    r.distance += h.x;
  }
  if (r.distance >= maxDist) {
    r.pos = ro + rd * r.distance;
    r.objectID = -1;
    r.distance = 1000.0f;
  }
}

RM_DEV float3 sky(const RmOpts& o, float3 d) { return lerp3(o.sky1, o.sky2, d.y * 0.5f + 0.5f); }  // :259-261

// renderer.cl:263-269
RM_DEV float3 light_pos(const Scene& s, const PixelState& st, int i) {
  const uint32_t seed = f2u_wrap(st.px * 1957.0f + st.py * 2173.0f + s.time * 4763.742f);
  return table_xyz(s, seed) * s.o.lightScatter + s.o.lightPos[i];
}

RM_DEV float3 reflect3(float3 v, float3 n) { return v - n * (2.0f * dot3(v, n)); }  // :271-273

// renderer.cl:275-290
RM_DEV float3 atmosphere(const Scene& s, const PixelState& st, float3 ro, float3 rd, float distance, float3 col) {
  const RmOpts& o = s.o;
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
RM_DEV float ambient_occlusion(Scene& s, float3 pos, float3 n0) {
  const RmOpts& o = s.o;
  float ao = 1.0f, d = 0.0f;
  uint32_t seed = f2u_wrap(pos.x * 3183.75f + pos.y * 1831.42f + pos.z * 2945.87f + s.time * 2671.918f);
  for (int i = 0; i <= o.aoIter && ao > 0.01f; ++i) {
    d += o.aoStepDist;
    seed += 37u;
    const float3 n = unit3(table_xyz(s, seed) * 0.2f + n0);
    float3 unused;
    const float2 h = scene_distance(s, n * d + pos, n, o.maxVoxelIter / 2, false, &unused);
    ao *= 1.0f - cl_max((d - h.x) * o.aoAmp / d, 0.0f);
  }
  return ao;
}

// renderer.cl:348-381 (shadow :292-301 inlined)
RM_DEV float3 object_lighting(Scene& s, const PixelState& st, float3 rd, float3 ipos, const RmMaterial& m,
                              float3 n, float3 reflectCol) {
  const RmOpts& o = s.o;
  const float ao = ambient_occlusion(s, ipos, n);
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
      Isec sh;
      sphere_trace(s, ipos + ldir * o.shadowBias, ldir, sh, lmax, o.shadowIter, false);
      const float sf = sh.distance < lmax ? 0.0f : 1.0f;
      if (sf > 0.0f) {
        const float3 inc = (o.lightColor[i] * sf) * att;
        diff = diff + inc * cl_max(0.0f, dot3(ldir, n));
        spec = spec + inc * blinn_phong(m.smoothness, rd, ldir, n);
      }
    }
    diff = diff * m.albedo;
    fin = fin + lerp3(diff, spec, schlick(m.r0, m.smoothness, n, rd));
  }
  return fin / (float)o.numLights;
}

RM_DEV int mat_index(int id) { return id < 0 ? 0 : (id > 3 ? 3 : id); }

// renderer.cl:383-405
RM_DEV float3 bounce_color(Scene& s, const PixelState& st, float3 ro, float3 rd, Isec& isec) {
  const RmOpts& o = s.o;
  sphere_trace(s, ro, rd, isec, o.maxDist, o.maxIter, false);
  float3 col;
  if (isec.objectID < 0) {
    col = sky(o, rd);
  } else {
    col = object_lighting(s, st, rd, isec.pos, o.mat[mat_index(isec.objectID)], isec.normal,
                          sky(o, reflect3(rd, isec.normal)));
  }
  return atmosphere(s, st, ro, rd, isec.distance, col);
}

// renderer.cl:407-446
RM_DEV float3 scene_color(Scene& s, const PixelState& st, float3 ro, float3 rd) {
  const RmOpts& o = s.o;
  Isec isec;
  sphere_trace(s, ro, rd, isec, o.maxDist, o.maxIter, true);
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
        reflectCol = reflectCol + bounce_color(s, st, bo, bd, ri);
        if (ri.objectID < 0) break;
        if (o.mat[mat_index(ri.objectID)].r0 < 0.001f) break;
      }
    } else {
      reflectCol = sky(o, reflect3(rd, n));
    }
    col = object_lighting(s, st, rd, isec.pos, m, n, reflectCol);
  }
  return atmosphere(s, st, ro, rd, isec.distance, col);
}

// renderer.cl:467-476 + :456-465
RM_DEV float3 setup_pixel(const Scene& s, int id, PixelState& st) {
  const RmOpts& o = s.o;
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
RM_DEV float3 render_pixel_sample(Scene& s, int id) {
  PixelState st;
  const float3 rd = setup_pixel(s, id, st);
  return scene_color(s, st, st.eye, rd) * s.o.exposure;
}

}  // namespace plain
