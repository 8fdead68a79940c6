// ref_driver.cpp -- TEST INFRASTRUCTURE ONLY (oracle/). Not part of the product path.
//
// Host driver around the reference's OWN kernel text. oracle/build_ref.py rewrites
// /root/reference/resources/renderer.cl (vector literals, swizzles, optional counters) into a
// temporary file whose path is passed as -DRM_REF_SOURCE="..."; it is #included below inside
// namespace refcl, under the OpenCL-C compatibility header oracle/clshim.h. The result is a
// shared library (oracle/_ref/libref_*.so) that runs the reference's RenderImage/TonemapImage
// (renderer.cl:478-508) work-item by work-item on the host cores -- which is what a CPU OpenCL
// device does. Nothing of the reference's text is stored in this repository.
#include <cstdint>
#include <cstring>
#include <cstddef>
#include <omp.h>

#include "clshim.h"

namespace refcl {
thread_local int g_global_id = 0;
#if RM_COUNTERS
thread_local uint64_t g_cnt_step = 0, g_cnt_tap = 0, g_cnt_outer = 0;
#define RM_CNT(which) (++g_cnt_##which)
#else
#define RM_CNT(which) ((void)0)
#endif

#include RM_REF_SOURCE

}  // namespace refcl

using namespace refcl;

static void add_counters(uint64_t* out, uint64_t s, uint64_t t, uint64_t o) {
  if (!out) return;
#pragma omp atomic
  out[0] += s;
#pragma omp atomic
  out[1] += t;
#pragma omp atomic
  out[2] += o;
}

extern "C" {

int ref_sizeof_opts(void) { return (int)sizeof(TRenderOpts); }
int ref_has_counters(void) { return RM_COUNTERS ? 1 : 0; }
int ref_num_threads(void) { return omp_get_max_threads(); }
void ref_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }

// Runs the reference kernel RenderImage (renderer.cl:478-494) for the work-items listed in
// `ids` (or for 0..n-1 when ids == NULL). `pixels` is the full W*H float4 accumulator, updated
// in place exactly as the kernel does. counters[3] += {inner steps, occupancy taps, outer iters}.
void ref_render_pixels(const uint8_t* voxels, const float* mc, const void* opts, float* pixels,
                       int n, const int* ids, int n_ids, uint64_t* counters) {
  const int count = ids ? n_ids : n;
#pragma omp parallel
  {
#if RM_COUNTERS
    g_cnt_step = g_cnt_tap = g_cnt_outer = 0;
#endif
#pragma omp for schedule(dynamic, 64)
    for (int k = 0; k < count; ++k) {
      g_global_id = ids ? ids[k] : k;
      RenderImage((const uchar*)voxels, (const float4*)mc, (const TRenderOpts*)opts,
                  (float4*)pixels, n);
    }
#if RM_COUNTERS
    add_counters(counters, g_cnt_step, g_cnt_tap, g_cnt_outer);
#endif
  }
}

// Reference kernel TonemapImage (renderer.cl:496-508) over all n work-items.
void ref_tonemap(const float* pixels, const void* opts, uint32_t* argb, int n) {
  TRenderOpts o;
  std::memcpy(&o, opts, sizeof(o));
#pragma omp parallel for schedule(static)
  for (int k = 0; k < n; ++k) {
    g_global_id = k;
    TonemapImage((const float4*)pixels, &o, (refcl::uint*)argb, n);
  }
}

// ---- per-function hooks (known-answer vectors for the restatement and the CUDA path) ----

// intersectsBox, renderer.cl:153-161
float ref_intersects_box(const float* bmin, const float* bmax, const float* p, const float* dir) {
  return intersectsBox(float3(bmin[0], bmin[1], bmin[2]), float3(bmax[0], bmax[1], bmax[2]),
                       float3(p[0], p[1], p[2]), float3(dir[0], dir[1], dir[2]));
}

// voxelLookup, renderer.cl:163-170
int ref_voxel_lookup(const uint8_t* voxels, const void* opts, const float* p) {
  TRenderOpts o;
  std::memcpy(&o, opts, sizeof(o));
  return voxelLookup((const uchar*)voxels, &o, float3(p[0], p[1], p[2]));
}

// voxelNormal / voxelNormalSmooth, renderer.cl:180-203 (un-normalised / normalised as in the text)
void ref_voxel_normal(const uint8_t* voxels, const void* opts, const int* q, int smooth, float* out) {
  TRenderOpts o;
  std::memcpy(&o, opts, sizeof(o));
  const int3 qq(q[0], q[1], q[2]);
  const float3 n = smooth ? voxelNormalSmooth((const uchar*)voxels, &o, qq)
                          : voxelNormal((const uchar*)voxels, &o, qq);
  out[0] = n.x; out[1] = n.y; out[2] = n.z;
}

// distanceToScene, renderer.cl:209-237. out = {dist, id, nx, ny, nz}
void ref_distance_to_scene(const uint8_t* voxels, const void* opts, const float* rpos,
                           const float* dir, int steps, int smooth, float* out) {
  TRenderOpts o;
  std::memcpy(&o, opts, sizeof(o));
  TIsec isec;
  std::memset(&isec, 0, sizeof(isec));
  const float2 r = distanceToScene((const uchar*)voxels, &o, &isec, float3(rpos[0], rpos[1], rpos[2]),
                                   float3(dir[0], dir[1], dir[2]), steps, smooth != 0);
  out[0] = r.x; out[1] = r.y; out[2] = isec.normal.x; out[3] = isec.normal.y; out[4] = isec.normal.z;
}

// raymarch, renderer.cl:239-257. out = {px,py,pz, nx,ny,nz, distance, objectID}
void ref_raymarch(const uint8_t* voxels, const void* opts, const float* pos, const float* dir,
                  float max_dist, int max_steps, int smooth, float* out) {
  TRenderOpts o;
  std::memcpy(&o, opts, sizeof(o));
  TRay ray;
  ray.pos = float3(pos[0], pos[1], pos[2]);
  ray.dir = float3(dir[0], dir[1], dir[2]);
  TIsec isec;
  std::memset(&isec, 0, sizeof(isec));
  raymarch((const uchar*)voxels, &o, &ray, &isec, max_dist, max_steps, smooth != 0);
  out[0] = isec.pos.x; out[1] = isec.pos.y; out[2] = isec.pos.z;
  out[3] = isec.normal.x; out[4] = isec.normal.y; out[5] = isec.normal.z;
  out[6] = isec.distance; out[7] = (float)isec.objectID;
}

// initRenderState + cameraRayLookat, renderer.cl:456-476. out = {eye xyz, dir xyz, pixelPos xy}
void ref_camera_ray(const void* opts, const float* mc, int id, float* out) {
  TRenderOpts o;
  std::memcpy(&o, opts, sizeof(o));
  TRenderState st = initRenderState(&o, (const float4*)mc, id);
  TRay r = cameraRayLookat(&o, &st);
  out[0] = r.pos.x; out[1] = r.pos.y; out[2] = r.pos.z;
  out[3] = r.dir.x; out[4] = r.dir.y; out[5] = r.dir.z;
  out[6] = st.pixelPos.x; out[7] = st.pixelPos.y;
}

// Byte offsets of every TRenderOpts field as the reference's struct text lays them out.
// Order: see oracle/refso.py OPTS_FIELDS.
int ref_opts_offsets(int* out, int cap) {
#define OFF(f) do { if (k < cap) out[k] = (int)offsetof(TRenderOpts, f); ++k; } while (0)
  int k = 0;
  OFF(eyePos); OFF(targetPos); OFF(up); OFF(voxelBounds); OFF(voxelBounds2); OFF(voxelBoundsMin);
  OFF(voxelBoundsMax); OFF(invVoxelScale); OFF(skyColor1); OFF(skyColor2); OFF(voxelRes);
  OFF(resolution); OFF(invAspect); OFF(time); OFF(fov); OFF(maxIter); OFF(maxVoxelIter);
  OFF(maxDist); OFF(startDist); OFF(eps); OFF(aoIter); OFF(aoStepDist); OFF(aoAmp); OFF(voxelSize);
  OFF(groundY); OFF(shadowIter); OFF(reflectIter); OFF(shadowBias); OFF(lightScatter);
  OFF(minLightAtt); OFF(gamma); OFF(exposure); OFF(dof); OFF(frameBlend); OFF(fogPow); OFF(flareAmp);
  OFF(mcTableLength); OFF(isoVal); OFF(numLights); OFF(lightPos); OFF(lightColor); OFF(materials);
#undef OFF
  return k;
}

}  // extern "C"
