/* meshvoxel_oracle.c -- TEST INFRASTRUCTURE ONLY. Plain-C restatement of the reference's mesh
 * point-splat voxeliser, /root/reference/src/thi/ng/raymarchcl/meshvoxel.clj:
 *   mesh-scale   :16-23   bounding box -> (v - p) * (res / md) + 0.5 * res * (1 - size / md), fp64
 *   voxelize-ks  :45-58   (2ks+1)^3 cube of value -1 (= 255) per vertex, ranges clamped to the grid
 *   voxelize     :60-69   one voxel per vertex, vertices outside the grid dropped
 * (voxelize-scatter :25-43 draws from an unseeded (rand) and is not reproducible; not restated.)
 *
 * PARITY UNPINNED: the reference ships no test, fixture or mesh asset for this path, and no JVM
 * exists in this image to run it, so this restatement is pinned only by reading the source. The
 * semantics it fixes: vertices are the STL's float32 coordinates widened to double (thi.ng/geom
 * vec3), (map int ..) truncates toward zero, (int NaN) = 0 (clojure.lang.RT.intCast), the voxel
 * index is z*res*res + y*res + x (meshvoxel.clj:57,68 -- the same layout the renderer reads).
 * Only tests/ may call into this file. */
#include <math.h>
#include <stdint.h>
#include <string.h>

static int clj_int(double v) {
  if (!(v == v)) return 0;
  if (v >= 2147483647.0) return 2147483647;
  if (v <= -2147483648.0) return (-2147483647 - 1);
  return (int)v;
}

/* returns 0, or -1 when a coordinate is NaN / infinite */
int orc_voxelize_points(const float* xyz, long long n, int res, int ks, uint8_t* vox) {
  double p[3], hi[3], size[3], off[3], md, s;
  const size_t rxy = (size_t)res * res;
  memset(vox, 0, rxy * res);
  for (int k = 0; k < 3; ++k) { p[k] = INFINITY; hi[k] = -INFINITY; }
  for (long long i = 0; i < n; ++i)
    for (int k = 0; k < 3; ++k) {
      const double v = (double)xyz[3 * i + k];
      if (!(fabs(v) <= 3.0e38)) return -1;
      if (v < p[k]) p[k] = v;
      if (v > hi[k]) hi[k] = v;
    }
  for (int k = 0; k < 3; ++k) size[k] = hi[k] - p[k];   /* gu/bounding-box -> [p [sx sy sz]]        :18 */
  md = size[0] > size[1] ? size[0] : size[1];
  md = md > size[2] ? md : size[2];                      /* (max sx sy sz)                            :19 */
  for (int k = 0; k < 3; ++k) off[k] = (0.5 * (double)res) * (1.0 - size[k] / md); /* (* 0.5 res (- 1.0 (/ % md))) :20 */
  s = (double)res / md;                                  /* (vec3 (/ res md))                         :21 */
  for (long long i = 0; i < n; ++i) {
    int c[3];
    for (int k = 0; k < 3; ++k) c[k] = clj_int(off[k] + ((double)xyz[3 * i + k] - p[k]) * s); /* :23, (map int ..) :52,65 */
    if (ks < 0) {                                        /* voxelize :60-69 */
      if (c[0] >= 0 && c[0] < res && c[1] >= 0 && c[1] < res && c[2] >= 0 && c[2] < res)
        vox[(size_t)c[2] * rxy + (size_t)c[1] * res + c[0]] = 255;
      continue;
    }
    {                                                    /* voxelize-ks :45-58 */
      const long long z0 = (long long)c[2] - ks > 0 ? (long long)c[2] - ks : 0, z1 = (long long)c[2] + ks + 1 < res ? (long long)c[2] + ks + 1 : res;
      const long long y0 = (long long)c[1] - ks > 0 ? (long long)c[1] - ks : 0, y1 = (long long)c[1] + ks + 1 < res ? (long long)c[1] + ks + 1 : res;
      const long long x0 = (long long)c[0] - ks > 0 ? (long long)c[0] - ks : 0, x1 = (long long)c[0] + ks + 1 < res ? (long long)c[0] + ks + 1 : res;
      for (long long z = z0; z < z1; ++z)
        for (long long y = y0; y < y1; ++y)
          for (long long x = x0; x < x1; ++x) vox[(size_t)z * rxy + (size_t)y * res + (size_t)x] = 255;
    }
  }
  return 0;
}
