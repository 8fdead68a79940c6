// rm_generate.cu -- input generators on the device: the reference builds its test
// volume and its per-pass random tables on the JVM and uploads them every frame
// (generators.clj:8-42, core.clj:81,84,137-138). Generating them where they are used removes the
// host work (the 1024^3 gyroid takes tens of seconds on one CPU core) and the PCIe upload.
// Both generators follow the host mirror raymarchcl_b200/generators.py operation for operation in
// fp64 (this library is compiled -fmad=false; fp64 sqrt and division are IEEE), and the tests
// compare them with it byte for byte.
#include "rm_kernels.h"
#include <cstring>

namespace {

// cos / sin of coord*scl + offset for every coordinate of one axis
__global__ void k_axis_trig(int n, double scl, double offset, double* __restrict__ c, double* __restrict__ s) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double a = (double)i * scl + offset;
  c[i] = cos(a);
  s[i] = sin(a);
}

// make-gyroid-volume (generators.clj:27-42): g = |cos x sin z + cos y sin x + cos z sin y| - 1 in
// the slabs (z & 63) >= 32; |0.2 - g| < 0.05 -> 64 / 128 by x stripe; else g > 0.35 -> 255.
// One thread per 4 voxels of a row (one 32-bit store).
__global__ void __launch_bounds__(256)
k_gyroid(int rx, int ry, int rz, const double* __restrict__ cx, const double* __restrict__ sx,
         const double* __restrict__ cy, const double* __restrict__ sy, const double* __restrict__ cz,
         const double* __restrict__ sz, uint8_t* __restrict__ vox) {
  const int qx = (rx + 3) >> 2;
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  if (t >= (long long)qx * ry * rz) return;
  const int x0 = (int)(t % qx) * 4;
  const int y = (int)((t / qx) % ry);
  const int z = (int)(t / ((long long)qx * ry));
  uint8_t out[4] = {0, 0, 0, 0};
  if ((z & 63) >= 32) {
    const double cyv = cy[y], syv = sy[y], czv = cz[z], szv = sz[z];
    for (int k = 0; k < 4; ++k) {
      const int x = x0 + k;
      if (x >= rx) break;
      const double g = fabs(cx[x] * szv + cyv * sx[x] + czv * syv) - 1.0;
      if (fabs(0.2 - g) < 0.05) out[k] = (x & 63) < 32 ? 64 : 128;
      else if (g > 0.35) out[k] = 255;
    }
  }
  uint8_t* row = vox + ((size_t)z * ry + y) * rx + x0;
  if (x0 + 3 < rx && (((uintptr_t)row) & 3) == 0) {
    *reinterpret_cast<uint32_t*>(row) = out[0] | (out[1] << 8) | (out[2] << 16) | ((uint32_t)out[3] << 24);
  } else {
    for (int k = 0; k < 4 && x0 + k < rx; ++k) row[k] = out[k];
  }
}

// java.util.Random: state' = (state * 0x5DEECE66D + 0xB) mod 2^48; k steps at once by squaring
__device__ unsigned long long lcg_jump(unsigned long long s, unsigned long long k) {
  const unsigned long long M = (1ull << 48) - 1;
  unsigned long long a = 0x5DEECE66Dull, c = 0xBull;  // one step: s -> a*s + c
  while (k) {
    if (k & 1) s = (a * s + c) & M;
    c = (a * c + c) & M;  // two steps: a*(a*s + c) + c
    a = (a * a) & M;
    k >>= 1;
  }
  return s;
}

// generate-scatter-offsets (generators.clj:8-16) with an explicit seed: entry i of table t takes
// nextDouble() number 4i..4i+3 of java.util.Random(seed0 + t); each component is
// float(2*d - 1), the vector is scaled by 1/sqrt(sum of squares) in double and stored as float.
__global__ void __launch_bounds__(256)
k_scatter_tables(long long seed0, int tables, float4* __restrict__ out) {
  const int i = blockIdx.x * 256 + threadIdx.x;  // entry within a table
  const int t = blockIdx.y;
  if (i > RM_TABLE_MASK || t >= tables) return;
  const unsigned long long M = (1ull << 48) - 1;
  unsigned long long s = ((unsigned long long)(seed0 + t) ^ 0x5DEECE66Dull) & M;
  s = lcg_jump(s, 8ull * (unsigned long long)i);  // two next() calls per double, four doubles per entry
  double v[4];
  for (int k = 0; k < 4; ++k) {
    s = (s * 0x5DEECE66Dull + 0xBull) & M;
    const long long hi = (long long)(s >> (48 - 26));
    s = (s * 0x5DEECE66Dull + 0xBull) & M;
    const long long lo = (long long)(s >> (48 - 27));
    const double d = (double)((hi << 27) + lo) * (1.0 / 9007199254740992.0);
    v[k] = (double)(float)(2.0 * d - 1.0);
  }
  const double m = 1.0 / sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3]);
  out[(size_t)t * (RM_TABLE_MASK + 1) + i] =
      make_float4((float)(v[0] * m), (float)(v[1] * m), (float)(v[2] * m), (float)(v[3] * m));
}


// make-terrain (generators.clj:44-60): two 4-voxel walls of value 64 up to y < int(ry*0.666), then
// bumpy pillars of value 255 (they overwrite the walls): column (x, z) with dx = 16 - x%32,
// dz = 16 - z%32, dx^2 + dz^2 <= 121 is filled for y <= int(ry*(0.25 + 0.125*sin(0.02 z)*cos(0.03 x))).
// One thread per 4 voxels of a row.
__global__ void __launch_bounds__(256)
k_terrain(int rx, int ry, int rz, int ymax, const double* __restrict__ cosx, const double* __restrict__ sinz,
          uint8_t* __restrict__ vox) {
  const int qx = (rx + 3) >> 2;
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  if (t >= (long long)qx * ry * rz) return;
  const int x0 = (int)(t % qx) * 4;
  const int y = (int)((t / qx) % ry);
  const int z = (int)(t / ((long long)qx * ry));
  const int dz = 16 - (z % 32);
  const double sz = sinz[z];
  for (int k = 0; k < 4; ++k) {
    const int x = x0 + k;
    if (x >= rx) break;
    int v = 0;
    if (y < ymax && (z < 4 || (x >= rx - 4 && z < rx))) v = 64;  // second wall: index x'*rxy + y*rx + (rx-1-z'), x' < rx, z' < 4
    const int dx = 16 - (x % 32);
    if (dx * dx + dz * dz <= 121) {
      const int h = (int)((double)ry * (0.25 + 0.125 * (sz * cosx[x])));
      if (y <= h) v = 255;
    }
    vox[((size_t)z * ry + y) * rx + x] = (uint8_t)v;
  }
}

// ---- mesh point-splat voxeliser (meshvoxel.clj:16-69) ----
// floats as order-preserving ints, so that atomicMin / atomicMax give the bounding box
__device__ __forceinline__ int f_ord(float f) { const int i = __float_as_int(f); return i ^ ((i >> 31) & 0x7fffffff); }

// bounding box of the points (gu/bounding-box, meshvoxel.clj:18); bb[0..2] = min, bb[3..5] = max
// as ordered ints; bb[6] |= 1 when a coordinate is NaN or infinite
__global__ void __launch_bounds__(256)
k_points_bbox(const float* __restrict__ xyz, long long n, int* __restrict__ bb) {
  float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
  bool bad = false;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256)
    for (int k = 0; k < 3; ++k) {
      const float v = __ldg(xyz + 3 * i + k);
      bad = bad || !(fabsf(v) <= 3.0e38f);
      lo[k] = fminf(lo[k], v);
      hi[k] = fmaxf(hi[k], v);
    }
  for (int k = 0; k < 3; ++k)
    for (int off = 16; off > 0; off >>= 1) {
      lo[k] = fminf(lo[k], __shfl_down_sync(0xffffffffu, lo[k], off));
      hi[k] = fmaxf(hi[k], __shfl_down_sync(0xffffffffu, hi[k], off));
    }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(bb + 6, 1);
  if ((threadIdx.x & 31) == 0)
    for (int k = 0; k < 3; ++k) {
      atomicMin(bb + k, f_ord(lo[k]));
      atomicMax(bb + 3 + k, f_ord(hi[k]));
    }
}

struct SplatParams {  // mesh-scale (meshvoxel.clj:16-23), evaluated once on the host in fp64
  double p[3], off[3], s;
  int res, ks;  // ks < 0: `voxelize` (one voxel, points outside the grid dropped); else `voxelize-ks`
};

// Clojure's (int x) on a double: truncation toward zero; NaN -> 0. (The JVM throws beyond the int
// range; here such values saturate, which only ever happens for degenerate clouds.)
__device__ __forceinline__ int clj_int(double v) {
  if (!(v == v)) return 0;
  if (v >= 2147483647.0) return 2147483647;
  if (v <= -2147483648.0) return (-2147483647 - 1);
  return (int)v;
}

// one thread per (point, z-slice of its splat): writes are idempotent (always 255), no atomics
__global__ void __launch_bounds__(256)
k_splat_points(const float* __restrict__ xyz, long long n, const __grid_constant__ SplatParams P, uint8_t* __restrict__ vox) {
  const int slices = P.ks < 0 ? 1 : 2 * P.ks + 1;
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  if (t >= n * slices) return;
  const long long i = t / slices;
  const int dz = (int)(t - i * slices);
  int c[3];
  for (int k = 0; k < 3; ++k) {
    const double v = (double)__ldg(xyz + 3 * i + k);
    c[k] = clj_int(P.off[k] + (v - P.p[k]) * P.s);  // (g/+ off (g/* (g/- v p) s)), then (map int ..)
  }
  const int res = P.res;
  const size_t rxy = (size_t)res * res;
  if (P.ks < 0) {  // voxelize (meshvoxel.clj:60-69)
    if (c[0] >= 0 && c[0] < res && c[1] >= 0 && c[1] < res && c[2] >= 0 && c[2] < res)
      vox[(size_t)c[2] * rxy + (size_t)c[1] * res + c[0]] = 255;
    return;
  }
  // voxelize-ks (meshvoxel.clj:45-58): ranges clamped to the grid
  const long long z = (long long)c[2] - P.ks + dz;
  if (z < 0 || z >= res) return;
  const long long y0 = max((long long)c[1] - P.ks, 0ll), y1 = min((long long)c[1] + P.ks + 1, (long long)res);
  const long long x0 = max((long long)c[0] - P.ks, 0ll), x1 = min((long long)c[0] + P.ks + 1, (long long)res);
  for (long long y = y0; y < y1; ++y)
    for (long long x = x0; x < x1; ++x) vox[(size_t)z * rxy + (size_t)y * res + (size_t)x] = 255;
}

}  // namespace

cudaError_t rm_launch_gyroid(int rx, int ry, int rz, double* d_trig, uint8_t* d_vox, cudaStream_t stream) {
  // d_trig: 2 * (rx + ry + rz) doubles of scratch
  const double scl = 0.01 * (512.0 / (double)rx);
  double *cx = d_trig, *sx = cx + rx, *cy = sx + rx, *sy = cy + ry, *cz = sy + ry, *sz = cz + rz;
  k_axis_trig<<<(rx + 127) / 128, 128, 0, stream>>>(rx, scl, 0.3875, cx, sx);
  k_axis_trig<<<(ry + 127) / 128, 128, 0, stream>>>(ry, scl, 0.0, cy, sy);
  k_axis_trig<<<(rz + 127) / 128, 128, 0, stream>>>(rz, scl, 0.0, cz, sz);
  const long long n = (long long)((rx + 3) >> 2) * ry * rz;
  k_gyroid<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(rx, ry, rz, cx, sx, cy, sy, cz, sz, d_vox);
  return cudaGetLastError();
}

cudaError_t rm_launch_terrain(int rx, int ry, int rz, double* d_trig, uint8_t* d_vox, cudaStream_t stream) {
  // d_trig: 2 * (rx + rz) doubles of scratch
  double *cx = d_trig, *sx = cx + rx, *cz = sx + rx, *sz = cz + rz;
  k_axis_trig<<<(rx + 127) / 128, 128, 0, stream>>>(rx, 0.03, 0.0, cx, sx);
  k_axis_trig<<<(rz + 127) / 128, 128, 0, stream>>>(rz, 0.02, 0.0, cz, sz);
  const long long n = (long long)((rx + 3) >> 2) * ry * rz;
  k_terrain<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(rx, ry, rz, (int)((double)ry * 0.666), cx, sz, d_vox);
  return cudaGetLastError();
}

cudaError_t rm_launch_scatter_tables(long long seed0, int tables, float4* d_tables, cudaStream_t stream) {
  if (tables <= 0) return cudaSuccess;
  const dim3 grid((RM_TABLE_MASK + 1 + 255) / 256, (unsigned)tables);
  k_scatter_tables<<<grid, 256, 0, stream>>>(seed0, tables, d_tables);
  return cudaGetLastError();
}

static float ord_to_float(int i) { i ^= (i >> 31) & 0x7fffffff; float f; memcpy(&f, &i, 4); return f; }

cudaError_t rm_launch_voxelize_points(const float* d_xyz, long long n, int res, int ks, int* d_bb, uint8_t* d_vox,
                                      int* bad_input, cudaStream_t stream) {
  *bad_input = 0;
  cudaError_t e;
  const int init[7] = {0x7fffffff, 0x7fffffff, 0x7fffffff, (int)0x80000000, (int)0x80000000, (int)0x80000000, 0};
  if ((e = cudaMemcpyAsync(d_bb, init, sizeof init, cudaMemcpyHostToDevice, stream)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(d_vox, 0, (size_t)res * res * res, stream)) != cudaSuccess) return e;
  const unsigned blocks = (unsigned)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  k_points_bbox<<<blocks, 256, 0, stream>>>(d_xyz, n, d_bb);
  int bb[7];
  if ((e = cudaMemcpyAsync(bb, d_bb, sizeof bb, cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
  if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
  if (bb[6]) { *bad_input = 1; return cudaSuccess; }
  // mesh-scale (meshvoxel.clj:16-23) in fp64, in the reference's order of operations
  SplatParams P;
  double size[3], md = 0.0;
  for (int k = 0; k < 3; ++k) {
    P.p[k] = (double)ord_to_float(bb[k]);
    size[k] = (double)ord_to_float(bb[3 + k]) - P.p[k];
  }
  md = size[0] > size[1] ? size[0] : size[1];  // (max sx sy sz)
  md = md > size[2] ? md : size[2];
  for (int k = 0; k < 3; ++k) P.off[k] = (0.5 * (double)res) * (1.0 - size[k] / md);
  P.s = (double)res / md;
  P.res = res;
  P.ks = ks;
  const long long threads = n * (ks < 0 ? 1 : 2 * ks + 1);
  k_splat_points<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(d_xyz, n, P, d_vox);
  return cudaGetLastError();
}
