// rm_types.h -- host/device types shared by the C-ABI layer and the kernels.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define RM_TABLE_MASK 0x3fffu  // renderer.cl:143

struct RmMaterial {  // TMaterial, renderer.cl:14-19 (dummy float2 dropped)
  float3 albedo;
  float r0, smoothness;
};

// TRenderOpts (renderer.cl:35-78) decoded from the 544-byte blob by rm_api.cu:decode_opts().
struct RmOpts {
  float3 eyePos, targetPos, up;
  float3 voxelBounds, voxelBounds2, boundsMin, boundsMax, invVoxelScale;
  float3 sky1, sky2;
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
  float3 lightPos[4], lightColor[4];
  RmMaterial mat[4];
  // derived by rm_derive_opts (not part of the 544-byte blob)
  float ao_k;  // AO probes: samples per world unit of reach, a conservative upper bound (rm_scene_fused.cuh:ambient_occlusion)
  float st_k;  // the same for sphere traces along a unit direction (maxVoxelIter steps)
  int window_ok;  // the voxel box is small enough for march_window's error bounds (|bounds| <= 64)
};

// Derived constants of a decoded TRenderOpts; called by every decoder (rm_api.cu, tests/hostsim).
inline void rm_derive_opts(RmOpts* o) {
  // An AO probe marches along a UNIT direction n with delta = (n / (msteps * 0.5)) * invVoxelScale, i.e. world steps of
  // |n (.) invVoxelScale (.) voxelBounds2| / (msteps * 0.5) >= 0.999 * min_i |invVoxelScale_i * voxelBounds2_i| / (msteps * 0.5)
  // (0.999 covers |n| = 1 +- 1e-6 and the rounding of delta): the number of samples within a reach r is at most r * ao_k.
  const float fx = o->invVoxelScale.x * o->voxelBounds2.x, fy = o->invVoxelScale.y * o->voxelBounds2.y, fz = o->invVoxelScale.z * o->voxelBounds2.z;
  float m = fx < 0.f ? -fx : fx;
  const float ay = fy < 0.f ? -fy : fy, az = fz < 0.f ? -fz : fz;
  m = ay < m ? ay : m;
  m = az < m ? az : m;
  const float c = (float)(o->maxVoxelIter / 2) * 0.5f;
  o->ao_k = (m > 1e-20f && m < 1e20f && c > 0.f) ? 1.01f * c / (0.999f * m) : 3.0e38f;  // (3e38: never cut)
  const float c2 = (float)o->maxVoxelIter * 0.5f;
  o->st_k = (m > 1e-20f && m < 1e20f && c2 > 0.f) ? 1.01f * c2 / (0.999f * m) : 3.0e38f;
  const float b6[6] = {o->boundsMin.x, o->boundsMin.y, o->boundsMin.z, o->boundsMax.x, o->boundsMax.y, o->boundsMax.z};
  o->window_ok = 1;
  for (int i = 0; i < 6; ++i)
    if (!((b6[i] < 0.f ? -b6[i] : b6[i]) <= 64.0f)) o->window_ok = 0;  // (NaN fails too)
}

// Interleaved tile ownership: the frame is cut into tile_w x tile_h pixel tiles and tile
// (i, j) belongs to rank (i + skew * j) mod world -- diagonal stripes. (Round 1 dealt tiles row-major,
// t mod world: with 60 or 120 tiles per row and 8 ranks that degenerates into vertical stripes, and
// the render cost is anything but uniform across columns: 4 % load imbalance by the cost model of
// tools/simt_model.py, 0.2 % with the skew at 16 x 8 tiles.) Every rank owns exactly tiles_per_rank_row
// tile columns per tile row; columns beyond the frame (when tiles_x is not a multiple of world) are
// padding, like the pixels of edge tiles beyond the frame.
struct RmShard {
  int rank, world;
  int tile_w, tile_h;     // tile extent in pixels (multiples of 8 x 4)
  int tiles_x, tiles_y;   // tile grid over the framebuffer
  int tiles_per_rank_row; // ceil(tiles_x / world)
  int skew;               // coprime to world
  int owned_tiles;        // tiles_per_rank_row * tiles_y (incl. padding tiles)
  long long slots;        // owned_tiles * tile_w * tile_h (includes padding)
};

// Fills the derived fields of a shard for a W x H frame. Shared by the host API, the unpack kernel and
// (restated) raymarchcl_b200/dist.py:ShardLayout.
__host__ __device__ inline void rm_shard_layout(RmShard& s, int W, int H) {
  s.tiles_x = W > 0 ? (W + s.tile_w - 1) / s.tile_w : 0;
  s.tiles_y = H > 0 ? (H + s.tile_h - 1) / s.tile_h : 0;
  s.tiles_per_rank_row = (s.tiles_x + s.world - 1) / s.world;
  const int cand[5] = {3, 5, 7, 2, 1};
  s.skew = 1;
  for (int k = 0; k < 5; ++k) {
    int a = cand[k], b = s.world;
    while (b) { const int t = a % b; a = b; b = t; }
    if (a == 1) { s.skew = cand[k]; break; }
  }
  s.owned_tiles = s.tiles_per_rank_row * s.tiles_y;
  s.slots = (long long)s.owned_tiles * s.tile_w * s.tile_h;
}

struct RmCounters {  // reference-equivalent work, see rm_stats in raymarch_b200.h
  unsigned long long steps, taps, outer;
};

#define RM_DIST_CAP 32  // saturation of the macro-cell distance map

// Occupancy acceleration data derived from the volume for one isoVal (rm_accel.cu). It only tells
// the fast kernel which voxel fetches of the reference's march have a known outcome.
struct RmAccel {
  const uint8_t* vox;         // the uploaded volume, x fastest (material band of a hit voxel)
  const uint64_t* solid;      // bit-bricks of (v >  isoVal): one 64-bit word per 4x4x4 voxels
  const uint64_t* occ;        // bit-bricks of (v >= isoVal) (aliases `solid` when no voxel == isoVal)
  const uint8_t* dist;        // per macro-cell Chebyshev distance (in cells) to the nearest cell
                              // holding a solid voxel, saturated at RM_DIST_CAP; 0 = occupied
  int bx, by, bz;             // brick grid extents = ceil(res / 4)
  int mx, my, mz;             // macro-cell grid extents = ceil(res / cell)
  int cell_shift;             // macro-cell edge = 1 << cell_shift voxels (>= 2)
  float cellf;                // (float)(1 << cell_shift)
  float rxf, ryf, rzf;        // (float) of the grid extents: the march multiplies by them at every lookup
  float inv_rxf, inv_ryf, inv_rzf;  // their reciprocals (exact when the extents are powers of two, see pow2)
  int pow2;                   // all three extents are powers of two: scaling by them is exact in fp32, so the march's
                              // recurrence p += delta can run in voxel units (x = trunc(q.x), no multiply per sample)
  const uint8_t* nib;         // the same map packed to 4 bits per cell (min(dist, 15)); cell c = nibble c&1 of byte c>>1.
                              // The default kernel stages it into shared memory with a bulk TMA copy
  unsigned nib_bytes;         // its size, padded to a multiple of 16 bytes (bulk-copy granularity)
};

struct RmAccelStorage {  // owner of the device arrays behind an RmAccel view
  RmAccel view{};
  uint64_t* d_solid = nullptr;
  uint64_t* d_occ = nullptr;
  uint8_t* d_dist = nullptr;
  uint8_t* d_tmp = nullptr;
  uint8_t* d_nib = nullptr;
  unsigned* d_flag = nullptr;
  size_t brick_capacity = 0, cell_capacity = 0;
  int iso = -1;
  bool valid = false;
  int launches = 0;
};
