// rm_kernels.h -- host-callable launchers of the device code (rm_kernels.cu), used by rm_api.cu.
#pragma once
#include "rm_types.h"

// Maps a work slot of this shard to a pixel id (or -1 for padding). Slots are tile-major (the rank's
// tiles row by row); inside a tile consecutive groups of 32 slots cover 8x4 pixel blocks so that a
// warp traces a compact bundle of primary rays.
__host__ __device__ inline int rm_slot_to_pixel(const RmShard& sh, long long slot, int W, int H) {
  const unsigned tile_px = (unsigned)(sh.tile_w * sh.tile_h);
  // 32-bit divisions (a 64-bit one is ~100 instructions): rm_clear_accum bounds the frame to 2^31 / 37 pixels,
  // so a shard's slots -- pixels plus the padding of edge tiles and columns -- stay far below 2^32
  const unsigned s32 = (unsigned)slot;
  const unsigned lt = s32 / tile_px;
  const unsigned r = s32 - lt * tile_px;
  const unsigned ty = lt / (unsigned)sh.tiles_per_rank_row, k = lt - ty * (unsigned)sh.tiles_per_rank_row;
  // the rank's first column in this tile row: (tx + skew * ty) mod world == rank
  int first = (sh.rank - (int)(((unsigned)sh.skew * ty) % (unsigned)sh.world)) % sh.world;
  if (first < 0) first += sh.world;
  const int tx = first + (int)k * sh.world;
  if (tx >= sh.tiles_x) return -1;
  const unsigned sb = r >> 5, l = r & 31u;
  const unsigned sbw = (unsigned)sh.tile_w >> 3;
  const unsigned sby = sb / sbw, sbx = sb - sby * sbw;
  const int x = tx * sh.tile_w + (int)(sbx * 8u + (l & 7u));
  const int y = (int)ty * sh.tile_h + (int)(sby * 4u + (l >> 3));
  if (x >= W || y >= H) return -1;
  return y * W + x;
}

// RenderImage-equivalent, plain kernel (one thread per pixel-sample over the raw volume).
cudaError_t rm_launch_render_plain(const uint8_t* d_vox, const float4* d_table, const RmOpts& opts,
                                   const RmShard& shard, float4* d_accum, RmCounters* d_counters,
                                   cudaStream_t stream);

// TonemapImage-equivalent. packed == 0: d_argb[id] for every pixel of the frame (W*H words).
// packed == 1: d_argb[slot] for the slots of this shard (shard.slots words, padding = 0).
// packed == 2: d_argb[id] for the pixels this shard OWNS only (several GPUs fill one frame).
cudaError_t rm_launch_tonemap(const float4* d_accum, float gamma, int W, int H, const RmShard& shard,
                              uint32_t* d_argb, int packed, cudaStream_t stream);

// Accumulator gather into a packed, slot-ordered buffer (shard.slots float4, padding = 0).
cudaError_t rm_launch_pack_accum(const float4* d_accum, int W, int H, const RmShard& shard,
                                 float4* d_packed, cudaStream_t stream);

// The one multi-GPU assembly step after the gather: parts[r][slot] (r < world, `stride_slots`
// elements apart, elements of 4 (ARGB) or 16 (float4 accumulator) bytes) -> frame[pixel].
cudaError_t rm_launch_unpack_shards(const void* d_parts, int world, long long stride_slots, int elem_bytes, int W, int H,
                                    const RmShard& shard, void* d_frame, cudaStream_t stream);

// ---- fast path (rm_accel.cu, rm_render_fast.cu) ----
#define RM_MAX_FUSED_PASSES 32

// Build the occupancy acceleration data of the resident volume for one isoVal (asynchronous on `stream`).
cudaError_t rm_accel_build(const uint8_t* d_vox, int rx, int ry, int rz, int iso, int cell_shift,
                           RmAccelStorage* st, cudaStream_t stream);
void rm_accel_free(RmAccelStorage* st);

// RenderImage-equivalent for `passes` (<= RM_MAX_FUSED_PASSES) consecutive passes whose opts differ
// only in `time` / `frameBlend`: one launch over (pixel, pass) items + one blend launch
// (passes == 1: the render kernel blends into d_accum itself). d_tables = the passes' tables,
// contiguous; times / blend = per-pass TRenderOpts.time / frameBlend (host arrays, passed by value).
cudaError_t rm_launch_render_fast(const RmOpts& opts, const RmShard& shard, const RmAccel& accel,
                                  const float4* d_tables, const float* times, const float* blend,
                                  int passes, float4* d_colour, float4* d_accum, RmCounters* d_counters,
                                  cudaStream_t stream);

// pixels = mix(pixels, colour_k, blend_k) for k = 0..passes-1, in order (renderer.cl:492), from the
// colour buffer of a fused launch (passes x slots float4).
cudaError_t rm_launch_blend_passes(const float4* d_colour, const float* blend, int passes, const RmShard& shard,
                                   int W, int H, float4* d_accum, cudaStream_t stream);

// ---- default kernel: persistent warps, distance map in shared memory, blend + tonemap folded in
// (rm_scene_fused.cuh, rm_render_persist.cu), RM_OPT_KERNEL = 0 ----
#define RM_PERSIST_MAX_SMEM (200 * 1024) // largest 4-bit distance map staged into shared memory
#define RM_PERSIST_DEFAULT_BLOCK 0        // 0 = chosen per launch (rm_launch_render_persist): 1024 x 1 + TMA-staged map for long launches, else 256 x 5
#define RM_PERSIST_DEFAULT_SMEM 2         // 0 never, 1 whenever the map fits the layout, 2 = with the 1024-thread layout only
#define RM_PERSIST_DEFAULT_ROUND 0        // free-running warps (1 = block-synchronous rounds: faster while the code was 58 KB, slower at 47 KB)
#define RM_PERSIST_AUTO_BUNDLES_PER_WARP 4   // "long launch": at least this many bundles per resident warp slot of the 1024 x 1 layout
                                            // (measured: 27 per slot, a 1/8 shard of C2: 4.08 vs 4.23 ms; 13.7, 960x540 x 4 passes: 2.45 vs
                                            //  2.55; 0.43, C1: 0.403 vs 0.386 -- 2 048 bundles fill only 64 of the 148 big blocks)
// Layout and map location of one launch of the default kernel (RM_OPT_PERSIST_BLOCK / _SMEM; block_threads 0 and
// smem_map 2 = pick here). One 1024-thread block per SM with the 4-bit distance map staged into its shared memory by TMA
// is the fastest form once a launch has enough bundles to fill 148 such blocks a few times over (B200, C2: 30.99 ms
// against 31.43 for 256 x 5 with the byte map in L1 / L2; C1's 2 048 bundles: 0.403 against 0.386); five 256-thread blocks
// per SM each with a copy of the map are not (128^3: 30.18 against 30.03), so the automatic choice couples the two.
// The counting kernels always run in the big layout. (Host-compiled by tests/hostsim: tests/test_host.py.)
struct RmPersistLayout { int threads, blocks_per_sm, use_nib; };
inline RmPersistLayout rm_persist_pick_layout(long long bundles, int num_sms, unsigned nib_bytes, int block_threads, int smem_map,
                                              int counting, int small_threads = 256, int small_blocks = 5, int big_threads = 1024) {
  const bool have_nib = nib_bytes > 0;
  const bool fits_big = have_nib && nib_bytes <= (unsigned)RM_PERSIST_MAX_SMEM;
  const bool long_launch = bundles >= (long long)RM_PERSIST_AUTO_BUNDLES_PER_WARP * num_sms * (big_threads / 32);
  RmPersistLayout l;
  if (counting) l.threads = big_threads;
  else if (block_threads == 0) l.threads = (smem_map != 0 && fits_big && long_launch) ? big_threads : small_threads;
  else l.threads = block_threads == 256 ? small_threads : big_threads;
  const bool big = l.threads == big_threads;
  l.blocks_per_sm = big ? 1 : small_blocks;
  const bool fits = have_nib && nib_bytes <= (unsigned)(RM_PERSIST_MAX_SMEM / l.blocks_per_sm);
  l.use_nib = (fits && (smem_map == 1 || (smem_map == 2 && big))) ? 1 : 0;
  return l;
}
// How many of `available` consecutive fusable passes one launch should take: all (<= 32) when they
// fill >= 80 % of a warp's lanes as whole pixels x passes groups, else the largest power of two.
inline int rm_persist_pick_passes(int available) {
  int m = available < RM_MAX_FUSED_PASSES ? available : RM_MAX_FUSED_PASSES;
  if (m < 1) return 0;
  if ((32 / m) * m * 5 >= 32 * 4) return m;  // >= 80 % of the lanes carry an item
  int p = 1;
  while (p * 2 <= m) p *= 2;
  return p;
}
// RenderImage for `passes` consecutive fusable passes, blended into d_accum in pass order inside the
// kernel. d_argb (optional): the ARGB words of the frame so far (TonemapImage with opts.gamma), indexed
// by pixel id or, argb_packed != 0, by shard slot (padding slots = 0). d_queue: one 64-bit ticket
// counter owned by the context (zeroed on the stream by this call).
// block_threads: layout of the resident blocks: 1024 (x 1 per SM, 64 registers) or 256 (x 5 per SM, 48 registers).
// smem_map: 0 = read the distance map from global memory even when the 4-bit copy would fit the SM's shared memory.
// bottom_up: hand the bundles out from the end of the slot list (the launch ends on the image's top rows).
// round_bundles: 0 = free-running warps; else the block draws one bundle per warp together and meets at its barrier per draw.
cudaError_t rm_launch_render_persist(const RmOpts& opts, const RmShard& shard, const RmAccel& accel,
                                     const float4* d_tables, const float* times, const float* blend, int passes,
                                     float4* d_accum, uint32_t* d_argb, int argb_packed, RmCounters* d_counters,
                                     unsigned long long* d_queue, int num_sms,
                                     int block_threads, int round_bundles, int smem_map, int bottom_up, cudaStream_t stream);

// ---- warp-scheduled state machine (rm_render_warp.cu), RM_OPT_KERNEL = 2 ----
int rm_warp_blocks_per_sm(int count);
// Persistent launch over (pixel, pass) items; colours go to d_colour when passes > 1 (blend with
// rm_launch_blend_passes afterwards), else straight into d_accum. d_watchdog: 16 words zeroed by
// the caller; word 0 != 0 afterwards = a warp exceeded trip_limit trips and gave up.
cudaError_t rm_launch_render_warp(const RmOpts& opts, const RmShard& shard, const RmAccel& accel,
                                  const float4* d_tables, const float* times, int passes, float4* d_colour,
                                  float4* d_accum, unsigned long long* d_queue, RmCounters* d_counters,
                                  unsigned* d_watchdog, unsigned trip_limit, int grid_blocks, cudaStream_t stream);

// ---- device-side input generators (rm_generate.cu) ----
// make-gyroid-volume (generators.clj:27-42) into d_vox (rx*ry*rz bytes); d_trig = 2*(rx+ry+rz) doubles of scratch.
cudaError_t rm_launch_gyroid(int rx, int ry, int rz, double* d_trig, uint8_t* d_vox, cudaStream_t stream);
// make-terrain (generators.clj:44-60) into d_vox; d_trig = 2*(rx+rz) doubles of scratch. The reference indexes its
// second wall as x*rx*ry + ..., x < rx: the caller must make sure rz >= rx (as the reference implicitly requires).
cudaError_t rm_launch_terrain(int rx, int ry, int rz, double* d_trig, uint8_t* d_vox, cudaStream_t stream);
// generate-scatter-offsets (generators.clj:8-16) for java.util.Random seeds seed0 .. seed0+tables-1.
cudaError_t rm_launch_scatter_tables(long long seed0, int tables, float4* d_tables, cudaStream_t stream);
// mesh-scale + voxelize / voxelize-ks (meshvoxel.clj:16-69) of n points (d_xyz: 3n floats) into d_vox (res^3 bytes,
// zero-filled first). ks < 0 = `voxelize`, else `voxelize-ks`. d_bb: 7 ints of scratch. Synchronises the stream.
// *bad_input is set when a coordinate is NaN / infinite (nothing is splatted then).
cudaError_t rm_launch_voxelize_points(const float* d_xyz, long long n, int res, int ks, int* d_bb, uint8_t* d_vox,
                                      int* bad_input, cudaStream_t stream);

// ---- wavefront path (rm_wave.cuh, rm_render_wave.cu), RM_OPT_KERNEL = 3 ----
struct RmWaveScratch {  // HBM scratch of one chunk of items, owned by the context
  void* d_rec = nullptr;    // (reflectIter + 1) x cap records of 64 B
  void* d_refl = nullptr;   // cap float4
  void* d_pxy = nullptr;    // cap float2
  void* d_jobs = nullptr;   // job_cap jobs of 64 B
  unsigned* d_ctr = nullptr;  // [0] jobs appended, [1] queue head of the trace kernel
  unsigned cap = 0, job_cap = 0;
  int levels = 0;
};
void rm_wave_free(RmWaveScratch* w);
// can the wavefront path render these options? (reflectIter < 8, numLights <= 4; otherwise use the fused kernel)
int rm_wave_supports(const RmOpts& opts);
// RenderImage-equivalent for `passes` fusable passes, same contract as rm_launch_render_fast; the items are
// processed in chunks of at most chunk_items; refill_min_idle (1..32) is the trace kernel's refill policy.
// *launches is incremented by the number of kernels launched.
cudaError_t rm_launch_render_wave(const RmOpts& opts, const RmShard& shard, const RmAccel& accel,
                                  const float4* d_tables, const float* times, const float* blend, int passes,
                                  float4* d_colour, float4* d_accum, RmCounters* d_counters, RmWaveScratch* w,
                                  int num_sms, unsigned chunk_items, int refill_min_idle, int* launches, cudaStream_t stream);
