// rm_render_fast.cu -- the production form of RenderImage (renderer.cl:478-494) for sm_100a.
//
// One thread per work item (pixel, pass) running the per-pixel-sample routine of
// rm_scene_plain.cuh over the BrickVolume policy (bit-bricks + macro-cell distance map,
// rm_accel.cu): the reference's fixed-step march keeps its fp32 recurrence, but samples known
// to be empty cost three adds and no fetch, and the surface predicate is read from 8x smaller
// bit-bricks. All passes of a frame (up to RM_MAX_FUSED_PASSES) are items of ONE launch: the
// pass colours go to a colour buffer and k_blend_passes folds them into the accumulator in pass
// order, so the hardware block scheduler balances the 10x spread in per-pixel cost over
// passes x pixels / 128 blocks instead of per pass.
//
// Compiled with -fmad=false like the rest of the library (pinned two-rounding evaluation order).
#include "rm_kernels.h"
#include "rm_scene_plain.cuh"

namespace {

#ifndef RM_FAST_BLOCK
#define RM_FAST_BLOCK 256  // measured on B200 (C2): 64 / 128 / 256 / 512 threads at 40 warps per SM: 42.4 / 42.3 / 40.8 / 42.4 ms
#endif
constexpr int kFastBlock = RM_FAST_BLOCK;

struct FastParams {
  const float4* tables;              // passes x 16384 float4
  float times[RM_MAX_FUSED_PASSES];  // TRenderOpts.time per pass
  float4* colour;                    // passes x slots (null when passes == 1: blend straight into accum)
  float4* accum;
  RmCounters* counters;
  int passes;
};

// 40 resident warps per SM (48 registers per thread, a few spills to L1): the kernel is latency-
// and issue-bound, not register-bound; measured on B200 (C2) with 128-thread blocks, 1 -> 6 -> 8 ->
// 10 -> 12 blocks: 178 -> 140 -> 124 -> 118 -> 116 ms per frame (early kernel); re-measured on the
// current one: 32 / 40 / 48 warps = 43.5 / 41.9 / 42.0 ms.
#ifndef RM_FAST_MINBLOCKS
#define RM_FAST_MINBLOCKS 5
#endif

template <bool kCount>
__global__ void __launch_bounds__(kFastBlock, RM_FAST_MINBLOCKS)
k_render_bricks(const __grid_constant__ RmShard sh, const __grid_constant__ FastParams P) {
  const RmOpts& o = plain::g_opts;
  const long long item = (long long)blockIdx.x * kFastBlock + threadIdx.x;
  const long long total = (long long)P.passes * sh.slots;
  plain::Scene s(plain::g_accel.vox, P.tables);
  if (item < total) {
    // pass-minor item order: the lanes of a warp render the SAME pixels in different passes
    // (32 / passes neighbouring pixels x all passes). Their rays differ only by the per-pass
    // jitter, so they follow nearly the same control flow: far less divergence than 32
    // different pixels of one pass.
    const long long slot = item / P.passes;
    const int pass = (int)(item - slot * P.passes);
    const int id = rm_slot_to_pixel(sh, slot, o.width, o.height);
    if (id >= 0) {
      s.time = P.times[pass];
      s.table = P.tables + (size_t)pass * (RM_TABLE_MASK + 1);
      const plain::BrickVolume V{};
      const float3 c = plain::render_pixel_sample<kCount>(s, V, id);
      if (P.colour) {
        P.colour[item] = make_float4(c.x, c.y, c.z, 1.0f);  // [slot][pass]
      } else {
        const float4 old = P.accum[id];
        const float3 m = lerp3(f3(old.x, old.y, old.z), c, o.frameBlend);  // mix(), renderer.cl:492
        P.accum[id] = make_float4(m.x, m.y, m.z, 1.0f);
      }
    }
  }
  if (kCount) {
    unsigned long long a = s.w.steps, b = s.w.taps, c = s.w.outer;
    for (int off = 16; off > 0; off >>= 1) {
      a += __shfl_down_sync(0xffffffffu, a, off);
      b += __shfl_down_sync(0xffffffffu, b, off);
      c += __shfl_down_sync(0xffffffffu, c, off);
    }
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(&P.counters->steps, a);
      atomicAdd(&P.counters->taps, b);
      atomicAdd(&P.counters->outer, c);
    }
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
    const float4 c = __ldcs(colour + (size_t)slot * passes + k);
    p = lerp3(p, f3(c.x, c.y, c.z), bw.w[k]);
  }
  accum[id] = make_float4(p.x, p.y, p.z, 1.0f);
}

}  // namespace

cudaError_t rm_launch_render_fast(const RmOpts& opts, const RmShard& shard, const RmAccel& accel,
                                  const float4* d_tables, const float* times, const float* blend,
                                  int passes, float4* d_colour, float4* d_accum, RmCounters* d_counters,
                                  cudaStream_t stream) {
  if (shard.slots <= 0 || passes <= 0) return cudaSuccess;
  if (passes > RM_MAX_FUSED_PASSES) return cudaErrorInvalidValue;
  FastParams P;
  P.tables = d_tables;
  for (int i = 0; i < RM_MAX_FUSED_PASSES; ++i) P.times[i] = i < passes ? times[i] : 0.0f;
  P.colour = passes > 1 ? d_colour : nullptr;
  P.accum = d_accum;
  P.counters = d_counters;
  P.passes = passes;
  const long long total = (long long)passes * shard.slots;
  const long long blocks = (total + kFastBlock - 1) / kFastBlock;
  if (blocks > 0x7fffffffLL) return cudaErrorInvalidValue;
#ifndef RM_NO_CARVEOUT
  // no shared memory is used: give the whole 256 KB of the SM's unified cache to L1 (the distance
  // map, the bit-bricks and the spill slots all live there)
  // (function attributes are per device: one flag per device, set on the first launch there)
  static bool carveout_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !carveout_set[dev]) {
    cudaFuncSetAttribute(k_render_bricks<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
    cudaFuncSetAttribute(k_render_bricks<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
    carveout_set[dev] = true;
  }
#endif
  cudaError_t e = cudaMemcpyToSymbolAsync(plain::g_opts, &opts, sizeof(RmOpts), 0, cudaMemcpyHostToDevice, stream);
  if (e != cudaSuccess) return e;
  if ((e = cudaMemcpyToSymbolAsync(plain::g_accel, &accel, sizeof(RmAccel), 0, cudaMemcpyHostToDevice, stream)) != cudaSuccess) return e;
  if (d_counters)
    k_render_bricks<true><<<(unsigned)blocks, kFastBlock, 0, stream>>>(shard, P);
  else
    k_render_bricks<false><<<(unsigned)blocks, kFastBlock, 0, stream>>>(shard, P);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (passes > 1) e = rm_launch_blend_passes(d_colour, blend, passes, shard, opts.width, opts.height, d_accum, stream);
  return e;
}

cudaError_t rm_launch_blend_passes(const float4* d_colour, const float* blend, int passes, const RmShard& shard,
                                   int W, int H, float4* d_accum, cudaStream_t stream) {
  if (shard.slots <= 0 || passes <= 0) return cudaSuccess;
  if (passes > RM_MAX_FUSED_PASSES) return cudaErrorInvalidValue;
  BlendWeights bw;
  for (int i = 0; i < RM_MAX_FUSED_PASSES; ++i) bw.w[i] = i < passes ? blend[i] : 0.0f;
  k_blend_passes<<<(unsigned)((shard.slots + 255) / 256), 256, 0, stream>>>(d_colour, bw, passes, shard, W, H, d_accum);
  return cudaGetLastError();
}
