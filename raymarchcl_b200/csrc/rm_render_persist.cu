// rm_render_persist.cu -- the DEFAULT RenderImage kernel (renderer.cl:478-494) for sm_100a, with
// the frame's blend (renderer.cl:492) and TonemapImage (renderer.cl:496-508) folded into it.
//
//   persistent   the blocks live for the whole launch and every WARP draws its next work bundle from a global
//                ticket counter as soon as it is done with the last one. Two layouts, picked per launch
//                (rm_launch_render_persist): one 1024-thread block per SM (64 registers, 32 warps) with the
//                distance map in shared memory (TMA, below) for launches long enough to pay for staging it, and
//                five 256-thread blocks per SM (48 registers, 40 warps) with the byte map in global memory for
//                short ones. (An option, RM_OPT_PERSIST_GROUP, makes the warps of a block draw together and meet
//                at the block barrier per draw: the warps then share the instruction lines they pull in. That
//                was 4 % faster while the kernel's code was 58-68 KB against a 32 KB instruction cache and is
//                3 % slower now that it is 47 KB: profiles/r02_scheduling_ab.md.)
//   bundles      a bundle is 32 / m neighbouring pixels x the m passes of the launch, pass-minor:
//                the lanes of a warp render the SAME pixels in different passes (rays that differ
//                only by jitter), which is what keeps their control flow together.
//   blend        the m passes of a pixel sit in m adjacent lanes: mix(pixels, colour_k, frameBlend_k)
//                is folded in pass order with warp shuffles -- the same operations in the same order
//                as m separate RenderImage launches, hence the same bits -- and written once. No colour
//                buffer (531 MB at C2), no blend kernel.
//   tonemap      the lane that writes the accumulator also writes the ARGB word of the frame so far
//                (gamma of the launch's opts) -- into the context's frame, a caller's gather buffer or,
//                in a multi-GPU group, straight into GPU 0's frame over NVLink peer memory;
//                rm_tonemap returns that buffer when nothing changed.
//   TMA          on block start one thread arms an mbarrier and issues cp.async.bulk copies of the 4-bit macro-cell
//                distance map (128 KiB at 64^3 cells) into shared memory; every thread waits on the barrier and the
//                march then reads the map with LDS (rm_scene_fused.cuh): no long-scoreboard stall on the march's
//                first load, a third less L2 traffic. One copy per resident block, so at 256^3 it needs the
//                1024-thread layout. B200, C2: 30.8 ms against 31.4 for 256 x 5 with the byte map in L1 / L2.
//   FADD2        float3 adds and the march recurrence use Blackwell's packed fp32 add on the (x, y) lanes
//                (rm_math.cuh): the same IEEE results in two thirds of the issue slots.
//
// Compiled with -fmad=false like the rest of the library (pinned two-rounding evaluation order).
#include <mutex>

// float3 adds / subtracts of this kernel as packed FADD2 on the (x, y) lanes (rm_math.cuh; Blackwell-only SASS). The
// other render kernels keep scalar arithmetic, so tests that compare kernels bit for bit compare the two forms as well.
#define RM_PACKED_F3 1
#include "rm_kernels.h"
#include "rm_scene_fused.cuh"

namespace {


struct PersistParams {
  const float4* tables;              // passes x 16384 float4
  float times[RM_MAX_FUSED_PASSES];  // TRenderOpts.time per pass
  float blend[RM_MAX_FUSED_PASSES];  // TRenderOpts.frameBlend per pass
  float4* accum;
  uint32_t* argb;                    // optional: TonemapImage output of the frame so far
  float gamma;
  int argb_packed;                   // argb is indexed by shard slot (else by pixel id)
  RmCounters* counters;
  unsigned long long* queue;         // bundle tickets (zeroed before the launch)
  long long bundles;
  int passes;                        // m
  int ppb;                           // pixels per bundle = 32 / m
  const uint8_t* nib;                // 4-bit distance map in global memory (source of the bulk copy)
  unsigned nib_bytes;                // multiple of 16
  int round_bundles;                 // 0 = free-running warps; else block-synchronous rounds of one bundle per warp
  int bottom_up;                     // hand the bundles out last-to-first
};

// tonemap + pack of one pixel (renderer.cl:448-454, :502-506); same expression as rm_kernels.cu
__device__ __forceinline__ uint32_t tonemap_pack3(float3 p, float gamma) {
  const float c[3] = {p.x, p.y, p.z};
  uint32_t ch[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float t = c[i] / (gamma + c[i]);
    t = t * t * 255.0f;
    ch[i] = (uint32_t)f2i_sat(cl_clamp(t, 0.0f, 255.0f));
  }
  return 0xff000000u | (ch[0] << 16) | (ch[1] << 8) | ch[2];
}

// Layouts (threads per block x resident blocks per SM), RM_OPT_PERSIST_BLOCK: 1024 x 1 (64 registers,
// 32 warps per SM, room for a 200 KB distance map) and 256 x 5 (48 registers, 40 warps, <= 40 KB map per
// block). Every block stages its own copy of the map, so the 128 KiB map of a 256^3 volume only fits
// the first layout. (Measured and dropped, C2 ms per frame, free-running warps, byte map: 256 x 5 33.19, 256 x 6
// 33.03, 192 x 6 33.15, 512 x 2 33.76, 256 x 4 33.87, 128 x 8 35.50, 128 x 10 36.90.)
// (the small layout as macros so that an experiment can rebuild with another one: build.py -DRM_PERSIST_SMALL_T=128 ...)
#ifndef RM_PERSIST_SMALL_T
#define RM_PERSIST_SMALL_T 256
#define RM_PERSIST_SMALL_B 5
#endif
#ifndef RM_PERSIST_BIG_T
#define RM_PERSIST_BIG_T 1024  // (measured, C2 ms: 1024 threads 30.78; 896 (72 registers) 32.05; 768 (80) 33.94; 640 (96) 36.63)
#endif
#ifdef RM_PERSIST_TIMELINE
__device__ unsigned long long g_tl_start[148 * 40], g_tl_first[148 * 40], g_tl_end[148 * 40];
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#endif
template <bool kCount, int kMap, int kThreads, int kBlocksPerSM>
__global__ void __launch_bounds__(kThreads, kBlocksPerSM)
k_render_persist(const __grid_constant__ RmShard sh, const __grid_constant__ PersistParams P) {
  const RmOpts& o = fused::g_opts;
#ifdef RM_PERSIST_TIMELINE
  const unsigned tl_id = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  if ((threadIdx.x & 31) == 0 && tl_id < 148 * 40) g_tl_start[tl_id] = gtimer();
#endif
  constexpr bool kNib = (kMap & fused::kMapNib) != 0;
  if (kNib) {
    // Stage the distance map: bulk async copies (TMA engine, no registers, no per-thread loads)
    // complete on an mbarrier that every thread of the block then waits on.
    __shared__ __align__(8) unsigned long long s_bar;
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&s_bar);
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(P.nib_bytes) : "memory");
      const unsigned dst = (unsigned)__cvta_generic_to_shared(fused::rm_smem_nib);
      for (unsigned off = 0; off < P.nib_bytes; off += 32768u) {
        const unsigned n = P.nib_bytes - off < 32768u ? P.nib_bytes - off : 32768u;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst + off), "l"(P.nib + off), "r"(n), "r"(bar) : "memory");
      }
    }
    unsigned done;
    do {
      asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                   : "=r"(done) : "r"(bar), "r"(0u) : "memory");
    } while (!done);
  }

  const unsigned lane = threadIdx.x & 31u;
#ifdef RM_PERSIST_TIMELINE
  if (lane == 0 && tl_id < 148 * 40) g_tl_first[tl_id] = gtimer();
#endif
  const int m = P.passes;
  const int sub = (int)lane / m, pass = (int)lane - sub * m;  // pixel of the bundle, pass of the launch
  const bool lane_used = sub < P.ppb;
  const int base = sub * m;  // first lane of this pixel's group
  fused::Cnt<kCount> cnt;
  fused::Lane s;
  s.table = P.tables + (size_t)(lane_used ? pass : 0) * (RM_TABLE_MASK + 1);
  s.time = P.times[lane_used ? pass : 0];

  // Scheduling granularity (round_bundles, RM_OPT_PERSIST_GROUP).
  //   0 (default): every warp draws its next bundle on its own -- no warp ever waits for another, but the warps of
  //      an SM drift through the routine independently and the kernel's code (larger than the 32 KB L1.5
  //      instruction cache) is fetched again and again: ncu stall_no_instruction 2.0 per issue at 256 x 5, 1.0
  //      at 1024 x 1.
  //   1: the block draws one bundle per warp together (warp w takes bundle t0 + w) and meets at its barrier
  //      before the next draw. Its warps then run the same phases of the routine at about the same time and
  //      share the instruction lines they pull in (stall_no_instruction 0.6), at the price of waiting for the
  //      slowest warp of a round (17 % of the warp-time). Which one wins is a matter of code size: rounds by 4 %
  //      at 58 KB (36.7 vs 38.2 ms, 34.8 vs 35.4), free-running by 3 % at 47 KB (31.4 vs 32.3).
  //   (Measured and removed: 2 / 4 bundles per warp per round, two half-block sync groups over named barriers,
  //    named-barrier groups inside a 1024-thread block: all slower, profiles/r02_scheduling_ab.md.)
  __shared__ unsigned long long s_ticket[2];
  const bool rounds = P.round_bundles != 0;
  constexpr int W = kThreads / 32;
  const int warp = (int)(threadIdx.x >> 5);
  unsigned round = 0;
  for (;;) {
    unsigned long long t = 0;
    if (!rounds) {
      if (lane == 0) t = atomicAdd(P.queue, 1ull);
      t = __shfl_sync(0xffffffffu, t, 0);
      if (t >= (unsigned long long)P.bundles) break;
    } else {
      if (threadIdx.x == 0) s_ticket[round & 1u] = atomicAdd(P.queue, (unsigned long long)W);
      __syncthreads();
      const unsigned long long t0 = s_ticket[round & 1u];
      ++round;
      if (t0 >= (unsigned long long)P.bundles) break;  // the whole block leaves together
      t = t0 + (unsigned)warp;
      if (t >= (unsigned long long)P.bundles) continue;  // ragged last draw: sit this one out
    }
    // Bundles are handed out from the END of the shard's slot list, i.e. the frame is walked bottom-up: the
    // launch then ends on the top rows of the image, which in this renderer's scenes are mostly sky -- the
    // cheapest bundles there are -- so that the SMs drain within microseconds of each other instead of
    // within one expensive bundle (a pure scheduling choice: every bundle is rendered exactly once either way).
    if (P.bottom_up) t = (unsigned long long)P.bundles - 1ull - t;
    const long long slot = (long long)t * P.ppb + sub;
    const bool in_shard = lane_used && slot < sh.slots;
    const int id = in_shard ? rm_slot_to_pixel(sh, slot, o.width, o.height) : -1;
    float3 c = f3s(0.0f);
    if (id >= 0) c = fused::render_pixel_sample<kCount, kMap>(cnt, s, id);
    __syncwarp();
    // pixels = mix(pixels, colour_k, frameBlend_k), k = 0 .. m-1 in pass order (renderer.cl:492)
    float3 p = f3s(0.0f);
    if (id >= 0) {
      const float4 old = P.accum[id];
      p = f3(old.x, old.y, old.z);
    }
#pragma unroll 1  // (once per bundle: not worth four copies of the body)
    for (int k = 0; k < m; ++k) {
      const float3 ck = f3(__shfl_sync(0xffffffffu, c.x, base + k), __shfl_sync(0xffffffffu, c.y, base + k),
                           __shfl_sync(0xffffffffu, c.z, base + k));
      p = lerp3(p, ck, P.blend[k]);
    }
    if (in_shard && pass == 0) {
      if (id >= 0) P.accum[id] = make_float4(p.x, p.y, p.z, 1.0f);
      if (P.argb && (id >= 0 || P.argb_packed)) {
        const uint32_t word = id >= 0 ? tonemap_pack3(p, P.gamma) : 0u;  // (padding slots of a packed shard read 0)
        P.argb[P.argb_packed ? slot : (long long)id] = word;
      }
    }
  }

#ifdef RM_PERSIST_TIMELINE
  if (lane == 0 && tl_id < 148 * 40) g_tl_end[tl_id] = gtimer();
#endif
  if constexpr (kCount) {
    unsigned long long a = cnt.steps, b = cnt.taps, c = cnt.outer;
    for (int off = 16; off > 0; off >>= 1) {
      a += __shfl_down_sync(0xffffffffu, a, off);
      b += __shfl_down_sync(0xffffffffu, b, off);
      c += __shfl_down_sync(0xffffffffu, c, off);
    }
    if (lane == 0) {
      atomicAdd(&P.counters->steps, a);
      atomicAdd(&P.counters->taps, b);
      atomicAdd(&P.counters->outer, c);
    }
  }
}

// function attributes are per device: set once per (device, instantiation)
std::once_flag g_attr_once[64][64];

template <bool kCount, int kMap, int kThreads, int kBlocksPerSM>
cudaError_t launch(const RmShard& shard, const PersistParams& P, int blocks, size_t smem, int dev, cudaStream_t stream) {
  cudaError_t attr = cudaSuccess;
  constexpr int variant = ((kThreads == RM_PERSIST_BIG_T && kBlocksPerSM == 1 ? 0 : 1) << 4) | (kCount ? 8 : 0) | kMap;
  std::call_once(g_attr_once[dev & 63][variant], [&] {
    if (kMap & fused::kMapNib) attr = cudaFuncSetAttribute(k_render_persist<kCount, kMap, kThreads, kBlocksPerSM>, cudaFuncAttributeMaxDynamicSharedMemorySize, RM_PERSIST_MAX_SMEM / kBlocksPerSM);
    else attr = cudaFuncSetAttribute(k_render_persist<kCount, kMap, kThreads, kBlocksPerSM>, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
  });
  if (attr != cudaSuccess) return attr;
  k_render_persist<kCount, kMap, kThreads, kBlocksPerSM><<<blocks, kThreads, smem, stream>>>(shard, P);
  return cudaGetLastError();
}

template <int kMap>
cudaError_t launch_any(bool count, int threads, const RmShard& shard, const PersistParams& P, int blocks, size_t smem, int dev,
                       cudaStream_t stream) {
  // (the counting kernels visit every sample anyway: only the map's location matters to them)
  if (count) return launch<true, kMap & fused::kMapNib, RM_PERSIST_BIG_T, 1>(shard, P, blocks, smem, dev, stream);
  if (threads == RM_PERSIST_SMALL_T) return launch<false, kMap, RM_PERSIST_SMALL_T, RM_PERSIST_SMALL_B>(shard, P, blocks, smem, dev, stream);
  return launch<false, kMap, RM_PERSIST_BIG_T, 1>(shard, P, blocks, smem, dev, stream);
}

}  // namespace

#ifdef RM_PERSIST_TIMELINE
extern "C" int rm_debug_timeline(unsigned long long* out /* 3 x 5920 */) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, g_tl_start, sizeof(g_tl_start));
  cudaMemcpyFromSymbol(out + 148 * 40, g_tl_first, sizeof(g_tl_first));
  cudaMemcpyFromSymbol(out + 2 * 148 * 40, g_tl_end, sizeof(g_tl_end));
  return 0;
}
#endif

cudaError_t rm_launch_render_persist(const RmOpts& opts, const RmShard& shard, const RmAccel& accel,
                                     const float4* d_tables, const float* times, const float* blend, int passes,
                                     float4* d_accum, uint32_t* d_argb, int argb_packed, RmCounters* d_counters,
                                     unsigned long long* d_queue, int num_sms,
                                     int block_threads, int round_bundles, int smem_map, int bottom_up, cudaStream_t stream) {
  if (shard.slots <= 0 || passes <= 0) return cudaSuccess;
  if (passes > RM_MAX_FUSED_PASSES) return cudaErrorInvalidValue;
  PersistParams P;
  P.tables = d_tables;
  for (int i = 0; i < RM_MAX_FUSED_PASSES; ++i) {
    P.times[i] = i < passes ? times[i] : 0.0f;
    P.blend[i] = i < passes ? blend[i] : 0.0f;
  }
  P.accum = d_accum;
  P.argb = d_argb;
  P.gamma = opts.gamma;
  P.argb_packed = argb_packed;
  P.counters = d_counters;
  P.queue = d_queue;
  P.passes = passes;
  P.ppb = 32 / passes;
  P.bundles = (shard.slots + P.ppb - 1) / P.ppb;
  P.nib = accel.nib;
  P.nib_bytes = accel.nib_bytes;
  const RmPersistLayout lay = rm_persist_pick_layout(P.bundles, num_sms, accel.nib != nullptr ? accel.nib_bytes : 0u, block_threads, smem_map,
                                                     d_counters != nullptr, RM_PERSIST_SMALL_T, RM_PERSIST_SMALL_B, RM_PERSIST_BIG_T);
  const int threads = lay.threads, blocks_per_sm = lay.blocks_per_sm;
  const bool use_nib = lay.use_nib != 0;
  const int warps_per_block = threads / 32;
  long long blocks = (P.bundles + warps_per_block - 1) / warps_per_block;
  if (blocks > (long long)num_sms * blocks_per_sm) blocks = (long long)num_sms * blocks_per_sm;
  P.round_bundles = round_bundles != 0;
  P.bottom_up = bottom_up != 0;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if ((e = cudaMemcpyToSymbolAsync(fused::g_opts, &opts, sizeof(RmOpts), 0, cudaMemcpyHostToDevice, stream)) != cudaSuccess) return e;
  if ((e = cudaMemcpyToSymbolAsync(fused::g_accel, &accel, sizeof(RmAccel), 0, cudaMemcpyHostToDevice, stream)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(d_queue, 0, sizeof(unsigned long long), stream)) != cudaSuccess) return e;
  const size_t smem = use_nib ? accel.nib_bytes : 0;
  const int map = (use_nib ? fused::kMapNib : 0) | (accel.cell_shift == 2 ? fused::kMapCell4 : 0) | (accel.pow2 ? fused::kMapPow2 : 0);
  const bool cnt = d_counters != nullptr;
  switch (map) {
    case 0: e = launch_any<0>(cnt, threads, shard, P, (int)blocks, smem, dev, stream); break;
    case 1: e = launch_any<1>(cnt, threads, shard, P, (int)blocks, smem, dev, stream); break;
    case 2: e = launch_any<2>(cnt, threads, shard, P, (int)blocks, smem, dev, stream); break;
    case 3: e = launch_any<3>(cnt, threads, shard, P, (int)blocks, smem, dev, stream); break;
    case 4: e = launch_any<4>(cnt, threads, shard, P, (int)blocks, smem, dev, stream); break;
    case 5: e = launch_any<5>(cnt, threads, shard, P, (int)blocks, smem, dev, stream); break;
    case 6: e = launch_any<6>(cnt, threads, shard, P, (int)blocks, smem, dev, stream); break;
    default: e = launch_any<7>(cnt, threads, shard, P, (int)blocks, smem, dev, stream); break;
  }
  if (e != cudaSuccess) return e;
  return cudaSuccess;
}
