// rm_kernels.cu -- device code of the render op for sm_100a.
// Compiled with -fmad=false: geometry must follow the pinned two-rounding evaluation order
// (DESIGN.md "Numerics"), and the shading code shares the translation unit.
#include "rm_kernels.h"
#include "rm_scene_plain.cuh"

namespace {

constexpr int kPlainBlock = 128;

// RenderImage (renderer.cl:478-494), plain form: one thread per work slot.
template <bool kCount>
__global__ void __launch_bounds__(kPlainBlock)
k_render_plain(const uint8_t* __restrict__ vox, const float4* __restrict__ table,
               const __grid_constant__ RmShard sh, float4* __restrict__ accum, RmCounters* __restrict__ counters) {
  const RmOpts& o = plain::g_opts;
  const long long slot = (long long)blockIdx.x * kPlainBlock + threadIdx.x;
  plain::Scene s(vox, table);
  if (slot < sh.slots) {
    const int id = rm_slot_to_pixel(sh, slot, o.width, o.height);
    if (id >= 0) {
      const plain::ByteVolume V{vox};
      const float3 c = plain::render_pixel_sample<kCount>(s, V, id);
      const float4 old = accum[id];
      const float3 m = lerp3(f3(old.x, old.y, old.z), c, o.frameBlend);  // mix(), renderer.cl:492
      accum[id] = make_float4(m.x, m.y, m.z, 1.0f);
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
      atomicAdd(&counters->steps, a);
      atomicAdd(&counters->taps, b);
      atomicAdd(&counters->outer, c);
    }
  }
}

// tonemap + pack of one pixel (renderer.cl:448-454, :502-506)
__device__ __forceinline__ uint32_t tonemap_pack(float4 p, float gamma) {
  float c[3] = {p.x, p.y, p.z};
  uint32_t ch[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float t = c[i] / (gamma + c[i]);
    t = t * t * 255.0f;
    ch[i] = (uint32_t)f2i_sat(cl_clamp(t, 0.0f, 255.0f));
  }
  return 0xff000000u | (ch[0] << 16) | (ch[1] << 8) | ch[2];
}

__global__ void __launch_bounds__(256)
k_tonemap_linear(const float4* __restrict__ accum, float gamma, int n, uint32_t* __restrict__ argb) {
  const int id = blockIdx.x * 256 + threadIdx.x;
  if (id < n) argb[id] = tonemap_pack(accum[id], gamma);
}

__global__ void __launch_bounds__(256)
k_tonemap_packed(const float4* __restrict__ accum, float gamma, int W, int H,
                 const __grid_constant__ RmShard sh, uint32_t* __restrict__ argb) {
  const long long slot = (long long)blockIdx.x * 256 + threadIdx.x;
  if (slot >= sh.slots) return;
  const int id = rm_slot_to_pixel(sh, slot, W, H);
  argb[slot] = id >= 0 ? tonemap_pack(accum[id], gamma) : 0u;
}

// only the pixels this shard owns, pixel-indexed: several GPUs fill ONE frame (possibly over NVLink)
__global__ void __launch_bounds__(256)
k_tonemap_owned(const float4* __restrict__ accum, float gamma, int W, int H,
                const __grid_constant__ RmShard sh, uint32_t* __restrict__ argb) {
  const long long slot = (long long)blockIdx.x * 256 + threadIdx.x;
  if (slot >= sh.slots) return;
  const int id = rm_slot_to_pixel(sh, slot, W, H);
  if (id >= 0) argb[id] = tonemap_pack(accum[id], gamma);
}

__global__ void __launch_bounds__(256)
k_pack_accum(const float4* __restrict__ accum, int W, int H, const __grid_constant__ RmShard sh,
             float4* __restrict__ packed) {
  const long long slot = (long long)blockIdx.x * 256 + threadIdx.x;
  if (slot >= sh.slots) return;
  const int id = rm_slot_to_pixel(sh, slot, W, H);
  packed[slot] = id >= 0 ? accum[id] : make_float4(0.f, 0.f, 0.f, 0.f);
}

// De-interleave the gathered per-rank packed buffers (parts[r][slot], r < world) into the frame.
template <typename T>
__global__ void __launch_bounds__(256)
k_unpack_shards(const T* __restrict__ parts, int world, long long stride, int W, int H,
                const __grid_constant__ RmShard sh0, T* __restrict__ frame) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  const int r = blockIdx.y;
  RmShard sh = sh0;
  sh.rank = r;
  sh.world = world;
  rm_shard_layout(sh, W, H);
  if (i >= sh.slots) return;
  const int id = rm_slot_to_pixel(sh, i, W, H);
  if (id >= 0) frame[id] = parts[(size_t)r * stride + i];
}

}  // namespace

cudaError_t rm_launch_unpack_shards(const void* d_parts, int world, long long stride_slots, int elem_bytes, int W, int H,
                                    const RmShard& shard, void* d_frame, cudaStream_t stream) {
  if (world <= 0 || stride_slots <= 0) return cudaErrorInvalidValue;
  const dim3 grid((unsigned)((stride_slots + 255) / 256), (unsigned)world);
  if (elem_bytes == 4)
    k_unpack_shards<uint32_t><<<grid, 256, 0, stream>>>(static_cast<const uint32_t*>(d_parts), world, stride_slots, W, H,
                                                        shard, static_cast<uint32_t*>(d_frame));
  else if (elem_bytes == 16)
    k_unpack_shards<float4><<<grid, 256, 0, stream>>>(static_cast<const float4*>(d_parts), world, stride_slots, W, H, shard,
                                                      static_cast<float4*>(d_frame));
  else
    return cudaErrorInvalidValue;
  return cudaGetLastError();
}

cudaError_t rm_launch_render_plain(const uint8_t* d_vox, const float4* d_table, const RmOpts& opts,
                                   const RmShard& shard, float4* d_accum, RmCounters* d_counters,
                                   cudaStream_t stream) {
  if (shard.slots <= 0) return cudaSuccess;
  const unsigned blocks = (unsigned)((shard.slots + kPlainBlock - 1) / kPlainBlock);
  cudaError_t e = cudaMemcpyToSymbolAsync(plain::g_opts, &opts, sizeof(RmOpts), 0, cudaMemcpyHostToDevice, stream);
  if (e != cudaSuccess) return e;
  if (d_counters)
    k_render_plain<true><<<blocks, kPlainBlock, 0, stream>>>(d_vox, d_table, shard, d_accum, d_counters);
  else
    k_render_plain<false><<<blocks, kPlainBlock, 0, stream>>>(d_vox, d_table, shard, d_accum, nullptr);
  return cudaGetLastError();
}

cudaError_t rm_launch_tonemap(const float4* d_accum, float gamma, int W, int H, const RmShard& shard,
                              uint32_t* d_argb, int packed, cudaStream_t stream) {
  if (packed) {
    if (shard.slots <= 0) return cudaSuccess;
    const unsigned blocks = (unsigned)((shard.slots + 255) / 256);
    if (packed == 2) k_tonemap_owned<<<blocks, 256, 0, stream>>>(d_accum, gamma, W, H, shard, d_argb);
    else k_tonemap_packed<<<blocks, 256, 0, stream>>>(d_accum, gamma, W, H, shard, d_argb);
  } else {
    const int n = W * H;
    k_tonemap_linear<<<(n + 255) / 256, 256, 0, stream>>>(d_accum, gamma, n, d_argb);
  }
  return cudaGetLastError();
}

cudaError_t rm_launch_pack_accum(const float4* d_accum, int W, int H, const RmShard& shard,
                                 float4* d_packed, cudaStream_t stream) {
  if (shard.slots <= 0) return cudaSuccess;
  const unsigned blocks = (unsigned)((shard.slots + 255) / 256);
  k_pack_accum<<<blocks, 256, 0, stream>>>(d_accum, W, H, shard, d_packed);
  return cudaGetLastError();
}
