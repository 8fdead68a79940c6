// rm_render_wave.cu -- RenderImage (renderer.cl:478-494) as a wavefront pipeline (RM_OPT_KERNEL = 3).
//
// The stages of rm_wave.cuh as kernels over chunks of (pixel, pass) items:
//
//   k_wave_primary                       one thread per item: ray setup + primary trace (coherent: the
//                                        32 lanes of a warp are the same pixels in different passes)
//   per level L = 0 .. reflectIter:
//     k_wave_prepare(L)                  one thread per item: [finish level L-1], AO of surface L,
//                                        shadow jobs + the ray of bounce L+1 appended to the job queue
//     k_wave_trace                       the queue's sphere traces. PERSISTENT: each lane keeps pulling
//                                        jobs (warp-aggregated atomic on the queue head) until the queue
//                                        is dry, so a lane whose ray ended early does not wait for the
//                                        longest ray of its warp -- the fused kernel's main loss
//   k_wave_final                         [finish last level], primary surface + atmosphere + exposure,
//                                        colour out exactly like k_render_bricks (colour buffer + blend)
//
// Same arithmetic, same order as the fused kernel (tests/hostsim runs these very stage bodies on the
// host bit for bit against the oracle); only where the work runs differs. Per item the pipeline keeps
// (reflectIter + 1) 64-byte records, 24 bytes of pixel state and at most numLights + 1 queued 32-byte
// jobs, all in HBM scratch sized for one chunk.
#include "rm_kernels.h"
#include "rm_wave.cuh"

namespace {

constexpr int kBlock = 256;

struct WaveParams {
  const float4* tables;              // passes x 16384 float4
  float times[RM_MAX_FUSED_PASSES];  // TRenderOpts.time per pass
  float4* colour;                    // passes x slots (null when passes == 1)
  float4* accum;
  RmCounters* counters;
  int passes;
  long long item0;                   // first item of this chunk
  unsigned nitems;                   // items in this chunk
};

template <bool kCount>
__device__ __forceinline__ void flush_counters(const plain::Scene& s, RmCounters* counters) {
  if (!kCount) return;
  unsigned long long a = s.w.steps, b = s.w.taps, c = s.w.outer;
  for (int off = 16; off > 0; off >>= 1) {
    a += __shfl_down_sync(0xffffffffu, a, off);
    b += __shfl_down_sync(0xffffffffu, b, off);
    c += __shfl_down_sync(0xffffffffu, c, off);
  }
  if ((threadIdx.x & 31) == 0 && (a | b | c)) {
    atomicAdd(&counters->steps, a);
    atomicAdd(&counters->taps, b);
    atomicAdd(&counters->outer, c);
  }
}

// item -> (pixel id, pass); the item order is the fused kernel's (pass-minor)
__device__ __forceinline__ int item_pixel(const RmShard& sh, const WaveParams& P, unsigned it, int& pass) {
  const long long item = P.item0 + it;
  const long long slot = item / P.passes;
  pass = (int)(item - slot * P.passes);
  return rm_slot_to_pixel(sh, slot, plain::g_opts.width, plain::g_opts.height);
}

__device__ __forceinline__ plain::Scene scene_of(const WaveParams& P, int pass) {
  plain::Scene s(plain::g_accel.vox, P.tables + (size_t)pass * (RM_TABLE_MASK + 1));
  s.time = P.times[pass];
  return s;
}

template <bool kCount>
__global__ void __launch_bounds__(kBlock)
k_wave_primary(const __grid_constant__ RmShard sh, const __grid_constant__ WaveParams P, const __grid_constant__ wave::WaveBuf B) {
  const unsigned it = blockIdx.x * kBlock + threadIdx.x;
  plain::Scene s(plain::g_accel.vox, P.tables);
  if (it < P.nitems) {
    int pass;
    const int id = item_pixel(sh, P, it, pass);
    if (id >= 0) {
      s = scene_of(P, pass);
      wave::wave_primary<kCount>(B, it, s, plain::BrickVolume{}, id);
    } else {
      wave::rec_at(B, 0, it).flags = 0;  // padding of an edge tile
    }
  }
  flush_counters<kCount>(s, P.counters);
}

template <bool kCount>
__global__ void __launch_bounds__(kBlock)
k_wave_prepare(const __grid_constant__ RmShard sh, const __grid_constant__ WaveParams P, const __grid_constant__ wave::WaveBuf B, int level) {
  const unsigned it = blockIdx.x * kBlock + threadIdx.x;
  plain::Scene s(plain::g_accel.vox, P.tables);
  if (it < P.nitems) {
    int pass;
    (void)item_pixel(sh, P, it, pass);
    s = scene_of(P, pass);
    if (level >= 2) wave::wave_finish<kCount>(B, it, s, plain::BrickVolume{}, level - 1);
    wave::wave_prepare<kCount>(B, it, s, plain::BrickVolume{}, level);
  }
  flush_counters<kCount>(s, P.counters);
}

// Persistent trace kernel. The grid is a fixed number of blocks; a warp takes batches of kBatch
// consecutive jobs from the queue head (one atomic per batch) and hands them to its lanes as they
// fall idle. One trip of the loop = at most one full distanceToScene evaluation per lane (trace_full), so
// the 32 lanes -- each on its own ray, at its own iteration -- meet once per evaluation; a lane whose
// ray has ended writes its result and takes the next job on the following trip.
constexpr unsigned kBatch = 128;
#ifndef RM_WAVE_CHEAP_PER_TRIP
#define RM_WAVE_CHEAP_PER_TRIP 8  // measured on B200 (C2, 16 M-item chunks): 2 / 4 / 8 / unlimited = 60.1 / 57.4 / 56.2 / 57.0 ms
#endif
constexpr int kCheapPerTrip = RM_WAVE_CHEAP_PER_TRIP;  // ground-only evaluations a lane may run per trip

template <bool kCount>
__global__ void __launch_bounds__(kBlock)
k_wave_trace(const __grid_constant__ wave::WaveBuf B, unsigned* __restrict__ head, RmCounters* counters, int min_idle) {
  const unsigned kFull = 0xffffffffu;
  plain::Scene s(plain::g_accel.vox, nullptr);
  const plain::BrickVolume V{};
  const unsigned n = min(*B.njobs, B.job_cap);
  const unsigned lane = threadIdx.x & 31;
  plain::TraceState t;
  unsigned info = 0;
  bool have = false, dry = false;   // dry: the queue head has passed the end (warp-uniform)
  unsigned wnext = 0, wend = 0;     // the warp's current batch [wnext, wend) (warp-uniform)
  for (;;) {
    // Refill policy: idle lanes take new rays only when at least min_idle lanes are idle. 1 = a lane
    // refills as soon as its ray ends (best balance, but the lanes of a warp drift to unrelated rays
    // and iterations); 32 = the warp starts 32 consecutive rays of the queue together and finishes
    // them together, which keeps neighbouring rays -- same pixels, different passes -- in step.
    unsigned want = __ballot_sync(kFull, !have);
    if ((int)__popc(want) < min_idle) want = 0;
    if (want && wnext >= wend && !dry) {
      unsigned base = 0;
      if (lane == 0) base = atomicAdd(head, kBatch);
      base = __shfl_sync(kFull, base, 0);
      if (base >= n) { dry = true; }
      else { wnext = base; wend = min(base + kBatch, n); }
    }
    if (!have && want && wnext < wend) {
      const unsigned k = wnext + __popc(want & ((1u << lane) - 1u));
      if (k < wend) {
        const wave::WaveJob j = B.jobs[k];
        wave::job_begin(j, t);
        info = j.info;
        have = true;
      }
    }
    wnext = min(wnext + (unsigned)__popc(want), wend);
    if (!__any_sync(kFull, have)) {
      if (dry) break;
      continue;
    }
    // each lane runs through its ground-only evaluations, then the lanes that need a full
    // distanceToScene call make it together
    bool done = false;
    int st = plain::kTraceDone;
    if (have) st = plain::trace_run_cheap<kCount>(s, t, kCheapPerTrip);
    if (have && st == plain::kTraceDone) done = true;
    if (have && st == plain::kTraceNeedsFull) done = plain::trace_full<kCount>(s, V, t);
    if (have && done) {
      wave::job_end<kCount>(B, info, s, V, t);
      have = false;
    }
  }
  flush_counters<kCount>(s, counters);
}

template <bool kCount>
__global__ void __launch_bounds__(kBlock)
k_wave_final(const __grid_constant__ RmShard sh, const __grid_constant__ WaveParams P, const __grid_constant__ wave::WaveBuf B, int lmax) {
  const RmOpts& o = plain::g_opts;
  const unsigned it = blockIdx.x * kBlock + threadIdx.x;
  if (it >= P.nitems) return;
  if (!(wave::rec_at(B, 0, it).flags & wave::kValid)) return;
  int pass;
  const int id = item_pixel(sh, P, it, pass);
  const plain::Scene s = scene_of(P, pass);
  if (lmax >= 1) wave::wave_finish<kCount>(B, it, s, plain::BrickVolume{}, lmax);
  const float3 c = wave::wave_final<kCount>(B, it, s);
  if (P.colour) {
    P.colour[P.item0 + it] = make_float4(c.x, c.y, c.z, 1.0f);  // [slot][pass]
  } else {
    const float4 old = P.accum[id];
    const float3 m = lerp3(f3(old.x, old.y, old.z), c, o.frameBlend);  // mix(), renderer.cl:492
    P.accum[id] = make_float4(m.x, m.y, m.z, 1.0f);
  }
}

}  // namespace

void rm_wave_free(RmWaveScratch* w) {
  cudaFree(w->d_rec); cudaFree(w->d_refl); cudaFree(w->d_pxy); cudaFree(w->d_jobs); cudaFree(w->d_ctr);
  *w = RmWaveScratch{};
}

static cudaError_t wave_reserve(RmWaveScratch* w, unsigned cap, int levels, unsigned job_cap) {
  cudaError_t e;
  if (w->cap < cap || w->levels < levels) {
    cudaFree(w->d_rec); cudaFree(w->d_refl); cudaFree(w->d_pxy);
    w->d_rec = nullptr; w->d_refl = nullptr; w->d_pxy = nullptr; w->cap = 0; w->levels = 0;
    if ((e = cudaMalloc(&w->d_rec, sizeof(wave::WaveRec) * (size_t)levels * cap)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&w->d_refl, sizeof(float4) * (size_t)cap)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&w->d_pxy, sizeof(float2) * (size_t)cap)) != cudaSuccess) return e;
    w->cap = cap; w->levels = levels;
  }
  if (w->job_cap < job_cap) {
    cudaFree(w->d_jobs);
    w->d_jobs = nullptr; w->job_cap = 0;
    if ((e = cudaMalloc(&w->d_jobs, sizeof(wave::WaveJob) * (size_t)job_cap)) != cudaSuccess) return e;
    w->job_cap = job_cap;
  }
  if (!w->d_ctr && (e = cudaMalloc(&w->d_ctr, 2 * sizeof(unsigned))) != cudaSuccess) return e;
  return cudaSuccess;
}

int rm_wave_supports(const RmOpts& opts) { return opts.reflectIter < wave::kMaxLevels && opts.numLights <= 4; }

cudaError_t rm_launch_render_wave(const RmOpts& opts, const RmShard& shard, const RmAccel& accel,
                                  const float4* d_tables, const float* times, const float* blend, int passes,
                                  float4* d_colour, float4* d_accum, RmCounters* d_counters, RmWaveScratch* w,
                                  int num_sms, unsigned chunk_items, int refill_min_idle, int* launches, cudaStream_t stream) {
  if (shard.slots <= 0 || passes <= 0) return cudaSuccess;
  if (passes > RM_MAX_FUSED_PASSES || !rm_wave_supports(opts)) return cudaErrorInvalidValue;
  const long long total = (long long)passes * shard.slots;
  if (chunk_items > (1u << 24)) chunk_items = 1u << 24;   // the item index of a job has 24 bits
  if (chunk_items < 1024) chunk_items = 1024;
  const unsigned cap = (unsigned)(total < (long long)chunk_items ? total : (long long)chunk_items);
  const int lmax = opts.reflectIter < 0 ? 0 : opts.reflectIter;
  const unsigned job_cap = cap * (unsigned)(opts.numLights + 1) + 1;
  const int levels = lmax + 2 < wave::kMaxLevels ? lmax + 2 : wave::kMaxLevels;  // prepare(lmax) clears the flags of level lmax + 1
  cudaError_t e = wave_reserve(w, cap, levels, job_cap);
  if (e != cudaSuccess) return e;
  if ((e = cudaMemcpyToSymbolAsync(plain::g_opts, &opts, sizeof(RmOpts), 0, cudaMemcpyHostToDevice, stream)) != cudaSuccess) return e;
  if ((e = cudaMemcpyToSymbolAsync(plain::g_accel, &accel, sizeof(RmAccel), 0, cudaMemcpyHostToDevice, stream)) != cudaSuccess) return e;

  WaveParams P;
  P.tables = d_tables;
  for (int i = 0; i < RM_MAX_FUSED_PASSES; ++i) P.times[i] = i < passes ? times[i] : 0.0f;
  P.colour = passes > 1 ? d_colour : nullptr;
  P.accum = d_accum;
  P.counters = d_counters;
  P.passes = passes;
  wave::WaveBuf B;
  B.rec = static_cast<wave::WaveRec*>(w->d_rec);
  B.refl = static_cast<float4*>(w->d_refl);
  B.pxy = static_cast<float2*>(w->d_pxy);
  B.jobs = static_cast<wave::WaveJob*>(w->d_jobs);
  B.njobs = w->d_ctr;
  unsigned* head = w->d_ctr + 1;
  B.cap = w->cap;
  B.job_cap = w->job_cap;
  B.levels = levels;
  const unsigned trace_blocks = (unsigned)(num_sms > 0 ? num_sms : 148) * 5u;  // 40 warps per SM

  for (long long item0 = 0; item0 < total; item0 += cap) {
    P.item0 = item0;
    P.nitems = (unsigned)((total - item0) < (long long)cap ? (total - item0) : (long long)cap);
    const unsigned blocks = (P.nitems + kBlock - 1) / kBlock;
    if (d_counters) k_wave_primary<true><<<blocks, kBlock, 0, stream>>>(shard, P, B);
    else k_wave_primary<false><<<blocks, kBlock, 0, stream>>>(shard, P, B);
    *launches += 1;
    for (int L = 0; L <= lmax; ++L) {
      if ((e = cudaMemsetAsync(w->d_ctr, 0, 2 * sizeof(unsigned), stream)) != cudaSuccess) return e;
      if (d_counters) {
        k_wave_prepare<true><<<blocks, kBlock, 0, stream>>>(shard, P, B, L);
        k_wave_trace<true><<<trace_blocks, kBlock, 0, stream>>>(B, head, d_counters, refill_min_idle);
      } else {
        k_wave_prepare<false><<<blocks, kBlock, 0, stream>>>(shard, P, B, L);
        k_wave_trace<false><<<trace_blocks, kBlock, 0, stream>>>(B, head, d_counters, refill_min_idle);
      }
      *launches += 2;
    }
    if (d_counters) k_wave_final<true><<<blocks, kBlock, 0, stream>>>(shard, P, B, lmax);
    else k_wave_final<false><<<blocks, kBlock, 0, stream>>>(shard, P, B, lmax);
    *launches += 1;
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
  }
  if (passes > 1) {
    e = rm_launch_blend_passes(d_colour, blend, passes, shard, opts.width, opts.height, d_accum, stream);
    *launches += 1;
  }
  return e;
}
