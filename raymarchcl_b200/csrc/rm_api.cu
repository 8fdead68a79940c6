// rm_api.cu -- the C ABI of libraymarch_b200.so (include/raymarch_b200.h).
// Replaces the simplecl/JOCL pipeline of /root/reference/src/thi/ng/raymarchcl/core.clj:76-148.
#include "../../include/raymarch_b200.h"

#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "rm_kernels.h"
#include "rm_types.h"

namespace {

thread_local std::string g_create_error;

struct EventPair {
  cudaEvent_t a = nullptr, b = nullptr;
  int kind = 0;  // 0 render, 1 tonemap, 2 h2d, 3 d2h
};

}  // namespace

struct rm_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;      // stream in use
  cudaStream_t own_stream = nullptr;  // the context's own stream
  std::string err;

  // volume ("v-buf")
  uint8_t* d_vox = nullptr;
  size_t vox_capacity = 0;
  int rx = 0, ry = 0, rz = 0;

  // framebuffer ("p-buf", "q-buf")
  float4* d_accum = nullptr;
  uint32_t* d_argb2[2] = {nullptr, nullptr};  // two ARGB frames: one is read back while the next is rendered
  int argb_cur = 0;                           // the one the current frame is written to
  int W = 0, H = 0;
  size_t fb_capacity = 0;  // pixels allocated
  // TonemapImage folded into the render launch (rm_render_persist.cu): where the default kernel
  // writes the ARGB words of the frame so far, and whether that buffer matches the accumulator
  uint32_t* argb_target = nullptr;   // caller-owned device buffer (rm_set_argb_target) or null = d_argb2[argb_cur]
  int argb_target_packed = 0;
  const void* argb_fresh_ptr = nullptr;  // buffer that holds tonemap(accum, argb_fresh_gamma) right now, or null
  int argb_fresh_packed = 0;
  float argb_fresh_gamma = 0.f;
  // asynchronous read-back (rm_tonemap_async / rm_wait)
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t frame_ready[2] = {nullptr, nullptr};  // main stream: ARGB buffer b complete
  cudaEvent_t copy_done[2] = {nullptr, nullptr};    // copy stream: read-back of buffer b complete
  bool copy_pending[2] = {false, false};
  int slot_buffer[2] = {-1, -1};                    // which ARGB buffer the read-back of slot s uses

  // per-pass inputs
  float4* d_tables = nullptr;  // resident scatter tables, 16384 float4 each
  int table_capacity = 0;
  std::vector<RmOpts> passes;  // decoded resident opts
  int resident = 0;            // passes uploaded by rm_upload_passes
  int generated_tables = 0;    // tables produced in place by rm_generate_scatter_tables

  RmShard shard{};
  int shard_rank = 0, shard_world = 1, shard_tw = 32, shard_th = 32;

  RmCounters* d_counters = nullptr;
  int count_work = 0;
  int kernel_kind = 0;

  // fast path state (rm_accel.cu, rm_render_fast.cu)
  RmAccelStorage accel;
  float4* d_colour = nullptr;       // per-pass colours of one fused launch
  size_t colour_capacity = 0;       // float4 elements
  int num_sms = 0;
  cudaEvent_t launch_done = nullptr;  // completion of this context's last render launch (DeviceGuard)
  unsigned long long* d_queue = nullptr;  // [0] work queue head of the warp kernel, [1] bundle tickets of the default kernel
  int persist_block = 0;                  // threads of the default kernel's block; 0 = library default
  int persist_bottom_up = 1;              // RM_OPT_PERSIST_ORDER
  int persist_smem = RM_PERSIST_DEFAULT_SMEM;  // stage the 4-bit distance map into shared memory by bulk TMA: 0 never, 1 when it fits, 2 auto (RM_OPT_PERSIST_SMEM);
  int persist_group = -1;                 // 1 = block-synchronous rounds of the default kernel; 0 = free-running warps; -1 = default
  unsigned* d_watchdog = nullptr;         // 16 words, see rm_launch_render_warp
  unsigned trip_limit = 1u << 28;
  int warp_blocks[2] = {0, 0};            // resident blocks per SM of the warp kernel [plain, counting]
  RmWaveScratch wave;                     // scratch of the wavefront path (kernel 3)
  unsigned wave_chunk = 1u << 24;         // items per chunk of the wavefront path
  int wave_refill = 16;                   // its trace kernel refills when this many lanes of a warp are idle (B200, C2: 1 / 16 / 28 / 32 = 56.3 / 54.4 / 55.2 / 54.6 ms)
  int cell_shift_opt = 0;           // 0 = auto
  int fuse_limit = RM_MAX_FUSED_PASSES;

  rm_stats stats{};
  std::vector<EventPair> pending, free_events;

  // rm_create_multi: a group context owns one member context per GPU and forwards every call
  std::vector<rm_ctx*> members;
  bool is_group = false;
};

// group (multi-GPU) forms of the public calls, defined at the end of this file
namespace grp {
int set_volume(rm_ctx* g, const uint8_t* voxels, int rx, int ry, int rz);
int set_volume_device(rm_ctx* g, const void* d_voxels, int rx, int ry, int rz);
int load_volume_file(rm_ctx* g, const char* path, int* rx, int* ry, int* rz);
int generate_volume(rm_ctx* g, int kind, int rx, int ry, int rz);
int voxelize_points(rm_ctx* g, const float* xyz, int64_t n, int res, int ks);
int clear_accum(rm_ctx* g, int w, int h);
int render_frame(rm_ctx* g, const void* const* opts, const float* const* mc, int iter);
int tonemap(rm_ctx* g, const void* opts, size_t len, uint32_t* out, int async_slot);
int wait(rm_ctx* g, int slot);
int read_accum(rm_ctx* g, float* out);
int get_stats(rm_ctx* g, rm_stats* out);
int set_tile(rm_ctx* g, int tile_w, int tile_h);
void destroy(rm_ctx* g);
int member_failed(rm_ctx* g, rm_ctx* m, int rc);
int unsupported(rm_ctx* g, const char* what);
}  // namespace grp

// Forwards a call to every member of a group; `m` names the member inside CALL. Calls that are
// asynchronous on one GPU stay asynchronous, so the GPUs of a group work concurrently.
#define RM_FORWARD(c, CALL)                                        \
  do {                                                             \
    if ((c)->is_group) {                                           \
      for (rm_ctx* m : (c)->members) {                             \
        const int rc__ = (CALL);                                   \
        if (rc__) return grp::member_failed((c), m, rc__);         \
      }                                                            \
      return RM_OK;                                                \
    }                                                              \
  } while (0)


namespace {

int fail(rm_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->err = msg; else g_create_error = msg;
  return code;
}

int cuda_fail(rm_ctx* ctx, cudaError_t e, const char* what) {
  return fail(ctx, RM_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

#define RM_CUDA(ctx, call)                                  \
  do {                                                      \
    cudaError_t e__ = (call);                               \
    if (e__ != cudaSuccess) return cuda_fail(ctx, e__, #call); \
  } while (0)

float rd_f(const uint8_t* b, int off) { float f; std::memcpy(&f, b + off, 4); return f; }
int rd_i(const uint8_t* b, int off) { int32_t i; std::memcpy(&i, b + off, 4); return i; }
float3 rd_f3(const uint8_t* b, int off) { return make_float3(rd_f(b, off), rd_f(b, off + 4), rd_f(b, off + 8)); }

// TRenderOpts blob -> RmOpts. Offsets: OpenCL layout of renderer.cl:35-78 (float3/int3 = 16 B).
void decode_opts(const void* blob, RmOpts* o) {
  const uint8_t* b = static_cast<const uint8_t*>(blob);
  o->eyePos = rd_f3(b, 0);          o->targetPos = rd_f3(b, 16);   o->up = rd_f3(b, 32);
  o->voxelBounds = rd_f3(b, 48);    o->voxelBounds2 = rd_f3(b, 64);
  o->boundsMin = rd_f3(b, 80);      o->boundsMax = rd_f3(b, 96);
  o->invVoxelScale = rd_f3(b, 112); o->sky1 = rd_f3(b, 128);       o->sky2 = rd_f3(b, 144);
  o->rx = rd_i(b, 160); o->ry = rd_i(b, 164); o->rz = rd_i(b, 168); o->rxy = rd_i(b, 172);
  o->width = rd_i(b, 176); o->height = rd_i(b, 180);
  o->invAspect = rd_f(b, 184); o->time = rd_f(b, 188); o->fov = rd_f(b, 192);
  o->maxIter = rd_i(b, 196); o->maxVoxelIter = rd_i(b, 200);
  o->maxDist = rd_f(b, 204); o->startDist = rd_f(b, 208); o->eps = rd_f(b, 212);
  o->aoIter = rd_i(b, 216);
  o->aoStepDist = rd_f(b, 220); o->aoAmp = rd_f(b, 224); o->voxelSize = rd_f(b, 228);
  o->groundY = rd_f(b, 232);
  o->shadowIter = rd_i(b, 236); o->reflectIter = rd_i(b, 240);
  o->shadowBias = rd_f(b, 244); o->lightScatter = rd_f(b, 248); o->minLightAtt = rd_f(b, 252);
  o->gamma = rd_f(b, 256); o->exposure = rd_f(b, 260); o->dof = rd_f(b, 264);
  o->frameBlend = rd_f(b, 268); o->fogPow = rd_f(b, 272); o->flareAmp = rd_f(b, 276);
  o->isoVal = b[284]; o->numLights = b[285];
  for (int i = 0; i < 4; ++i) {
    o->lightPos[i] = rd_f3(b, 288 + 16 * i);
    o->lightColor[i] = rd_f3(b, 352 + 16 * i);
    o->mat[i].albedo = rd_f3(b, 416 + 32 * i);
    o->mat[i].r0 = rd_f(b, 416 + 32 * i + 16);
    o->mat[i].smoothness = rd_f(b, 416 + 32 * i + 20);
  }
  rm_derive_opts(o);
}

int check_opts(rm_ctx* ctx, const RmOpts& o) {
  char msg[256];
  if (o.rx != ctx->rx || o.ry != ctx->ry || o.rz != ctx->rz || o.rxy != ctx->rx * ctx->ry) {
    std::snprintf(msg, sizeof msg, "TRenderOpts.voxelRes (%d,%d,%d,%d) does not match the uploaded volume %dx%dx%d",
                  o.rx, o.ry, o.rz, o.rxy, ctx->rx, ctx->ry, ctx->rz);
    return fail(ctx, RM_ERR_BAD_OPTS, msg);
  }
  if (o.width != ctx->W || o.height != ctx->H) {
    std::snprintf(msg, sizeof msg, "TRenderOpts.resolution (%d,%d) does not match the framebuffer %dx%d",
                  o.width, o.height, ctx->W, ctx->H);
    return fail(ctx, RM_ERR_BAD_OPTS, msg);
  }
  if (o.numLights > 4) return fail(ctx, RM_ERR_BAD_OPTS, "TRenderOpts.numLights > 4");
  if (o.maxVoxelIter < 0 || o.maxIter < 0 || o.shadowIter < 0)
    return fail(ctx, RM_ERR_BAD_OPTS, "TRenderOpts iteration limits must be >= 0");
  return RM_OK;
}

void update_shard(rm_ctx* c) {
  RmShard& s = c->shard;
  s.rank = c->shard_rank; s.world = c->shard_world;
  s.tile_w = c->shard_tw; s.tile_h = c->shard_th;
  rm_shard_layout(s, c->W, c->H);
}

int ensure_tables(rm_ctx* c, int n) {
  if (c->table_capacity >= n) return RM_OK;
  if (c->d_tables) cudaFree(c->d_tables);
  c->d_tables = nullptr; c->table_capacity = 0;
  c->generated_tables = 0;  // whatever was generated in place lived in the old allocation
  c->resident = 0;
  RM_CUDA(c, cudaMalloc(&c->d_tables, (size_t)n * RM_TABLE_FLOATS * sizeof(float)));
  c->table_capacity = n;
  return RM_OK;
}

EventPair begin_timed(rm_ctx* c, int kind) {
  EventPair p;
  if (!c->free_events.empty()) { p = c->free_events.back(); c->free_events.pop_back(); }
  else { cudaEventCreate(&p.a); cudaEventCreate(&p.b); }
  p.kind = kind;
  cudaEventRecord(p.a, c->stream);
  return p;
}
void end_timed(rm_ctx* c, EventPair p) {
  cudaEventRecord(p.b, c->stream);
  c->pending.push_back(p);
}
void resolve_timers(rm_ctx* c) {
  for (EventPair& p : c->pending) {
    float ms = 0.f;
    if (cudaEventSynchronize(p.b) == cudaSuccess && cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
      switch (p.kind) {
        case 0: c->stats.render_ms += ms; break;
        case 1: c->stats.tonemap_ms += ms; break;
        case 2: c->stats.h2d_ms += ms; break;
        default: c->stats.d2h_ms += ms; break;
      }
    }
    c->free_events.push_back(p);
  }
  c->pending.clear();
}

// The render kernels read their per-launch constants from __constant__ memory (one symbol per
// device). Launches from one stream are ordered by the stream; launches from DIFFERENT streams
// on the same device (two contexts, or a context whose stream was swapped) are ordered here:
// the newcomer's stream first waits for the previous launch to finish.
struct DeviceGuard {
  std::mutex mu;
  rm_ctx* owner = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t done = nullptr;
};
DeviceGuard g_guards[64];

void guard_before_launch(rm_ctx* c) {
  DeviceGuard& g = g_guards[c->device & 63];
  g.mu.lock();
  if (g.owner && g.done && g.stream != c->stream) cudaStreamWaitEvent(c->stream, g.done, 0);
}

void guard_after_launch(rm_ctx* c) {
  DeviceGuard& g = g_guards[c->device & 63];
  if (!c->launch_done) cudaEventCreateWithFlags(&c->launch_done, cudaEventDisableTiming);
  if (c->launch_done) cudaEventRecord(c->launch_done, c->stream);
  g.owner = c;
  g.stream = c->stream;
  g.done = c->launch_done;
  g.mu.unlock();
}

void guard_forget(rm_ctx* c) {
  DeviceGuard& g = g_guards[c->device & 63];
  std::lock_guard<std::mutex> lock(g.mu);
  if (g.owner == c) { g.owner = nullptr; g.done = nullptr; g.stream = nullptr; }
}

// Macro-cell edge of the distance map: one 4 x 4 x 4 brick (the cell index is then the brick index, the cheapest form of the
// march loop, and skips are decided at the finest grain) as long as the byte map stays <= 32 MiB, i.e. up to 1024^3 voxels;
// beyond that the smallest cell that fits. Round 1 kept the map at ~64^3 cells (8-voxel cells at 512^3, 16 at 1024^3) so that it
// stayed L1-resident; with the leaner march loop of round 2 the finer map wins although it lives in L2 and cannot be staged
// into shared memory: C3 (512^3) 29.67 vs 31.47 ms, C5 (1024^3) 38.46 vs 44.05 (profiles/r02_scheduling_ab.md 13).
int auto_cell_shift(int rx, int ry, int rz) {
  int shift = 2;
  for (; shift < 6; ++shift) {
    const long long mx = (rx + (1 << shift) - 1) >> shift, my = (ry + (1 << shift) - 1) >> shift, mz = (rz + (1 << shift) - 1) >> shift;
    if (mx * my * mz <= (1LL << 25)) break;
  }
  return shift;
}

int ensure_accel(rm_ctx* c, int iso) {
  if (c->accel.valid && c->accel.iso == iso) return RM_OK;
  const int shift = c->cell_shift_opt > 0 ? c->cell_shift_opt : auto_cell_shift(c->rx, c->ry, c->rz);
  cudaError_t e = rm_accel_build(c->d_vox, c->rx, c->ry, c->rz, iso, shift, &c->accel, c->stream);
  if (e != cudaSuccess) return cuda_fail(c, e, "acceleration data build");
  c->stats.kernel_launches += c->accel.launches;
  return RM_OK;
}

// opts of two passes may share one fused launch when they differ only in time / frameBlend
bool fusable(const RmOpts& a, const RmOpts& b) {
  RmOpts x = a, y = b;
  x.time = y.time = 0.0f;
  x.frameBlend = y.frameBlend = 0.0f;
  return std::memcmp(&x, &y, sizeof(RmOpts)) == 0;
}

// Scope of one timed, guarded launch: takes DeviceGuard.mu and an event pair on construction and
// releases both on EVERY way out of the scope.
struct LaunchScope {
  rm_ctx* c;
  EventPair t;
  explicit LaunchScope(rm_ctx* ctx) : c(ctx), t(begin_timed(ctx, 0)) { guard_before_launch(c); }
  ~LaunchScope() {
    guard_after_launch(c);
    end_timed(c, t);
  }
  LaunchScope(const LaunchScope&) = delete;
  LaunchScope& operator=(const LaunchScope&) = delete;
};

// the ARGB buffer the default kernel writes while it renders (null: none)
uint32_t* fused_argb_target(rm_ctx* c, int* packed) {
  if (c->argb_target) { *packed = c->argb_target_packed; return c->argb_target; }
  *packed = 0;
  // the context's own frame is pixel-indexed: only complete when this context owns every pixel
  return c->shard.world == 1 ? c->d_argb2[c->argb_cur] : nullptr;
}

// RenderImage for passes [0, n) whose tables are contiguous at d_tables, in submission order.
int launch_passes(rm_ctx* c, const RmOpts* passes, int n, const float4* d_tables) {
  const size_t tstride = RM_TABLE_FLOATS / 4;
  RmCounters* cnt = c->count_work ? c->d_counters : nullptr;
  c->argb_fresh_ptr = nullptr;  // the accumulator is about to change
  if (c->kernel_kind == 1) {
    for (int i = 0; i < n; ++i) {
      cudaError_t e;
      {
        LaunchScope scope(c);
        e = rm_launch_render_plain(c->d_vox, d_tables + i * tstride, passes[i], c->shard, c->d_accum, cnt, c->stream);
      }
      if (e != cudaSuccess) return cuda_fail(c, e, "render kernel launch");
      c->stats.kernel_launches += 1;
      c->stats.render_launches += 1;
    }
  } else {
    int warp_blocks = 0;
    if (c->kernel_kind == 2) {  // resolved before any lock or event is taken
      const int variant = cnt ? 1 : 0;
      if (!c->warp_blocks[variant]) c->warp_blocks[variant] = rm_warp_blocks_per_sm(variant);
      if (c->warp_blocks[variant] <= 0) return fail(c, RM_ERR_CUDA, "warp render kernel does not fit on an SM");
      warp_blocks = c->warp_blocks[variant];
    }
    int i = 0;
    while (i < n) {
      int m = 1;
      while (i + m < n && m < c->fuse_limit && fusable(passes[i], passes[i + m])) ++m;
      if (c->kernel_kind == 0) m = rm_persist_pick_passes(m);
      int rc = ensure_accel(c, passes[i].isoVal);
      if (rc) return rc;
      if (m > 1 && c->kernel_kind != 0) {
        const size_t need = (size_t)m * (size_t)c->shard.slots;
        if (need > c->colour_capacity) {
          RM_CUDA(c, cudaStreamSynchronize(c->stream));
          cudaFree(c->d_colour);
          c->d_colour = nullptr; c->colour_capacity = 0;
          RM_CUDA(c, cudaMalloc(&c->d_colour, need * sizeof(float4)));
          c->colour_capacity = need;
        }
      }
      float times[RM_MAX_FUSED_PASSES], blend[RM_MAX_FUSED_PASSES];
      for (int k = 0; k < m; ++k) { times[k] = passes[i + k].time; blend[k] = passes[i + k].frameBlend; }
      cudaError_t e;
      int launched = m > 1 ? 2 : 1;
      {
        LaunchScope scope(c);
        if (c->kernel_kind == 2) {
          e = rm_launch_render_warp(passes[i], c->shard, c->accel.view, d_tables + i * tstride, times, m, c->d_colour,
                                    c->d_accum, c->d_queue, cnt, c->d_watchdog, c->trip_limit, warp_blocks * c->num_sms, c->stream);
          if (e == cudaSuccess && m > 1)
            e = rm_launch_blend_passes(c->d_colour, blend, m, c->shard, c->W, c->H, c->d_accum, c->stream);
        } else if (c->kernel_kind == 3 && rm_wave_supports(passes[i])) {
          launched = 0;
          e = rm_launch_render_wave(passes[i], c->shard, c->accel.view, d_tables + i * tstride, times, blend, m,
                                    c->d_colour, c->d_accum, cnt, &c->wave, c->num_sms, c->wave_chunk, c->wave_refill, &launched, c->stream);
        } else if (c->kernel_kind == 0) {
          int packed = 0;
          uint32_t* argb = fused_argb_target(c, &packed);
          // defaults (measured on B200): layout and map location picked per launch, free-running warps
          const int persist_block = c->persist_block ? c->persist_block : RM_PERSIST_DEFAULT_BLOCK;
          const int persist_group = c->persist_group >= 0 ? c->persist_group : RM_PERSIST_DEFAULT_ROUND;
          e = rm_launch_render_persist(passes[i], c->shard, c->accel.view, d_tables + i * tstride, times, blend, m, c->d_accum,
                                       argb, packed, cnt, c->d_queue + 1, c->num_sms, persist_block, persist_group, c->persist_smem, c->persist_bottom_up, c->stream);
          launched = 1;
          if (e == cudaSuccess && argb) {
            c->argb_fresh_ptr = argb;
            c->argb_fresh_packed = packed;
            c->argb_fresh_gamma = passes[i].gamma;
          }
        } else {
          e = rm_launch_render_fast(passes[i], c->shard, c->accel.view, d_tables + i * tstride, times, blend, m,
                                    c->d_colour, c->d_accum, cnt, c->stream);
        }
      }
      if (e != cudaSuccess) return cuda_fail(c, e, "render kernel launch");
      c->stats.kernel_launches += launched;
      c->stats.render_launches += 1;
      i += m;
    }
  }
  c->stats.pixel_samples += (uint64_t)rm_shard_pixels(c) * (uint64_t)n;
  if (c->pending.size() > 512) resolve_timers(c);
  return RM_OK;
}

// After a synchronisation point: did a warp of the warp kernel give up (watchdog)?
int check_watchdog(rm_ctx* c) {
  if (c->kernel_kind != 2) return RM_OK;
  unsigned w[16];
  RM_CUDA(c, cudaMemcpy(w, c->d_watchdog, sizeof w, cudaMemcpyDeviceToHost));
  if (!w[0]) return RM_OK;
  cudaMemset(c->d_watchdog, 0, sizeof w);
  char msg[384];
  std::snprintf(msg, sizeof msg,
                "render kernel watchdog: a warp exceeded %u trips (lane state %u sub %u trace %u consumer %u rem %d iters %d "
                "pixel %u item %u masks idle %08x march %08x job %08x shade %08x exhausted %u block %u thread %u)",
                c->trip_limit, w[1], w[2], w[3], w[4], (int)w[5], (int)w[6], w[7], w[8], w[9], w[10], w[11], w[12], w[13],
                w[14], w[15]);
  return fail(c, RM_ERR_CUDA, msg);
}

// Make room for a volume of `bytes` bytes that is about to be (re)written. From here until
// commit_volume() the context has no usable volume: the occupancy data is invalid and the extents
// are cleared, so a failure half way leaves nothing stale behind (a render then fails with
// RM_ERR_NO_VOLUME / RM_ERR_BAD_OPTS instead of reading freed or half-written memory).
int begin_volume(rm_ctx* c, size_t bytes, const char* who) {
  if (bytes / 64 > 0x7fffffffULL) return fail(c, RM_ERR_INVALID_ARG, std::string(who) + ": more than 2^37 voxels");
  c->accel.valid = false;
  c->rx = c->ry = c->rz = 0;
  if (bytes > c->vox_capacity) {  // the allocation is kept across re-uploads (the reference re-uploads every frame)
    RM_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->d_vox) { cudaFree(c->d_vox); c->d_vox = nullptr; c->vox_capacity = 0; }
    RM_CUDA(c, cudaMalloc(&c->d_vox, bytes));
    c->vox_capacity = bytes;
  }
  return RM_OK;
}

void commit_volume(rm_ctx* c, int rx, int ry, int rz) {
  c->rx = rx; c->ry = ry; c->rz = rz;
  c->accel.valid = false;
}

int require_ready(rm_ctx* c) {
  if (!c->d_vox || c->rx <= 0) return fail(c, RM_ERR_NO_VOLUME, "no volume uploaded (rm_set_volume)");
  if (!c->d_accum) return fail(c, RM_ERR_NO_FRAMEBUFFER, "no framebuffer (rm_clear_accum)");
  return RM_OK;
}

}  // namespace

extern "C" {

int rm_abi_version(void) { return RM_ABI_VERSION; }

int rm_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

const char* rm_last_error(const rm_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int rm_create(int device_id, rm_ctx** out_ctx) {
  if (!out_ctx) return fail(nullptr, RM_ERR_INVALID_ARG, "rm_create: out_ctx is null");
  *out_ctx = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(nullptr, RM_ERR_NO_DEVICE, std::string("no CUDA device: ") + cudaGetErrorString(e));
  }
  if (device_id < 0 || device_id >= n) return fail(nullptr, RM_ERR_INVALID_ARG, "rm_create: device_id out of range");
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device_id)) != cudaSuccess) return cuda_fail(nullptr, e, "cudaGetDeviceProperties");
  if (prop.major != 10) {
    char msg[400];
    std::snprintf(msg, sizeof msg, "device %d (%s) is sm_%d%d; this library carries sm_100a code only",
                  device_id, prop.name, prop.major, prop.minor);
    return fail(nullptr, RM_ERR_NO_DEVICE, msg);
  }
  rm_ctx* c = new (std::nothrow) rm_ctx();
  if (!c) return fail(nullptr, RM_ERR_INVALID_ARG, "out of host memory");
  c->device = device_id;
  if ((e = cudaSetDevice(device_id)) != cudaSuccess ||
      (e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = cudaMalloc(&c->d_counters, sizeof(RmCounters))) != cudaSuccess ||
      (e = cudaMemset(c->d_counters, 0, sizeof(RmCounters))) != cudaSuccess ||
      (e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = cudaMalloc(&c->d_queue, 2 * sizeof(unsigned long long))) != cudaSuccess ||
      (e = cudaMemset(c->d_queue, 0, 2 * sizeof(unsigned long long))) != cudaSuccess ||
      (e = cudaMalloc(&c->d_watchdog, 16 * sizeof(unsigned))) != cudaSuccess ||
      (e = cudaMemset(c->d_watchdog, 0, 16 * sizeof(unsigned))) != cudaSuccess) {
    int rc = cuda_fail(nullptr, e, "rm_create");
    rm_destroy(c);
    return rc;
  }
  for (int b = 0; b < 2; ++b) {
    cudaEventCreateWithFlags(&c->frame_ready[b], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->copy_done[b], cudaEventDisableTiming);
  }
  c->stream = c->own_stream;
  c->num_sms = prop.multiProcessorCount;
  update_shard(c);
  *out_ctx = c;
  return RM_OK;
}

void rm_destroy(rm_ctx* c) {
  if (!c) return;
  if (c->is_group) { grp::destroy(c); return; }
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
  for (int b = 0; b < 2; ++b) {
    if (c->frame_ready[b]) cudaEventDestroy(c->frame_ready[b]);
    if (c->copy_done[b]) cudaEventDestroy(c->copy_done[b]);
  }
  guard_forget(c);
  if (c->launch_done) cudaEventDestroy(c->launch_done);
  resolve_timers(c);
  for (EventPair& p : c->free_events) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
  cudaFree(c->d_vox); cudaFree(c->d_accum); cudaFree(c->d_argb2[0]); cudaFree(c->d_argb2[1]); cudaFree(c->d_tables); cudaFree(c->d_counters);
  cudaFree(c->d_colour); cudaFree(c->d_queue); cudaFree(c->d_watchdog);
  rm_wave_free(&c->wave);
  rm_accel_free(&c->accel);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  delete c;
}

int rm_set_volume(rm_ctx* c, const uint8_t* voxels, int rx, int ry, int rz) {
  if (!c) return RM_ERR_INVALID_ARG;
  if (c->is_group) return grp::set_volume(c, voxels, rx, ry, rz);
  if (!voxels || rx <= 0 || ry <= 0 || rz <= 0) return fail(c, RM_ERR_INVALID_ARG, "rm_set_volume: null volume or non-positive extent");
  if ((long long)rx * ry > 0x7fffffffLL) return fail(c, RM_ERR_INVALID_ARG, "rm_set_volume: rx*ry overflows int (voxelRes.w)");
  RM_CUDA(c, cudaSetDevice(c->device));
  const size_t bytes = (size_t)rx * ry * rz;
  int rc = begin_volume(c, bytes, "rm_set_volume");
  if (rc) return rc;
  EventPair t = begin_timed(c, 2);
  cudaError_t e = cudaMemcpyAsync(c->d_vox, voxels, bytes, cudaMemcpyHostToDevice, c->stream);
  end_timed(c, t);
  if (e != cudaSuccess) return cuda_fail(c, e, "volume upload");
  RM_CUDA(c, cudaStreamSynchronize(c->stream));
  c->stats.h2d_bytes += bytes;
  commit_volume(c, rx, ry, rz);
  return RM_OK;
}

// The same from DEVICE memory (e.g. a volume assembled by an all-gather over NVLink, or produced by
// another kernel): one device-to-device copy on the context's stream, no host traffic.
int rm_set_volume_device(rm_ctx* c, const void* d_voxels, int rx, int ry, int rz) {
  if (!c) return RM_ERR_INVALID_ARG;
  if (c->is_group) return grp::set_volume_device(c, d_voxels, rx, ry, rz);
  if (!d_voxels || rx <= 0 || ry <= 0 || rz <= 0) return fail(c, RM_ERR_INVALID_ARG, "rm_set_volume_device: null volume or non-positive extent");
  if ((long long)rx * ry > 0x7fffffffLL) return fail(c, RM_ERR_INVALID_ARG, "rm_set_volume_device: rx*ry overflows int (voxelRes.w)");
  RM_CUDA(c, cudaSetDevice(c->device));
  const size_t bytes = (size_t)rx * ry * rz;
  int rc = begin_volume(c, bytes, "rm_set_volume_device");
  if (rc) return rc;
  RM_CUDA(c, cudaMemcpyAsync(c->d_vox, d_voxels, bytes, cudaMemcpyDefault, c->stream));
  commit_volume(c, rx, ry, rz);
  return RM_OK;
}

// .vox reader: "VOXEL", 3 x int32 big-endian, 1 byte element size, raw bytes x-fastest
// (save-volume / load-volume, io.clj:9-33).
int rm_load_volume_file(rm_ctx* c, const char* path, int* out_rx, int* out_ry, int* out_rz) {
  if (!c) return RM_ERR_INVALID_ARG;
  if (c->is_group) return grp::load_volume_file(c, path, out_rx, out_ry, out_rz);
  if (!path) return fail(c, RM_ERR_INVALID_ARG, "rm_load_volume_file: null path");
  std::FILE* f = std::fopen(path, "rb");
  if (!f) return fail(c, RM_ERR_IO, std::string("rm_load_volume_file: cannot open ") + path);
  unsigned char head[18];
  if (std::fread(head, 1, sizeof head, f) != sizeof head || std::memcmp(head, "VOXEL", 5) != 0) {
    std::fclose(f);
    return fail(c, RM_ERR_IO, std::string(path) + ": not a VOXEL file");
  }
  auto be32 = [&](int off) {
    return (int)(((uint32_t)head[off] << 24) | ((uint32_t)head[off + 1] << 16) | ((uint32_t)head[off + 2] << 8) | head[off + 3]);
  };
  const int rx = be32(5), ry = be32(9), rz = be32(13), esize = head[17];
  if (esize != 1 || rx <= 0 || ry <= 0 || rz <= 0 || (long long)rx * ry > 0x7fffffffLL) {
    std::fclose(f);
    return fail(c, RM_ERR_IO, std::string(path) + ": unsupported header (element size must be 1, extents positive)");
  }
  const size_t bytes = (size_t)rx * ry * rz;
  void* host = nullptr;
  cudaError_t e = cudaSetDevice(c->device);
  if (e == cudaSuccess) e = cudaMallocHost(&host, bytes);  // pinned: the upload then runs at full PCIe rate
  if (e != cudaSuccess) { std::fclose(f); return cuda_fail(c, e, "rm_load_volume_file: pinned staging"); }
  const size_t got = std::fread(host, 1, bytes, f);
  std::fclose(f);
  int rc = RM_OK;
  if (got != bytes) rc = fail(c, RM_ERR_IO, std::string(path) + ": truncated voxel data");
  else rc = rm_set_volume(c, static_cast<const uint8_t*>(host), rx, ry, rz);
  cudaFreeHost(host);
  if (rc == RM_OK) {
    if (out_rx) *out_rx = rx;
    if (out_ry) *out_ry = ry;
    if (out_rz) *out_rz = rz;
  }
  return rc;
}

// make-gyroid-volume on the device (generators.clj:27-42); replaces rm_set_volume for that volume.
int rm_generate_gyroid_volume(rm_ctx* c, int rx, int ry, int rz) {
  if (!c) return RM_ERR_INVALID_ARG;
  if (c->is_group) return grp::generate_volume(c, 0, rx, ry, rz);
  if (rx <= 0 || ry <= 0 || rz <= 0 || (long long)rx * ry > 0x7fffffffLL)
    return fail(c, RM_ERR_INVALID_ARG, "rm_generate_gyroid_volume: bad extents");
  RM_CUDA(c, cudaSetDevice(c->device));
  const size_t bytes = (size_t)rx * ry * rz;
  int rc = begin_volume(c, bytes, "rm_generate_gyroid_volume");
  if (rc) return rc;
  double* d_trig = nullptr;
  RM_CUDA(c, cudaMalloc(&d_trig, sizeof(double) * 2 * ((size_t)rx + ry + rz)));
  cudaError_t e = rm_launch_gyroid(rx, ry, rz, d_trig, c->d_vox, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  cudaFree(d_trig);
  if (e != cudaSuccess) return cuda_fail(c, e, "gyroid generator");
  c->stats.kernel_launches += 4;
  commit_volume(c, rx, ry, rz);
  return RM_OK;
}

// gen/make-terrain (generators.clj:44-60) on the device.
int rm_generate_terrain_volume(rm_ctx* c, int rx, int ry, int rz) {
  if (!c) return RM_ERR_INVALID_ARG;
  if (c->is_group) return grp::generate_volume(c, 1, rx, ry, rz);
  if (rx <= 0 || ry <= 0 || rz < rx || (long long)rx * ry > 0x7fffffffLL)  // the reference's second wall needs rz >= rx
    return fail(c, RM_ERR_INVALID_ARG, "rm_generate_terrain_volume: bad extents");
  RM_CUDA(c, cudaSetDevice(c->device));
  const size_t bytes = (size_t)rx * ry * rz;
  int rc = begin_volume(c, bytes, "rm_generate_terrain_volume");
  if (rc) return rc;
  double* d_trig = nullptr;
  RM_CUDA(c, cudaMalloc(&d_trig, sizeof(double) * 2 * ((size_t)rx + ry + rz)));
  cudaError_t e = rm_launch_terrain(rx, ry, rz, d_trig, c->d_vox, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  cudaFree(d_trig);
  if (e != cudaSuccess) return cuda_fail(c, e, "terrain generator");
  c->stats.kernel_launches += 3;
  commit_volume(c, rx, ry, rz);
  return RM_OK;
}


// meshvoxel/voxelize and voxelize-ks (meshvoxel.clj:45-69) on the device: the points (mesh vertices)
// are scaled into the res^3 grid like mesh-scale (:16-23) and splatted with value 255.
int rm_voxelize_points(rm_ctx* c, const float* xyz, int64_t n_points, int res, int ks) {
  if (!c) return RM_ERR_INVALID_ARG;
  if (c->is_group) return grp::voxelize_points(c, xyz, n_points, res, ks);
  if (!xyz || n_points <= 0 || n_points > (1ll << 40)) return fail(c, RM_ERR_INVALID_ARG, "rm_voxelize_points: null points or bad count");
  if (res <= 0 || res > 2048 || ks > 64) return fail(c, RM_ERR_INVALID_ARG, "rm_voxelize_points: res must be 1..2048 and ks <= 64");
  // the splat kernel runs one thread per (point, x-offset of the dilation cube): its 256-thread block count must fit a grid
  if ((long long)n_points * (2 * (ks < 0 ? 0 : ks) + 1) > 256ll * 0x7fffffffLL)
    return fail(c, RM_ERR_INVALID_ARG, "rm_voxelize_points: n_points * (2*ks + 1) exceeds 2^39 splat threads");
  RM_CUDA(c, cudaSetDevice(c->device));
  const size_t bytes = (size_t)res * res * res;
  int rc = begin_volume(c, bytes, "rm_voxelize_points");
  if (rc) return rc;
  float* d_xyz = nullptr;
  int* d_bb = nullptr;
  const size_t pbytes = (size_t)n_points * 3 * sizeof(float);
  RM_CUDA(c, cudaMalloc(&d_xyz, pbytes));
  cudaError_t e = cudaMalloc(&d_bb, 8 * sizeof(int));
  int bad = 0;
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_xyz, xyz, pbytes, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) e = rm_launch_voxelize_points(d_xyz, (long long)n_points, res, ks, d_bb, c->d_vox, &bad, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  cudaFree(d_xyz);
  cudaFree(d_bb);
  if (e != cudaSuccess) return cuda_fail(c, e, "point voxeliser");
  if (bad) return fail(c, RM_ERR_INVALID_ARG, "rm_voxelize_points: NaN or infinite coordinate");
  commit_volume(c, res, res, res);
  c->stats.h2d_bytes += pbytes;
  c->stats.kernel_launches += 2;
  return RM_OK;
}

// Parity hook: read the resident volume back.
int rm_read_volume(rm_ctx* c, uint8_t* voxels_out) {
  if (!c) return RM_ERR_INVALID_ARG;
  if (c->is_group) return rm_read_volume(c->members[0], voxels_out);
  if (!voxels_out) return fail(c, RM_ERR_INVALID_ARG, "rm_read_volume: null output");
  if (!c->d_vox || c->rx <= 0) return fail(c, RM_ERR_NO_VOLUME, "no volume uploaded (rm_set_volume)");
  RM_CUDA(c, cudaSetDevice(c->device));
  const size_t bytes = (size_t)c->rx * c->ry * c->rz;
  RM_CUDA(c, cudaMemcpyAsync(voxels_out, c->d_vox, bytes, cudaMemcpyDeviceToHost, c->stream));
  RM_CUDA(c, cudaStreamSynchronize(c->stream));
  c->stats.d2h_bytes += bytes;
  return RM_OK;
}

// generate-scatter-offsets on the device for seeds seed0 .. seed0+count-1 into the resident table
// slots 0 .. count-1; rm_upload_passes(ctx, opts, NULL, iter) then uses them.
int rm_generate_scatter_tables(rm_ctx* c, int64_t seed0, int count) {
  if (!c) return RM_ERR_INVALID_ARG;
  RM_FORWARD(c, rm_generate_scatter_tables(m, seed0, count));
  if (count <= 0) return fail(c, RM_ERR_INVALID_ARG, "rm_generate_scatter_tables: count <= 0");
  RM_CUDA(c, cudaSetDevice(c->device));
  if (c->table_capacity < count) RM_CUDA(c, cudaStreamSynchronize(c->stream));
  int rc = ensure_tables(c, count);
  if (rc) return rc;
  cudaError_t e = rm_launch_scatter_tables((long long)seed0, count, c->d_tables, c->stream);
  if (e != cudaSuccess) return cuda_fail(c, e, "scatter table generator");
  c->stats.kernel_launches += 1;
  c->generated_tables = count;
  c->resident = 0;
  return RM_OK;
}

int rm_clear_accum(rm_ctx* c, int width, int height) {
  if (!c) return RM_ERR_INVALID_ARG;
  if (c->is_group) return grp::clear_accum(c, width, height);
  if (width <= 0 || height <= 0 || (long long)width * height > 0x7fffffffLL / 37)
    return fail(c, RM_ERR_INVALID_ARG, "rm_clear_accum: bad framebuffer extent");
  RM_CUDA(c, cudaSetDevice(c->device));
  const size_t n = (size_t)width * height;
  if (n > c->fb_capacity) {
    RM_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->copy_stream) RM_CUDA(c, cudaStreamSynchronize(c->copy_stream));
    cudaFree(c->d_accum); cudaFree(c->d_argb2[0]); cudaFree(c->d_argb2[1]);
    c->d_accum = nullptr; c->d_argb2[0] = c->d_argb2[1] = nullptr; c->fb_capacity = 0;
    RM_CUDA(c, cudaMalloc(&c->d_accum, n * sizeof(float4)));
    RM_CUDA(c, cudaMalloc(&c->d_argb2[0], n * sizeof(uint32_t)));
    RM_CUDA(c, cudaMalloc(&c->d_argb2[1], n * sizeof(uint32_t)));
    c->fb_capacity = n;
  }
  c->W = width; c->H = height;
  c->argb_fresh_ptr = nullptr;
  update_shard(c);
  RM_CUDA(c, cudaMemsetAsync(c->d_accum, 0, n * sizeof(float4), c->stream));
  return RM_OK;
}

int rm_render_pass(rm_ctx* c, const void* opts, size_t opts_len, const float* mc, size_t mc_floats) {
  if (!c) return RM_ERR_INVALID_ARG;
  if (!opts || opts_len != RM_OPTS_BYTES) return fail(c, RM_ERR_INVALID_ARG, "rm_render_pass: opts must be a 544-byte TRenderOpts blob");
  if (!mc || mc_floats != RM_TABLE_FLOATS) return fail(c, RM_ERR_INVALID_ARG, "rm_render_pass: mc must hold 65536 floats (16384 float4)");
  const void* o[1] = {opts};
  const float* m[1] = {mc};
  return rm_render_frame(c, o, m, 1);
}

// Queue all passes of a frame on the context's stream (no synchronisation). The tables come from the
// caller's host buffers, or -- peer_src != null, multi-GPU groups -- from the table slots of another
// member over NVLink once `src_ready` has fired there.
static int queue_frame(rm_ctx* c, const void* const* opts, const float* const* mc, int iter, const rm_ctx* peer_src,
                       cudaEvent_t src_ready, cudaEvent_t uploaded = nullptr) {
  if (!opts || !mc || iter <= 0) return fail(c, RM_ERR_INVALID_ARG, "rm_render_frame: null arrays or iter <= 0");
  int rc = require_ready(c);
  if (rc) return rc;
  RM_CUDA(c, cudaSetDevice(c->device));
  std::vector<RmOpts> dec((size_t)iter);
  for (int i = 0; i < iter; ++i) {
    if (!opts[i] || !mc[i]) return fail(c, RM_ERR_INVALID_ARG, "rm_render_frame: null per-pass pointer");
    std::memset(&dec[i], 0, sizeof(RmOpts));
    decode_opts(opts[i], &dec[i]);
    if ((rc = check_opts(c, dec[i]))) return rc;
  }
  const size_t tbytes = (size_t)RM_TABLE_FLOATS * sizeof(float);
  if ((rc = ensure_tables(c, iter > c->resident ? iter : c->resident))) return rc;
  // a frame rendered from host buffers replaces any resident passes and overwrites table slots
  // 0 .. iter-1, including tables generated in place (rm_generate_scatter_tables)
  c->resident = 0;
  c->generated_tables = 0;
  EventPair t = begin_timed(c, 2);
  cudaError_t e = cudaSuccess;
  if (peer_src) {
    e = cudaStreamWaitEvent(c->stream, src_ready, 0);
    if (e == cudaSuccess)
      e = cudaMemcpyPeerAsync(c->d_tables, c->device, peer_src->d_tables, peer_src->device, tbytes * iter, c->stream);
  } else {
    for (int i = 0; i < iter && e == cudaSuccess; ++i)  // straight from the caller's (ideally pinned) buffers
      e = cudaMemcpyAsync(c->d_tables + (size_t)i * (RM_TABLE_FLOATS / 4), mc[i], tbytes, cudaMemcpyHostToDevice, c->stream);
    c->stats.h2d_bytes += tbytes * iter;
  }
  end_timed(c, t);
  if (e == cudaSuccess && uploaded) e = cudaEventRecord(uploaded, c->stream);  // the tables are in place from here on
  if (e != cudaSuccess) return cuda_fail(c, e, "table upload");
  c->stats.h2d_bytes += (size_t)RM_OPTS_BYTES * iter;
  return launch_passes(c, dec.data(), iter, c->d_tables);
}

int rm_render_frame(rm_ctx* c, const void* const* opts, const float* const* mc, int iter) {
  if (!c) return RM_ERR_INVALID_ARG;
  if (c->is_group) return grp::render_frame(c, opts, mc, iter);
  int rc = queue_frame(c, opts, mc, iter, nullptr, nullptr);
  if (rc) return rc;
  RM_CUDA(c, cudaStreamSynchronize(c->stream));
  return check_watchdog(c);
}

int rm_upload_passes(rm_ctx* c, const void* const* opts, const float* const* mc, int iter) {
  if (!c) return RM_ERR_INVALID_ARG;
  RM_FORWARD(c, rm_upload_passes(m, opts, mc, iter));
  if (!opts || iter <= 0) return fail(c, RM_ERR_INVALID_ARG, "rm_upload_passes: null opts array or iter <= 0");
  if (!mc && iter > c->generated_tables)
    return fail(c, RM_ERR_INVALID_ARG, "rm_upload_passes: mc == NULL needs rm_generate_scatter_tables(count >= iter) first");
  int rc = require_ready(c);
  if (rc) return rc;
  RM_CUDA(c, cudaSetDevice(c->device));
  std::vector<RmOpts> dec((size_t)iter);
  for (int i = 0; i < iter; ++i) {
    if (!opts[i] || (mc && !mc[i])) return fail(c, RM_ERR_INVALID_ARG, "rm_upload_passes: null per-pass pointer");
    std::memset(&dec[i], 0, sizeof(RmOpts));
    decode_opts(opts[i], &dec[i]);
    if ((rc = check_opts(c, dec[i]))) return rc;
  }
  const size_t tbytes = (size_t)RM_TABLE_FLOATS * sizeof(float);
  if (mc) {
    if ((rc = ensure_tables(c, iter))) return rc;
    c->generated_tables = 0;
    RM_CUDA(c, cudaStreamSynchronize(c->stream));
    for (int i = 0; i < iter; ++i)
      RM_CUDA(c, cudaMemcpyAsync(c->d_tables + (size_t)i * (RM_TABLE_FLOATS / 4), mc[i], tbytes, cudaMemcpyHostToDevice, c->stream));
    RM_CUDA(c, cudaStreamSynchronize(c->stream));
    c->stats.h2d_bytes += tbytes * iter;
  }
  c->stats.h2d_bytes += (size_t)RM_OPTS_BYTES * iter;
  c->passes = dec;
  c->resident = iter;
  return RM_OK;
}

// update-render-option-buffer (core.clj:108-117) for resident passes: replace the opts of the `iter`
// uploaded passes (a new camera, say); the tables stay where they are. 544 bytes per pass of host traffic.
int rm_update_opts(rm_ctx* c, const void* const* opts, int iter) {
  if (!c) return RM_ERR_INVALID_ARG;
  RM_FORWARD(c, rm_update_opts(m, opts, iter));
  if (!opts || iter <= 0) return fail(c, RM_ERR_INVALID_ARG, "rm_update_opts: null opts array or iter <= 0");
  if (iter != c->resident) return fail(c, RM_ERR_INVALID_ARG, "rm_update_opts: iter differs from the passes uploaded by rm_upload_passes");
  int rc = require_ready(c);
  if (rc) return rc;
  std::vector<RmOpts> dec((size_t)iter);
  for (int i = 0; i < iter; ++i) {
    if (!opts[i]) return fail(c, RM_ERR_INVALID_ARG, "rm_update_opts: null per-pass pointer");
    std::memset(&dec[i], 0, sizeof(RmOpts));
    decode_opts(opts[i], &dec[i]);
    if ((rc = check_opts(c, dec[i]))) return rc;
  }
  c->passes = dec;
  c->stats.h2d_bytes += (size_t)RM_OPTS_BYTES * iter;
  return RM_OK;
}

int rm_render_resident(rm_ctx* c, int first, int count) {
  if (!c) return RM_ERR_INVALID_ARG;
  RM_FORWARD(c, rm_render_resident(m, first, count));
  int rc = require_ready(c);
  if (rc) return rc;
  if (first < 0 || count <= 0 || first + count > c->resident)
    return fail(c, RM_ERR_INVALID_ARG, "rm_render_resident: pass range outside the uploaded passes");
  RM_CUDA(c, cudaSetDevice(c->device));
  for (int i = first; i < first + count; ++i)
    if ((rc = check_opts(c, c->passes[i]))) return rc;
  return launch_passes(c, c->passes.data() + first, count, c->d_tables + (size_t)first * (RM_TABLE_FLOATS / 4));
}

// Is `buf` (indexed like `packed` says) already tonemap(accumulator, gamma)? True when the default
// kernel wrote it while rendering the last launch and nothing has touched the accumulator since.
static bool argb_is_fresh(const rm_ctx* c, const void* buf, int packed, float gamma) {
  return c->argb_fresh_ptr && c->argb_fresh_ptr == buf && (c->argb_fresh_packed != 0) == (packed != 0) &&
         std::memcmp(&c->argb_fresh_gamma, &gamma, sizeof(float)) == 0;
}

// TonemapImage into the context's current ARGB frame (skipped when the render launch already wrote it).
static int tonemap_into_frame(rm_ctx* c, const RmOpts& o) {
  uint32_t* frame = c->d_argb2[c->argb_cur];
  if (argb_is_fresh(c, frame, 0, o.gamma)) return RM_OK;
  EventPair t = begin_timed(c, 1);
  cudaError_t e = rm_launch_tonemap(c->d_accum, o.gamma, c->W, c->H, c->shard, frame, 0, c->stream);
  end_timed(c, t);
  if (e != cudaSuccess) return cuda_fail(c, e, "tonemap kernel launch");
  c->stats.kernel_launches += 1;
  c->argb_fresh_ptr = frame;  // the frame now holds TonemapImage(accumulator, this gamma)
  c->argb_fresh_packed = 0;
  c->argb_fresh_gamma = o.gamma;
  return RM_OK;
}

int rm_tonemap(rm_ctx* c, const void* opts, size_t opts_len, uint32_t* argb_out) {
  if (!c) return RM_ERR_INVALID_ARG;
  if (c->is_group) return grp::tonemap(c, opts, opts_len, argb_out, -1);
  if (!opts || opts_len != RM_OPTS_BYTES || !argb_out) return fail(c, RM_ERR_INVALID_ARG, "rm_tonemap: bad opts blob or null output");
  if (!c->d_accum) return fail(c, RM_ERR_NO_FRAMEBUFFER, "no framebuffer (rm_clear_accum)");
  RM_CUDA(c, cudaSetDevice(c->device));
  RmOpts o;
  decode_opts(opts, &o);
  if (o.width != c->W || o.height != c->H) return fail(c, RM_ERR_BAD_OPTS, "rm_tonemap: TRenderOpts.resolution does not match the framebuffer");
  const size_t n = (size_t)c->W * c->H;
  int rc = tonemap_into_frame(c, o);
  if (rc) return rc;
  EventPair t2 = begin_timed(c, 3);
  cudaError_t e = cudaMemcpyAsync(argb_out, c->d_argb2[c->argb_cur], n * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream);
  end_timed(c, t2);
  if (e != cudaSuccess) return cuda_fail(c, e, "argb read-back");
  RM_CUDA(c, cudaStreamSynchronize(c->stream));
  if ((rc = check_watchdog(c))) return rc;
  c->stats.d2h_bytes += n * sizeof(uint32_t);
  return RM_OK;
}

// TonemapImage + an ASYNCHRONOUS read-back into (ideally pinned) host memory: returns as soon as the
// work is queued. The copy runs on a second stream out of one of two ARGB frames, so the next frame
// (rm_clear_accum / rm_render_* after this call) renders into the other one while this one travels.
// slot (0 or 1) names the transfer for rm_wait.
int rm_tonemap_async(rm_ctx* c, const void* opts, size_t opts_len, uint32_t* argb_out, int slot) {
  if (!c) return RM_ERR_INVALID_ARG;
  if (c->is_group) return (slot < 0 || slot > 1) ? fail(c, RM_ERR_INVALID_ARG, "rm_tonemap_async: slot must be 0 or 1") : grp::tonemap(c, opts, opts_len, argb_out, slot);
  if (!opts || opts_len != RM_OPTS_BYTES || !argb_out) return fail(c, RM_ERR_INVALID_ARG, "rm_tonemap_async: bad opts blob or null output");
  if (slot < 0 || slot > 1) return fail(c, RM_ERR_INVALID_ARG, "rm_tonemap_async: slot must be 0 or 1");
  if (!c->d_accum) return fail(c, RM_ERR_NO_FRAMEBUFFER, "no framebuffer (rm_clear_accum)");
  RM_CUDA(c, cudaSetDevice(c->device));
  RmOpts o;
  decode_opts(opts, &o);
  if (o.width != c->W || o.height != c->H) return fail(c, RM_ERR_BAD_OPTS, "rm_tonemap_async: TRenderOpts.resolution does not match the framebuffer");
  if (c->copy_pending[slot]) {  // the caller reuses a slot it never waited for
    RM_CUDA(c, cudaEventSynchronize(c->copy_done[c->slot_buffer[slot]]));
    c->copy_pending[slot] = false;
  }
  int rc = tonemap_into_frame(c, o);
  if (rc) return rc;
  const int b = c->argb_cur;
  const size_t n = (size_t)c->W * c->H;
  RM_CUDA(c, cudaEventRecord(c->frame_ready[b], c->stream));
  RM_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->frame_ready[b], 0));
  RM_CUDA(c, cudaMemcpyAsync(argb_out, c->d_argb2[b], n * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->copy_stream));
  RM_CUDA(c, cudaEventRecord(c->copy_done[b], c->copy_stream));
  c->copy_pending[slot] = true;
  c->slot_buffer[slot] = b;
  c->stats.d2h_bytes += n * sizeof(uint32_t);
  // the next frame goes to the other buffer -- once the transfer that last used it has drained
  c->argb_cur = b ^ 1;
  c->argb_fresh_ptr = nullptr;
  for (int s2 = 0; s2 < 2; ++s2)
    if (c->copy_pending[s2] && c->slot_buffer[s2] == c->argb_cur)
      RM_CUDA(c, cudaStreamWaitEvent(c->stream, c->copy_done[c->argb_cur], 0));
  return RM_OK;
}

// Block until the transfer started by rm_tonemap_async(.., slot) has landed in host memory.
int rm_wait(rm_ctx* c, int slot) {
  if (!c) return RM_ERR_INVALID_ARG;
  if (c->is_group) return grp::wait(c, slot);
  if (slot < 0 || slot > 1) return fail(c, RM_ERR_INVALID_ARG, "rm_wait: slot must be 0 or 1");
  if (!c->copy_pending[slot]) return RM_OK;
  RM_CUDA(c, cudaSetDevice(c->device));
  RM_CUDA(c, cudaEventSynchronize(c->copy_done[c->slot_buffer[slot]]));
  c->copy_pending[slot] = false;
  return check_watchdog(c);
}

int rm_host_alloc(rm_ctx* c, size_t bytes, void** out_ptr) {
  if (!c) return RM_ERR_INVALID_ARG;
  if (c->is_group) return rm_host_alloc(c->members[0], bytes, out_ptr);
  if (!out_ptr || bytes == 0) return fail(c, RM_ERR_INVALID_ARG, "rm_host_alloc: null out pointer or zero size");
  *out_ptr = nullptr;
  RM_CUDA(c, cudaSetDevice(c->device));
  RM_CUDA(c, cudaHostAlloc(out_ptr, bytes, cudaHostAllocPortable));
  return RM_OK;
}

int rm_host_free(rm_ctx* c, void* ptr) {
  if (!c) return RM_ERR_INVALID_ARG;
  if (c->is_group) return rm_host_free(c->members[0], ptr);
  if (!ptr) return RM_OK;
  RM_CUDA(c, cudaSetDevice(c->device));
  for (int s2 = 0; s2 < 2; ++s2)  // a transfer into it may still be in flight
    if (c->copy_pending[s2]) { cudaEventSynchronize(c->copy_done[c->slot_buffer[s2]]); c->copy_pending[s2] = false; }
  RM_CUDA(c, cudaFreeHost(ptr));
  return RM_OK;
}

int rm_tonemap_device(rm_ctx* c, const void* opts, size_t opts_len, void* d_argb, int packed) {
  if (!c) return RM_ERR_INVALID_ARG;
  if (c->is_group) return grp::unsupported(c, "rm_tonemap_device");
  if (!opts || opts_len != RM_OPTS_BYTES || !d_argb) return fail(c, RM_ERR_INVALID_ARG, "rm_tonemap_device: bad opts blob or null output");
  if (!c->d_accum) return fail(c, RM_ERR_NO_FRAMEBUFFER, "no framebuffer (rm_clear_accum)");
  RM_CUDA(c, cudaSetDevice(c->device));
  RmOpts o;
  decode_opts(opts, &o);
  if (o.width != c->W || o.height != c->H) return fail(c, RM_ERR_BAD_OPTS, "rm_tonemap_device: TRenderOpts.resolution does not match the framebuffer");
  if (argb_is_fresh(c, d_argb, packed, o.gamma)) return RM_OK;  // the render launch wrote it (rm_set_argb_target)
  EventPair t = begin_timed(c, 1);
  cudaError_t e = rm_launch_tonemap(c->d_accum, o.gamma, c->W, c->H, c->shard, static_cast<uint32_t*>(d_argb), packed, c->stream);
  end_timed(c, t);
  if (e != cudaSuccess) return cuda_fail(c, e, "tonemap kernel launch");
  c->stats.kernel_launches += 1;
  c->argb_fresh_ptr = d_argb;  // (whatever the render launch left in this or another buffer is no longer "the" frame)
  c->argb_fresh_packed = packed != 0;
  c->argb_fresh_gamma = o.gamma;
  return RM_OK;
}

// Register a device buffer the caller owns (width*height words, or rm_shard_slots words when packed)
// as the place where the default render kernel leaves the ARGB words of the frame while it renders:
// a later rm_tonemap_device(opts, same buffer, same packing) with the same gamma is then free. The
// buffer may live on a PEER device (NVLink): every GPU of a box can store its tiles straight into
// one frame. NULL restores the context's own frame.
int rm_set_argb_target(rm_ctx* c, void* d_argb, int packed) {
  if (!c) return RM_ERR_INVALID_ARG;
  if (c->is_group) return grp::unsupported(c, "rm_set_argb_target");
  c->argb_target = static_cast<uint32_t*>(d_argb);
  c->argb_target_packed = d_argb ? (packed != 0) : 0;
  c->argb_fresh_ptr = nullptr;
  return RM_OK;
}

int rm_copy_accum_device(rm_ctx* c, void* d_rgba, int packed) {
  if (!c) return RM_ERR_INVALID_ARG;
  if (c->is_group) return grp::unsupported(c, "rm_copy_accum_device");
  if (!d_rgba) return fail(c, RM_ERR_INVALID_ARG, "rm_copy_accum_device: null output");
  if (!c->d_accum) return fail(c, RM_ERR_NO_FRAMEBUFFER, "no framebuffer (rm_clear_accum)");
  RM_CUDA(c, cudaSetDevice(c->device));
  if (packed) {
    cudaError_t e = rm_launch_pack_accum(c->d_accum, c->W, c->H, c->shard, static_cast<float4*>(d_rgba), c->stream);
    if (e != cudaSuccess) return cuda_fail(c, e, "pack kernel launch");
    c->stats.kernel_launches += 1;
  } else {
    RM_CUDA(c, cudaMemcpyAsync(d_rgba, c->d_accum, (size_t)c->W * c->H * sizeof(float4), cudaMemcpyDeviceToDevice, c->stream));
  }
  return RM_OK;
}

int rm_read_accum(rm_ctx* c, float* rgba_out) {
  if (!c) return RM_ERR_INVALID_ARG;
  if (c->is_group) return grp::read_accum(c, rgba_out);
  if (!rgba_out) return fail(c, RM_ERR_INVALID_ARG, "rm_read_accum: null output");
  if (!c->d_accum) return fail(c, RM_ERR_NO_FRAMEBUFFER, "no framebuffer (rm_clear_accum)");
  RM_CUDA(c, cudaSetDevice(c->device));
  const size_t bytes = (size_t)c->W * c->H * sizeof(float4);
  RM_CUDA(c, cudaMemcpyAsync(rgba_out, c->d_accum, bytes, cudaMemcpyDeviceToHost, c->stream));
  RM_CUDA(c, cudaStreamSynchronize(c->stream));
  c->stats.d2h_bytes += bytes;
  return check_watchdog(c);
}

int rm_sync(rm_ctx* c) {
  if (!c) return RM_ERR_INVALID_ARG;
  RM_FORWARD(c, rm_sync(m));
  RM_CUDA(c, cudaSetDevice(c->device));
  RM_CUDA(c, cudaStreamSynchronize(c->stream));
  return check_watchdog(c);
}

int rm_set_stream(rm_ctx* c, void* cuda_stream) {
  if (!c) return RM_ERR_INVALID_ARG;
  if (c->is_group) return grp::unsupported(c, "rm_set_stream");
  RM_CUDA(c, cudaSetDevice(c->device));
  RM_CUDA(c, cudaStreamSynchronize(c->stream));
  resolve_timers(c);
  c->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : c->own_stream;
  return RM_OK;
}

int rm_set_tile_shard(rm_ctx* c, int rank, int world, int tile_w, int tile_h) {
  if (!c) return RM_ERR_INVALID_ARG;
  if (c->is_group) return (rank != 0 || world != 1) ? fail(c, RM_ERR_INVALID_ARG, "rm_set_tile_shard on a multi-GPU context: rank 0, world 1 (only the tile size is taken)") : grp::set_tile(c, tile_w, tile_h);
  if (world <= 0 || rank < 0 || rank >= world || tile_w <= 0 || tile_h <= 0 || (tile_w & 7) || (tile_h & 3))
    return fail(c, RM_ERR_INVALID_ARG, "rm_set_tile_shard: need 0 <= rank < world, tile_w % 8 == 0, tile_h % 4 == 0");
  c->shard_rank = rank; c->shard_world = world; c->shard_tw = tile_w; c->shard_th = tile_h;
  c->argb_fresh_ptr = nullptr;
  update_shard(c);
  return RM_OK;
}

int64_t rm_shard_slots(const rm_ctx* c, int rank, int world) {
  if (c && c->is_group) c = c->members[0];
  if (!c || c->W <= 0 || world <= 0 || rank < 0 || rank >= world) return 0;
  RmShard s = c->shard;
  s.rank = rank; s.world = world;
  rm_shard_layout(s, c->W, c->H);
  return s.slots;  // the same for every rank of a world: tiles_per_rank_row x tiles_y tiles, padding included
}

int rm_unpack_shards(rm_ctx* c, const void* d_parts, int world, int64_t stride_slots, int elem_bytes, void* d_frame) {
  if (!c) return RM_ERR_INVALID_ARG;
  if (c->is_group) return grp::unsupported(c, "rm_unpack_shards");
  if (!d_parts || !d_frame || world <= 0 || (elem_bytes != 4 && elem_bytes != 16))
    return fail(c, RM_ERR_INVALID_ARG, "rm_unpack_shards: null buffer, world <= 0 or element size not 4 / 16");
  if (c->W <= 0) return fail(c, RM_ERR_NO_FRAMEBUFFER, "no framebuffer (rm_clear_accum)");
  if (stride_slots < rm_shard_slots(c, 0, world))
    return fail(c, RM_ERR_INVALID_ARG, "rm_unpack_shards: stride_slots smaller than the largest shard");
  RM_CUDA(c, cudaSetDevice(c->device));
  cudaError_t e = rm_launch_unpack_shards(d_parts, world, stride_slots, elem_bytes, c->W, c->H, c->shard, d_frame, c->stream);
  if (e != cudaSuccess) return cuda_fail(c, e, "unpack kernel launch");
  c->stats.kernel_launches += 1;
  return RM_OK;
}

int64_t rm_shard_pixels(const rm_ctx* c) {
  if (c && c->is_group) {
    int64_t px = 0;
    for (const rm_ctx* m : c->members) px += rm_shard_pixels(m);
    return px;
  }
  if (!c || c->W <= 0) return 0;
  const RmShard& s = c->shard;
  int64_t px = 0;
  for (int ty = 0; ty < s.tiles_y; ++ty) {
    int first = (s.rank - (int)(((long long)s.skew * ty) % s.world)) % s.world;
    if (first < 0) first += s.world;
    const int h = (ty + 1) * s.tile_h <= c->H ? s.tile_h : c->H - ty * s.tile_h;
    for (int tx = first; tx < s.tiles_x; tx += s.world) {
      const int w = (tx + 1) * s.tile_w <= c->W ? s.tile_w : c->W - tx * s.tile_w;
      px += (int64_t)w * h;
    }
  }
  return px;
}

int rm_set_option(rm_ctx* c, int option, int64_t value) {
  if (!c) return RM_ERR_INVALID_ARG;
  RM_FORWARD(c, rm_set_option(m, option, value));
  switch (option) {
    case RM_OPT_COUNT_WORK: c->count_work = value != 0; return RM_OK;
    case RM_OPT_KERNEL:
      if (value < 0 || value > 4) return fail(c, RM_ERR_INVALID_ARG, "RM_OPT_KERNEL: 0 (persistent, default), 1 (plain), 2 (warp), 3 (wavefront) or 4 (per-item bricks)");
      c->kernel_kind = (int)value;
      return RM_OK;
    case RM_OPT_CELL_SHIFT:
      if (value < 0 || value > 6 || value == 1) return fail(c, RM_ERR_INVALID_ARG, "RM_OPT_CELL_SHIFT: 0 (auto) or 2..6");
      c->cell_shift_opt = (int)value;
      c->accel.valid = false;
      return RM_OK;
    case RM_OPT_WAVE_CHUNK:
      if (value < 1024 || value > (1ll << 24)) return fail(c, RM_ERR_INVALID_ARG, "RM_OPT_WAVE_CHUNK: 1024..2^24 items");
      c->wave_chunk = (unsigned)value;
      return RM_OK;
    case RM_OPT_WAVE_REFILL:
      if (value < 1 || value > 32) return fail(c, RM_ERR_INVALID_ARG, "RM_OPT_WAVE_REFILL: 1..32 idle lanes");
      c->wave_refill = (int)value;
      return RM_OK;
    case RM_OPT_TRIP_LIMIT:
      if (value < 1 || value > 0xffffffffLL) return fail(c, RM_ERR_INVALID_ARG, "RM_OPT_TRIP_LIMIT: 1..2^32-1");
      c->trip_limit = (unsigned)value;
      return RM_OK;
    case RM_OPT_PERSIST_BLOCK:
      if (value != 0 && value != 256 && value != 1024)
        return fail(c, RM_ERR_INVALID_ARG, "RM_OPT_PERSIST_BLOCK: 0 (default), 1024 (x 1 block per SM) or 256 (x 5) threads");
      c->persist_block = (int)value;
      return RM_OK;
    case RM_OPT_PERSIST_ORDER:
      c->persist_bottom_up = value != 0;
      return RM_OK;
    case RM_OPT_PERSIST_SMEM:
      if (value < 0 || value > 2)
        return fail(c, RM_ERR_INVALID_ARG, "RM_OPT_PERSIST_SMEM: 0 (never), 1 (whenever the map fits the layout) or 2 (default: with the 1024-thread layout)");
      c->persist_smem = (int)value;
      return RM_OK;
    case RM_OPT_PERSIST_GROUP:
      if (value < -1 || value > 1) return fail(c, RM_ERR_INVALID_ARG, "RM_OPT_PERSIST_GROUP: -1 (default), 0 (free-running warps) or 1 (block-synchronous rounds)");
      c->persist_group = (int)value;
      return RM_OK;
    case RM_OPT_FUSE_LIMIT:
      if (value < 1 || value > RM_MAX_FUSED_PASSES) return fail(c, RM_ERR_INVALID_ARG, "RM_OPT_FUSE_LIMIT: 1..32");
      c->fuse_limit = (int)value;
      return RM_OK;
    default: return fail(c, RM_ERR_UNSUPPORTED, "rm_set_option: unknown option");
  }
}

int rm_get_stats(const rm_ctx* cc, rm_stats* out) {
  rm_ctx* c = const_cast<rm_ctx*>(cc);
  if (!c || !out) return RM_ERR_INVALID_ARG;
  if (c->is_group) return grp::get_stats(c, out);
  RM_CUDA(c, cudaSetDevice(c->device));
  RM_CUDA(c, cudaStreamSynchronize(c->stream));
  resolve_timers(c);
  RmCounters h{};
  RM_CUDA(c, cudaMemcpy(&h, c->d_counters, sizeof h, cudaMemcpyDeviceToHost));
  c->stats.steps = h.steps; c->stats.taps = h.taps; c->stats.outer_iters = h.outer;
  *out = c->stats;
  return RM_OK;
}

int rm_reset_stats(rm_ctx* c) {
  if (!c) return RM_ERR_INVALID_ARG;
  RM_FORWARD(c, rm_reset_stats(m));
  RM_CUDA(c, cudaSetDevice(c->device));
  RM_CUDA(c, cudaStreamSynchronize(c->stream));
  resolve_timers(c);
  RM_CUDA(c, cudaMemset(c->d_counters, 0, sizeof(RmCounters)));
  c->stats = rm_stats{};
  return RM_OK;
}


int rm_create_multi(const int* device_ids, int n, rm_ctx** out_ctx) {
  if (!out_ctx) return fail(nullptr, RM_ERR_INVALID_ARG, "rm_create_multi: out_ctx is null");
  *out_ctx = nullptr;
  if (!device_ids || n <= 0 || n > 64) return fail(nullptr, RM_ERR_INVALID_ARG, "rm_create_multi: need 1..64 device ids");
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < i; ++j)
      if (device_ids[i] == device_ids[j]) return fail(nullptr, RM_ERR_INVALID_ARG, "rm_create_multi: duplicate device id");
  rm_ctx* g = new (std::nothrow) rm_ctx();
  if (!g) return fail(nullptr, RM_ERR_INVALID_ARG, "out of host memory");
  g->is_group = true;
  g->device = device_ids[0];
  for (int i = 0; i < n; ++i) {
    rm_ctx* m = nullptr;
    const int rc = rm_create(device_ids[i], &m);
    if (rc) { grp::destroy(g); return rc; }
    g->members.push_back(m);
  }
  // every GPU stores its tiles of the ARGB frame straight into device_ids[0]'s memory (NVLink / NVSwitch)
  for (int i = 1; i < n; ++i) {
    int can = 0;
    cudaError_t e = cudaSetDevice(device_ids[i]);
    if (e == cudaSuccess) e = cudaDeviceCanAccessPeer(&can, device_ids[i], device_ids[0]);
    if (e != cudaSuccess || !can) {
      grp::destroy(g);
      return fail(nullptr, RM_ERR_UNSUPPORTED, "rm_create_multi: device " + std::to_string(device_ids[i]) +
                                                   " has no peer access to device " + std::to_string(device_ids[0]));
    }
    e = cudaDeviceEnablePeerAccess(device_ids[0], 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
    if (e == cudaSuccess) {  // and back, so that the volume / table broadcasts go directly as well
      cudaSetDevice(device_ids[0]);
      cudaError_t e2 = cudaDeviceEnablePeerAccess(device_ids[i], 0);
      if (e2 != cudaSuccess) cudaGetLastError();
    }
    if (e != cudaSuccess) { grp::destroy(g); return cuda_fail(nullptr, e, "cudaDeviceEnablePeerAccess"); }
  }
  g->shard_tw = 16; g->shard_th = 8;  // small tiles: load balance; the skewed ownership keeps them apart
  grp::set_tile(g, g->shard_tw, g->shard_th);
  *out_ctx = g;
  return RM_OK;
}

int rm_member_count(const rm_ctx* c) { return !c ? 0 : (c->is_group ? (int)c->members.size() : 1); }

int rm_get_member_stats(const rm_ctx* c, int member, rm_stats* out) {
  if (!c || !out) return RM_ERR_INVALID_ARG;
  if (!c->is_group) return member == 0 ? rm_get_stats(c, out) : RM_ERR_INVALID_ARG;
  if (member < 0 || member >= (int)c->members.size()) return RM_ERR_INVALID_ARG;
  return rm_get_stats(c->members[(size_t)member], out);
}

}  // extern "C"

// ---- multi-GPU groups ------------------------------------------------------------------------------
// One process, one member context (own stream) per GPU, image tiles dealt to the members by
// rm_set_tile_shard(i, n). Inputs are uploaded ONCE (to member 0) and broadcast device-to-device; every
// member renders its tiles and its kernel stores the ARGB words of those tiles directly into member
// 0's frame over NVLink (rm_render_persist.cu writes through argb_target): there is no gather step,
// no packing and no unpack kernel -- the "collective" is the render kernel's own epilogue.
namespace grp {

int member_failed(rm_ctx* g, rm_ctx* m, int rc) {
  g->err = "device " + std::to_string(m->device) + ": " + m->err;
  return rc;
}

int unsupported(rm_ctx* g, const char* what) {
  return fail(g, RM_ERR_UNSUPPORTED, std::string(what) + " is not available on a multi-GPU context (use the members' single-GPU form)");
}

void destroy(rm_ctx* g) {
  for (rm_ctx* m : g->members) rm_destroy(m);
  delete g;
}

// every member writes its ARGB words into member 0's current frame
static void retarget(rm_ctx* g) {
  rm_ctx* m0 = g->members[0];
  for (rm_ctx* m : g->members) {
    m->argb_target = m0->d_argb2[m0->argb_cur];
    m->argb_target_packed = 0;
    m->argb_fresh_ptr = nullptr;
  }
}

int set_tile(rm_ctx* g, int tile_w, int tile_h) {
  const int n = (int)g->members.size();
  for (int i = 0; i < n; ++i) {
    const int rc = rm_set_tile_shard(g->members[(size_t)i], i, n, tile_w, tile_h);
    if (rc) return member_failed(g, g->members[(size_t)i], rc);
  }
  g->shard_tw = tile_w; g->shard_th = tile_h;
  return RM_OK;
}

// member 0 holds the new volume (possibly still being written on its stream): copy it to the others
static int broadcast_volume(rm_ctx* g) {
  rm_ctx* m0 = g->members[0];
  const size_t bytes = (size_t)m0->rx * m0->ry * m0->rz;
  RM_CUDA(g, cudaSetDevice(m0->device));
  RM_CUDA(g, cudaEventRecord(m0->frame_ready[0], m0->stream));
  for (size_t i = 1; i < g->members.size(); ++i) {
    rm_ctx* m = g->members[i];
    RM_CUDA(g, cudaSetDevice(m->device));
    int rc = begin_volume(m, bytes, "volume broadcast");
    if (rc) return member_failed(g, m, rc);
    RM_CUDA(g, cudaStreamWaitEvent(m->stream, m0->frame_ready[0], 0));
    RM_CUDA(g, cudaMemcpyPeerAsync(m->d_vox, m->device, m0->d_vox, m0->device, bytes, m->stream));
    commit_volume(m, m0->rx, m0->ry, m0->rz);
  }
  return RM_OK;
}

int set_volume(rm_ctx* g, const uint8_t* voxels, int rx, int ry, int rz) {
  int rc = rm_set_volume(g->members[0], voxels, rx, ry, rz);  // ONE upload over PCIe ...
  if (rc) return member_failed(g, g->members[0], rc);
  return broadcast_volume(g);                                    // ... then NVLink
}

int set_volume_device(rm_ctx* g, const void* d_voxels, int rx, int ry, int rz) {
  int rc = rm_set_volume_device(g->members[0], d_voxels, rx, ry, rz);
  if (rc) return member_failed(g, g->members[0], rc);
  return broadcast_volume(g);
}

int load_volume_file(rm_ctx* g, const char* path, int* rx, int* ry, int* rz) {
  int rc = rm_load_volume_file(g->members[0], path, rx, ry, rz);
  if (rc) return member_failed(g, g->members[0], rc);
  return broadcast_volume(g);
}

int generate_volume(rm_ctx* g, int kind, int rx, int ry, int rz) {
  rm_ctx* m0 = g->members[0];
  int rc = kind == 0 ? rm_generate_gyroid_volume(m0, rx, ry, rz) : rm_generate_terrain_volume(m0, rx, ry, rz);
  if (rc) return member_failed(g, m0, rc);
  return broadcast_volume(g);
}

int voxelize_points(rm_ctx* g, const float* xyz, int64_t n, int res, int ks) {
  int rc = rm_voxelize_points(g->members[0], xyz, n, res, ks);
  if (rc) return member_failed(g, g->members[0], rc);
  return broadcast_volume(g);
}

int clear_accum(rm_ctx* g, int w, int h) {
  for (rm_ctx* m : g->members) {
    const int rc = rm_clear_accum(m, w, h);
    if (rc) return member_failed(g, m, rc);
  }
  g->W = w; g->H = h;
  retarget(g);
  return RM_OK;
}

int render_frame(rm_ctx* g, const void* const* opts, const float* const* mc, int iter) {
  rm_ctx* m0 = g->members[0];
  // tables: host -> member 0 over PCIe once, then device to device over NVLink (an event marks the upload)
  int rc = queue_frame(m0, opts, mc, iter, nullptr, nullptr, m0->frame_ready[1]);
  if (rc) return member_failed(g, m0, rc);
  for (size_t i = 1; i < g->members.size(); ++i) {
    rm_ctx* m = g->members[i];
    rc = queue_frame(m, opts, mc, iter, m0, m0->frame_ready[1]);
    if (rc) return member_failed(g, m, rc);
  }
  for (rm_ctx* m : g->members) {
    RM_CUDA(g, cudaSetDevice(m->device));
    RM_CUDA(g, cudaStreamSynchronize(m->stream));
    if ((rc = check_watchdog(m))) return member_failed(g, m, rc);
  }
  return RM_OK;
}

// TonemapImage of the group's frame + read-back from member 0 (async_slot < 0: blocking)
int tonemap(rm_ctx* g, const void* opts, size_t len, uint32_t* out, int async_slot) {
  rm_ctx* m0 = g->members[0];
  if (!opts || len != RM_OPTS_BYTES || !out) return fail(g, RM_ERR_INVALID_ARG, "rm_tonemap: bad opts blob or null output");
  if (!m0->d_accum) return fail(g, RM_ERR_NO_FRAMEBUFFER, "no framebuffer (rm_clear_accum)");
  RmOpts o;
  decode_opts(opts, &o);
  if (o.width != m0->W || o.height != m0->H) return fail(g, RM_ERR_BAD_OPTS, "rm_tonemap: TRenderOpts.resolution does not match the framebuffer");
  const int b = m0->argb_cur;
  uint32_t* frame = m0->d_argb2[b];
  const size_t n = (size_t)m0->W * m0->H;
  if (async_slot >= 0 && m0->copy_pending[async_slot]) {
    RM_CUDA(g, cudaEventSynchronize(m0->copy_done[m0->slot_buffer[async_slot]]));
    m0->copy_pending[async_slot] = false;
  }
  for (rm_ctx* m : g->members) {
    RM_CUDA(g, cudaSetDevice(m->device));
    if (!argb_is_fresh(m, frame, 0, o.gamma)) {  // not written by the render launch: tonemap the tiles this member owns
      EventPair t = begin_timed(m, 1);
      cudaError_t e = rm_launch_tonemap(m->d_accum, o.gamma, m->W, m->H, m->shard, frame, 2, m->stream);
      end_timed(m, t);
      if (e != cudaSuccess) return cuda_fail(g, e, "tonemap kernel launch");
      m->stats.kernel_launches += 1;
      m->argb_fresh_ptr = frame;
      m->argb_fresh_packed = 0;
      m->argb_fresh_gamma = o.gamma;
    }
    if (m != m0) {  // member 0's stream (which does the read-back) waits for every member's tiles
      RM_CUDA(g, cudaEventRecord(m->frame_ready[0], m->stream));
      RM_CUDA(g, cudaStreamWaitEvent(m0->stream, m->frame_ready[0], 0));
    }
  }
  RM_CUDA(g, cudaSetDevice(m0->device));
  if (async_slot < 0) {
    EventPair t2 = begin_timed(m0, 3);
    cudaError_t e = cudaMemcpyAsync(out, frame, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, m0->stream);
    end_timed(m0, t2);
    if (e != cudaSuccess) return cuda_fail(g, e, "argb read-back");
    RM_CUDA(g, cudaStreamSynchronize(m0->stream));
    m0->stats.d2h_bytes += n * sizeof(uint32_t);
    for (rm_ctx* m : g->members) {
      const int rc = check_watchdog(m);
      if (rc) return member_failed(g, m, rc);
    }
    return RM_OK;
  }
  RM_CUDA(g, cudaEventRecord(m0->frame_ready[b], m0->stream));
  RM_CUDA(g, cudaStreamWaitEvent(m0->copy_stream, m0->frame_ready[b], 0));
  RM_CUDA(g, cudaMemcpyAsync(out, frame, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, m0->copy_stream));
  RM_CUDA(g, cudaEventRecord(m0->copy_done[b], m0->copy_stream));
  m0->copy_pending[async_slot] = true;
  m0->slot_buffer[async_slot] = b;
  m0->stats.d2h_bytes += n * sizeof(uint32_t);
  // the next frame goes to the other buffer -- on every member, once the transfer that last used it has drained
  m0->argb_cur = b ^ 1;
  retarget(g);
  for (int s2 = 0; s2 < 2; ++s2)
    if (m0->copy_pending[s2] && m0->slot_buffer[s2] == m0->argb_cur)
      for (rm_ctx* m : g->members) RM_CUDA(g, cudaStreamWaitEvent(m->stream, m0->copy_done[m0->argb_cur], 0));
  // ... and nobody may start writing into it while another member's stores of the PREVIOUS use are in flight:
  // those were ordered before member 0's read-back by the events above.
  return RM_OK;
}

int wait(rm_ctx* g, int slot) {
  const int rc = rm_wait(g->members[0], slot);
  return rc ? member_failed(g, g->members[0], rc) : RM_OK;
}

// parity hook: every pixel is rendered by exactly one member (the others hold zeros there)
int read_accum(rm_ctx* g, float* out) {
  rm_ctx* m0 = g->members[0];
  if (!out) return fail(g, RM_ERR_INVALID_ARG, "rm_read_accum: null output");
  if (!m0->d_accum) return fail(g, RM_ERR_NO_FRAMEBUFFER, "no framebuffer (rm_clear_accum)");
  const size_t n = (size_t)m0->W * m0->H;
  std::vector<float> tmp(n * 4);
  std::memset(out, 0, n * 4 * sizeof(float));
  for (rm_ctx* m : g->members) {
    const int rc = rm_read_accum(m, tmp.data());
    if (rc) return member_failed(g, m, rc);
    for (size_t i = 0; i < n; ++i)
      if (tmp[4 * i + 3] != 0.0f) std::memcpy(out + 4 * i, tmp.data() + 4 * i, 4 * sizeof(float));
  }
  return RM_OK;
}

// work summed over the members; times = the slowest member (they run concurrently)
int get_stats(rm_ctx* g, rm_stats* out) {
  rm_stats a{};
  for (rm_ctx* m : g->members) {
    rm_stats s{};
    const int rc = rm_get_stats(m, &s);
    if (rc) return member_failed(g, m, rc);
    a.steps += s.steps; a.taps += s.taps; a.outer_iters += s.outer_iters; a.pixel_samples += s.pixel_samples;
    a.kernel_launches += s.kernel_launches; a.render_launches += s.render_launches;
    a.h2d_bytes += s.h2d_bytes; a.d2h_bytes += s.d2h_bytes;
    a.render_ms = s.render_ms > a.render_ms ? s.render_ms : a.render_ms;
    a.tonemap_ms = s.tonemap_ms > a.tonemap_ms ? s.tonemap_ms : a.tonemap_ms;
    a.h2d_ms = s.h2d_ms > a.h2d_ms ? s.h2d_ms : a.h2d_ms;
    a.d2h_ms = s.d2h_ms > a.d2h_ms ? s.d2h_ms : a.d2h_ms;
  }
  *out = a;
  return RM_OK;
}

}  // namespace grp
