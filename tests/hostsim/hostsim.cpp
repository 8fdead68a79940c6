// hostsim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// Compiles the device routine of raymarchcl_b200/csrc/rm_scene_plain.cuh for the host through
// tests/hostsim/stubs/cuda_runtime.h, together with a CPU restatement of the occupancy tables of
// rm_accel.cu (bit-bricks + macro-cell Chebyshev distance map). Built with
// `g++ -O2 -ffp-contract=off` the fp32 arithmetic is the kernel's (-fmad=false, IEEE div/sqrt), and
// exp/exp2/pow come from the same libm the C oracle uses, so the PRODUCTION algorithm (fetch
// elision, irrelevance culling, cut marches, closed-form recurrence jumps) can be compared with the
// oracle bit for bit on the CPU -- a much sharper check than the 2e-5 GPU tolerance, and one that
// needs no GPU time. It also counts what the production march actually does (skip lengths,
// lookups), which is what the kernel's cost model in DESIGN.md is built on.
#include <cstdio>
#include <vector>
#include <omp.h>

// statistics hooks of the production march (empty in the CUDA build)
struct SimStats {
  unsigned long long lookups, skips, skipped_samples, jumps, jump_samples, seq_adds, marches, traces, hist[16], events[32];
};
static thread_local SimStats t_stats;
// per call site cost model of one pixel-sample (instruction estimates: cheap evaluation 15, full
// distanceToScene call 150, march lookup 50, skipped sample 3)
// per pixel-sample log of the full distanceToScene evaluations: (call site << 8 | table lookups), in order
static thread_local unsigned short t_eval_log[256];
static thread_local int t_eval_n;
static thread_local float t_site_cost[64];
static thread_local int t_site, t_level;
#define RM_STAT_SITE(id) (t_site = (id) & 63)
#define RM_STAT_LEVEL(l) (t_level = (l))
#define RM_STAT_LEVEL_GET() t_level
#define RM_STAT_LOOKUP() (t_stats.lookups++, t_site_cost[t_site] += 50.0f, (t_eval_n > 0 && t_eval_n <= 256 && (t_eval_log[t_eval_n - 1] & 255) < 255) ? (void)++t_eval_log[t_eval_n - 1] : (void)0)
#define RM_STAT_SKIP(n) (t_stats.skips++, t_stats.skipped_samples += (n), t_stats.hist[stat_bucket(n)]++, t_site_cost[t_site] += 3.0f * (n))
#define RM_STAT_JUMP(n) (t_stats.jumps++, t_stats.jump_samples += (n))
#define RM_STAT_SEQ(n) (t_stats.seq_adds += (n))
#define RM_STAT_MARCH() (t_stats.marches++)
#define RM_STAT_TRACE() (t_stats.traces++)
#define RM_STAT_EVENT(id) (t_stats.events[(id)]++, t_site_cost[t_site] += ((id) == 0 ? 150.0f : ((id) == 12 ? 15.0f : 0.0f)), ((id) == 0 && t_eval_n < 256) ? (void)(t_eval_log[t_eval_n++] = (unsigned short)(t_site << 8)) : (void)0)
static inline int stat_bucket(int n) {
  int b = 0;
  while (n > 1 && b < 15) { n >>= 1; ++b; }
  return b;
}

#include "rm_kernels.h"  // rm_slot_to_pixel / rm_shard_layout: the tile ownership the kernels and raymarchcl_b200/dist.py share
#include "rm_scene_plain.cuh"
#include "rm_scene_fused.cuh"
#include "rm_wave.cuh"

namespace {

float rd_f(const uint8_t* b, int off) { float f; std::memcpy(&f, b + off, 4); return f; }
int rd_i(const uint8_t* b, int off) { int i; std::memcpy(&i, b + off, 4); return i; }
float3 rd_f3(const uint8_t* b, int off) { return make_float3(rd_f(b, off), rd_f(b, off + 4), rd_f(b, off + 8)); }

// the 544-byte TRenderOpts layout (renderer.cl:35-78; same table as rm_api.cu:decode_opts)
void decode_opts(const void* blob, RmOpts* o) {
  const uint8_t* b = static_cast<const uint8_t*>(blob);
  o->eyePos = rd_f3(b, 0); o->targetPos = rd_f3(b, 16); o->up = rd_f3(b, 32);
  o->voxelBounds = rd_f3(b, 48); o->voxelBounds2 = rd_f3(b, 64);
  o->boundsMin = rd_f3(b, 80); o->boundsMax = rd_f3(b, 96);
  o->invVoxelScale = rd_f3(b, 112); o->sky1 = rd_f3(b, 128); o->sky2 = rd_f3(b, 144);
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

// CPU restatement of rm_accel.cu
struct HostAccel {
  std::vector<uint64_t> solid, occ;
  std::vector<uint8_t> dist, nib;  // nib: rm_accel.cu:k_pack_nibbles restated
  RmAccel view{};
};

void build_accel(const uint8_t* vox, int rx, int ry, int rz, int iso, int cell_shift, HostAccel& A) {
  RmAccel& a = A.view;
  a.vox = vox;
  a.bx = (rx + 3) >> 2; a.by = (ry + 3) >> 2; a.bz = (rz + 3) >> 2;
  a.cell_shift = cell_shift < 2 ? 2 : cell_shift;
  const int cell = 1 << a.cell_shift;
  a.cellf = (float)cell;
  a.rxf = (float)rx; a.ryf = (float)ry; a.rzf = (float)rz;
  a.inv_rxf = 1.0f / a.rxf; a.inv_ryf = 1.0f / a.ryf; a.inv_rzf = 1.0f / a.rzf;
  a.pow2 = ((rx & (rx - 1)) == 0 && (ry & (ry - 1)) == 0 && (rz & (rz - 1)) == 0) ? 1 : 0;
  a.mx = (rx + cell - 1) >> a.cell_shift; a.my = (ry + cell - 1) >> a.cell_shift; a.mz = (rz + cell - 1) >> a.cell_shift;
  A.solid.assign((size_t)a.bx * a.by * a.bz, 0);
  A.occ.assign(A.solid.size(), 0);
  for (int z = 0; z < rz; ++z)
    for (int y = 0; y < ry; ++y)
      for (int x = 0; x < rx; ++x) {
        const int v = vox[((size_t)z * ry + y) * rx + x];
        const size_t b = ((size_t)(z >> 2) * a.by + (y >> 2)) * a.bx + (x >> 2);
        const unsigned bit = (x & 3) | ((y & 3) << 2) | ((z & 3) << 4);
        if (v > iso) A.solid[b] |= 1ull << bit;
        if (v >= iso) A.occ[b] |= 1ull << bit;
      }
  const size_t nc = (size_t)a.mx * a.my * a.mz;
  std::vector<uint8_t> seed(nc, RM_DIST_CAP);
  const int bpc = cell >> 2;
  for (int bz = 0; bz < a.bz; ++bz)
    for (int by = 0; by < a.by; ++by)
      for (int bx = 0; bx < a.bx; ++bx)
        if (A.solid[((size_t)bz * a.by + by) * a.bx + bx])
          seed[((size_t)(bz / bpc) * a.my + by / bpc) * a.mx + bx / bpc] = 0;
  // brute-force-free separable Chebyshev transform (same recurrence as k_cheb_axis)
  std::vector<uint8_t> tmp(nc);
  const int ext[3] = {a.mx, a.my, a.mz};
  const long long str[3] = {1, a.mx, (long long)a.mx * a.my};
  std::vector<uint8_t>* in = &seed; std::vector<uint8_t>* out = &tmp;
  for (int axis = 0; axis < 3; ++axis) {
    for (int cz = 0; cz < a.mz; ++cz)
      for (int cy = 0; cy < a.my; ++cy)
        for (int cx = 0; cx < a.mx; ++cx) {
          const long long c = ((long long)cz * a.my + cy) * a.mx + cx;
          const int pos = axis == 0 ? cx : (axis == 1 ? cy : cz);
          int best = (*in)[c];
          for (int k = 1; k < best; ++k) {
            int lo = RM_DIST_CAP, hi = RM_DIST_CAP;
            if (pos - k >= 0) lo = (*in)[c - k * str[axis]];
            if (pos + k < ext[axis]) hi = (*in)[c + k * str[axis]];
            const int m = std::max(k, std::min(lo, hi));
            best = std::min(best, m);
          }
          (*out)[c] = (uint8_t)best;
        }
    std::swap(in, out);
  }
  A.dist = *in;
  a.solid = A.solid.data();
  a.occ = A.occ.data();
  a.dist = A.dist.data();
  A.nib.assign(((nc + 1) / 2 + 15) & ~(size_t)15, 0);
  for (size_t c = 0; c < nc; ++c) A.nib[c >> 1] |= (uint8_t)(std::min<int>(A.dist[c], 15) << ((c & 1) * 4));
  a.nib = A.nib.data();
  a.nib_bytes = (unsigned)A.nib.size();
}

float* g_cost_out = nullptr;  // optional: count x 64 floats, per-site cost of every rendered pixel-sample
unsigned short* g_eval_out = nullptr;  // optional: count x 256 entries, the evaluation log of every pixel-sample (0xffff = end)
HostAccel g_host_accel;
const uint8_t* g_accel_vox = nullptr;
int g_accel_key[5] = {0, 0, 0, -1, -1};
SimStats g_total;

}  // namespace

extern "C" {

// One RenderImage pass over the listed pixel ids (or all n pixels), in place on `pixels`, exactly
// like the oracle's orc_render_pixels. mode: 0 = production routine over the occupancy tables,
// 1 = counting routine over the occupancy tables, 2 = routine over the raw byte volume.
// counters (3 x u64, optional) accumulate the reference-equivalent work in modes 1 and 2.
void sim_render_pixels(const uint8_t* vox, const float* mc, const void* opts544, float* pixels, int n,
                       const int* ids, int nids, unsigned long long* counters, int mode, int cell_shift) {
  RmOpts o;
  std::memset(&o, 0, sizeof o);
  decode_opts(opts544, &o);
  if (mode != 2) {
    const int key[5] = {o.rx, o.ry, o.rz, o.isoVal, cell_shift};
    if (g_accel_vox != vox || std::memcmp(key, g_accel_key, sizeof key) != 0) {
      build_accel(vox, o.rx, o.ry, o.rz, o.isoVal, cell_shift, g_host_accel);
      g_accel_vox = vox;
      std::memcpy(g_accel_key, key, sizeof key);
    }
    plain::g_accel = g_host_accel.view;
    fused::g_accel = g_host_accel.view;
    fused::rm_host_nib = g_host_accel.nib.data();
  }
  plain::g_opts = o;
  fused::g_opts = o;
  const int count = ids ? nids : n;
  unsigned long long cs = 0, ct = 0, co = 0;
#pragma omp parallel reduction(+ : cs, ct, co)
  {
    std::memset(&t_stats, 0, sizeof t_stats);
#pragma omp for schedule(dynamic, 64)
    for (int k = 0; k < count; ++k) {
      const int id = ids ? ids[k] : k;
      plain::Scene s(vox, reinterpret_cast<const float4*>(mc));
      std::memset(t_site_cost, 0, sizeof t_site_cost);
      t_site = 0; t_level = 0; t_eval_n = 0;
      float3 c;
      // modes 3..6: the default kernel's routine (rm_scene_fused.cuh): production / counting over the
      // 4-bit map (what the kernel reads from shared memory), production / counting over the byte map
      const fused::Lane lane{reinterpret_cast<const float4*>(mc), o.time};
      fused::Cnt<true> fc;
      if (mode == 0) c = plain::render_pixel_sample<false>(s, plain::BrickVolume{}, id);
      else if (mode == 1) c = plain::render_pixel_sample<true>(s, plain::BrickVolume{}, id);
      else if (mode == 2) c = plain::render_pixel_sample<true>(s, plain::ByteVolume{vox}, id);
      else if (mode == 3 && cell_shift == 2 && fused::g_accel.pow2) c = fused::render_pixel_sample<false, 7>(fused::Cnt<false>{}, lane, id);
      else if (mode == 3 && fused::g_accel.pow2) c = fused::render_pixel_sample<false, 5>(fused::Cnt<false>{}, lane, id);
      else if (mode == 3 && cell_shift == 2) c = fused::render_pixel_sample<false, 3>(fused::Cnt<false>{}, lane, id);
      else if (mode == 3) c = fused::render_pixel_sample<false, 1>(fused::Cnt<false>{}, lane, id);
      else if (mode == 4 && cell_shift == 2) c = fused::render_pixel_sample<true, 3>(fc, lane, id);
      else if (mode == 4) c = fused::render_pixel_sample<true, 1>(fc, lane, id);
      else if (mode == 5 && cell_shift == 2 && fused::g_accel.pow2) c = fused::render_pixel_sample<false, 6>(fused::Cnt<false>{}, lane, id);
      else if (mode == 5 && fused::g_accel.pow2) c = fused::render_pixel_sample<false, 4>(fused::Cnt<false>{}, lane, id);
      else if (mode == 5 && cell_shift == 2) c = fused::render_pixel_sample<false, 2>(fused::Cnt<false>{}, lane, id);
      else if (mode == 5) c = fused::render_pixel_sample<false, 0>(fused::Cnt<false>{}, lane, id);
      else if (cell_shift == 2) c = fused::render_pixel_sample<true, 2>(fc, lane, id);
      else c = fused::render_pixel_sample<true, 0>(fc, lane, id);
      if (mode == 4 || mode == 6) { s.w.steps = fc.steps; s.w.taps = fc.taps; s.w.outer = fc.outer; }
      float* px = pixels + 4 * (size_t)id;
      const float3 m = lerp3(make_float3(px[0], px[1], px[2]), c, o.frameBlend);  // mix(), renderer.cl:492
      px[0] = m.x; px[1] = m.y; px[2] = m.z; px[3] = 1.0f;
      cs += s.w.steps; ct += s.w.taps; co += s.w.outer;
      if (g_cost_out) std::memcpy(g_cost_out + 64 * (size_t)k, t_site_cost, sizeof t_site_cost);
      if (g_eval_out) {
        unsigned short* dst = g_eval_out + 256 * (size_t)k;
        for (int e = 0; e < 256; ++e) dst[e] = e < t_eval_n ? t_eval_log[e] : (unsigned short)0xffff;
      }
    }
#pragma omp critical
    {
      unsigned long long* dst = reinterpret_cast<unsigned long long*>(&g_total);
      const unsigned long long* src = reinterpret_cast<const unsigned long long*>(&t_stats);
      for (size_t i = 0; i < sizeof(SimStats) / 8; ++i) dst[i] += src[i];
    }
  }
  if (counters) { counters[0] += cs; counters[1] += ct; counters[2] += co; }
}

void sim_set_cost_buffer(float* p) { g_cost_out = p; }
void sim_set_eval_buffer(unsigned short* p) { g_eval_out = p; }
// The wavefront stages of rm_wave.cuh run on the host in the launcher's order (rm_render_wave.cu):
// primary; per level L: [finish(L-1)], prepare(L), all queued traces; finish(last); final.
// mode: 0 = production routine, 1 = counting routine.
void sim_render_pixels_wave(const uint8_t* vox, const float* mc, const void* opts544, float* pixels, int n,
                            const int* ids, int nids, unsigned long long* counters, int mode, int cell_shift) {
  RmOpts o;
  std::memset(&o, 0, sizeof o);
  decode_opts(opts544, &o);
  const int key[5] = {o.rx, o.ry, o.rz, o.isoVal, cell_shift};
  if (g_accel_vox != vox || std::memcmp(key, g_accel_key, sizeof key) != 0) {
    build_accel(vox, o.rx, o.ry, o.rz, o.isoVal, cell_shift, g_host_accel);
    g_accel_vox = vox;
    std::memcpy(g_accel_key, key, sizeof key);
  }
  plain::g_accel = g_host_accel.view;
  plain::g_opts = o;
  const int count = ids ? nids : n;
  const int lmax = std::min(o.reflectIter < 0 ? 0 : o.reflectIter, wave::kMaxLevels - 1);
  std::vector<wave::WaveRec> rec((size_t)wave::kMaxLevels * count);
  std::vector<float4> refl(count);
  std::vector<float2> pxy(count);
  std::vector<wave::WaveJob> jobs((size_t)(o.numLights + 1) * count + 1);
  unsigned njobs = 0;
  wave::WaveBuf B{rec.data(), refl.data(), pxy.data(), jobs.data(), &njobs, (unsigned)count, (unsigned)jobs.size(), lmax + 2 < wave::kMaxLevels ? lmax + 2 : wave::kMaxLevels};
  const float4* table = reinterpret_cast<const float4*>(mc);
  unsigned long long cs = 0, ct = 0, co = 0;
  const plain::BrickVolume V{};
#define SIM_STAGE(BODY)                                                        \
  _Pragma("omp parallel for schedule(dynamic, 64) reduction(+ : cs, ct, co)") \
  for (int it = 0; it < count; ++it) {                                         \
    plain::Scene s(vox, table);                                                \
    BODY;                                                                      \
    cs += s.w.steps; ct += s.w.taps; co += s.w.outer;                          \
  }
  if (mode == 0) { SIM_STAGE(wave::wave_primary<false>(B, (unsigned)it, s, V, ids ? ids[it] : it)) }
  else { SIM_STAGE(wave::wave_primary<true>(B, (unsigned)it, s, V, ids ? ids[it] : it)) }
  for (int L = 0; L <= lmax; ++L) {
    njobs = 0;
    if (mode == 0) { SIM_STAGE(if (L >= 2) wave::wave_finish<false>(B, (unsigned)it, s, V, L - 1); wave::wave_prepare<false>(B, (unsigned)it, s, V, L)) }
    else { SIM_STAGE(if (L >= 2) wave::wave_finish<true>(B, (unsigned)it, s, V, L - 1); wave::wave_prepare<true>(B, (unsigned)it, s, V, L)) }
    const int nj = (int)std::min<size_t>(njobs, jobs.size());
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : cs, ct, co)
    for (int k = 0; k < nj; ++k) {
      plain::Scene s(vox, table);
      if (mode == 0) wave::wave_trace<false>(B, jobs[k], s, V);
      else wave::wave_trace<true>(B, jobs[k], s, V);
      cs += s.w.steps; ct += s.w.taps; co += s.w.outer;
    }
  }
#pragma omp parallel for schedule(dynamic, 64)
  for (int it = 0; it < count; ++it) {
    plain::Scene s(vox, table);
    const int id = ids ? ids[it] : it;
    float3 c;
    if (mode == 0) { if (lmax >= 1) wave::wave_finish<false>(B, (unsigned)it, s, V, lmax); c = wave::wave_final<false>(B, (unsigned)it, s); }
    else { if (lmax >= 1) wave::wave_finish<true>(B, (unsigned)it, s, V, lmax); c = wave::wave_final<true>(B, (unsigned)it, s); }
    float* px = pixels + 4 * (size_t)id;
    const float3 m = lerp3(make_float3(px[0], px[1], px[2]), c, o.frameBlend);
    px[0] = m.x; px[1] = m.y; px[2] = m.z; px[3] = 1.0f;
  }
#undef SIM_STAGE
  if (counters) { counters[0] += cs; counters[1] += ct; counters[2] += co; }
}

// The shard layout exactly as the kernels compute it (rm_types.h:rm_shard_layout, rm_kernels.h:rm_slot_to_pixel):
// fills out[0 .. slots) with the pixel id of every work slot of `rank` (-1 = padding), returns the slot count.
long long sim_shard_slots(int W, int H, int rank, int world, int tile_w, int tile_h, int* out, long long cap) {
  RmShard s{};
  s.rank = rank; s.world = world; s.tile_w = tile_w; s.tile_h = tile_h;
  rm_shard_layout(s, W, H);
  for (long long i = 0; i < s.slots && i < cap; ++i) out[i] = rm_slot_to_pixel(s, i, W, H);
  return s.slots;
}

int sim_pick_passes(int available) { return rm_persist_pick_passes(available); }
// out[3] = threads, blocks per SM, map in shared memory
void sim_pick_layout(long long bundles, int num_sms, unsigned nib_bytes, int block_threads, int smem_map, int counting, int* out) {
  const RmPersistLayout l = rm_persist_pick_layout(bundles, num_sms, nib_bytes, block_threads, smem_map, counting);
  out[0] = l.threads; out[1] = l.blocks_per_sm; out[2] = l.use_nib;
}

int sim_stats_words(void) { return (int)(sizeof(SimStats) / 8); }
void sim_get_stats(unsigned long long* out, int reset) {
  std::memcpy(out, &g_total, sizeof g_total);
  if (reset) std::memset(&g_total, 0, sizeof g_total);
}
void sim_set_num_threads(int n) { omp_set_num_threads(n); }

}  // extern "C"
