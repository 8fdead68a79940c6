/* raymarch_b200.h -- C ABI of libraymarch_b200.so
 *
 * Drop-in boundary for the one hot path of thi-ng/raymarchcl: the OpenCL pipeline
 *   {:write [p-buf v-buf]} -> per pass {:write [o-buf mc-buf]} + kernel "RenderImage"
 *   -> {:write q-buf} + kernel "TonemapImage" + blocking read of q-buf
 * that /root/reference/src/thi/ng/raymarchcl/core.clj:76-97 (make-pipeline) describes and
 * core.clj:171 / :204 (ops/execute-pipeline) runs through thi.ng/simplecl -> JOCL.
 * A JVM host keeps render-options + structgen unchanged and swaps execute-pipeline for these
 * calls via JNA. Plain pointers and sizes only; no C++/torch types.
 *
 * Data conventions (identical to what the reference uploads):
 *   voxels  uint8[rz][ry][rx], index z*rx*ry + y*rx + x           (renderer.cl:167, io.clj:19-33)
 *   opts    the 544-byte TRenderOpts blob produced by sg/encode    (renderer.cl:35-78, core.clj:105)
 *   mc      16384 float4 = 65536 floats, the scatter table         (renderer.cl:143, core.clj:138)
 *   accum   float4[W*H] RGBA accumulator ("p-buf")                 (core.clj:144)
 *   argb    uint32[W*H] 0xFFRRGGBB ("q-buf", what :read [:out] returns) (renderer.cl:503-506)
 *
 * Errors: every call returns RM_OK (0) or a negative rm_status; nothing throws or aborts across
 * the ABI; rm_last_error() returns the message of the last failure on that context (or of the
 * last failed rm_create when ctx == NULL).
 * Threading: a context may be used from one thread at a time; distinct contexts are independent.
 * Ownership: the caller owns every host pointer; the library has copied what it needs when a
 * call returns. All device memory belongs to the context.
 */
#ifndef RAYMARCH_B200_H
#define RAYMARCH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RM_OPTS_BYTES   544    /* sizeof(TRenderOpts), renderer.cl:35-78 */
#define RM_TABLE_FLOATS 65536  /* 0x4000 float4, renderer.cl:143 / core.clj:138 */
#define RM_ABI_VERSION  2

typedef struct rm_ctx rm_ctx;

typedef enum rm_status {
  RM_OK = 0,
  RM_ERR_INVALID_ARG = -1,   /* null pointer, wrong blob size, non-positive extent ... */
  RM_ERR_BAD_OPTS = -2,      /* TRenderOpts disagrees with the uploaded volume / framebuffer */
  RM_ERR_NO_VOLUME = -3,     /* render before rm_set_volume */
  RM_ERR_NO_FRAMEBUFFER = -4,/* render / tonemap / read before rm_clear_accum */
  RM_ERR_CUDA = -5,          /* a CUDA runtime call failed; see rm_last_error */
  RM_ERR_NO_DEVICE = -6,     /* no usable sm_100 device */
  RM_ERR_UNSUPPORTED = -7,
  RM_ERR_IO = -8             /* rm_load_volume_file: unreadable / malformed .vox file */
} rm_status;

/* Work and timing of the calls since the last rm_reset_stats (or rm_create).
 * steps / taps / outer_iters are REFERENCE-EQUIVALENT counts: iterations of the inner march loop
 * (renderer.cl:219-234), calls of voxelLookupI (renderer.cl:172-178) and iterations of the
 * sphere-trace loop (renderer.cl:243-251) that the reference would execute on the same inputs.
 * They are gathered only while counting is enabled (rm_set_option RM_OPT_COUNT_WORK). */
typedef struct rm_stats {
  uint64_t steps;
  uint64_t taps;
  uint64_t outer_iters;
  uint64_t pixel_samples;   /* pixels x passes rendered */
  uint64_t kernel_launches; /* CUDA kernels launched by this context */
  uint64_t render_launches; /* ... of which RenderImage-equivalent kernels (render_ms covers these) */
  double   render_ms;       /* device time of RenderImage-equivalent kernels (CUDA events) */
  double   tonemap_ms;      /* device time of TonemapImage-equivalent kernels */
  double   h2d_ms, d2h_ms;  /* device time of copies issued by this context */
  uint64_t h2d_bytes, d2h_bytes;
} rm_stats;

typedef enum rm_option {
  RM_OPT_COUNT_WORK = 1,   /* 0 (default) | 1: gather reference-equivalent work counters */
  RM_OPT_KERNEL = 2,       /* which RenderImage kernel: 0 = default: persistent blocks over the bit-brick volume,
                              blend + tonemap folded in (distance map optionally staged into shared memory by bulk TMA);
                              1 = plain, over the raw byte volume (comparison kernel);
                              2 = warp-scheduled persistent state machine over the bit-brick volume;
                              3 = wavefront pipeline (stages + a persistent, refilling trace kernel);
                              4 = one thread per (pixel, pass) over the bit-brick volume + blend kernel
                              (round 1's default). All five produce identical results. */
  /* tuning knobs of the fast kernel; none of them changes results */
  RM_OPT_CELL_SHIFT = 3,   /* macro-cell edge of the distance map = 1<<value voxels; 0 = auto (4 voxels up to 1024^3) */
  RM_OPT_FUSE_LIMIT = 6,   /* max passes rendered by one launch (1..32) */
  RM_OPT_TRIP_LIMIT = 7,   /* watchdog of kernel 2: scheduling trips a warp may take per launch before
                              the launch is abandoned with RM_ERR_CUDA (default 2^28) */
  RM_OPT_WAVE_CHUNK = 8,   /* kernel 3: (pixel, pass) items per chunk of the pipeline (1024..2^24, default 2^24) */
  RM_OPT_WAVE_REFILL = 9,  /* kernel 3: the trace kernel hands new rays to idle lanes once this many lanes of a
                              warp are idle (1 = at once ... 32 = the warp starts 32 rays together) */
  RM_OPT_PERSIST_BLOCK = 10, /* kernel 0: resident block layout: 1024 (x 1 block per SM, 64 registers) or 256 (x 5, 48 registers)
                               threads; 0 (default) = per launch: 1024 when the distance map is staged into shared memory
                               (next option) and the launch is long enough to pay for it, else 256 */
  RM_OPT_PERSIST_SMEM = 12,  /* kernel 0: stage the 4-bit distance map into shared memory by bulk TMA (cp.async.bulk + mbarrier)
                               when a copy per resident block fits the SM (128 KiB at 256^3: 1024-thread layout only):
                               0 = never (byte map from global memory / L1), 1 = whenever it fits the layout in use,
                               2 (default) = with the 1024-thread layout (measured: DESIGN.md 4) */
  RM_OPT_PERSIST_ORDER = 13, /* kernel 0: 1 (default) = walk the frame bottom-up so that the launch ends on the (cheap) top rows; 0 = top-down */
  RM_OPT_PERSIST_GROUP = 11  /* kernel 0: 0 = every warp draws its next work bundle on its own; 1 = the warps of a block draw
                               one bundle each together and meet at the block barrier per draw; -1 = default (0) */
} rm_option;

/* ---- lifetime (replaces cl/select-platform .. cl/init-state, core.clj:121-128; cl/release :213) ---- */
int  rm_abi_version(void);
int  rm_device_count(void);
int  rm_create(int device_id, rm_ctx** out_ctx);
/* One context over the n GPUs device_ids[0..n-1] of a box (the reference is single-device: cl/max-device,
 * core.clj:121-123 -- a host changes exactly that one call site). The context takes every call of this
 * header that is not marked single-GPU: inputs are uploaded once and broadcast device-to-device, image tiles
 * are dealt to the GPUs (diagonal stripes of 16x8 tiles), every GPU's render kernel stores the ARGB words of
 * its tiles directly into device_ids[0]'s frame over NVLink, and rm_tonemap reads that frame back. Needs peer
 * access from every device to device_ids[0] (NVSwitch boxes have it); RM_ERR_UNSUPPORTED otherwise.
 * Single-GPU only: rm_set_stream, rm_tonemap_device, rm_copy_accum_device, rm_set_argb_target,
 * rm_unpack_shards; rm_set_tile_shard(ctx, 0, 1, w, h) sets the tile size. */
int  rm_create_multi(const int* device_ids, int n, rm_ctx** out_ctx);
/* GPUs behind a context (1 for rm_create), and the stats of one of them (rm_get_stats on a multi-GPU
 * context sums the work and takes the slowest member's times). */
int  rm_member_count(const rm_ctx* ctx);
int  rm_get_member_stats(const rm_ctx* ctx, int member, rm_stats* out);
void rm_destroy(rm_ctx* ctx);
const char* rm_last_error(const rm_ctx* ctx);

/* ---- inputs ---- */
/* v-buf upload (vio/load-volume, io.clj:19-33 + {:write [.. v-buf]}, core.clj:81). */
int rm_set_volume(rm_ctx* ctx, const uint8_t* voxels, int rx, int ry, int rz);
/* The same from DEVICE memory (a volume assembled over NVLink or written by another kernel): one
 * device-to-device copy, queued on the context's stream -- the source must stay valid (and unchanged) until
 * that copy has run: rm_sync, or any later blocking call of this context. */
int rm_set_volume_device(rm_ctx* ctx, const void* d_voxels, int rx, int ry, int rz);
/* The same from a .vox file as vio/save-volume writes it (io.clj:9-17: "VOXEL", 3 x int32 big-endian,
 * element-size byte, raw bytes): reads through pinned memory and uploads. The extents are returned
 * through the optional out pointers (the caller needs them for TRenderOpts.voxelRes). */
int rm_load_volume_file(rm_ctx* ctx, const char* path, int* out_rx, int* out_ry, int* out_rz);
/* p-buf / q-buf allocation + zero fill (ops/init-buffers core.clj:140-145, {:write [p-buf ..]} :81). */
int rm_clear_accum(rm_ctx* ctx, int width, int height);

/* ---- the hot path ---- */
/* One {:write [o-buf mc-buf]} + "RenderImage" launch (core.clj:84-89): host opts + table in,
 * accumulator updated in place on the device. */
int rm_render_pass(rm_ctx* ctx, const void* opts, size_t opts_len, const float* mc, size_t mc_floats);
/* All `iter` passes of a frame in submission order, from host buffers (core.clj:82-90).
 * opts[i] / mc[i] are per-pass pointers, as make-render-option-buffer / :mc-buffers hold them. */
int rm_render_frame(rm_ctx* ctx, const void* const* opts, const float* const* mc, int iter);
/* "TonemapImage" with pass-0 opts + :read [:out] (core.clj:91-97): ARGB words to host memory. */
int rm_tonemap(rm_ctx* ctx, const void* opts, size_t opts_len, uint32_t* argb_out);
/* Parity hook: copy the fp32 RGBA accumulator to the host (the reference never reads p-buf). */
int rm_read_accum(rm_ctx* ctx, float* rgba_out);
/* "TonemapImage" + an ASYNCHRONOUS :read [:out] for animation loops (test-anim, core.clj:199-212): returns
 * once the work is queued; the ARGB words land in argb_out (ideally pinned host memory) some time before
 * rm_wait(ctx, slot) returns. Two frames are kept on the device, so the render calls that follow write
 * the next frame while this one is still travelling. slot is 0 or 1. */
int rm_tonemap_async(rm_ctx* ctx, const void* opts, size_t opts_len, uint32_t* argb_out, int slot);
int rm_wait(rm_ctx* ctx, int slot);
/* Page-locked host memory for rm_tonemap_async / rm_render_frame inputs (a JVM host wraps it in a direct
 * ByteBuffer): copies to and from it are truly asynchronous and run at full PCIe rate. */
int rm_host_alloc(rm_ctx* ctx, size_t bytes, void** out_ptr);
int rm_host_free(rm_ctx* ctx, void* ptr);

/* ---- resident-input variants (inputs already in HBM when the timed region starts) ---- */
/* Store the per-pass opts blobs and tables on the device once (test-anim keeps them across frames,
 * core.clj:189-208) ... */
int rm_upload_passes(rm_ctx* ctx, const void* const* opts, const float* const* mc, int iter);
/* (mc == NULL: use the tables produced in place by rm_generate_scatter_tables.) */
/* update-render-option-buffer (core.clj:108-117) for the resident passes: new opts blobs (a new camera),
 * same tables. iter must equal the uploaded pass count. 544 bytes per pass cross the bus. */
int rm_update_opts(rm_ctx* ctx, const void* const* opts, int iter);
/* ... then render passes [first, first+count) from the resident copies; no host traffic. */
int rm_render_resident(rm_ctx* ctx, int first, int count);
/* Tonemap into device memory the caller owns (e.g. a torch tensor used for the NCCL gather);
 * `packed` != 0 writes only this context's tile shard, tile-major (see rm_set_tile_shard). */
int rm_tonemap_device(rm_ctx* ctx, const void* opts, size_t opts_len, void* d_argb, int packed);
/* Register caller-owned device memory (width*height words, or rm_shard_slots() words when packed) that
 * the default render kernel fills with the ARGB words of the frame WHILE it renders; a later
 * rm_tonemap_device(opts, same pointer, same packing) with the launch's gamma then costs nothing. The
 * memory may belong to a peer GPU with access enabled (every GPU stores its tiles into one frame over
 * NVLink). NULL restores the context's own frame. */
int rm_set_argb_target(rm_ctx* ctx, void* d_argb, int packed);
/* Copy the accumulator into device memory the caller owns (same packing rule). */
int rm_copy_accum_device(rm_ctx* ctx, void* d_rgba, int packed);
int rm_sync(rm_ctx* ctx);
/* Issue all further work of this context on a CUDA stream the caller owns (a cudaStream_t passed
 * as void*; NULL restores the context's own stream), so that the caller's events time it. */
int rm_set_stream(rm_ctx* ctx, void* cuda_stream);

/* ---- inputs generated on the device (the reference builds them on the JVM and uploads them) ---- */
/* gen/make-gyroid-volume (generators.clj:27-42) straight into the context's volume; byte-identical
 * to the host generator raymarchcl_b200/generators.py:make_gyroid_volume. Replaces rm_set_volume. */
int rm_generate_gyroid_volume(rm_ctx* ctx, int rx, int ry, int rz);
/* gen/make-terrain (generators.clj:44-60) likewise; byte-identical to generators.py:make_terrain.
 * Needs rz >= rx (the reference indexes its second wall with x as the slice). */
int rm_generate_terrain_volume(rm_ctx* ctx, int rx, int ry, int rz);
/* gen/generate-scatter-offsets (generators.clj:8-16) for java.util.Random seeds seed0 .. seed0+count-1
 * (the reference seeds from nanoTime) into the resident table slots 0 .. count-1. */
int rm_generate_scatter_tables(rm_ctx* ctx, int64_t seed0, int count);
/* meshvoxel/voxelize (ks < 0; meshvoxel.clj:60-69) and meshvoxel/voxelize-ks (ks >= 0; :45-58) on the
 * device: `xyz` = n_points mesh vertices (3 floats each, e.g. from mio/read-stl :12-14) are mapped
 * into a res^3 grid by mesh-scale (:16-23, fp64 like the JVM) and splatted with value 255 (single
 * voxel, or a (2ks+1)^3 cube clamped to the grid). The result replaces the context's volume.
 * NaN / infinite coordinates are rejected with RM_ERR_INVALID_ARG. */
int rm_voxelize_points(rm_ctx* ctx, const float* xyz, int64_t n_points, int res, int ks);
/* Parity hook: copy the resident volume (rx*ry*rz bytes) to the host. */
int rm_read_volume(rm_ctx* ctx, uint8_t* voxels_out);

/* ---- multi-GPU: interleaved tile ownership ---- */
/* This context renders only tiles t with t % world == rank, tiles being tile_w x tile_h pixel
 * rectangles numbered row-major. world == 1 (default) renders everything. */
int rm_set_tile_shard(rm_ctx* ctx, int rank, int world, int tile_w, int tile_h);
/* Number of pixels this context owns under the current shard and framebuffer. */
int64_t rm_shard_pixels(const rm_ctx* ctx);
/* Work slots (pixels + padding of edge tiles) of shard `rank` of `world` under the current
 * framebuffer and tile size = the element count of that rank's packed buffer. Rank 0's is the largest. */
int64_t rm_shard_slots(const rm_ctx* ctx, int rank, int world);
/* After the one gather: de-interleave the `world` packed per-rank buffers (device memory,
 * `stride_slots` elements apart, elem_bytes = 4 for ARGB words or 16 for float4 accumulators)
 * into a full frame in device memory (width*height elements). */
int rm_unpack_shards(rm_ctx* ctx, const void* d_parts, int world, int64_t stride_slots, int elem_bytes, void* d_frame);

/* ---- options, stats ---- */
int rm_set_option(rm_ctx* ctx, int option, int64_t value);
int rm_get_stats(const rm_ctx* ctx, rm_stats* out);
int rm_reset_stats(rm_ctx* ctx);


#ifdef __cplusplus
}
#endif
#endif /* RAYMARCH_B200_H */
