/* rtr_b200 — C ABI of the B200-native point-projection path for RTRenderer.
 *
 * Drop-in boundary (SURVEY.md §8 b).  The reference exposes a C++ class, not an FFI:
 *     class ProjectCloud            /root/reference/src/RTRenderer/include/project_cloud.h:11-60
 *       ProjectCloud(grid, model)                 project_cloud.cu:189-251
 *       computeRGBD(calib, w2c, color*, depth*)   project_cloud.cu:268-312
 *       computeFilteredRGBD(...)                  project_cloud.cu:394-434
 *       computeFull(...)  [projection+filter part, tensor hand-off]  project_cloud.cu:437-471
 * Each entry point below names the reference member it replaces.  The header-only C++ adapter
 * include/rtr_b200/project_cloud.hpp wraps these back into a class with the reference's method
 * names; INTEGRATION.md shows the binding a maintainer of the reference would add.
 *
 * Conventions (same as the reference):
 *   - pose / extrinsics = WORLD->CAMERA 4x4, row-major doubles (cv::Matx44d; example passes pose.inv()).
 *   - colour is B,G,R interleaved uint8 (CV_8UC3), depth is float32 (CV_32F), both H*W, contiguous,
 *     caller-allocated; either pointer may be NULL.
 *   - return 1 on success (the reference's `return 1`), negative on error (never exit()).
 *   - one renderer = one GPU = one CUDA stream; distinct renderers may be driven from distinct
 *     threads/processes (frame sharding).  A renderer is not re-entrant.
 * No CPU fallback exists: every call fails with RTR_ERR_CUDA if no sm_100 device is usable.
 */
#ifndef RTR_B200_H
#define RTR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

typedef struct rtr_renderer rtr_renderer;

enum {
    RTR_OK = 1,
    RTR_ERR_ARG = -1,      /* bad argument (computeRGBD returns -1 when both outputs are NULL) */
    RTR_ERR_CUDA = -2,     /* CUDA runtime / launch failure; see rtr_last_error() */
    RTR_ERR_STATE = -3,    /* no cloud uploaded / no camera set */
    RTR_ERR_UNSUPPORTED = -4,
    RTR_ERR_COMM = -5      /* NCCL failure, or a peer of the point-sharded merge did not answer (the frame is invalid) */
};

/* render stages for rtr_render_device() */
enum {
    RTR_STAGE_RGBD = 0,      /* clear + z-min + blend + resolve        == computeRGBDInternal   */
    RTR_STAGE_FILTERED = 1   /* ... + depth prefilter + tensor write   == + applyDepthFilter     */
};

/* ---- lifetime.  Replaces ProjectCloud::ProjectCloud / ~ProjectCloud (project_cloud.cu:189-266). */
int rtr_create(int device, rtr_renderer** out);
void rtr_destroy(rtr_renderer* r);
/* Last error text of this renderer (or of the failed rtr_create when r == NULL). */
const char* rtr_last_error(const rtr_renderer* r);

/* ---- cloud upload.  Replaces OctreeGrid::getVertexPositions/getVertexColors + the two cudaMemcpy
 * in the constructor (Octreegrid.h:162-180, project_cloud.cu:191-206).  The cloud is repacked into
 * 16-byte records {x, y, z, b|g<<8|r<<16|255<<24}; point order is irrelevant to every output. */
int rtr_upload_cloud_xyz_bgr(rtr_renderer* r, const float* xyz, const uint8_t* bgr, uint64_t n_points);
int rtr_upload_cloud_packed16(rtr_renderer* r, const void* host_records, uint64_t n_points);
/* Zero-copy: use records already resident on this renderer's device (caller keeps ownership). */
int rtr_adopt_device_cloud_packed16(rtr_renderer* r, void* device_records, uint64_t n_points);
/* Fill the cloud on the device with the deterministic synthetic hall (bench/test support; same
 * generator as oracle/rtr_oracle.c:rtro_synth_packed).  Points [first, first+count) of n_total. */
int rtr_synth_cloud(rtr_renderer* r, uint64_t seed, uint64_t n_total, uint64_t first, uint64_t count,
                    int lx_q, int ly_q, int lz_q, int n_boxes);
uint64_t rtr_cloud_size(const rtr_renderer* r);
/* Copy records [first, first+count) back to the host (tests). */
int rtr_download_cloud_packed16(rtr_renderer* r, uint64_t first, uint64_t count, void* host_records);

/* ---- camera.  Replaces the CameraCalibration argument (CameraCalibration.h:8-54: W, H, K; the
 * reference ignores its distortion vector) and the extrinsics argument. */
int rtr_set_intrinsics(rtr_renderer* r, int width, int height, double fx, double fy, double cx, double cy,
                       double skew, const double* dist5 /* k1,k2,p1,p2,k3 or NULL */);
int rtr_set_intrinsics_matrix(rtr_renderer* r, int width, int height, const double* K9_row_major,
                              const double* dist5);
int rtr_set_pose_w2c(rtr_renderer* r, const double* E16_row_major);
/* Parity hook: bypass K*E and use this row-major float[16] camProj verbatim (project_cloud.cu:318-320). */
int rtr_set_cam_proj_raw(rtr_renderer* r, const float* m16_row_major);
/* The camProj the next frame will use (row-major float[16]). */
int rtr_get_cam_proj(const rtr_renderer* r, float* m16_row_major);

/* ---- render one frame, results to HOST buffers (synchronous, like the reference). */
int rtr_render_rgbd(rtr_renderer* r, uint8_t* bgr, float* depth);      /* computeRGBD          */
int rtr_render_filtered(rtr_renderer* r, uint8_t* bgr, float* depth);  /* computeFilteredRGBD  */
/* computeFull's projection + prefilter; *device_fp16 receives the device pointer of the
 * 1x5xHxW fp16 U-Net input (torch::from_blob it, project_cloud.cu:471).  Stream-synchronised. */
int rtr_render_tensor(rtr_renderer* r, void** device_fp16);

/* ---- device-resident / asynchronous variants (no host copies, no sync).  Back-to-back frames form a SEQUENCE
 * (options "pipeline" and "fuse", default 1): frame k's blend shares one stream of chunks with frame k+1's z-min
 * (consecutive poses see nearly the same part of the cloud, so every chunk is read from HBM once per frame instead of
 * twice), the image passes run on a second stream and the clears on a third, over three frame-buffer sets.  A frame's
 * blend and image passes are therefore enqueued with the NEXT rtr_render_device call — or by rtr_sync,
 * rtr_get_device_buffers, rtr_read_buffer, rtr_set_option and every blocking render call, all of which complete the
 * outstanding frame first.  Frames are byte-identical to the blocking calls'.  rtr_get_device_buffers /
 * rtr_read_buffer refer to the frame enqueued last; rtr_get_device_buffers makes `stream` wait for every frame in
 * flight.  "fuse": 1 (default) = fused sequences for clouds of >= 40 000 chunks (41 M points)
 * (below that a frame is a few short kernels and two passes per frame, whole frames rotating through three frame sets
 * on three streams — option "pipeline_depth" — are faster), 0 = never, 2 = always. */
int rtr_render_device(rtr_renderer* r, int stage);
int rtr_sync(rtr_renderer* r);
/* Render n_frames poses (n_frames x 16 doubles, world->camera) back to back.  bgr/depth, when not
 * NULL, receive n_frames images each (pinned or pageable host memory); copies overlap rendering.
 * `stage` as above.  Returns after everything completed. */
int rtr_render_trajectory(rtr_renderer* r, int stage, const double* poses_w2c, int n_frames, uint8_t* bgr,
                          float* depth);

typedef struct {
    void* points;        /* n x 16 B records                                     */
    uint32_t* zbuf;      /* W*H u32 depth bits (float view = depth; filtered: -1 where masked) */
    uint32_t* accum;     /* W*H*4 u32 {sum b, sum g, sum r, count}               */
    uint8_t* image;      /* W*H*3 u8 BGR                                         */
    uint16_t* tensor;    /* 5*W*H fp16; planes packed at stride tensor_plane     */
    uint32_t* minmax;    /* {min, max} depth bits over valid pixels              */
    float* level[5];     /* pyramid levels (level[0] == zbuf)                    */
    uint8_t* mask[4];    /* up-pass masks when option keep_masks=1, else NULL    */
    int width, height;
    int level_w[5], level_h[5];   /* true pyramid dims */
    int up_w[5], up_h[5];         /* dims the up-pass uses (reference truncation) */
    uint64_t tensor_plane;        /* up_w[0]*up_h[0] */
    void* stream;                 /* cudaStream_t the renderer launches on.  Work enqueued here after rtr_get_device_buffers
                                     sees the frame.  (The library's kernels trigger programmatic dependents early: a kernel
                                     launched on this stream WITH the programmatic-stream-serialization attribute must
                                     execute griddepcontrol.wait / cudaGridDependencySynchronize before reading a buffer;
                                     plainly launched kernels — torch, cuDNN, TensorRT — are ordered as usual.) */
} rtr_device_buffers;
int rtr_get_device_buffers(rtr_renderer* r, rtr_device_buffers* out);

/* Copy a device buffer of the current frame to the host after syncing the stream (tests/taps).
 * what: 0 zbuf  1 accum  2 image  3 tensor  4 minmax  5..8 level 1..4  9..12 mask 0..3 */
int rtr_read_buffer(rtr_renderer* r, int what, void* dst, size_t bytes);
/* Per-point projection tap: pix (int32, -1 = culled) and depth bits for every point (tests). */
int rtr_project_points(rtr_renderer* r, int32_t* pix_host, uint32_t* zbits_host);

/* ---- options / introspection.  An option applies to the frames enqueued after the call.  Known keys: "zmin_variant"
 * (bit0 early test, bit1 warp aggregation, bit2 L1-cached test, bit6 = 64: shared-memory tile pre-reduction in the
 * z-min ring pass of two-pass frames; the measurement-only bits 8/16/32, whose frames are
 * wrong by design, are rejected unless the library was built with -DRTR_EXPERIMENTS), "zmin_unroll", "blend_variant", "blend_unroll",
 * "force_generic", "keep_masks", "timing", "key64", "chunk_cull", "ring", "ring_dynamic" (default 8: the ring kernels'
 * list passes claim their tiles from this many counters; 0 = round-robin), "ring_claim_min" (default 12: passes with no more tiles per CTA than this stay
 * round-robin), "ring_ctas" (ring-kernel CTAs per SM, 2 or 1),
 * "clear_lean", "fused_up", "pipeline", "fuse", "pipeline_depth" (3 (default) or 2: whole frames of a two-pass, non-fused frame
 * sequence in flight at once, each in its own frame set on its own stream),
 * "bands" (the frame's visible-chunk list ordered by horizontal screen band, so that the tiles in flight share a band
 * of the z-buffer / colour sums: 1 = list order, 2 ... 8 = that many bands, 0 (default) = 8 bands for frames whose
 * z-buffer + colour sums exceed the 126 MB L2, e.g. 3840x2160, list order otherwise; frames are identical either way;
 * rtr_get_option "bands_active" = what the next frame will use),
 * "sort_on_upload" (default 1: every upload
 * re-orders the cloud along a Morton curve on the GPU — no output depends on point order; set 0 BEFORE uploading
 * to keep the input order, e.g. when the point index of the 64-bit key must be the caller's index). */
int rtr_set_option(rtr_renderer* r, const char* key, int64_t value);
int64_t rtr_get_option(const rtr_renderer* r, const char* key);
/* With option timing=1: CUDA-event ms of the last frame's stages
 * {clear, zmin, blend, resolve+pyramid, up-pass, total}. */
int rtr_get_stage_ms(rtr_renderer* r, float* ms6);
/* With option timing=2 every frame records its own six events (pooled); this returns the per-stage
 * SUMS in ms over all frames rendered since the last reset, and how many frames that was.
 * timing=3 does the same for fused sequences, per point pass that carries both halves:
 * {chunk classification for two cameras, fused pass: blend k-1 + z-min k (both on the point stream), wait,
 *  resolve + fix-up gate of frame k-1, up-pass of frame k-1 (image stream), first to last event} — the two streams
 * overlap one another across frames.  (timing 1 / 2 render whole frames one after the other.) */
int rtr_get_stage_ms_sum(rtr_renderer* r, double* ms6_sum, uint64_t* n_frames, int reset);
/* Chunk-level frustum culling statistics since the last reset: frames rendered with culling, the sum
 * over those frames of the frame's visible chunks (1024 consecutive points), and the cloud's
 * chunk count.  Option "chunk_cull" (default 1) switches the culling; results are identical. */
int rtr_get_cull_stats(rtr_renderer* r, uint64_t* frames, uint64_t* visible_chunks_total, uint64_t* n_chunks, int reset);
/* What the point passes really read from HBM since the last reset: the number of passes over a visible-chunk list
 * and the chunks they streamed (16 KB each).  A frame rendered on its own walks its list twice (z-min, blend); a
 * fused sequence walks the union of two consecutive frames' lists once per frame.  Resets the same counters as
 * rtr_get_cull_stats. */
int rtr_get_stream_stats(rtr_renderer* r, uint64_t* passes, uint64_t* chunks_streamed, int reset);
/* Statistics of the shared-memory tile pre-reduction (option zmin_variant bit 6 = 64, two-pass frames) since the last
 * reset of the culling statistics: {tiles whose pixels fitted the 32 x 32 shared-memory window, tiles that went
 * straight to global memory, window pixels flushed (one RED candidate each), records that entered a window}. */
int rtr_get_smem_tile_stats(rtr_renderer* r, uint64_t* stats4);
/* Number of kernel launches issued by this renderer since creation. */
uint64_t rtr_launch_count(const rtr_renderer* r);

/* Measurement support: RED.MIN throughput into a W*H L2-resident buffer (u32, or u64 with key64).
 * mode 0: n_ops uniformly random addresses generated in registers; mode 1: the pixel ids the
 * current cloud + camera project to (n_ops = cloud size; *live_ops = in-frustum points, the number
 * of REDs really issued); modes 2 / 3: the colour sums' update at a random pixel as two RED.ADD.64 / one
 * RED.ADD.F32x4; modes 4 + 4*same + log2(L) (RED.MIN.U32) and 12 + 4*same + log2(L) (RED.ADD.F32x4), L = 1, 2, 4, 8:
 * groups of L consecutive lanes of a warp instruction share one random 32-byte sector — distinct words of it
 * (same = 0) or one word (same = 1).  Returns the mean CUDA-event time of one launch. */
int rtr_bench_red_min(rtr_renderer* r, int mode, uint64_t n_ops, int key64, int iters, float* ms_per_launch,
                      uint64_t* live_ops);

/* Host-only helper (no GPU needed; what the renderer uses internally for a camera with distortion coefficients):
 * r2_max = squared normalised radius beyond which the distorted projection culls a point (fold-back guard),
 * rstar  = normalised, undistorted radius beyond which no point can reach the W x H image (0: unknown) — the bound
 * the chunk-level culling tests against.  K9 row-major 3x3, dist5 = k1, k2, p1, p2, k3. */
int rtr_host_distortion_bounds(int width, int height, const double* K9, const double* dist5, double* r2_max, double* rstar);

/* Host-only helper: the multiplier m of the order in which a stream-all pass of the ring kernels (option ring = 2)
 * visits the cloud's 1024-point chunks, tile t -> chunk (t * m) mod n_chunks; coprime with n_chunks, i.e. a permutation. */
uint32_t rtr_host_ring_stride(uint64_t n_points);

/* Host-only helper: the copy that ends a band-ordered classification (option "bands"), replayed on the host with the
 * kernel's own index arithmetic: `threads` threads walk ONE flat loop over the 16-byte vectors of the n_bands segments
 * (cap entries each, cap % 4 == 0; counts8[b] valid entries in segment b) and write the segments, band after band,
 * into list_out; *n_out = entries written. */
int rtr_host_band_compact(const uint32_t* scratch, uint32_t cap, const uint32_t* counts8, uint32_t n_bands, uint32_t threads,
                          uint32_t* list_out, uint32_t* n_out);

/* Host-only helper: how the ring kernels' list passes hand out tiles (option ring_dynamic = n_queues > 0).  A launch of
 * `grid` CTAs, each with *groups_per_cta consumer groups and a ring of *stages stages: the CTA's first *stages tiles are
 * tiles block + k * grid; every tile from *stages * grid on is entry `claim` of one of n_queues queues, and consumer
 * group `group` of CTA `block` claims from *queue only (one atomic counter per queue; n_queues is first clamped to the
 * number of groups, grid * *groups_per_cta, so that no queue is left without a group).  Returns the queue of that group
 * and the tile its claim number `claim` stands for. */
int rtr_host_ring_claim(uint32_t grid, uint32_t n_queues, uint32_t block, uint32_t group, uint32_t claim, uint32_t* queue,
                        uint32_t* tile, uint32_t* stages, uint32_t* groups_per_cta);

/* Device self-test of the ring kernels' perspective divide: for n_pairs random bit patterns (a, b) (every class of
 * float: NaN, inf, denormal, huge) checks on the GPU that whenever !(|b| < 2^-126) the directly issued
 * MUFU.RCP + FMUL gives the same bits as __fdividef(a, b) (what the reference compiles, render.cu:65-66), and the same
 * rounded pixel coordinate.  *mismatches = number of pairs that differ (must be 0). */
int rtr_selftest_fast_divide(rtr_renderer* r, uint64_t n_pairs, uint64_t seed, uint64_t* mismatches);

/* ---- point-sharded multi-GPU (one process per GPU; plumbing by the caller, e.g. torch.distributed).
 * rtr_comm_unique_id fills a 128-byte NCCL id on rank 0; broadcast it, then every rank calls
 * rtr_comm_init.  With a communicator attached, rendering merges the per-GPU z-buffers with
 * ncclAllReduce(min) and the colour sums with ncclAllReduce(sum) so every rank holds the result a
 * single GPU with all points would produce (bit-identical: integer min / integer add). */
int rtr_comm_unique_id(void* id128);
int rtr_comm_init(rtr_renderer* r, const void* id128, int rank, int n_ranks);
int rtr_comm_destroy(rtr_renderer* r);

/* The same merge without a library collective: every rank maps the others' frame buffers (CUDA IPC) and one
 * kernel per buffer does a two-shot all-reduce (min / integer sum) straight over NVLink peer memory, with epoch
 * flags for the cross-GPU barriers (csrc/rtr_peer.cu).  One process per GPU on one node:
 *   every rank: set the intrinsics, rtr_peer_export(r, blob)  ->  all-gather the 512-byte blobs by any means
 *   every rank: rtr_peer_attach(r, all_blobs, rank, n_ranks)  ->  every render call now merges across the ranks.
 * Resolution changes need detach / export / attach again.  Takes precedence over rtr_comm_init. */
#define RTR_PEER_BLOB_BYTES 512
int rtr_peer_export(rtr_renderer* r, void* blob512);
int rtr_peer_attach(rtr_renderer* r, const void* blobs, int rank, int n_ranks);
int rtr_peer_detach(rtr_renderer* r);

const char* rtr_version(void);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* RTR_B200_H */
