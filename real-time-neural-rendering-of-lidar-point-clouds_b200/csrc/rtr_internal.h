// Internal definition of the renderer object shared by rtr_renderer.cu (hot path + C ABI) and
// rtr_io.cu (loaders / post-process, SURVEY.md §8 f).  Not installed.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../include/rtr_b200.h"
#include "rtr_kernels.h"

typedef struct ncclComm* ncclComm_t;

namespace rtr {
constexpr int kFrameSets = 4;  // a fused frame sequence: z-min of frame k, blend of k-1, image passes / D2H of k-2, and the set being cleared for k+1
struct FrameSet {
    FrameBuffers fb{};
    // accum | zbuf | image live in ONE allocation (one CUDA IPC handle per set for the point-sharded merge)
    uint8_t* arena = nullptr;
    size_t arena_bytes = 0, zbuf_off = 0, image_off = 0;
    bool f32acc = false;  // the last frame rendered into this set accumulated colour sums as floats
    cudaEvent_t rendered = nullptr, copied = nullptr;
    // fused frame sequences: `points_done` = the point pass that z-min'ed this set's frame has finished
    cudaEvent_t points_done = nullptr, cleared = nullptr;
    bool clean = false;  // a point pass has cleared this set for the next frame of the sequence (no clear launch needed)
    // chunk-level frustum culling state of the frame rendered into this set (rtr_cull.cu); per set so that two frames
    // can be in flight on two streams
    uint32_t* vis_list = nullptr;
    uint32_t* band_scratch = nullptr;  // kMaxBands segments of band_cap entries: screen-band ordering of the list (BandSort)
    uint32_t band_cap = 0;
    CullState* cull_state = nullptr;
    uint32_t cull_parity = 0;  // alternates per culled frame (CullState::n_visible double buffer)
};

// What a frame needs besides its buffers: fixed when the frame is enqueued (the pose may change before its blend runs).
struct FramePlan {
    ProjParams pp;
    CullParams cp;
    bool cull = false;      // chunk-level frustum culling applies
    bool use_ring = false;  // the point passes run through the TMA-fed kernels
    uint32_t bands = 1;     // screen bands the visible list is ordered by (1: list order)
};
// A frame of a fused sequence whose z-min pass has been enqueued and whose blend / image passes have not: they are
// enqueued together with the NEXT frame's z-min (one stream of chunks for both) or by flush_pending().
struct PendingFrame {
    bool active = false;
    int si = 0, stage = 0;
    FramePlan plan;
    uint8_t* bgr = nullptr;  // host destinations of the frame's D2H copies (trajectory call), or null
    float* depth = nullptr;
};
}  // namespace rtr

struct rtr_renderer {
    int device = 0;
    int sm_count = 148;
    // `stream`: everything; `stream2`: every other frame of an asynchronous frame sequence (option "pipeline")
    // `stream`: everything; `stream2`: every other frame of a two-pass sequence (option "pipeline"); `image_stream`: the image
    // passes of fused sequences (higher priority); `copy_stream`: D2H
    cudaStream_t stream = nullptr, stream2 = nullptr, stream3 = nullptr, image_stream = nullptr, copy_stream = nullptr;  // stream3: third frame of a two-pass sequence (option "pipeline_depth" = 3)
    // cloud
    rtr::PointRecord* points = nullptr;
    uint64_t n_points = 0;
    bool owns_points = false;
    uint64_t index_base = 0;  // global index of local point 0 (point sharding)
    // chunk-level frustum culling (rtr_cull.cu)
    rtr::ChunkBounds* bounds = nullptr;
    uint32_t n_chunks = 0;
    rtr::RingSchedule ring_sched{};  // stream-all tile order of the ring kernels (rtr_point_ring.cu), fixed at upload
    // camera
    int W = 0, H = 0;
    double K[9] = {0};
    double dist[5] = {0};
    double E[16] = {0};
    bool have_K = false, have_E = false, raw_proj = false;
    float cam_proj[16] = {0};
    double cull_rstar = 0;  // distortion: normalised radius beyond which nothing reaches the image (make_params)
    // rtr_host_distortion_bounds of the current intrinsics (12 K samples of the radial polynomial): computed once per
    // rtr_set_intrinsics*, not once per frame
    bool dist_bounds_valid = false;
    double dist_r2_max = 0, dist_rstar = 0;
    // frame buffers (two sets, see header)
    rtr::FrameSet set[rtr::kFrameSets];
    int cur = 0;
    rtr::PendingFrame pending;
    int fuse = 1;  // asynchronous frame sequences stream each chunk once per frame (blend of frame k-1 + z-min of frame k in one pass):
                   // 0 never, 2 always, 1 = for clouds of >= 40 000 chunks (where it pays, see fused_sequence)
    int alloc_W = 0, alloc_H = 0;
    rtr::PyramidDims dims{};
    bool masks_allocated = false, key64_allocated = false;
    // options
    int zmin_variant = 5, zmin_unroll = 4, blend_variant = 4, blend_unroll = 4;
    int force_generic = 0, keep_masks = 0, timing = 0, key64 = 0, chunk_cull = 1, sort_on_upload = 1;
    int fused_up = 1;  // the four up-pass levels in one launch (needs W % 16 == 0 and keep_masks = 0)
    // Views in which a pixel collects more than 65 793 points make every frame pay for the exact re-run of the colour
    // sums.  The re-run leaves a note in a mapped host word; the next `int_sum_frames` frames then use integer sums from
    // the start (results identical either way), after which float sums are tried again.
    uint32_t* overflow_note = nullptr;      // pinned, mapped host word
    uint32_t* overflow_note_dev = nullptr;  // its device alias
    int int_sum_frames = 0;
    int pipeline_depth = 3;  // whole frames in flight in a two-pass (non-fused) sequence: 3 (default) or 2 frame sets / streams in rotation (profiles/r02T_exp_pipeline_depth.json)
    int pipeline = 1;  // asynchronous frame sequences (rtr_render_device, rtr_render_trajectory) rotate through pipeline_depth frame
                       // sets AND streams, so frame i+1's point passes overlap frame i's image passes (off with peers / NCCL / timing)
    int ring_early = 1;  // blend ring pass requests its first chunks before the PDL wait (0: measurement only)
    int ring_dynamic = 8;  // list passes of the ring kernels: tiles beyond a CTA's first ring-full are claimed from this many counters (0: round-robin)
    int ring_claim_min = 12;  // tiles are claimed only in list passes with more tiles per CTA than this (0: always)
    int fixup_launches = 1;  // fused sequences: the exact re-run gate as one cluster of 8 CTAs (1), a small cooperative grid (2) or three gated launches (3)
    int fused_tiles_per_cta = 0;  // fused pass: 0 = persistent grid (2 CTAs per SM claim tiles until the list is empty); t > 0 = about t tiles
                                  // per CTA, i.e. many short-lived CTAs between which the block scheduler slots the image stream's CTAs
    uint32_t* tiles_hint = nullptr;      // mapped pinned word: tiles of the last fused pass (sizes the next launch when fused_tiles_per_cta > 0)
    uint32_t* tiles_hint_dev = nullptr;
    int ring_ctas = 2;  // ring-kernel CTAs per SM: 2 fill the SM's shared memory, 1 leaves room for the other frame's ring kernel
    int clear_lean = 1;  // clear + classify compiled for <= 64 registers, so that it fits beside the other frame's ring kernel (0: 80 registers)
    int ring_perm = 1;  // stream-all ring passes visit the chunks in a low-discrepancy order (0: storage order)
    int ring = 1;  // point passes through the TMA-fed persistent kernels: 1 = for culled frames, 2 = always, 0 = never
    int bands = 0;  // visible list ordered by screen band (BandSort): 1 = off, 2 ... 8 = that many bands, 0 = by the size of the frame
                    // buffers (off while z-buffer + colour sums of the frames in flight fit the L2)
    cudaEvent_t ev[6] = {nullptr};
    // timing == 2: per-frame event sextets from a pool, summed on demand (bench roofline leg)
    std::vector<cudaEvent_t> ev_pool;
    int ev_frames = 0;
    double ev_sum[6] = {0, 0, 0, 0, 0, 0};
    uint64_t ev_count = 0;
    uint64_t launches = 0;
    // point-sharded merge over peer memory (rtr_peer.cu): device pointers of every rank's buffers
    struct Peer {
        bool attached = false;
        bool exported = false;               // a fresh rtr_peer_export (flags zeroed, epochs reset) has not been attached yet
        int rank = 0, n = 1;
        uint32_t* flags = nullptr;           // local flag array + [64] local barrier counter
        uint32_t* peer_flags[rtr::kMaxPeers] = {nullptr};
        // every rank's buffers of frame sets 0 / 1: [what][set][rank], what = 0 z-buffer, 1 colour sums, 2 64-bit keys, 3 image
        void* peer_buf[4][2][rtr::kMaxPeers] = {{{nullptr}}};
        std::vector<void*> opened;           // cudaIpcOpenMemHandle results to close
        uint32_t epoch = 1, local_base = 0;
        int W = 0, H = 0;
        uint32_t* err_host = nullptr;        // mapped pinned word a timed-out wait raises (checked after every synchronisation)
        uint32_t* err_dev = nullptr;
        int timeout_ms = 10000;              // option "peer_timeout_ms"
    } peer;
    // persistent scratch of rtr_postprocess_unet_output (no per-call cudaMalloc/cudaFree)
    uint8_t* post_scratch = nullptr;
    size_t post_scratch_bytes = 0;
    // comm
    ncclComm_t comm = nullptr;
    int rank = 0, n_ranks = 1;
    std::string err;
};

namespace rtr {
int renderer_fail(rtr_renderer* r, int code, const std::string& msg);
// Frees the chunk bounds and both frame sets' visible-list storage.
void free_cull_storage(rtr_renderer* r);
// Waits until the compute streams are idle (does not enqueue anything: see flush_pending).
cudaError_t sync_compute(rtr_renderer* r);
// Enqueues the blend and image passes of a fused sequence's last frame, if one is outstanding.
int flush_pending(rtr_renderer* r);
// Frees the current cloud and allocates room for n records (r->points, owned).
int replace_cloud(rtr_renderer* r, uint64_t n);
// (Re)builds the chunk bounds for the cloud in r->points; every upload path ends with it.
int build_chunk_bounds(rtr_renderer* r);
// Morton re-ordering of the owned resident cloud (rtr_io.cu); ends with build_chunk_bounds.
int reorder_morton(rtr_renderer* r);
// What every upload path calls last: Morton-sort when option sort_on_upload is set, then the chunk bounds.
int finish_upload(rtr_renderer* r);
}  // namespace rtr
