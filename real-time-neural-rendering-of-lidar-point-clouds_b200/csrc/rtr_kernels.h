// Host-callable launchers of the sm_100a kernels (definitions in rtr_point_kernels.cu,
// rtr_image_kernels.cu, rtr_synth.cu).  Every launcher only enqueues on `s` and returns the launch
// status; nothing here synchronises.
#pragma once
#include "rtr_common.cuh"

namespace rtr {

constexpr int kPointBlock = 256;    // threads per CTA of the two point passes
constexpr int kChunkPoints = 1024;  // points per culling chunk (= one CTA tile at unroll 4)
// Float accumulators ({b,g,r,count} as 4 x f32) hold exact integers while count <= 65793
// (255 * 65793 = 2^24 - 1).  A pixel beyond that raises the frame's overflow flag in resolve.
constexpr float kF32ExactCount = 65793.0f;

// Axis-aligned bounds of one chunk of kChunkPoints consecutive records (rtr_cull.cu).
struct ChunkBounds {
    float lo[3], hi[3];
    uint32_t always_visible;  // a coordinate is NaN/inf or huge: never cull (NaN points are live in the reference)
    uint32_t pad;
};
// Per-renderer culling state in device memory.
struct CullState {
    uint32_t n_visible[2];         // [parity]: chunks in vis_list for the frame being rendered; [parity ^ 1] is
                                   // zeroed meanwhile for the next frame (classification shares a launch with the clear)
    uint32_t parity;               // which counter the frame in flight uses (written by clear_classify_kernel)
    uint32_t armed;                // a classified frame has not been folded into the totals yet
    uint32_t frames;               // frames folded into total_visible
    uint32_t kind;                 // what the list in flight is: kListPair | have flags (classify_pair_kernel) or kListWhole
    unsigned long long total_visible;  // sum over the folded frames of the frame's visible chunks
    // two-camera lists (classify_pair_kernel): how many entries of the list in flight carry each flag
    uint32_t n_blend[2], n_zmin[2];
    unsigned long long total_streamed;  // chunk reads from HBM: entries of every pair list, 2 x entries of every whole-frame list
    uint32_t passes, pad;               // point passes folded into total_streamed
    // screen-band ordering of the list (BandSort below): entries per band of the classification in flight, and the
    // ticket that tells the classification's last CTA that it is the last
    uint32_t band_count[8];
    uint32_t band_ticket, pad2[3];
};
static_assert(sizeof(CullState) <= 256, "the tile-claim counters start 256 bytes behind the CullState");
constexpr uint32_t kListHasBlend = 1u, kListHasZmin = 2u, kListWhole = 4u;
// Fold the list in flight (counter `old`) into the totals — device (next classification) and host (statistics read-out).
__host__ __device__ inline void cull_fold(CullState* c, uint32_t old) {
    if (c->kind & kListWhole) {  // one camera, walked by the z-min pass and again by the blend pass
        c->total_visible += c->n_visible[old];
        c->frames += 1u;
        c->total_streamed += 2ull * c->n_visible[old];
        c->passes += 2u;
    } else {
        if (c->kind & kListHasZmin) { c->total_visible += c->n_zmin[old]; c->frames += 1u; }
        c->total_streamed += c->n_visible[old];
        c->passes += 1u;
    }
}
// A visible-list entry is a chunk id (< 2^30) plus, in lists built for TWO cameras (the fused point pass: frame k-1's
// blend and frame k's z-min over one stream of chunks), which of the two passes the chunk takes part in.  Lists built
// for one camera carry no flags; every consumer masks the id.
constexpr uint32_t kTileBlend = 1u << 30;   // the chunk can hold a point of the PREVIOUS frame's frustum: blend it
constexpr uint32_t kTileZmin = 1u << 31;    // the chunk can hold a point of THIS frame's frustum: z-min it
constexpr uint32_t kTileIdMask = kTileBlend - 1u;
// The allocation that holds a CullState continues with the ring kernels' tile-claim counters of the frame in flight:
// kMaxTileQueues counters per point pass ([0] z-min, [1] blend), one per 128-byte line (same-address atomics
// serialise in L2), zeroed by clear_classify_kernel.
constexpr int kMaxTileQueues = 64;
constexpr int kTileQueueStride = 32;  // uint32 words between two counters
constexpr size_t kCullStateAlloc = 256 + size_t(2) * kMaxTileQueues * kTileQueueStride * sizeof(uint32_t) + 64;
__host__ __device__ inline uint32_t* tile_counters(CullState* c, int pass) {
    return reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(c) + 256) + size_t(pass) * kMaxTileQueues * kTileQueueStride;
}
// ... and ends with the statistics of the shared-memory tile pre-reduction (zmin_variant bit 6): [0] tiles whose pixels
// fitted the shared-memory window, [1] tiles that went straight to global memory, [2] pixels flushed from windows,
// [3] records that entered a window.  Never reset by the frame's kernels.
__host__ __device__ inline unsigned long long* smem_tile_stats(CullState* c) {
    return reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(c) + 256 + size_t(2) * kMaxTileQueues * kTileQueueStride * sizeof(uint32_t));
}
__host__ __device__ inline uint32_t cull_count(const CullState* c) { return c->n_visible[c->parity & 1u]; }
// The frame's camera for the chunk test, in double (exact images of the float camProj rows).
struct CullParams {
    double r0[4], r1[4], r2[4];
    double W, H;
};
// Screen-band ordering of the visible list.  The order in which a pass walks its tiles changes no result (min and
// integer / exact-float sums are order-free), but it decides WHICH z-buffer / accumulator lines the CTAs in flight are
// hitting at any one time.  In list order (= Morton order of the cloud) that is the whole frame; once the frame
// buffers outgrow the L2 (3840x2160: 33 MB z-buffer + 133 MB colour sums per frame against 126 MB) every reduction
// and gather then goes to DRAM.  With n_bands > 1 the classification files every visible chunk under the horizontal
// screen band its box centre projects into (per-band segments of `scratch`, `cap` entries each) and its last CTA
// copies the segments band after band into the list, so the tiles in flight — a window of the list — share a band
// and the band's lines stay in L2 while they are hot.  n_bands <= 1: the list is written directly, as before.
constexpr int kMaxBands = 8;
struct BandSort {
    uint32_t n_bands;
    uint32_t cap;        // entries per band segment (a multiple of 4, >= the cloud's chunk count)
    uint32_t* scratch;   // kMaxBands * cap entries, one allocation per frame set
};
// The arithmetic of the segments -> list copy (band_compact in rtr_cull.cu; replayed on the host by
// rtr_host_band_compact for the CPU tests).  The copy is ONE flat loop over the 16-byte vectors of all segments:
// voff[b] / eoff[b] = vectors / entries in front of band b.
struct BandLayout {
    uint32_t voff[kMaxBands + 1], eoff[kMaxBands + 1];
};
__host__ __device__ inline void band_layout(const uint32_t (&cnt)[kMaxBands], uint32_t n_bands, BandLayout& L) {
    L.voff[0] = 0u;
    L.eoff[0] = 0u;
#pragma unroll
    for (int b = 0; b < kMaxBands; ++b) {
        const uint32_t c = uint32_t(b) < n_bands ? cnt[b] : 0u;
        L.voff[b + 1] = L.voff[b] + ((c + 3u) >> 2);
        L.eoff[b + 1] = L.eoff[b] + c;
    }
}
// Flat vector v (< voff[kMaxBands]) -> its band, its index lv inside the band's segment, the list position dst0 of its
// first entry and how many of its four entries exist (the rest of the vector is stale scratch).  Select chain with
// compile-time indices only: the arrays stay in registers on the device.
__host__ __device__ inline void band_locate(const BandLayout& L, const uint32_t (&cnt)[kMaxBands], uint32_t v, uint32_t& band, uint32_t& lv,
                                            uint32_t& dst0, uint32_t& left) {
    uint32_t vb = 0u, eb = 0u, cb = cnt[0];
    band = 0u;
#pragma unroll
    for (int k = 1; k < kMaxBands; ++k)
        if (v >= L.voff[k]) { vb = L.voff[k]; eb = L.eoff[k]; cb = cnt[k]; band = uint32_t(k); }
    lv = v - vb;
    dst0 = eb + lv * 4u;
    const uint32_t rest = cb - lv * 4u;
    left = rest < 4u ? rest : 4u;
}

// Pyramid geometry exactly as applyDepthFilter derives it (project_cloud.cu:336-362): true level
// dims are halved (floor) four times on the way down, the up-pass re-doubles the level-4 dims.
struct PyramidDims {
    int w[5], h[5];    // true dims of L_0..L_4 (w[0] = W, h[0] = H); buffers are w[i]*h[i] floats
    int uw[5], uh[5];  // dims the up-pass indexes level i with: uw[4] = w[4], uw[i] = 2*uw[i+1]
};
inline PyramidDims make_pyramid_dims(int W, int H) {
    PyramidDims d;
    d.w[0] = W; d.h[0] = H;
    for (int i = 1; i <= 4; ++i) { d.w[i] = d.w[i - 1] / 2; d.h[i] = d.h[i - 1] / 2; }
    d.uw[4] = d.w[4]; d.uh[4] = d.h[4];
    for (int i = 3; i >= 0; --i) { d.uw[i] = d.uw[i + 1] * 2; d.uh[i] = d.uh[i + 1] * 2; }
    return d;
}
inline uint64_t clear_coverage(int W, int H) {  // fillBuffer/resolvePass grid: (W/16)*(H/16) blocks of 256
    uint64_t c = uint64_t(W / 16) * uint64_t(H / 16) * 256u, p = uint64_t(W) * H;
    return c < p ? c : p;
}

// Launch with the PDL attribute (see pdl_prologue in rtr_common.cuh).  RTR_PDL=0 in the environment
// falls back to plain stream serialisation (A/B measurements).
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline void launch_pdl_smem(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem_bytes, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// For kernels with a grid-wide spin barrier (exact_fixup_kernel, peer_allreduce_kernel): the cooperative attribute makes
// the driver guarantee that all CTAs are co-resident — or run the grid after whatever occupies the SMs — also when
// another renderer shares the GPU.  If the driver rejects the cooperative + PDL pair the launch is retried cooperative
// only; it is NEVER downgraded to a plain launch (a spin barrier without the co-residency guarantee can deadlock):
// the error is returned instead.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_cooperative(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = 0;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
    if (e != cudaSuccess && cfg.numAttrs == 2) {
        (void)cudaGetLastError();
        cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
    }
    return e;
}
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t s, Args&&... args) {
    launch_pdl_smem(kernel, grid, block, 0, s, static_cast<Args&&>(args)...);
}

// ---- point passes (rtr_point_kernels.cu)
cudaError_t launch_clear(cudaStream_t s, int sm_count, uint32_t* zbuf, uint64_t cov, uint32_t* accum, uint64_t n_px,
                         uint32_t* minmax, CullState* cull);
// Point passes over the frame's visible chunks only (persistent grid over the compacted list).
cudaError_t launch_zmin_list(cudaStream_t s, int sm_count, int variant, const PointRecord* pts, uint64_t n,
                             uint64_t index_base, const ProjParams& pp, const CullState* cull, const uint32_t* vis_list,
                             uint32_t* zbuf, unsigned long long* zkey);
cudaError_t launch_blend_list(cudaStream_t s, int sm_count, int variant, const PointRecord* pts, uint64_t n,
                              const ProjParams& pp, const CullState* cull, const uint32_t* vis_list, const uint32_t* zbuf,
                              uint32_t* accum, const uint32_t* gate, uint32_t need_flag = 0u);

// ---- the point passes as persistent TMA-fed kernels (rtr_point_ring.cu) — the default (option "ring")
// Tile t of a launch is chunk vis_list[t] (list = true: the frame's visible chunks, count read on the device from
// `cull`) or chunk (t * perm_mul) mod n_chunks (list = false: every chunk, in a low-discrepancy order).
struct RingSchedule {
    const CullState* cull;
    const uint32_t* vis_list;
    uint32_t n_chunks;
    uint32_t perm_mul;
    uint32_t early;  // blend pass over the list: request the first chunks before the PDL wait (0: measurement only)
    uint32_t* tile_counter;  // list passes: consumer groups claim their tiles beyond the CTA's first ring-full from
                             // n_queues counters (tile_counters()), so that faster SMs take more tiles; null = round-robin
    uint32_t n_queues;       // 1 ... kMaxTileQueues
    uint32_t claim_min_tiles_per_cta;  // list passes with no more tiles per CTA than this stay round-robin (default 12)
    uint32_t ctas_per_sm;    // host side: CTAs launched per SM (2 fill an SM's shared memory; 1 leaves room for another frame's ring kernel)
    uint32_t grid_override;  // host side, fused pass: launch this many CTAs instead of a persistent grid (0: persistent); tiles are
                             // then dealt round-robin, a CTA takes a handful of tiles and leaves, and the SM's block scheduler can
                             // slot the image stream's CTAs in between (option fused_tiles_per_cta)
    uint32_t* tiles_hint;    // mapped host word: the fused pass reports how many tiles its list held (sizes the next launch)
    uint32_t zero;           // always 0, and not known to the compiler: (loaded word & zero) is how ring_walk makes an instruction
                             // wait for a load without changing its operands
};
// Claimed tiles (list passes): a CTA's first `stages` tiles are tiles blockIdx.x + k * grid of the launch; the tiles from
// stages * grid on are dealt round-robin into n_queues queues, entry c of queue q being this tile.  Every tile index
// >= stages * grid is entry (t - stages * grid) / n_queues of queue (t - stages * grid) % n_queues: a bijection.
__host__ __device__ inline uint32_t ring_claimed_tile(uint32_t grid, uint32_t stages, uint32_t n_queues, uint32_t queue, uint32_t claim) {
    return stages * grid + claim * n_queues + queue;
}
// Every queue needs a group that claims from it: with fewer consumer groups than queues the tiles dealt to the
// orphaned queues would never be streamed.  The launchers (and rtr_host_ring_claim) clamp the queue count with this.
__host__ __device__ inline uint32_t ring_effective_queues(uint32_t grid, uint32_t groups_per_cta, uint32_t n_queues) {
    const uint32_t groups = grid * groups_per_cta;
    const uint32_t q = n_queues < groups ? n_queues : groups;
    return q < 1u ? 1u : q;
}
// The queue a consumer group claims from.
__host__ __device__ inline uint32_t ring_queue_of(uint32_t block, uint32_t group, uint32_t groups_per_cta, uint32_t n_queues) {
    return (block * groups_per_cta + group) % n_queues;
}
RingSchedule make_ring_schedule(uint64_t n_points, const CullState* cull, const uint32_t* vis_list);
void ring_geometry(uint32_t* stages, uint32_t* groups_per_cta, uint32_t* ctas_per_sm);  // the ring kernels' compile-time shape
cudaError_t launch_zmin_ring(cudaStream_t s, int sm_count, int variant, const PointRecord* pts, uint64_t n,
                             uint64_t index_base, const ProjParams& pp, const RingSchedule& sc, bool list, uint32_t* zbuf,
                             unsigned long long* zkey);
cudaError_t launch_blend_ring(cudaStream_t s, int sm_count, int variant, const PointRecord* pts, uint64_t n,
                              const ProjParams& pp, const RingSchedule& sc, bool list, const uint32_t* zbuf, uint32_t* accum);
// One stream of chunks for two frames: tiles flagged kTileBlend are blended for the camera pp_blend against
// zbuf_blend (complete) into accum_blend, tiles flagged kTileZmin are z-min'ed for the camera pp_zmin into zbuf_zmin.
// The list (sc.vis_list / sc.cull) comes from launch_classify_pair.  blend_variant bit 2: float colour sums.
// `clear`: frame buffers the pass leaves cleared for the frame after next (fillBuffer + cudaMemset + min/max reset of
// a set nobody reads any more) — every CTA clears its slice once it has run out of tiles; all-null = nothing to clear.
struct ClearTarget {
    uint32_t* zbuf;      // [0, cov) <- FLT_MAX bits
    uint64_t cov;
    uint4* accum;        // [0, n_px) <- 0
    uint64_t n_px;
    uint32_t* minmax;    // {UINT_MAX, 0, 0, 0}
};
cudaError_t launch_fused_ring(cudaStream_t s, int sm_count, int zmin_variant, int blend_variant, const PointRecord* pts,
                              uint64_t n, const ProjParams& pp_blend, const ProjParams& pp_zmin, const RingSchedule& sc,
                              const uint32_t* zbuf_blend, uint32_t* accum_blend, uint32_t* zbuf_zmin, const ClearTarget& clear);

// ---- chunk-level frustum culling (rtr_cull.cu)
cudaError_t launch_chunk_bounds(cudaStream_t s, const PointRecord* pts, uint64_t n, ChunkBounds* bounds);
// clear (zbuf coverage, accum, minmax) + per-frame chunk classification in ONE launch; parity alternates per frame.
cudaError_t launch_clear_classify(cudaStream_t s, int sm_count, uint32_t* zbuf, uint64_t cov, uint32_t* accum,
                                  uint64_t n_px, uint32_t* minmax, const ChunkBounds* bounds, uint32_t n_chunks,
                                  const CullParams& cp, uint32_t* vis_list, CullState* cull, uint32_t parity, bool lean = true,
                                  const BandSort& bands = BandSort{1u, 0u, nullptr});
// Chunk classification for two cameras at once, no clear: entry = chunk | kTileBlend (visible for cp_blend, if
// have_blend) | kTileZmin (visible for cp_zmin, if have_zmin); chunks visible for neither are dropped.
cudaError_t launch_classify_pair(cudaStream_t s, int sm_count, const ChunkBounds* bounds, uint32_t n_chunks,
                                 const CullParams& cp_blend, bool have_blend, const CullParams& cp_zmin, bool have_zmin,
                                 uint32_t* vis_list, CullState* cull, uint32_t parity, bool late_wait = false,
                                 const BandSort& bands = BandSort{1u, 0u, nullptr});
cudaError_t launch_zmin(cudaStream_t s, int variant, int unroll, const PointRecord* pts, uint64_t n,
                        uint64_t index_base, const ProjParams& pp, uint32_t* zbuf, unsigned long long* zkey);
cudaError_t launch_blend(cudaStream_t s, int variant, int unroll, const PointRecord* pts, uint64_t n,
                         const ProjParams& pp, const uint32_t* zbuf, uint32_t* accum, const uint32_t* gate);
// The in-stream exact re-run after a float-accumulator overflow (clear + integer blend + resolve in one launch
// that returns at once when minmax[2] == 0).  cull/vis_list null = all tiles.
constexpr unsigned kFixupClusterCtas = 8;  // launch_exact_fixup(grid_ctas = kFixupClusterCtas): one thread-block cluster, no cooperative launch
// need_flag: 0 for one-camera lists; kTileBlend for two-camera lists (only the entries carrying the flag are redone).
cudaError_t launch_exact_fixup(cudaStream_t s, int sm_count, const PointRecord* pts, uint64_t n, const ProjParams& pp,
                               const CullState* cull, const uint32_t* vis_list, const uint32_t* zbuf, uint32_t* accum,
                               uint64_t n_px, uint8_t* image, uint64_t cov, uint32_t* minmax, uint32_t* host_note,
                               uint32_t need_flag = 0u, unsigned grid_ctas = 0u);  // grid_ctas 0: two CTAs per SM
// The exact re-run as three gated launches without a grid barrier (fused sequences: a cooperative grid on the image
// stream would have to wait for the point stream's persistent kernel to leave the SMs): clear -> launch_blend_list
// (integer sums, gate) -> launch_resolve_gated.  Each returns at once unless *gate != 0.
cudaError_t launch_clear_accum_gated(cudaStream_t s, int sm_count, uint32_t* accum, uint64_t n_px, const uint32_t* gate, uint32_t* host_note);
cudaError_t launch_project_dump(cudaStream_t s, const PointRecord* pts, uint64_t n, const ProjParams& pp,
                                int32_t* pix_out, uint32_t* zbits_out);

// ---- image-space passes (rtr_image_kernels.cu)
struct FrameBuffers {
    uint32_t* zbuf;      // P u32 depth bits; viewed as float = pyramid level 0; filtered in place
    uint32_t* accum;     // 4P u32
    uint8_t* image;      // 3P u8 BGR interleaved
    uint16_t* tensor;    // 5P fp16, planes of stride uw[0]*uh[0]
    float* level[5];     // level[0] aliases zbuf; level[1..4] persistent scratch
    uint8_t* mask[4];    // optional taps: mask[i-1] produced by up-pass iteration i (nullptr = not kept)
    uint32_t* minmax;    // {min, max} of the valid depth bits, [2] = float-accumulator overflow flag, [3] = fix-up barrier
    unsigned long long* zkey;  // P u64, only in key64 mode
};
// accum -> image over [0, cov) (resolvePass), fused with the 4-level min pyramid (reduce x4) and the
// depth min/max (find_*_minmax_kernel).  `pyramid` = false for plain computeRGBD.
// f32acc: accum holds floats (blend variant bit 2); raises minmax[2] when a pixel's count exceeds
// kF32ExactCount.  gated: run (resolve only) iff minmax[2] != 0 — the exact re-resolve.
cudaError_t launch_resolve_pyramid(cudaStream_t s, const FrameBuffers& fb, int W, int H, const PyramidDims& d,
                                   bool pyramid, bool resolve, bool force_generic, bool f32acc = false);
cudaError_t launch_resolve_gated(cudaStream_t s, const FrameBuffers& fb, int W, int H);
// 64-bit key mode: clear / split keys into depth bits + nearest-point colour.
cudaError_t launch_clear_key64(cudaStream_t s, int sm_count, unsigned long long* zkey, uint64_t cov);
cudaError_t launch_resolve_key64(cudaStream_t s, const unsigned long long* zkey, const PointRecord* pts,
                                 uint64_t index_base, uint64_t n_local, uint32_t* zbuf, uint8_t* image, uint64_t n_px,
                                 uint64_t cov);
// up-pass: laplacian + compare (+ resize | + removeMask).  fused: all four levels in ONE launch (up_fused_kernel) when
// W % 16 == 0 and no mask taps are requested; otherwise one launch per level.
cudaError_t launch_up_pass(cudaStream_t s, const FrameBuffers& fb, const PyramidDims& d, bool force_generic, bool fused);
int up_pass_launches(const FrameBuffers& fb, const PyramidDims& d, bool force_generic, bool fused);

// ---- point-sharded merge over peer memory (rtr_peer.cu)
constexpr int kMaxPeers = 16;
struct PeerMergeParams {
    uint4* buf[kMaxPeers];       // the buffer to all-reduce on every rank (peer-mapped; [rank] is the local one)
    uint32_t* flags[kMaxPeers];  // every rank's flag array: flags[p][s] = last epoch rank s signalled to rank p
    int rank, n_ranks;
    uint64_t n_vec;              // 16-byte elements in the buffer
    uint32_t epoch;              // this launch uses epoch, epoch + 1, epoch + 2
    uint32_t* local_bar;         // local grid-barrier counter (grows by 2 * gridDim.x per launch)
    uint32_t local_base;         // its value before this launch
    volatile uint32_t* err;      // mapped host word, raised when a wait timed out
    unsigned long long timeout_ns;
};
// op 0: min of u32, op 1: sum of u32, op 2: min of u64.  Grid = 2 CTAs per SM (cooperative launch, see kernel).
cudaError_t launch_peer_allreduce(cudaStream_t s, int sm_count, int op, const PeerMergeParams& pm);
// Plainly launched empty kernel: what is enqueued behind it cannot start before what is in front of it has completed,
// whatever launch attributes a library (NCCL) gives its kernels (see rtr_peer.cu).
cudaError_t launch_stream_fence(cudaStream_t s);

// ---- synthetic cloud on the device (rtr_synth.cu; bench/test support, same generator as the oracle)
cudaError_t launch_synth(cudaStream_t s, uint64_t seed, uint64_t n_total, uint64_t first, uint64_t count, int lx,
                         int ly, int lz, int nbox, PointRecord* out);

// ---- L2 atomic micro-benchmark (rtr_microbench.cu; measurement support)
cudaError_t launch_red_bench(cudaStream_t s, int sm_count, int mode, bool key64, const int32_t* pix, uint64_t n_ops,
                             uint32_t n_px, uint32_t* z32, unsigned long long* z64);

cudaError_t launch_fast_divide_selftest(cudaStream_t s, int sm_count, uint64_t n_pairs, uint64_t seed, unsigned long long* mismatches);

}  // namespace rtr
