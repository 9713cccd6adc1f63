// Host side of the B200-native renderer + its C ABI (include/rtr_b200.h).
//
// Replaces ProjectCloud (project_cloud.cu:189-493) for the projection / z-buffer / blend / prefilter
// path.  Differences in mechanism, not in results:
//   - one non-blocking stream per renderer, no cudaDeviceSynchronize between stages
//     (the reference issues 21 per filtered frame, project_cloud.cu:322-390);
//   - persistent per-resolution buffers incl. pyramid scratch (the reference cudaMalloc/cudaFree's
//     12 buffers per frame, project_cloud.cu:346-390), zero-initialised once at (re)allocation,
//     which is the parity definition for non-multiple-of-16 resolutions;
//   - camProj travels as a kernel parameter (the reference does a blocking 64-byte H2D per frame,
//     project_cloud.cu:320);
//   - two frame-buffer sets so a trajectory's D2H copies overlap the next frame's kernels.
// There is no CPU fallback anywhere in this file.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <unistd.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/rtr_b200.h"
#include "rtr_internal.h"
#include "rtr_kernels.h"

using namespace rtr;

namespace rtr {
bool pdl_enabled() {
    static const bool on = [] {
        const char* e = std::getenv("RTR_PDL");
        return !(e && e[0] == '0');
    }();
    return on;
}
}  // namespace rtr

namespace {

thread_local std::string g_create_error;

// ---- minimal NCCL binding, resolved at run time (libnccl.so.2 — torch's bundled copy when the
// process already loaded it, else the system one).  Only what the point-sharded path needs.
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclUint32_ = 3, ncclUint64_ = 5 };  // ncclDataType_t values (nccl.h)
enum { ncclSum_ = 0, ncclMin_ = 3 };        // ncclRedOp_t values
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool load(std::string& err) {
        if (lib) return true;
        lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) { err = std::string("dlopen libnccl.so.2 failed: ") + dlerror(); return false; }
        GetUniqueId = reinterpret_cast<decltype(GetUniqueId)>(dlsym(lib, "ncclGetUniqueId"));
        CommInitRank = reinterpret_cast<decltype(CommInitRank)>(dlsym(lib, "ncclCommInitRank"));
        CommDestroy = reinterpret_cast<decltype(CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
        AllReduce = reinterpret_cast<decltype(AllReduce)>(dlsym(lib, "ncclAllReduce"));
        GetErrorString = reinterpret_cast<decltype(GetErrorString)>(dlsym(lib, "ncclGetErrorString"));
        if (!GetUniqueId || !CommInitRank || !CommDestroy || !AllReduce) { err = "libnccl lacks required symbols"; return false; }
        return true;
    }
};
NcclApi g_nccl;
std::mutex g_nccl_mu;

}  // namespace


namespace {

int fail(rtr_renderer* r, int code, const std::string& msg) {
    if (r) r->err = msg; else g_create_error = msg;
    return code;
}
}  // namespace
namespace rtr {
int renderer_fail(rtr_renderer* r, int code, const std::string& msg) { return fail(r, code, msg); }
}
namespace {
int cuda_fail(rtr_renderer* r, cudaError_t e, const char* what) {
    return fail(r, RTR_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define RTR_CUDA(r, call)                                   \
    do {                                                    \
        cudaError_t e__ = (call);                           \
        if (e__ != cudaSuccess) return cuda_fail((r), e__, #call); \
    } while (0)

}  // namespace
namespace rtr {
// Frees the chunk bounds and both frame sets' visible-list storage (the cloud is about to change).
void free_cull_storage(rtr_renderer* r) {
    cudaFree(r->bounds);
    r->bounds = nullptr;
    for (auto& s : r->set) {
        cudaFree(s.vis_list); cudaFree(s.cull_state); cudaFree(s.band_scratch);
        s.vis_list = nullptr; s.cull_state = nullptr; s.band_scratch = nullptr; s.band_cap = 0; s.cull_parity = 0;
    }
}
// The compute streams idle (stream2 / image_stream only ever hold work of a pipelined frame sequence).
cudaError_t sync_compute(rtr_renderer* r) {
    cudaError_t e = cudaStreamSynchronize(r->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(r->stream2);
    if (e == cudaSuccess) e = cudaStreamSynchronize(r->stream3);
    if (e == cudaSuccess) e = cudaStreamSynchronize(r->image_stream);
    return e;
}
}  // namespace rtr
namespace {
// timing 3 = per-stage events of the FUSED sequence (point pass on one stream, image passes on the other)
bool pipelined(const rtr_renderer* r) { return r->pipeline && !r->peer.attached && !r->comm && (!r->timing || r->timing == 3); }
// the set a non-fused pipelined sequence (and the trajectory call's D2H overlap) moves on to: sets 0 and 1, or 0, 1 and 2
// with option pipeline_depth = 3
int sequence_depth(const rtr_renderer* r) { return (pipelined(r) && r->pipeline_depth == 3) ? 3 : 2; }
int next_set(const rtr_renderer* r, int cur) { return (cur + 1) % sequence_depth(r); }

void free_frame_sets(rtr_renderer* r) {
    for (auto& s : r->set) {
        cudaFree(s.arena); cudaFree(s.fb.tensor); cudaFree(s.fb.minmax);
        s.arena = nullptr;
        cudaFree(s.fb.zkey);
        for (int i = 1; i <= 4; ++i) cudaFree(s.fb.level[i]);
        for (int i = 0; i < 4; ++i) cudaFree(s.fb.mask[i]);
        s.fb = FrameBuffers{};
        s.clean = false;
    }
    r->alloc_W = r->alloc_H = 0;
    r->masks_allocated = r->key64_allocated = false;
}

template <typename T> cudaError_t zalloc(T** p, size_t bytes, cudaStream_t s) {
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(p), bytes ? bytes : 16);
    if (e != cudaSuccess) return e;
    return cudaMemsetAsync(*p, 0, bytes ? bytes : 16, s);
}

// (Re)allocate per-resolution buffers, zero-initialised — the reference reallocates six buffers
// when W x H changes (project_cloud.cu:275-298).
int ensure_buffers(rtr_renderer* r) {
    const int W = r->W, H = r->H;
    if (r->alloc_W != W || r->alloc_H != H) {
        const int frc = flush_pending(r);  // a frame of the old resolution may still lack its blend / image passes
        if (frc != RTR_OK) return frc;
        RTR_CUDA(r, sync_compute(r));
        RTR_CUDA(r, cudaStreamSynchronize(r->copy_stream));
        free_frame_sets(r);
        r->dims = make_pyramid_dims(W, H);
        const size_t P = size_t(W) * H;
        auto up256 = [](size_t b) { return (b + 255) & ~size_t(255); };
        for (auto& s : r->set) {
            s.zbuf_off = up256(P * 16);
            s.image_off = s.zbuf_off + up256(P * 4);
            s.arena_bytes = s.image_off + up256(P * 3 + 16);
            RTR_CUDA(r, zalloc(&s.arena, s.arena_bytes, r->stream));
            s.fb.accum = reinterpret_cast<uint32_t*>(s.arena);
            s.fb.zbuf = reinterpret_cast<uint32_t*>(s.arena + s.zbuf_off);
            s.fb.image = s.arena + s.image_off;
            RTR_CUDA(r, zalloc(&s.fb.tensor, P * 5 * 2, r->stream));
            RTR_CUDA(r, zalloc(&s.fb.minmax, 16, r->stream));
            s.fb.level[0] = reinterpret_cast<float*>(s.fb.zbuf);
            for (int i = 1; i <= 4; ++i) RTR_CUDA(r, zalloc(&s.fb.level[i], size_t(r->dims.w[i]) * r->dims.h[i] * 4, r->stream));
        }
        r->alloc_W = W; r->alloc_H = H;
        RTR_CUDA(r, sync_compute(r));  // the zero-fills ran on `stream`; frames may start on either stream
    }
    if (r->keep_masks && !r->masks_allocated) {
        for (auto& s : r->set)
            for (int i = 0; i < 4; ++i) RTR_CUDA(r, zalloc(&s.fb.mask[i], size_t(r->dims.uw[i]) * r->dims.uh[i], r->stream));
        r->masks_allocated = true;
        RTR_CUDA(r, sync_compute(r));
    }
    if (r->key64 && !r->key64_allocated) {
        for (auto& s : r->set) RTR_CUDA(r, zalloc(&s.fb.zkey, size_t(W) * H * 8, r->stream));
        r->key64_allocated = true;
        RTR_CUDA(r, sync_compute(r));
    }
    return RTR_OK;
}

// camProj = K4 * E in float, the way the reference's glm expression evaluates it
// (project_cloud.cu:318, project_cloud.h:50-59, CameraCalibration.cpp:17-27): every factor is
// cast to float first, products are summed left to right, no FMA.
void build_cam_proj(const double* K9, const double* E16, float* out16) {
    float K[4][4] = {{0.f}}, E[4][4];
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) K[a][b] = static_cast<float>(K9[a * 3 + b]);
    K[3][3] = 1.0f;
    for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b) E[a][b] = static_cast<float>(E16[a * 4 + b]);
    for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b) {
            volatile float t = K[a][0] * E[0][b];
            volatile float p1 = K[a][1] * E[1][b];
            t = t + p1;
            volatile float p2 = K[a][2] * E[2][b];
            t = t + p2;
            volatile float p3 = K[a][3] * E[3][b];
            t = t + p3;
            out16[a * 4 + b] = t;
        }
}

int make_params(rtr_renderer* r, ProjParams& pp) {
    if (r->W < 16 || r->H < 16) return fail(r, RTR_ERR_STATE, "intrinsics not set (width/height must be >= 16)");
    std::memset(&pp, 0, sizeof(pp));
    pp.W = r->W; pp.H = r->H;
    if (r->raw_proj) {
        std::memcpy(pp.m, r->cam_proj, sizeof(float) * 12);
        return RTR_OK;
    }
    if (!r->have_K || !r->have_E) return fail(r, RTR_ERR_STATE, "camera not set: call rtr_set_intrinsics and rtr_set_pose_w2c");
    build_cam_proj(r->K, r->E, r->cam_proj);
    std::memcpy(pp.m, r->cam_proj, sizeof(float) * 12);
    bool any = false;
    for (double d : r->dist) any = any || (d != 0.0);
    if (any) {  // new feature (the reference never applies distortion): dist == 0 keeps the exact path
        pp.distort = 1;
        for (int i = 0; i < 12; ++i) pp.e[i] = static_cast<float>(r->E[i]);
        pp.fx = float(r->K[0]); pp.skew = float(r->K[1]); pp.cx = float(r->K[2]);
        pp.fy = float(r->K[4]); pp.cy = float(r->K[5]);
        pp.k1 = float(r->dist[0]); pp.k2 = float(r->dist[1]); pp.p1 = float(r->dist[2]);
        pp.p2 = float(r->dist[3]); pp.k3 = float(r->dist[4]);
        if (!r->dist_bounds_valid) {  // (sampling the polynomial takes tens of microseconds: as long as a small frame)
            rtr_host_distortion_bounds(r->W, r->H, r->K, r->dist, &r->dist_r2_max, &r->dist_rstar);
            r->dist_bounds_valid = true;
        }
        pp.r2_max = float(r->dist_r2_max);
        r->cull_rstar = r->dist_rstar;
    }
    return RTR_OK;
}

int comm_allreduce(rtr_renderer* r, const void* src, void* dst, size_t count, int dtype, int op) {
    // our kernels trigger their programmatic dependents early; NCCL's kernels never execute griddepcontrol.wait (rtr_peer.cu)
    RTR_CUDA(r, launch_stream_fence(r->stream));
    r->launches += 1;
    const int rc = g_nccl.AllReduce(src, dst, count, dtype, op, r->comm, r->stream);
    if (rc != 0) return fail(r, RTR_ERR_COMM, std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error"));
    return RTR_OK;
}

constexpr int kEvPoolFrames = 256;
constexpr double kBandsAutoMinMB = 126.0;  // option bands = 0: frames whose z-buffer + colour sums exceed the L2 (126 MB) are walked band by band

// Fold the pooled per-frame events into ev_sum (blocks until the last recorded frame finished).
int drain_event_pool(rtr_renderer* r) {
    if (r->ev_frames == 0) return RTR_OK;
    RTR_CUDA(r, cudaEventSynchronize(r->ev_pool[size_t(r->ev_frames - 1) * 6 + 5]));
    for (int f = 0; f < r->ev_frames; ++f) {
        cudaEvent_t* e = &r->ev_pool[size_t(f) * 6];
        float ms = 0.f;
        for (int i = 0; i < 5; ++i) {
            RTR_CUDA(r, cudaEventElapsedTime(&ms, e[i], e[i + 1]));
            r->ev_sum[i] += ms;
        }
        RTR_CUDA(r, cudaEventElapsedTime(&ms, e[0], e[5]));
        r->ev_sum[5] += ms;
    }
    r->ev_count += uint64_t(r->ev_frames);
    r->ev_frames = 0;
    return RTR_OK;
}

// Two-shot all-reduce of one frame buffer over the peers' memory (rtr_peer.cu).
// what: 0 z-buffer, 1 colour sums, 2 64-bit keys, 3 image bytes;  op: 0 min u32, 1 sum u32, 2 min u64.
int peer_merge(rtr_renderer* r, int si, int what, int op) {
    rtr_renderer::Peer& pe = r->peer;
    if (si > 1) return fail(r, RTR_ERR_STATE, "point-sharded frames use frame sets 0 and 1");
    const uint64_t P = uint64_t(r->W) * r->H;
    PeerMergeParams pm;
    std::memset(&pm, 0, sizeof(pm));
    for (int p = 0; p < pe.n; ++p) {
        pm.buf[p] = static_cast<uint4*>(pe.peer_buf[what][si][p]);
        pm.flags[p] = pe.peer_flags[p];
        if (!pm.buf[p]) return fail(r, RTR_ERR_STATE, what == 2 ? "key64 frames need option key64 = 1 on every rank BEFORE rtr_peer_export" : "peer buffer not mapped");
    }
    pm.rank = pe.rank;
    pm.n_ranks = pe.n;
    pm.n_vec = what == 0 ? P / 4 : (what == 1 ? P : (what == 2 ? P / 2 : (P * 3 + 15) / 16));
    pm.epoch = pe.epoch;
    pm.local_bar = pe.flags + 64;
    pm.local_base = pe.local_base;
    pm.err = pe.err_dev;
    pm.timeout_ns = uint64_t(pe.timeout_ms < 1 ? 1 : pe.timeout_ms) * 1000000ull;
    pe.epoch += 3;
    pe.local_base += 2u * unsigned(r->sm_count) * 2u;
    RTR_CUDA(r, launch_peer_allreduce(r->stream, r->sm_count, op, pm));
    r->launches += 1;
    return RTR_OK;
}

// After a synchronisation: did a cross-GPU wait of the peer merge give up?  The frame is then invalid.
int peer_status(rtr_renderer* r) {
    if (r->peer.err_host && *reinterpret_cast<volatile uint32_t*>(r->peer.err_host)) {
        *reinterpret_cast<volatile uint32_t*>(r->peer.err_host) = 0u;
        return fail(r, RTR_ERR_COMM, "point-sharded merge: a peer did not reach the barrier within peer_timeout_ms; the frame is invalid "
                                     "(is every rank rendering the same frames? detach / export / attach again before the next frame)");
    }
    return RTR_OK;
}

// Screen bands the visible list is ordered by (BandSort, rtr_kernels.h).  A point pass gathers from a z-buffer and
// reduces into a z-buffer / the colour sums: 24 B per pixel that should stay in L2 while the pass runs.  1920x1080:
// 50 MB, fits, list order is kept.  3840x2160: 199 MB — every RED and gather would go to DRAM; ordered by 8 bands the
// tiles in flight share about 25 MB of it.
uint32_t frame_bands(const rtr_renderer* r) {
    if (r->bands >= 1) return uint32_t(r->bands > kMaxBands ? kMaxBands : r->bands);
    const double live_mb = double(r->W) * double(r->H) * 24.0 / 1048576.0;
    if (live_mb <= kBandsAutoMinMB) return 1u;
    const double want = std::ceil(live_mb / 25.0);
    return uint32_t(want > double(kMaxBands) ? double(kMaxBands) : (want < 2.0 ? 2.0 : want));
}

// Everything a frame needs besides its buffers, from the renderer's current camera and options.
int plan_frame(rtr_renderer* r, FramePlan& pl) {
    int rc = make_params(r, pl.pp);
    if (rc != RTR_OK) return rc;
    const ProjParams& pp = pl.pp;
    // chunk-level frustum culling: exact (conservative) for the pinhole path.  Under distortion (ring kernels only)
    // the chunk test runs against the square [-r*, r*]^2 of normalised coordinates that contains every point able to
    // reach the image (make_params), written as a pinhole camera of 2001 x 2001 "pixels" of r*/1000 each.
    pl.cull = r->chunk_cull && r->bounds && (!pp.distort || (r->ring && r->cull_rstar > 0));
    // ring = 1: the TMA-fed kernels for culled frames, the per-thread LDG.128 kernels when every chunk is streamed
    // (measured 3 % faster there, profiles/r01h_exp_ring_c3.json); ring = 2: always; ring = 0: never
    pl.use_ring = r->ring == 2 || (r->ring == 1 && pl.cull);
    pl.bands = frame_bands(r);
    CullParams& cp = pl.cp;
    std::memset(&cp, 0, sizeof(cp));
    if (pl.cull && !pp.distort) {
        for (int k = 0; k < 4; ++k) { cp.r0[k] = pp.m[k]; cp.r1[k] = pp.m[4 + k]; cp.r2[k] = pp.m[8 + k]; }
        cp.W = r->W; cp.H = r->H;
    } else if (pl.cull) {
        const double a = 1000.0 / r->cull_rstar;
        for (int k = 0; k < 4; ++k) {
            cp.r0[k] = a * double(pp.e[k]) + 1000.0 * double(pp.e[8 + k]);
            cp.r1[k] = a * double(pp.e[4 + k]) + 1000.0 * double(pp.e[8 + k]);
            cp.r2[k] = double(pp.e[8 + k]);
        }
        cp.W = 2001.0; cp.H = 2001.0;
    }
    return RTR_OK;
}

// The ring kernels' schedule for a pass over `fs`'s visible list (or every chunk when cull is false).
RingSchedule ring_schedule_for(const rtr_renderer* r, const FrameSet& fs, bool cull) {
    RingSchedule sched = r->ring_sched;  // chunk permutation of the stream-all order, fixed at upload
    if (!r->ring_perm) sched.perm_mul = 1;  // measurement: stream-all tiles in storage order
    sched.early = r->ring_early ? 1u : 0u;
    sched.ctas_per_sm = uint32_t(r->ring_ctas);
    sched.claim_min_tiles_per_cta = uint32_t(r->ring_claim_min < 0 ? 0 : r->ring_claim_min);
    sched.n_queues = uint32_t(r->ring_dynamic < 1 ? 1 : (r->ring_dynamic > kMaxTileQueues ? kMaxTileQueues : r->ring_dynamic));
    sched.cull = cull ? fs.cull_state : nullptr;
    sched.vis_list = cull ? fs.vis_list : nullptr;
    sched.grid_override = 0;
    sched.tiles_hint = nullptr;
    return sched;
}

// Float colour sums (one 16-byte RED per point) unless the sums must be all-reduced as integers or an earlier frame
// of this view overflowed them (the exact re-run left a note in a mapped host word).
int blend_variant_now(rtr_renderer* r, bool merged_across_ranks) {
    if (r->overflow_note && *reinterpret_cast<volatile uint32_t*>(r->overflow_note)) {
        *reinterpret_cast<volatile uint32_t*>(r->overflow_note) = 0u;
        r->int_sum_frames = 64;
    }
    const bool int_sums = merged_across_ranks || r->int_sum_frames > 0;
    if (r->int_sum_frames > 0) r->int_sum_frames -= 1;
    return int_sums ? (r->blend_variant & ~4) : r->blend_variant;
}

cudaEvent_t* timing_events(rtr_renderer* r, int* rc) {
    *rc = RTR_OK;
    if (r->timing < 2) return r->ev;
    if (r->ev_pool.empty()) {
        r->ev_pool.resize(size_t(kEvPoolFrames) * 6);
        for (auto& e : r->ev_pool)
            if (cudaEventCreate(&e) != cudaSuccess) { *rc = fail(r, RTR_ERR_CUDA, "cudaEventCreate"); return r->ev; }
    }
    if (r->ev_frames == kEvPoolFrames && (*rc = drain_event_pool(r)) != RTR_OK) return r->ev;
    return &r->ev_pool[size_t(r->ev_frames) * 6];
}

// Enqueue one whole frame into frame set `si` (two point passes): on `stream`, or — second set of a pipelined
// non-fused sequence — on `stream2`.
int enqueue_frame(rtr_renderer* r, int stage, int si, bool allow_pipeline = false) {
    const bool peer = r->peer.attached;
    if (peer && (r->peer.W != r->W || r->peer.H != r->H)) return fail(r, RTR_ERR_STATE, "resolution changed while peers are attached: rtr_peer_detach first");
    if (!r->points || r->n_points == 0) return fail(r, RTR_ERR_STATE, "no cloud uploaded");
    int rc = flush_pending(r);
    if (rc != RTR_OK) return rc;
    FramePlan pl;
    if ((rc = plan_frame(r, pl)) != RTR_OK) return rc;
    const ProjParams& pp = pl.pp;
    rc = ensure_buffers(r);
    if (rc != RTR_OK) return rc;
    FrameSet& fs = r->set[si];
    FrameBuffers fb = fs.fb;
    if (!r->keep_masks) for (auto& m : fb.mask) m = nullptr;
    const uint64_t P = uint64_t(r->W) * r->H, cov = clear_coverage(r->W, r->H);
    const bool filtered = stage == RTR_STAGE_FILTERED;
    cudaStream_t s = (allow_pipeline && si >= 1 && si <= 2 && pipelined(r)) ? (si == 1 ? r->stream2 : r->stream3) : r->stream;
    RTR_CUDA(r, cudaStreamWaitEvent(s, fs.rendered, 0));               // this set's previous frame (may have run on the other stream)
    if (fs.copied) RTR_CUDA(r, cudaStreamWaitEvent(s, fs.copied, 0));  // previous D2H of this set must be done
    fs.clean = false;
    cudaEvent_t* ev = timing_events(r, &rc);
    if (rc != RTR_OK) return rc;
    const bool cull = pl.cull, use_ring = pl.use_ring;
    const CullParams& cp = pl.cp;
    RingSchedule sched = ring_schedule_for(r, fs, cull);

    if (r->timing) cudaEventRecord(ev[0], s);
    if (r->key64) {
        fs.f32acc = false;
        if (cull) {
            fs.cull_parity ^= 1u;
            RTR_CUDA(r, launch_clear_classify(s, r->sm_count, fb.zbuf, 0, nullptr, 0, fb.minmax, r->bounds, r->n_chunks, cp, fs.vis_list, fs.cull_state, fs.cull_parity, r->clear_lean != 0, BandSort{pl.bands, fs.band_cap, fs.band_scratch}));
        } else {
            RTR_CUDA(r, launch_clear(s, r->sm_count, fb.zbuf, 0, nullptr, 0, fb.minmax, nullptr));
        }
        RTR_CUDA(r, launch_clear_key64(s, r->sm_count, fb.zkey, cov));
        r->launches += 2;
        if (r->timing) cudaEventRecord(ev[1], s);
        sched.tile_counter = (cull && r->ring_dynamic > 0) ? tile_counters(fs.cull_state, 0) : nullptr;
        if (use_ring) RTR_CUDA(r, launch_zmin_ring(s, r->sm_count, r->zmin_variant & 5, r->points, r->n_points, r->index_base, pp, sched, cull, fb.zbuf, fb.zkey));
        else if (cull) RTR_CUDA(r, launch_zmin_list(s, r->sm_count, r->zmin_variant & 5, r->points, r->n_points, r->index_base, pp, fs.cull_state, fs.vis_list, fb.zbuf, fb.zkey));
        else RTR_CUDA(r, launch_zmin(s, r->zmin_variant & 5, r->zmin_unroll, r->points, r->n_points, r->index_base, pp, fb.zbuf, fb.zkey));
        r->launches += 1;
        if (peer) {  // north_star's merge: min over the ranks of the 64-bit (depth bits << 32 | point index) keys
            if ((rc = peer_merge(r, si, 2, 2)) != RTR_OK) return rc;
        } else if (r->comm) {
            rc = comm_allreduce(r, fb.zkey, fb.zkey, P, ncclUint64_, ncclMin_);
            if (rc != RTR_OK) return rc;
        }
        if (r->timing) { cudaEventRecord(ev[2], s); cudaEventRecord(ev[3], s); }
        RTR_CUDA(r, launch_resolve_key64(s, fb.zkey, r->points, r->index_base, r->n_points, fb.zbuf, fb.image, P, cov));
        r->launches += 1;
        // colour of a winning point lives on exactly one rank; the others wrote 0 -> the sum over the ranks is that colour.
        // The image bytes are summed as u32 words; no carry can occur because at most one rank is non-zero per byte.
        if (peer) {
            if ((rc = peer_merge(r, si, 3, 1)) != RTR_OK) return rc;
        } else if (r->comm) {
            rc = comm_allreduce(r, fb.image, fb.image, (P * 3 + 3) / 4, ncclUint32_, ncclSum_);
            if (rc != RTR_OK) return rc;
        }
        RTR_CUDA(r, launch_resolve_pyramid(s, fb, r->W, r->H, r->dims, filtered, false, r->force_generic != 0));
        r->launches += filtered ? (((r->W % 16) == 0 && !r->force_generic) ? 1 : 5) : 0;
    } else {
        if (cull) {
            fs.cull_parity ^= 1u;
            RTR_CUDA(r, launch_clear_classify(s, r->sm_count, fb.zbuf, cov, fb.accum, P, fb.minmax, r->bounds, r->n_chunks, cp, fs.vis_list, fs.cull_state, fs.cull_parity, r->clear_lean != 0, BandSort{pl.bands, fs.band_cap, fs.band_scratch}));
        } else {
            RTR_CUDA(r, launch_clear(s, r->sm_count, fb.zbuf, cov, fb.accum, P, fb.minmax, nullptr));
        }
        r->launches += 1;
        if (r->timing) cudaEventRecord(ev[1], s);
        sched.tile_counter = (cull && r->ring_dynamic > 0) ? tile_counters(fs.cull_state, 0) : nullptr;
        if (use_ring) RTR_CUDA(r, launch_zmin_ring(s, r->sm_count, r->zmin_variant, r->points, r->n_points, r->index_base, pp, sched, cull, fb.zbuf, nullptr));
        else if (cull) RTR_CUDA(r, launch_zmin_list(s, r->sm_count, r->zmin_variant, r->points, r->n_points, r->index_base, pp, fs.cull_state, fs.vis_list, fb.zbuf, nullptr));
        else RTR_CUDA(r, launch_zmin(s, r->zmin_variant, r->zmin_unroll, r->points, r->n_points, r->index_base, pp, fb.zbuf, nullptr));
        r->launches += 1;
        if (peer) {
            if ((rc = peer_merge(r, si, 0, 0)) != RTR_OK) return rc;
        } else if (r->comm) {
            rc = comm_allreduce(r, fb.zbuf, fb.zbuf, P, ncclUint32_, ncclMin_);
            if (rc != RTR_OK) return rc;
        }
        if (r->timing) cudaEventRecord(ev[2], s);
        const int bv = blend_variant_now(r, r->comm || peer);
        const bool f32acc = (bv & 4) != 0;
        fs.f32acc = f32acc;
        sched.tile_counter = (cull && r->ring_dynamic > 0) ? tile_counters(fs.cull_state, 1) : nullptr;
        if (use_ring) RTR_CUDA(r, launch_blend_ring(s, r->sm_count, bv, r->points, r->n_points, pp, sched, cull, fb.zbuf, fb.accum));
        else if (cull) RTR_CUDA(r, launch_blend_list(s, r->sm_count, bv, r->points, r->n_points, pp, fs.cull_state, fs.vis_list, fb.zbuf, fb.accum, nullptr));
        else RTR_CUDA(r, launch_blend(s, bv, r->blend_unroll, r->points, r->n_points, pp, fb.zbuf, fb.accum, nullptr));
        r->launches += 1;
        if (peer) {
            if ((rc = peer_merge(r, si, 1, 1)) != RTR_OK) return rc;
        } else if (r->comm) {
            rc = comm_allreduce(r, fb.accum, fb.accum, P * 4, ncclUint32_, ncclSum_);
            if (rc != RTR_OK) return rc;
        }
        if (r->timing) cudaEventRecord(ev[3], s);
        RTR_CUDA(r, launch_resolve_pyramid(s, fb, r->W, r->H, r->dims, filtered, true, r->force_generic != 0, f32acc));
        r->launches += ((r->W % 16) == 0 && !r->force_generic) ? 1 : (filtered ? 6 : 1);
        if (f32acc) {
            // A pixel that collected more than 65793 points leaves the exact range of the float sums; resolve
            // raised minmax[2] for it.  This launch returns at once unless that happened, in which
            // case it redoes the colour sums of the frame with the integer REDs and resolve again.
            RTR_CUDA(r, launch_exact_fixup(s, r->sm_count, r->points, r->n_points, pp, cull ? fs.cull_state : nullptr,
                                           cull ? fs.vis_list : nullptr, fb.zbuf, fb.accum, P, fb.image, cov, fb.minmax, r->overflow_note_dev));
            r->launches += 1;
        }
    }
    if (r->timing) cudaEventRecord(ev[4], s);
    if (filtered) {
        RTR_CUDA(r, launch_up_pass(s, fb, r->dims, r->force_generic != 0, r->fused_up != 0));
        r->launches += uint64_t(up_pass_launches(fb, r->dims, r->force_generic != 0, r->fused_up != 0));
    }
    if (r->timing) cudaEventRecord(ev[5], s);
    if (r->timing >= 2) r->ev_frames += 1;
    RTR_CUDA(r, cudaEventRecord(fs.rendered, s));
    return RTR_OK;
}

int enqueue_copy(rtr_renderer* r, int si, uint8_t* bgr, float* depth);

// ---- fused frame sequences: one stream of chunks per frame
// Consecutive frames of an asynchronous sequence (rtr_render_device back to back, rtr_render_trajectory) see nearly the
// same chunks.  Frame k is therefore enqueued as
//     point stream:  classify_pair(k-1, k)  ->  fused pass: blend(k-1) + z-min(k) over the union of the two lists,
//                    then every CTA clears its slice of the frame set frame k+1 will use
//     image stream:  [wait fused pass]  resolve(k-1) -> exact fix-up gate -> up-pass(k-1)  [-> D2H(k-1) on the copy stream]
// so every chunk is read from HBM once per frame instead of twice, the point stream holds nothing but the passes and
// their (overlapped) classification, and frame k stays "pending" (z-min done, blend outstanding) until frame k+1
// arrives or flush_pending() runs its blend alone.  Four frame sets: k (z-min), k-1 (blend, then image passes),
// k-2 (image passes / D2H), k+1 (being cleared).  Frames are byte-identical to the two-pass path (tests/test_gpu_fused.py).
// Option fuse: 0 never, 2 always, 1 (default) when the cloud is large enough for the shared stream to pay: the fused
// sequence trades one of the two chunk streams per frame for fewer, longer kernel boundaries in which the image passes
// can run, which wins once a point pass is long (C3, 100 M points: 9 100 vs 8 020 frames/s) and loses on small clouds
// whose frames are a few short kernels (C1, 1 M points: 46 200 vs 52 800; C2, 20 M points: 21 900 vs 23 400;
// profiles/r02m_exp_fixup_gate.json).  The switch is the cloud's chunk count, known on the host without a read-back.
// (Distorted cameras used to switch it too — 17 800 vs 12 400 frames/s at C2's size, profiles/r02k_bench_c2_distort.json —
// but that two-pass number was the HOST's: make_params sampled the distortion polynomial 12 K times per frame.  With the
// bounds cached per set of intrinsics the two-pass sequence does 20 700 frames/s there against 16 800 fused,
// profiles/r02F_exp_host_enqueue.json.)
constexpr uint32_t kFuseAutoMinChunks = 40000;  // 41 M points
bool fused_sequence(const rtr_renderer* r, const FramePlan& pl) {
    const bool want = r->fuse == 2 || (r->fuse == 1 && r->n_chunks >= kFuseAutoMinChunks);
    return want && pipelined(r) && pl.cull && pl.use_ring && !r->key64 && r->ring >= 1;
}

#ifdef RTR_EXPERIMENTS
// Measurement only (wrong frames): RTR_EXP_NO_IMAGES=1 skips the image passes of fused sequences, RTR_EXP_NO_EVENTS=1 the
// event records / waits between the point stream's kernels — what the point stream costs on its own.
static bool exp_flag(const char* name) { const char* e = std::getenv(name); return e && e[0] == '1'; }
#endif

// resolve -> fix-up gate -> up-pass (+ D2H) of the pending frame, on the image stream, once the point pass that
// blended it (recorded in `pass_set`'s points_done; the pass's list lives in pass_set too) has finished.
int finish_images(rtr_renderer* r, PendingFrame& pf, int pass_set, bool f32acc, cudaEvent_t* ev) {
    FrameSet& fs = r->set[pf.si];
    FrameSet& ps = r->set[pass_set];
    FrameBuffers fb = fs.fb;
    if (!r->keep_masks) for (auto& m : fb.mask) m = nullptr;
    const int W = r->alloc_W, H = r->alloc_H;
    const uint64_t P = uint64_t(W) * H, cov = clear_coverage(W, H);
    const bool filtered = pf.stage == RTR_STAGE_FILTERED;
    cudaStream_t s = r->image_stream;
#ifdef RTR_EXPERIMENTS
    if (exp_flag("RTR_EXP_NO_IMAGES")) { pf.active = false; return RTR_OK; }
#endif
    RTR_CUDA(r, cudaStreamWaitEvent(s, ps.points_done, 0));
    if (ev) cudaEventRecord(ev[3], s);
    RTR_CUDA(r, launch_resolve_pyramid(s, fb, W, H, r->dims, filtered, true, r->force_generic != 0, f32acc));
    r->launches += ((W % 16) == 0 && !r->force_generic) ? 1 : (filtered ? 6 : 1);
    if (f32acc) {
        // Exact re-run of the colour sums if a pixel left the float sums' exact range (resolve raised minmax[2]): ONE launch
        // that returns at once otherwise.  Its two barriers need every CTA resident at the same time, and the SMs belong to the
        // point stream's persistent pass, so a cooperative grid would wait for a drain window (measured: -4 % frames/s on C3);
        // the grid is therefore ONE thread-block cluster of 8 CTAs, which the hardware co-schedules wherever 8 SMs have a free
        // slot.  fixup_launches = 3 (option): the same work as three gated launches without a barrier; 2: small cooperative grid
        // (A/B in profiles/r02l_exp_fixup_gate.json).
        if (r->fixup_launches == 3) {
            RTR_CUDA(r, launch_clear_accum_gated(s, r->sm_count, fb.accum, P, fb.minmax + 2, r->overflow_note_dev));
            RTR_CUDA(r, launch_blend_list(s, r->sm_count, 0, r->points, r->n_points, pf.plan.pp, ps.cull_state, ps.vis_list, fb.zbuf, fb.accum,
                                          fb.minmax + 2, kTileBlend));
            RTR_CUDA(r, launch_resolve_gated(s, fb, W, H));
            r->launches += 3;
        } else {
            RTR_CUDA(r, launch_exact_fixup(s, r->sm_count, r->points, r->n_points, pf.plan.pp, ps.cull_state, ps.vis_list, fb.zbuf, fb.accum, P,
                                           fb.image, cov, fb.minmax, r->overflow_note_dev, kTileBlend,
                                           r->fixup_launches == 2 ? unsigned(r->sm_count) / 2u : kFixupClusterCtas));
            r->launches += 1;
        }
        (void)cov;
    }
    if (ev) cudaEventRecord(ev[4], s);
    if (filtered) {
        RTR_CUDA(r, launch_up_pass(s, fb, r->dims, r->force_generic != 0, r->fused_up != 0));
        r->launches += uint64_t(up_pass_launches(fb, r->dims, r->force_generic != 0, r->fused_up != 0));
    }
    if (ev) cudaEventRecord(ev[5], s);
    RTR_CUDA(r, cudaEventRecord(fs.rendered, s));
    if (pf.bgr || pf.depth) {
        const int rc = enqueue_copy(r, pf.si, pf.bgr, pf.depth);
        if (rc != RTR_OK) return rc;
    }
    pf.active = false;
    return RTR_OK;
}

int enqueue_fused(rtr_renderer* r, int stage, const FramePlan& pl, uint8_t* bgr, float* depth) {
    int rc = ensure_buffers(r);  // (flushes a pending frame of another resolution first)
    if (rc != RTR_OK) return rc;
    PendingFrame& pf = r->pending;
    if (pf.active && pf.plan.pp.distort != pl.pp.distort && (rc = flush_pending(r)) != RTR_OK) return rc;  // one kernel, one projection model
    const int si = (pf.active ? pf.si + 1 : r->cur + 1) % kFrameSets;
    FrameSet& fs = r->set[si];
    FrameSet& nx = r->set[(si + 1) % kFrameSets];  // the set the NEXT frame will use: this pass leaves it cleared
    const uint64_t P = uint64_t(r->W) * r->H, cov = clear_coverage(r->W, r->H);
    cudaStream_t s = r->stream;
    cudaEvent_t* ev = nullptr;
    if (r->timing == 3 && pf.active) {  // only passes that carry both halves are timed
        ev = timing_events(r, &rc);
        if (rc != RTR_OK) return rc;
    }
    if (!fs.clean) {  // first frame of a sequence: nobody has cleared this set ahead of time
        RTR_CUDA(r, cudaStreamWaitEvent(s, fs.rendered, 0));
        if (fs.copied) RTR_CUDA(r, cudaStreamWaitEvent(s, fs.copied, 0));
        RTR_CUDA(r, launch_clear(s, r->sm_count, fs.fb.zbuf, cov, fs.fb.accum, P, fs.fb.minmax, nullptr));
        r->launches += 1;
    }
    // the set this pass clears must be free: its last frame's image passes (which were also the last readers of that
    // set's visible list) and D2H copy done — three frames back, long finished unless the copies are the bottleneck
    ClearTarget clr{nullptr, 0, nullptr, 0, nullptr};
    bool no_events = false;
#ifdef RTR_EXPERIMENTS
    no_events = exp_flag("RTR_EXP_NO_EVENTS");
#endif
    if (!nx.clean) {
        if (!no_events) {
            RTR_CUDA(r, cudaStreamWaitEvent(s, nx.rendered, 0));
            if (nx.copied) RTR_CUDA(r, cudaStreamWaitEvent(s, nx.copied, 0));
        }
        clr = ClearTarget{nx.fb.zbuf, cov, reinterpret_cast<uint4*>(nx.fb.accum), P, nx.fb.minmax};
    }
    // point stream: classification for the two cameras (it starts while the previous pass drains and waits for it only
    // at its end), then the pass
    if (ev) cudaEventRecord(ev[0], s);
    fs.cull_parity ^= 1u;
    RTR_CUDA(r, launch_classify_pair(s, r->sm_count, r->bounds, r->n_chunks, pf.active ? pf.plan.cp : pl.cp, pf.active, pl.cp, true,
                                     fs.vis_list, fs.cull_state, fs.cull_parity, true, BandSort{pl.bands, fs.band_cap, fs.band_scratch}));
    if (ev) cudaEventRecord(ev[1], s);
    RingSchedule sched = ring_schedule_for(r, fs, true);
    sched.tile_counter = r->ring_dynamic > 0 ? tile_counters(fs.cull_state, 0) : nullptr;
    int bv = r->blend_variant;
    if (pf.active) {
        bv = blend_variant_now(r, false);
        r->set[pf.si].f32acc = (bv & 4) != 0;
    }
    if (r->fused_tiles_per_cta > 0 && r->tiles_hint) {
        sched.tiles_hint = r->tiles_hint_dev;
        const uint32_t last = *reinterpret_cast<volatile uint32_t*>(r->tiles_hint);  // a few passes old: a hint, any grid is correct
        sched.grid_override = last / uint32_t(r->fused_tiles_per_cta);
    }
    const FrameSet& prev = r->set[pf.active ? pf.si : si];
    RTR_CUDA(r, launch_fused_ring(s, r->sm_count, r->zmin_variant, bv, r->points, r->n_points, pf.active ? pf.plan.pp : pl.pp, pl.pp, sched,
                                  prev.fb.zbuf, prev.fb.accum, fs.fb.zbuf, clr));
    r->launches += 2;
    fs.clean = false;
    nx.clean = true;
    if (ev) cudaEventRecord(ev[2], s);
    if (!no_events) RTR_CUDA(r, cudaEventRecord(fs.points_done, s));
    if (pf.active) {
        if ((rc = finish_images(r, pf, si, (bv & 4) != 0, ev)) != RTR_OK) return rc;
        if (ev) r->ev_frames += 1;
    }
    pf.active = true;
    pf.si = si;
    pf.stage = stage;
    pf.plan = pl;
    pf.bgr = bgr;
    pf.depth = depth;
    r->cur = si;
    return RTR_OK;
}

}  // namespace

namespace rtr {
int flush_pending(rtr_renderer* r) {
    PendingFrame& pf = r->pending;
    if (!pf.active) return RTR_OK;
    FrameSet& fs = r->set[pf.si];
    cudaStream_t s = r->stream;
    // the frame's blend alone: a list of its own chunks (all flagged kTileBlend) through the same kernel.  The list is
    // rebuilt in place: the previous frame's gated fix-up (image stream) may still have to read the old one.
    RTR_CUDA(r, cudaStreamWaitEvent(s, r->set[(pf.si + kFrameSets - 1) % kFrameSets].rendered, 0));
    fs.cull_parity ^= 1u;
    RTR_CUDA(r, launch_classify_pair(s, r->sm_count, r->bounds, r->n_chunks, pf.plan.cp, true, pf.plan.cp, false, fs.vis_list, fs.cull_state,
                                     fs.cull_parity, false, BandSort{pf.plan.bands, fs.band_cap, fs.band_scratch}));
    RingSchedule sched = ring_schedule_for(r, fs, true);
    sched.tile_counter = r->ring_dynamic > 0 ? tile_counters(fs.cull_state, 0) : nullptr;
    const int bv = blend_variant_now(r, false);
    fs.f32acc = (bv & 4) != 0;
    RTR_CUDA(r, launch_fused_ring(s, r->sm_count, r->zmin_variant, bv, r->points, r->n_points, pf.plan.pp, pf.plan.pp, sched, fs.fb.zbuf,
                                  fs.fb.accum, fs.fb.zbuf, ClearTarget{nullptr, 0, nullptr, 0, nullptr}));
    r->launches += 2;
    RTR_CUDA(r, cudaEventRecord(fs.points_done, s));
    return finish_images(r, pf, pf.si, (bv & 4) != 0, nullptr);
}
}  // namespace rtr

namespace {

// D2H of one frame set's outputs on the copy stream (after its render event).
int enqueue_copy(rtr_renderer* r, int si, uint8_t* bgr, float* depth) {
    FrameSet& fs = r->set[si];
    const size_t P = size_t(r->W) * r->H;
    RTR_CUDA(r, cudaStreamWaitEvent(r->copy_stream, fs.rendered, 0));
    if (depth) RTR_CUDA(r, cudaMemcpyAsync(depth, fs.fb.zbuf, P * 4, cudaMemcpyDeviceToHost, r->copy_stream));
    if (bgr) RTR_CUDA(r, cudaMemcpyAsync(bgr, fs.fb.image, P * 3, cudaMemcpyDeviceToHost, r->copy_stream));
    RTR_CUDA(r, cudaEventRecord(fs.copied, r->copy_stream));
    return RTR_OK;
}

int render_to_host(rtr_renderer* r, int stage, uint8_t* bgr, float* depth) {
    if (!r) return RTR_ERR_ARG;
    if (!bgr && !depth) return fail(r, RTR_ERR_ARG, "both output pointers are NULL");  // project_cloud.cu:270-273
    RTR_CUDA(r, cudaSetDevice(r->device));
    int rc = enqueue_frame(r, stage, r->cur);
    if (rc != RTR_OK) return rc;
    rc = enqueue_copy(r, r->cur, bgr, depth);
    if (rc != RTR_OK) return rc;
    RTR_CUDA(r, cudaStreamSynchronize(r->copy_stream));
    return peer_status(r);
}

}  // namespace

extern "C" int rtr_host_ring_claim(uint32_t grid, uint32_t n_queues, uint32_t block, uint32_t group, uint32_t claim, uint32_t* queue,
                                   uint32_t* tile, uint32_t* stages, uint32_t* groups_per_cta) {
    uint32_t st = 0, gr = 0, cps = 0;
    ring_geometry(&st, &gr, &cps);
    if (n_queues < 1 || n_queues > uint32_t(kMaxTileQueues) || group >= gr || !queue || !tile) return RTR_ERR_ARG;
    n_queues = ring_effective_queues(grid, gr, n_queues);   // what the launchers do: no queue without a group
    *queue = ring_queue_of(block, group, gr, n_queues);
    *tile = ring_claimed_tile(grid, st, n_queues, *queue, claim);
    if (stages) *stages = st;
    if (groups_per_cta) *groups_per_cta = gr;
    return RTR_OK;
}
extern "C" uint32_t rtr_host_ring_stride(uint64_t n_points) { return make_ring_schedule(n_points, nullptr, nullptr).perm_mul; }

extern "C" int rtr_host_distortion_bounds(int W, int H, const double* K9, const double* dist5, double* r2_max, double* rstar) {
    if (!K9 || !dist5 || !r2_max || !rstar || W < 1 || H < 1) return RTR_ERR_ARG;
    const double* K = K9;
    // r2_max: points are culled beyond 3x the farthest image corner in normalised coordinates, or where the radial
    // polynomial stops being monotone (fold-back of far off-axis points).
    double rmax2 = 0;
    const double xs[2] = {(0 - K[2]) / K[0], (W - 1 - K[2]) / K[0]};
    const double ys[2] = {(0 - K[5]) / K[4], (H - 1 - K[5]) / K[4]};
    for (double x : xs) for (double y : ys) rmax2 = std::fmax(rmax2, x * x + y * y);
    double lim = rmax2 * 2.25 * 4.0;  // undistorted radius can exceed the distorted one; generous
    const double k1 = dist5[0], k2 = dist5[1], k3 = dist5[4];
    for (int i = 1; i <= 4096; ++i) {  // first r2 where d/dr [ r (1 + k1 r2 + k2 r4 + k3 r6) ] <= 0
        const double r2 = lim * i / 4096.0;
        const double deriv = 1 + 3 * k1 * r2 + 5 * k2 * r2 * r2 + 7 * k3 * r2 * r2 * r2;
        if (deriv <= 0) { lim = lim * (i - 1) / 4096.0; break; }
    }
    *r2_max = double(float(lim));  // the kernel compares in float
    // r*: radius (normalised, undistorted) beyond which no point can land in the image — what chunk culling uses under
    // distortion (enqueue_frame).  A visible point has (xd, yd) inside the image's parallelogram in normalised
    // distorted coordinates, i.e. |(xd, yd)| <= R_d, and |(xd, yd)| >= r |radial(r^2)| - T r^2, where
    // T = 4.3 (|p1| + |p2|) >= sqrt(10) (|p1| + |p2|) bounds the tangential terms.  r* = the largest sampled r in
    // [0, sqrt(r2_max)] that still satisfies g(r) = r |radial| - T r^2 <= R_d (0.1 % slack), plus one step and 1 %.
    const double p1 = dist5[2], p2 = dist5[3], T = 4.3 * (std::fabs(p1) + std::fabs(p2));
    double Rd = 0;
    const double us[2] = {-0.5, W - 0.5}, vs[2] = {-0.5, H - 0.5};
    for (double v : vs) for (double u : us) {
        const double yd = (v - K[5]) / K[4], xd = (u - K[2] - K[1] * yd) / K[0];
        Rd = std::fmax(Rd, std::sqrt(xd * xd + yd * yd));
    }
    const double rmax = std::sqrt(*r2_max);
    double best = 0;
    for (int i = 0; i <= 8192; ++i) {
        const double rr = rmax * i / 8192.0, q = rr * rr;
        const double g = rr * std::fabs(1 + k1 * q + k2 * q * q + k3 * q * q * q) - T * q;
        if (g <= Rd * 1.001) best = rr;
    }
    *rstar = std::fmin(best + rmax / 8192.0, rmax) * 1.01;
    if (!(Rd > 0) || !std::isfinite(*rstar)) *rstar = 0;  // 0 = no culling under this camera
    return RTR_OK;
}

extern "C" {

const char* rtr_version(void) { return "rtr_b200 0.1 (sm_100a)"; }

int rtr_create(int device, rtr_renderer** out) {
    if (!out) return fail(nullptr, RTR_ERR_ARG, "out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(nullptr, RTR_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) + " (rtr_b200 has no CPU fallback)");
    if (device < 0 || device >= n) return fail(nullptr, RTR_ERR_ARG, "device index out of range");
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return cuda_fail(nullptr, e, "cudaGetDeviceProperties");
    if (prop.major != 10)
        return fail(nullptr, RTR_ERR_UNSUPPORTED, "rtr_b200 is built for sm_100a only; device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor));
    rtr_renderer* r = new rtr_renderer;
    r->device = device;
    r->sm_count = prop.multiProcessorCount;
    // Fused sequences keep two compute streams busy: point passes on `stream`, image passes on `image_stream`, which gets
    // the higher priority: its short CTAs fill an SM the moment a CTA of the persistent point pass leaves it, instead of
    // queueing behind the next pass (+4 % frames/s, profiles/r02e_exp_fused_ab.json; RTR_STREAM_PRIORITY=0: equal).
    // Two-pass sequences alternate whole frames between `stream` and `stream2`, which must stay equal in priority.
    int prio_lo = 0, prio_hi = 0;
    const char* pe = std::getenv("RTR_STREAM_PRIORITY");
    if (!(pe && pe[0] == '0')) (void)cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if ((e = cudaSetDevice(device)) != cudaSuccess ||
        (e = cudaStreamCreateWithPriority(&r->stream, cudaStreamNonBlocking, prio_lo)) != cudaSuccess ||
        (e = cudaStreamCreateWithPriority(&r->stream2, cudaStreamNonBlocking, prio_lo)) != cudaSuccess ||
        (e = cudaStreamCreateWithPriority(&r->stream3, cudaStreamNonBlocking, prio_lo)) != cudaSuccess ||
        (e = cudaStreamCreateWithPriority(&r->image_stream, cudaStreamNonBlocking, prio_hi)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&r->copy_stream, cudaStreamNonBlocking)) != cudaSuccess) {
        delete r;
        return cuda_fail(nullptr, e, "stream creation");
    }
    for (auto& s : r->set) {
        cudaEventCreateWithFlags(&s.rendered, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&s.copied, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&s.points_done, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&s.cleared, cudaEventDisableTiming);
    }
    for (auto& ev : r->ev) cudaEventCreate(&ev);
    if (cudaHostAlloc(reinterpret_cast<void**>(&r->tiles_hint), 64, cudaHostAllocMapped) == cudaSuccess) {
        *r->tiles_hint = 0u;
        if (cudaHostGetDevicePointer(reinterpret_cast<void**>(&r->tiles_hint_dev), r->tiles_hint, 0) != cudaSuccess) { cudaFreeHost(r->tiles_hint); r->tiles_hint = nullptr; }
    } else {
        (void)cudaGetLastError();
        r->tiles_hint = nullptr;
    }
    if (cudaHostAlloc(reinterpret_cast<void**>(&r->overflow_note), 64, cudaHostAllocMapped) == cudaSuccess) {
        *r->overflow_note = 0u;
        if (cudaHostGetDevicePointer(reinterpret_cast<void**>(&r->overflow_note_dev), r->overflow_note, 0) != cudaSuccess) r->overflow_note_dev = nullptr;
    } else {
        (void)cudaGetLastError();
        r->overflow_note = nullptr;  // only the adaptive switch is lost
    }
    *out = r;
    return RTR_OK;
}

void rtr_destroy(rtr_renderer* r) {
    if (!r) return;
    cudaSetDevice(r->device);
    r->pending.active = false;  // an outstanding blend is simply dropped
    sync_compute(r);
    cudaStreamSynchronize(r->copy_stream);
    if (r->comm) g_nccl.CommDestroy(r->comm);
    rtr_peer_detach(r);
    cudaFree(r->peer.flags);
    if (r->peer.err_host) cudaFreeHost(r->peer.err_host);
    cudaFree(r->post_scratch);
    if (r->overflow_note) cudaFreeHost(r->overflow_note);
    if (r->tiles_hint) cudaFreeHost(r->tiles_hint);
    free_frame_sets(r);
    if (r->owns_points) cudaFree(r->points);
    free_cull_storage(r);
    for (auto& s : r->set) { cudaEventDestroy(s.rendered); cudaEventDestroy(s.copied); cudaEventDestroy(s.points_done); cudaEventDestroy(s.cleared); }
    for (auto& ev : r->ev) cudaEventDestroy(ev);
    for (auto& ev : r->ev_pool) cudaEventDestroy(ev);
    cudaStreamDestroy(r->stream);
    cudaStreamDestroy(r->stream2);
    cudaStreamDestroy(r->stream3);
    cudaStreamDestroy(r->copy_stream);
    cudaStreamDestroy(r->image_stream);
    delete r;
}

const char* rtr_last_error(const rtr_renderer* r) { return r ? r->err.c_str() : g_create_error.c_str(); }

}  // extern "C"

namespace rtr {
int replace_cloud(rtr_renderer* r, uint64_t n) {
    RTR_CUDA(r, cudaSetDevice(r->device));
    r->pending.active = false;  // a frame of the old cloud whose blend is outstanding is dropped with the cloud
    RTR_CUDA(r, sync_compute(r));
    if (r->owns_points) cudaFree(r->points);
    free_cull_storage(r);
    r->index_base = 0;
    r->n_chunks = 0;
    r->points = nullptr; r->n_points = 0; r->owns_points = false;
    if (n == 0) return RTR_OK;
    RTR_CUDA(r, cudaMalloc(reinterpret_cast<void**>(&r->points), n * sizeof(PointRecord)));
    r->owns_points = true;
    r->n_points = n;
    return RTR_OK;
}

// Chunk bounds + visible-list storage for the cloud now in r->points (every upload path ends here).
int build_chunk_bounds(rtr_renderer* r) {
    if (r->n_points == 0) return RTR_OK;
    r->n_chunks = uint32_t((r->n_points + kChunkPoints - 1) / kChunkPoints);
    r->ring_sched = make_ring_schedule(r->n_points, nullptr, nullptr);
    RTR_CUDA(r, cudaMalloc(reinterpret_cast<void**>(&r->bounds), size_t(r->n_chunks) * sizeof(ChunkBounds)));
    for (auto& fs : r->set) {
        RTR_CUDA(r, cudaMalloc(reinterpret_cast<void**>(&fs.vis_list), size_t(r->n_chunks) * sizeof(uint32_t)));
        fs.band_cap = (r->n_chunks + 3u) & ~3u;
        RTR_CUDA(r, cudaMalloc(reinterpret_cast<void**>(&fs.band_scratch), size_t(kMaxBands) * fs.band_cap * sizeof(uint32_t)));
        RTR_CUDA(r, cudaMalloc(reinterpret_cast<void**>(&fs.cull_state), kCullStateAlloc));  // + the tile-claim counters
        RTR_CUDA(r, cudaMemsetAsync(fs.cull_state, 0, kCullStateAlloc, r->stream));
        fs.cull_parity = 0;
    }
    RTR_CUDA(r, launch_chunk_bounds(r->stream, r->points, r->n_points, r->bounds));
    r->launches += 1;
    RTR_CUDA(r, sync_compute(r));
    return RTR_OK;
}

int finish_upload(rtr_renderer* r) {
    if (r->sort_on_upload && r->owns_points && r->n_points > 1) return reorder_morton(r);
    return build_chunk_bounds(r);
}
}  // namespace rtr

extern "C" {

int rtr_upload_cloud_xyz_bgr(rtr_renderer* r, const float* xyz, const uint8_t* bgr, uint64_t n) {
    if (!r) return RTR_ERR_ARG;
    if (n && (!xyz || !bgr)) return fail(r, RTR_ERR_ARG, "xyz/bgr is NULL");
    if (n > 0xFFFFFFFFull) return fail(r, RTR_ERR_UNSUPPORTED, "more than 2^32 points per renderer");
    int rc = replace_cloud(r, n);
    if (rc != RTR_OK || n == 0) return rc;
    // repack on the host through a pinned staging ring, 4 Mi points per chunk
    const uint64_t chunk = 1ull << 22;
    PointRecord* stage[2] = {nullptr, nullptr};
    cudaEvent_t done[2];
    for (int i = 0; i < 2; ++i) {
        RTR_CUDA(r, cudaMallocHost(reinterpret_cast<void**>(&stage[i]), std::min(chunk, n) * sizeof(PointRecord)));
        cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming);
    }
    int b = 0;
    for (uint64_t off = 0; off < n; off += chunk, b ^= 1) {
        const uint64_t m = std::min(chunk, n - off);
        cudaEventSynchronize(done[b]);
        for (uint64_t i = 0; i < m; ++i) {
            const float* p = xyz + (off + i) * 3;
            const uint8_t* c = bgr + (off + i) * 3;
            stage[b][i].x = p[0]; stage[b][i].y = p[1]; stage[b][i].z = p[2];
            stage[b][i].bgra = uint32_t(c[0]) | (uint32_t(c[1]) << 8) | (uint32_t(c[2]) << 16) | 0xFF000000u;  // Octreegrid.h:176
        }
        RTR_CUDA(r, cudaMemcpyAsync(r->points + off, stage[b], m * sizeof(PointRecord), cudaMemcpyHostToDevice, r->stream));
        cudaEventRecord(done[b], r->stream);
    }
    RTR_CUDA(r, sync_compute(r));
    for (int i = 0; i < 2; ++i) { cudaFreeHost(stage[i]); cudaEventDestroy(done[i]); }
    return finish_upload(r);
}

int rtr_upload_cloud_packed16(rtr_renderer* r, const void* host_records, uint64_t n) {
    if (!r) return RTR_ERR_ARG;
    if (n && !host_records) return fail(r, RTR_ERR_ARG, "records is NULL");
    if (n > 0xFFFFFFFFull) return fail(r, RTR_ERR_UNSUPPORTED, "more than 2^32 points per renderer");
    int rc = replace_cloud(r, n);
    if (rc != RTR_OK || n == 0) return rc;
    RTR_CUDA(r, cudaMemcpyAsync(r->points, host_records, n * sizeof(PointRecord), cudaMemcpyHostToDevice, r->stream));
    RTR_CUDA(r, sync_compute(r));
    return finish_upload(r);
}

int rtr_adopt_device_cloud_packed16(rtr_renderer* r, void* device_records, uint64_t n) {
    if (!r) return RTR_ERR_ARG;
    if (n && !device_records) return fail(r, RTR_ERR_ARG, "records is NULL");
    if (reinterpret_cast<uintptr_t>(device_records) & 15) return fail(r, RTR_ERR_ARG, "records must be 16-byte aligned");
    int rc = replace_cloud(r, 0);
    if (rc != RTR_OK) return rc;
    r->points = static_cast<PointRecord*>(device_records);
    r->n_points = n;
    r->owns_points = false;
    return build_chunk_bounds(r);
}

int rtr_synth_cloud(rtr_renderer* r, uint64_t seed, uint64_t n_total, uint64_t first, uint64_t count, int lx, int ly,
                    int lz, int nbox) {
    if (!r) return RTR_ERR_ARG;
    if (lx < 8 || ly < 8 || lz < 4 || nbox < 0 || nbox > 20) return fail(r, RTR_ERR_ARG, "scene dims out of range");
    if (count > 0xFFFFFFFFull) return fail(r, RTR_ERR_UNSUPPORTED, "more than 2^32 points per renderer");
    int rc = replace_cloud(r, count);
    if (rc != RTR_OK || count == 0) return rc;
    RTR_CUDA(r, launch_synth(r->stream, seed, n_total, first, count, lx, ly, lz, nbox, r->points));
    r->launches += 1;
    r->index_base = first;
    RTR_CUDA(r, sync_compute(r));
    return finish_upload(r);
}

uint64_t rtr_cloud_size(const rtr_renderer* r) { return r ? r->n_points : 0; }

int rtr_download_cloud_packed16(rtr_renderer* r, uint64_t first, uint64_t count, void* host_records) {
    if (!r || !host_records) return RTR_ERR_ARG;
    if (first + count > r->n_points) return fail(r, RTR_ERR_ARG, "range exceeds cloud");
    RTR_CUDA(r, cudaSetDevice(r->device));
    RTR_CUDA(r, cudaMemcpyAsync(host_records, r->points + first, count * sizeof(PointRecord), cudaMemcpyDeviceToHost, r->stream));
    RTR_CUDA(r, sync_compute(r));
    return RTR_OK;
}

int rtr_set_intrinsics_matrix(rtr_renderer* r, int width, int height, const double* K9, const double* dist5) {
    if (!r || !K9) return RTR_ERR_ARG;
    if (width < 16 || height < 16 || width > 32768 || height > 32768) return fail(r, RTR_ERR_ARG, "width/height must be in [16, 32768]");
    double d5[5];
    for (int i = 0; i < 5; ++i) d5[i] = dist5 ? dist5[i] : 0.0;
    if (r->W != width || r->H != height || std::memcmp(r->K, K9, sizeof(r->K)) || std::memcmp(r->dist, d5, sizeof(d5))) r->dist_bounds_valid = false;
    r->W = width; r->H = height;
    std::memcpy(r->K, K9, sizeof(r->K));
    std::memcpy(r->dist, d5, sizeof(d5));
    r->have_K = true;
    return RTR_OK;
}

int rtr_set_intrinsics(rtr_renderer* r, int width, int height, double fx, double fy, double cx, double cy, double skew,
                       const double* dist5) {
    const double K9[9] = {fx, skew, cx, 0, fy, cy, 0, 0, 1};
    return rtr_set_intrinsics_matrix(r, width, height, K9, dist5);
}

int rtr_set_pose_w2c(rtr_renderer* r, const double* E16) {
    if (!r || !E16) return RTR_ERR_ARG;
    std::memcpy(r->E, E16, sizeof(r->E));
    r->have_E = true;
    r->raw_proj = false;
    return RTR_OK;
}

int rtr_set_cam_proj_raw(rtr_renderer* r, const float* m16) {
    if (!r || !m16) return RTR_ERR_ARG;
    std::memcpy(r->cam_proj, m16, sizeof(r->cam_proj));
    r->raw_proj = true;
    return RTR_OK;
}

int rtr_get_cam_proj(const rtr_renderer* r, float* m16) {
    if (!r || !m16) return RTR_ERR_ARG;
    if (r->raw_proj) std::memcpy(m16, r->cam_proj, 64);
    else if (r->have_K && r->have_E) build_cam_proj(r->K, r->E, m16);
    else return RTR_ERR_STATE;
    return RTR_OK;
}

int rtr_render_rgbd(rtr_renderer* r, uint8_t* bgr, float* depth) { return render_to_host(r, RTR_STAGE_RGBD, bgr, depth); }
int rtr_render_filtered(rtr_renderer* r, uint8_t* bgr, float* depth) { return render_to_host(r, RTR_STAGE_FILTERED, bgr, depth); }

int rtr_render_tensor(rtr_renderer* r, void** device_fp16) {
    if (!r || !device_fp16) return RTR_ERR_ARG;
    RTR_CUDA(r, cudaSetDevice(r->device));
    int rc = enqueue_frame(r, RTR_STAGE_FILTERED, r->cur);
    if (rc != RTR_OK) return rc;
    RTR_CUDA(r, sync_compute(r));
    *device_fp16 = r->set[r->cur].fb.tensor;
    return peer_status(r);
}

// One asynchronous frame of a sequence: fused with its neighbours when the sequence qualifies (see enqueue_fused),
// else a whole frame alternating between two frame sets / streams (pipeline) or into the current set.
static int enqueue_sequence_frame(rtr_renderer* r, int stage, uint8_t* bgr, float* depth) {
    if (!r->points || r->n_points == 0) return fail(r, RTR_ERR_STATE, "no cloud uploaded");
    FramePlan pl;
    int rc = plan_frame(r, pl);
    if (rc != RTR_OK) return rc;
    if (fused_sequence(r, pl)) return enqueue_fused(r, stage, pl, bgr, depth);
    if ((rc = flush_pending(r)) != RTR_OK) return rc;
    if (r->cur >= sequence_depth(r)) r->cur = 0;  // whole frames rotate through sets 0, 1 (and 2: option pipeline_depth = 3)
    const int si = (bgr || depth || pipelined(r)) ? next_set(r, r->cur) : r->cur;  // another set: this one may still drain over PCIe
    rc = enqueue_frame(r, stage, si, true);
    if (rc != RTR_OK) return rc;
    r->cur = si;
    if (bgr || depth) rc = enqueue_copy(r, si, bgr, depth);
    return rc;
}

int rtr_render_device(rtr_renderer* r, int stage) {
    if (!r) return RTR_ERR_ARG;
    if (stage != RTR_STAGE_RGBD && stage != RTR_STAGE_FILTERED) return fail(r, RTR_ERR_ARG, "bad stage");
    RTR_CUDA(r, cudaSetDevice(r->device));
    // back-to-back asynchronous frames: r->cur = the set of the frame enqueued last
    return enqueue_sequence_frame(r, stage, nullptr, nullptr);
}

int rtr_sync(rtr_renderer* r) {
    if (!r) return RTR_ERR_ARG;
    RTR_CUDA(r, cudaSetDevice(r->device));
    const int rc = flush_pending(r);
    if (rc != RTR_OK) return rc;
    RTR_CUDA(r, sync_compute(r));
    RTR_CUDA(r, cudaStreamSynchronize(r->copy_stream));
    return peer_status(r);
}

int rtr_render_trajectory(rtr_renderer* r, int stage, const double* poses, int n_frames, uint8_t* bgr, float* depth) {
    if (!r || !poses || n_frames < 0) return RTR_ERR_ARG;
    if (stage != RTR_STAGE_RGBD && stage != RTR_STAGE_FILTERED) return fail(r, RTR_ERR_ARG, "bad stage");
    RTR_CUDA(r, cudaSetDevice(r->device));
    const size_t P = size_t(r->W) * r->H;
    for (int f = 0; f < n_frames; ++f) {
        int rc = rtr_set_pose_w2c(r, poses + size_t(f) * 16);
        if (rc != RTR_OK) return rc;
        rc = enqueue_sequence_frame(r, stage, bgr ? bgr + size_t(f) * P * 3 : nullptr, depth ? depth + size_t(f) * P : nullptr);
        if (rc != RTR_OK) return rc;
    }
    return rtr_sync(r);
}

int rtr_get_device_buffers(rtr_renderer* r, rtr_device_buffers* out) {
    if (!r || !out) return RTR_ERR_ARG;
    if (r->alloc_W == 0) return fail(r, RTR_ERR_STATE, "no frame rendered yet");
    RTR_CUDA(r, cudaSetDevice(r->device));
    const int frc = flush_pending(r);
    if (frc != RTR_OK) return frc;
    const FrameBuffers& fb = r->set[r->cur].fb;
    std::memset(out, 0, sizeof(*out));
    out->points = r->points; out->zbuf = fb.zbuf; out->accum = fb.accum; out->image = fb.image; out->tensor = fb.tensor;
    out->minmax = fb.minmax;
    for (int i = 0; i < 5; ++i) {
        out->level[i] = fb.level[i];
        out->level_w[i] = r->dims.w[i]; out->level_h[i] = r->dims.h[i];
        out->up_w[i] = r->dims.uw[i]; out->up_h[i] = r->dims.uh[i];
    }
    for (int i = 0; i < 4; ++i) out->mask[i] = r->keep_masks ? fb.mask[i] : nullptr;
    out->width = r->alloc_W; out->height = r->alloc_H;
    out->tensor_plane = uint64_t(r->dims.uw[0]) * r->dims.uh[0];
    // work the caller orders on `stream` after this call must see the frame(s) in flight, whichever stream they run on
    for (auto& fs : r->set) RTR_CUDA(r, cudaStreamWaitEvent(r->stream, fs.rendered, 0));
    out->stream = r->stream;
    return RTR_OK;
}

int rtr_read_buffer(rtr_renderer* r, int what, void* dst, size_t bytes) {
    if (!r || !dst) return RTR_ERR_ARG;
    if (r->alloc_W == 0) return fail(r, RTR_ERR_STATE, "no frame rendered yet");
    RTR_CUDA(r, cudaSetDevice(r->device));
    const int frc = flush_pending(r);
    if (frc != RTR_OK) return frc;
    const FrameBuffers& fb = r->set[r->cur].fb;
    const size_t P = size_t(r->alloc_W) * r->alloc_H;
    const void* src = nullptr;
    size_t cap = 0;
    if (what == 0) { src = fb.zbuf; cap = P * 4; }
    else if (what == 1) { src = fb.accum; cap = P * 16; }
    else if (what == 2) { src = fb.image; cap = P * 3; }
    else if (what == 3) { src = fb.tensor; cap = P * 10; }
    else if (what == 4) { src = fb.minmax; cap = 12; }
    else if (what >= 5 && what <= 8) { const int i = what - 4; src = fb.level[i]; cap = size_t(r->dims.w[i]) * r->dims.h[i] * 4; }
    else if (what >= 9 && what <= 12) { const int i = what - 9; src = r->keep_masks ? fb.mask[i] : nullptr; cap = size_t(r->dims.uw[i]) * r->dims.uh[i]; }
    if (!src) return fail(r, RTR_ERR_ARG, "buffer not available");
    if (bytes > cap) return fail(r, RTR_ERR_ARG, "read exceeds buffer");
    RTR_CUDA(r, cudaSetDevice(r->device));
    RTR_CUDA(r, sync_compute(r));
    RTR_CUDA(r, cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    if (what == 1 && r->set[r->cur].f32acc) {  // always hand out the reference's layout: {b,g,r,count} as u32
        uint32_t flag = 0;
        RTR_CUDA(r, cudaMemcpy(&flag, fb.minmax + 2, 4, cudaMemcpyDeviceToHost));
        if (!flag) {  // (after an overflow the exact re-run left integers in the buffer)
            uint32_t* u = static_cast<uint32_t*>(dst);
            for (size_t i = 0; i < bytes / 4; ++i) {
                float f;
                std::memcpy(&f, u + i, 4);
                u[i] = static_cast<uint32_t>(f);
            }
        }
    }
    return RTR_OK;
}

int rtr_project_points(rtr_renderer* r, int32_t* pix_host, uint32_t* zbits_host) {
    if (!r || !pix_host || !zbits_host) return RTR_ERR_ARG;
    if (!r->points || r->n_points == 0) return fail(r, RTR_ERR_STATE, "no cloud uploaded");
    RTR_CUDA(r, cudaSetDevice(r->device));
    ProjParams pp;
    int rc = make_params(r, pp);
    if (rc != RTR_OK) return rc;
    int32_t* d_pix = nullptr;
    uint32_t* d_z = nullptr;
    RTR_CUDA(r, cudaMalloc(reinterpret_cast<void**>(&d_pix), r->n_points * 4));
    RTR_CUDA(r, cudaMalloc(reinterpret_cast<void**>(&d_z), r->n_points * 4));
    cudaError_t e = launch_project_dump(r->stream, r->points, r->n_points, pp, d_pix, d_z);
    r->launches += 1;
    if (e == cudaSuccess) e = cudaMemcpyAsync(pix_host, d_pix, r->n_points * 4, cudaMemcpyDeviceToHost, r->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(zbits_host, d_z, r->n_points * 4, cudaMemcpyDeviceToHost, r->stream);
    if (e == cudaSuccess) e = sync_compute(r);
    cudaFree(d_pix); cudaFree(d_z);
    if (e != cudaSuccess) return cuda_fail(r, e, "project dump");
    return RTR_OK;
}

int rtr_bench_red_min(rtr_renderer* r, int mode, uint64_t n_ops, int key64, int iters, float* ms_per_launch,
                      uint64_t* live_ops) {
    if (!r || !ms_per_launch || iters < 1 || mode < 0 || mode > 19) return RTR_ERR_ARG;
    RTR_CUDA(r, cudaSetDevice(r->device));
    ProjParams pp;
    int rc = make_params(r, pp);
    if (rc != RTR_OK) return rc;
    const uint64_t P = uint64_t(r->W) * r->H;
    if (mode == 1) {
        if (!r->points || r->n_points == 0) return fail(r, RTR_ERR_STATE, "no cloud uploaded");
        n_ops = r->n_points;
    }
    void* zb = nullptr;
    int32_t* d_pix = nullptr;
    uint32_t* d_z = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    const bool wide = mode == 2 || mode == 3 || mode >= 12;  // 16-byte accumulators
    const size_t zb_bytes = P * (wide ? 16 : (key64 ? 8 : 4));
    cudaError_t e = cudaMalloc(&zb, zb_bytes);
    if (e == cudaSuccess) e = cudaMemsetAsync(zb, wide ? 0 : 0xFF, zb_bytes, r->stream);
    if (e == cudaSuccess && mode == 1) {
        e = cudaMalloc(reinterpret_cast<void**>(&d_pix), n_ops * 4);
        if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&d_z), n_ops * 4);
        if (e == cudaSuccess) e = launch_project_dump(r->stream, r->points, n_ops, pp, d_pix, d_z);
    }
    if (e == cudaSuccess) e = cudaEventCreate(&e0);
    if (e == cudaSuccess) e = cudaEventCreate(&e1);
    for (int it = 0; it < iters + 1 && e == cudaSuccess; ++it) {  // first launch = warm-up
        if (it == 1) e = cudaEventRecord(e0, r->stream);
        if (e == cudaSuccess)
            e = launch_red_bench(r->stream, r->sm_count, mode, key64 != 0, d_pix, n_ops, uint32_t(P),
                                 static_cast<uint32_t*>(zb), static_cast<unsigned long long*>(zb));
        r->launches += 1;
    }
    if (e == cudaSuccess) e = cudaEventRecord(e1, r->stream);
    if (e == cudaSuccess) e = cudaEventSynchronize(e1);
    float ms = 0.f;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
    *ms_per_launch = ms / float(iters);
    if (live_ops) {
        *live_ops = n_ops;
        if (mode == 1 && e == cudaSuccess) {  // count the in-frustum points on the host (measurement support only)
            std::vector<int32_t> h(n_ops);
            e = cudaMemcpy(h.data(), d_pix, n_ops * 4, cudaMemcpyDeviceToHost);
            uint64_t c = 0;
            for (int32_t v : h) c += (v >= 0);
            *live_ops = c;
        }
    }
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaFree(zb); cudaFree(d_pix); cudaFree(d_z);
    if (e != cudaSuccess) return cuda_fail(r, e, "rtr_bench_red_min");
    return RTR_OK;
}

int rtr_selftest_fast_divide(rtr_renderer* r, uint64_t n_pairs, uint64_t seed, uint64_t* mismatches) {
    if (!r || !mismatches) return RTR_ERR_ARG;
    RTR_CUDA(r, cudaSetDevice(r->device));
    unsigned long long* d = nullptr;
    RTR_CUDA(r, cudaMalloc(reinterpret_cast<void**>(&d), 8));
    cudaError_t e = cudaMemsetAsync(d, 0, 8, r->stream);
    if (e == cudaSuccess) e = launch_fast_divide_selftest(r->stream, r->sm_count, n_pairs, seed, d);
    r->launches += 1;
    unsigned long long h = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, r->stream);
    if (e == cudaSuccess) e = sync_compute(r);
    cudaFree(d);
    if (e != cudaSuccess) return cuda_fail(r, e, "rtr_selftest_fast_divide");
    *mismatches = h;
    return RTR_OK;
}

static int* option_slot(rtr_renderer* r, const char* key) {
    if (!std::strcmp(key, "zmin_variant")) return &r->zmin_variant;
    if (!std::strcmp(key, "zmin_unroll")) return &r->zmin_unroll;
    if (!std::strcmp(key, "blend_variant")) return &r->blend_variant;
    if (!std::strcmp(key, "blend_unroll")) return &r->blend_unroll;
    if (!std::strcmp(key, "force_generic")) return &r->force_generic;
    if (!std::strcmp(key, "keep_masks")) return &r->keep_masks;
    if (!std::strcmp(key, "timing")) return &r->timing;
    if (!std::strcmp(key, "key64")) return &r->key64;
    if (!std::strcmp(key, "chunk_cull")) return &r->chunk_cull;
    if (!std::strcmp(key, "sort_on_upload")) return &r->sort_on_upload;
    if (!std::strcmp(key, "ring")) return &r->ring;
    if (!std::strcmp(key, "bands")) return &r->bands;
    if (!std::strcmp(key, "fused_up")) return &r->fused_up;
    if (!std::strcmp(key, "ring_perm")) return &r->ring_perm;
    if (!std::strcmp(key, "ring_early")) return &r->ring_early;
    if (!std::strcmp(key, "ring_dynamic")) return &r->ring_dynamic;
    if (!std::strcmp(key, "clear_lean")) return &r->clear_lean;
    if (!std::strcmp(key, "ring_ctas")) return &r->ring_ctas;
    if (!std::strcmp(key, "ring_claim_min")) return &r->ring_claim_min;
    if (!std::strcmp(key, "pipeline")) return &r->pipeline;
    if (!std::strcmp(key, "pipeline_depth")) return &r->pipeline_depth;
    if (!std::strcmp(key, "fuse")) return &r->fuse;
    if (!std::strcmp(key, "fused_tiles_per_cta")) return &r->fused_tiles_per_cta;
    if (!std::strcmp(key, "fixup_launches")) return &r->fixup_launches;
    if (!std::strcmp(key, "peer_timeout_ms")) return &r->peer.timeout_ms;
    return nullptr;
}

int rtr_set_option(rtr_renderer* r, const char* key, int64_t value) {
    if (!r || !key) return RTR_ERR_ARG;
    // options apply to frames enqueued from now on: a frame whose blend is still outstanding is completed first
    RTR_CUDA(r, cudaSetDevice(r->device));
    const int frc = flush_pending(r);
    if (frc != RTR_OK) return frc;
    if (!std::strcmp(key, "index_base")) {
        // the 64-bit key holds the global point index in its low word
        if (value < 0 || uint64_t(value) + r->n_points > (1ull << 32)) return fail(r, RTR_ERR_ARG, "index_base + cloud size must not exceed 2^32");
        r->index_base = uint64_t(value);
        return RTR_OK;
    }
    int* slot = option_slot(r, key);
    if (!slot) return fail(r, RTR_ERR_ARG, std::string("unknown option: ") + key);
    if ((!std::strcmp(key, "zmin_unroll") || !std::strcmp(key, "blend_unroll")) && value != 1 && value != 2 && value != 4 && value != 8)
        return fail(r, RTR_ERR_ARG, "unroll must be 1, 2, 4 or 8");
    // Bits 8 / 16 / 32 select measurement-only kernels (no RED issued, ATOMG builtin, no in-register merge) whose frames
    // are WRONG or slower by design: they exist only in -DRTR_EXPERIMENTS builds (RTR_EXPERIMENTS=1 python build.py),
    // never in the library a caller links.
    // (bit 64 = shared-memory tile pre-reduction of the z-min ring pass: exact, a supported variant)
#ifdef RTR_EXPERIMENTS
    const int64_t zmask = 7 | 8 | 16 | 32 | 64, bmask = 6 | 32;
#else
    const int64_t zmask = 7 | 64, bmask = 6;
#endif
    if (!std::strcmp(key, "zmin_variant") &&
        (value < 0 || (value & ~zmask) || !((value & 7) == 0 || (value & 7) == 1 || (value & 7) == 2 || (value & 7) == 3 || (value & 7) == 5 || (value & 7) == 7)))
        return fail(r, RTR_ERR_ARG, "zmin_variant must be one of 0,1,2,3,5,7 (measurement bits 8/16/32 only in RTR_EXPERIMENTS builds)");
    if (!std::strcmp(key, "blend_variant") && (value < 0 || (value & ~bmask)))
        return fail(r, RTR_ERR_ARG, "blend_variant must be 0, 2, 4 or 6 (measurement bit 32 only in RTR_EXPERIMENTS builds)");
    if (!std::strcmp(key, "timing") && (value < 0 || value > 3)) return fail(r, RTR_ERR_ARG, "timing must be 0 ... 3");
    if (!std::strcmp(key, "bands") && (value < 0 || value > kMaxBands)) return fail(r, RTR_ERR_ARG, "bands must be 0 (by frame size), 1 (off) or 2 ... 8");
    if (!std::strcmp(key, "fuse") && (value < 0 || value > 2)) return fail(r, RTR_ERR_ARG, "fuse must be 0 (never), 1 (large clouds) or 2 (always)");
    *slot = int(value);
    return RTR_OK;
}

int64_t rtr_get_option(const rtr_renderer* r, const char* key) {
    if (!r || !key) return RTR_ERR_ARG;
    if (!std::strcmp(key, "index_base")) return int64_t(r->index_base);
    if (!std::strcmp(key, "sm_count")) return r->sm_count;
    if (!std::strcmp(key, "int_sum_frames")) return r->int_sum_frames;  // frames left that start with integer colour sums
    if (!std::strcmp(key, "fuse_active")) {  // would the next frame of a sequence take the fused path (camera set, options as they are)?
        FramePlan pl;
        rtr_renderer* m = const_cast<rtr_renderer*>(r);
        return (m->points && plan_frame(m, pl) == RTR_OK && fused_sequence(r, pl)) ? 1 : 0;
    }
    if (!std::strcmp(key, "bands_active")) return (r->W >= 16 && r->H >= 16) ? int64_t(frame_bands(r)) : 1;  // bands the next frame's list is ordered by
    if (!std::strcmp(key, "pending")) return r->pending.active ? 1 : 0;  // a fused sequence's last frame still lacks its blend
    if (!std::strcmp(key, "experiments")) {
#ifdef RTR_EXPERIMENTS
        return 1;
#else
        return 0;
#endif
    }
    if (!std::strcmp(key, "peer_error"))  // 1 while a timed-out wait of the peer merge has not been reported by a render / sync call yet
        return r->peer.err_host ? int64_t(*reinterpret_cast<volatile uint32_t*>(r->peer.err_host)) : 0;
    const int* slot = option_slot(const_cast<rtr_renderer*>(r), key);
    return slot ? *slot : RTR_ERR_ARG;
}

int rtr_get_stage_ms(rtr_renderer* r, float* ms6) {
    if (!r || !ms6) return RTR_ERR_ARG;
    if (r->timing != 1) return fail(r, RTR_ERR_STATE, "option timing is not 1");
    RTR_CUDA(r, cudaSetDevice(r->device));
    RTR_CUDA(r, cudaEventSynchronize(r->ev[5]));
    for (int i = 0; i < 5; ++i) RTR_CUDA(r, cudaEventElapsedTime(&ms6[i], r->ev[i], r->ev[i + 1]));
    RTR_CUDA(r, cudaEventElapsedTime(&ms6[5], r->ev[0], r->ev[5]));
    return RTR_OK;
}

int rtr_get_stage_ms_sum(rtr_renderer* r, double* ms6_sum, uint64_t* n_frames, int reset) {
    if (!r || !ms6_sum || !n_frames) return RTR_ERR_ARG;
    RTR_CUDA(r, cudaSetDevice(r->device));
    int rc = drain_event_pool(r);
    if (rc != RTR_OK) return rc;
    for (int i = 0; i < 6; ++i) ms6_sum[i] = r->ev_sum[i];
    *n_frames = r->ev_count;
    if (reset) {
        for (double& v : r->ev_sum) v = 0.0;
        r->ev_count = 0;
    }
    return RTR_OK;
}

static int read_cull_totals(rtr_renderer* r, CullState* sum, int reset) {
    std::memset(sum, 0, sizeof(*sum));
    if (!r->set[0].cull_state) return RTR_OK;
    RTR_CUDA(r, cudaSetDevice(r->device));
    const int frc = flush_pending(r);
    if (frc != RTR_OK) return frc;
    RTR_CUDA(r, sync_compute(r));
    for (auto& fs : r->set) {  // frames of a sequence rotate through the sets
        CullState st;
        RTR_CUDA(r, cudaMemcpy(&st, fs.cull_state, sizeof(st), cudaMemcpyDeviceToHost));
        if (st.armed) cull_fold(&st, st.parity & 1u);  // the list in flight is folded on the device by the next classification
        sum->frames += st.frames;
        sum->total_visible += st.total_visible;
        sum->total_streamed += st.total_streamed;
        sum->passes += st.passes;
        if (reset) {
            RTR_CUDA(r, cudaMemset(fs.cull_state, 0, sizeof(CullState)));
            RTR_CUDA(r, cudaMemset(smem_tile_stats(fs.cull_state), 0, 64));
            fs.cull_parity = 0;
        }
    }
    return RTR_OK;
}

int rtr_get_smem_tile_stats(rtr_renderer* r, uint64_t* stats4) {
    if (!r || !stats4) return RTR_ERR_ARG;
    for (int i = 0; i < 4; ++i) stats4[i] = 0;
    if (!r->set[0].cull_state) return RTR_OK;
    RTR_CUDA(r, cudaSetDevice(r->device));
    const int frc = flush_pending(r);
    if (frc != RTR_OK) return frc;
    RTR_CUDA(r, sync_compute(r));
    for (auto& fs : r->set) {
        unsigned long long v[4];
        RTR_CUDA(r, cudaMemcpy(v, smem_tile_stats(fs.cull_state), sizeof(v), cudaMemcpyDeviceToHost));
        for (int i = 0; i < 4; ++i) stats4[i] += v[i];
    }
    return RTR_OK;
}

int rtr_get_cull_stats(rtr_renderer* r, uint64_t* frames, uint64_t* visible_chunks_total, uint64_t* n_chunks, int reset) {
    if (!r || !frames || !visible_chunks_total || !n_chunks) return RTR_ERR_ARG;
    CullState t;
    const int rc = read_cull_totals(r, &t, reset);
    *frames = t.frames;
    *visible_chunks_total = t.total_visible;
    *n_chunks = r->n_chunks;
    return rc;
}

int rtr_get_stream_stats(rtr_renderer* r, uint64_t* passes, uint64_t* chunks_streamed, int reset) {
    if (!r || !passes || !chunks_streamed) return RTR_ERR_ARG;
    CullState t;
    const int rc = read_cull_totals(r, &t, reset);
    *passes = t.passes;
    *chunks_streamed = t.total_streamed;
    return rc;
}

uint64_t rtr_launch_count(const rtr_renderer* r) { return r ? r->launches : 0; }

int rtr_comm_unique_id(void* id128) {
    if (!id128) return RTR_ERR_ARG;
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    std::string err;
    if (!g_nccl.load(err)) return fail(nullptr, RTR_ERR_COMM, err);
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != 0) return fail(nullptr, RTR_ERR_COMM, "ncclGetUniqueId failed");
    std::memcpy(id128, &id, 128);
    return RTR_OK;
}

int rtr_comm_init(rtr_renderer* r, const void* id128, int rank, int n_ranks) {
    if (!r || !id128 || n_ranks < 1 || rank < 0 || rank >= n_ranks) return RTR_ERR_ARG;
    {
        std::lock_guard<std::mutex> lk(g_nccl_mu);
        std::string err;
        if (!g_nccl.load(err)) return fail(r, RTR_ERR_COMM, err);
    }
    RTR_CUDA(r, cudaSetDevice(r->device));
    ncclUniqueId id;
    std::memcpy(&id, id128, 128);
    const int rc = g_nccl.CommInitRank(&r->comm, n_ranks, id, rank);
    if (rc != 0) { r->comm = nullptr; return fail(r, RTR_ERR_COMM, std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error")); }
    r->rank = rank; r->n_ranks = n_ranks;
    return RTR_OK;
}

namespace {
struct PeerBlob {
    uint32_t magic;
    int32_t W, H;
    int32_t pid, device, has_zkey;
    uint64_t zbuf_off, image_off;
    void* raw_flags;             // the addresses themselves: peers inside ONE process (one thread per GPU) use them directly
    void* raw_arena[2];
    void* raw_zkey[2];
    cudaIpcMemHandle_t flags, arena[2], zkey[2];
};
static_assert(sizeof(PeerBlob) <= 512, "blob must fit RTR_PEER_BLOB_BYTES");
}  // namespace

int rtr_peer_export(rtr_renderer* r, void* blob512) {
    if (!r || !blob512) return RTR_ERR_ARG;
    RTR_CUDA(r, cudaSetDevice(r->device));
    if (r->W < 16 || r->H < 16) return fail(r, RTR_ERR_STATE, "set the intrinsics before rtr_peer_export");
    if ((uint64_t(r->W) * r->H) % 4) return fail(r, RTR_ERR_UNSUPPORTED, "peer merge needs W*H divisible by 4");
    int rc = rtr_peer_detach(r);  // (completes an outstanding frame, waits for the streams)
    if (rc != RTR_OK) return rc;
    rc = ensure_buffers(r);
    if (rc != RTR_OK) return rc;
    rtr_renderer::Peer& pe = r->peer;
    if (!pe.flags) RTR_CUDA(r, cudaMalloc(reinterpret_cast<void**>(&pe.flags), 4096));
    if (!pe.err_host) {
        RTR_CUDA(r, cudaHostAlloc(reinterpret_cast<void**>(&pe.err_host), 64, cudaHostAllocMapped));
        RTR_CUDA(r, cudaHostGetDevicePointer(reinterpret_cast<void**>(&pe.err_dev), pe.err_host, 0));
    }
    // A fresh protocol state with every export: flag words and epochs start from zero on every rank, so a re-attach
    // (after a resolution change, a lost peer, a different set of ranks) can never meet stale epochs.  Nobody writes
    // this rank's flags any more: its last merge kernel only finished after every peer's last signal had arrived.
    RTR_CUDA(r, cudaMemset(pe.flags, 0, 4096));
    *reinterpret_cast<volatile uint32_t*>(pe.err_host) = 0u;
    pe.epoch = 1;
    pe.local_base = 0;
    pe.exported = true;
    RTR_CUDA(r, sync_compute(r));
    PeerBlob b;
    std::memset(&b, 0, sizeof(b));
    b.magic = 0x52545251u;
    b.W = r->W; b.H = r->H;
    b.pid = int32_t(getpid());
    b.device = r->device;
    b.has_zkey = r->key64 ? 1 : 0;
    b.zbuf_off = r->set[0].zbuf_off;
    b.image_off = r->set[0].image_off;
    b.raw_flags = pe.flags;
    RTR_CUDA(r, cudaIpcGetMemHandle(&b.flags, pe.flags));
    for (int i = 0; i < 2; ++i) {
        b.raw_arena[i] = r->set[i].arena;
        RTR_CUDA(r, cudaIpcGetMemHandle(&b.arena[i], r->set[i].arena));
        if (r->key64) {
            b.raw_zkey[i] = r->set[i].fb.zkey;
            RTR_CUDA(r, cudaIpcGetMemHandle(&b.zkey[i], r->set[i].fb.zkey));
        }
    }
    std::memset(blob512, 0, 512);
    std::memcpy(blob512, &b, sizeof(b));
    return RTR_OK;
}

int rtr_peer_detach(rtr_renderer* r) {
    if (!r) return RTR_ERR_ARG;
    cudaSetDevice(r->device);
    flush_pending(r);
    sync_compute(r);
    for (void* p : r->peer.opened) cudaIpcCloseMemHandle(p);
    r->peer.opened.clear();
    r->peer.attached = false;
    return RTR_OK;
}

int rtr_peer_attach(rtr_renderer* r, const void* blobs, int rank, int n_ranks) {
    if (!r || !blobs || n_ranks < 1 || n_ranks > kMaxPeers || rank < 0 || rank >= n_ranks) return RTR_ERR_ARG;
    rtr_renderer::Peer& pe = r->peer;
    if (!pe.flags || !pe.exported) return fail(r, RTR_ERR_STATE, "call rtr_peer_export first (every attach needs a fresh export on every rank)");
    RTR_CUDA(r, cudaSetDevice(r->device));
    for (void* p : pe.opened) cudaIpcCloseMemHandle(p);
    pe.opened.clear();
    pe.attached = false;
    std::memset(pe.peer_buf, 0, sizeof(pe.peer_buf));
    for (int p = 0; p < n_ranks; ++p) {
        PeerBlob b;
        std::memcpy(&b, static_cast<const char*>(blobs) + size_t(p) * 512, sizeof(b));
        if (b.magic != 0x52545251u) return fail(r, RTR_ERR_ARG, "bad peer blob");
        if (b.W != r->W || b.H != r->H) return fail(r, RTR_ERR_ARG, "peers render different resolutions");
        void* flags = nullptr;
        void* arena[2] = {nullptr, nullptr};
        void* zkey[2] = {nullptr, nullptr};
        if (p == rank) {
            flags = pe.flags;
            for (int i = 0; i < 2; ++i) { arena[i] = r->set[i].arena; zkey[i] = r->set[i].fb.zkey; }
        } else if (b.pid == int32_t(getpid())) {  // a renderer of this process (one thread per GPU): plain peer access
            if (b.device != r->device) {
                const cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cuda_fail(r, e, "cudaDeviceEnablePeerAccess");
                (void)cudaGetLastError();
            }
            flags = b.raw_flags;
            for (int i = 0; i < 2; ++i) { arena[i] = b.raw_arena[i]; zkey[i] = b.has_zkey ? b.raw_zkey[i] : nullptr; }
        } else {
            auto open = [&](const cudaIpcMemHandle_t& h, void** out) -> cudaError_t {
                cudaError_t e = cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess);
                if (e == cudaSuccess) pe.opened.push_back(*out);
                return e;
            };
            RTR_CUDA(r, open(b.flags, &flags));
            for (int i = 0; i < 2; ++i) {
                RTR_CUDA(r, open(b.arena[i], &arena[i]));
                if (b.has_zkey) RTR_CUDA(r, open(b.zkey[i], &zkey[i]));
            }
        }
        pe.peer_flags[p] = static_cast<uint32_t*>(flags);
        for (int i = 0; i < 2; ++i) {
            char* a = static_cast<char*>(arena[i]);
            pe.peer_buf[1][i][p] = a;                  // colour sums at the start of the arena
            pe.peer_buf[0][i][p] = a + b.zbuf_off;
            pe.peer_buf[3][i][p] = a + b.image_off;
            pe.peer_buf[2][i][p] = zkey[i];
        }
    }
    pe.rank = rank; pe.n = n_ranks; pe.W = r->W; pe.H = r->H;
    pe.attached = n_ranks > 1;
    pe.exported = false;
    r->cur = 0;
    return RTR_OK;
}

int rtr_comm_destroy(rtr_renderer* r) {
    if (!r) return RTR_ERR_ARG;
    if (r->comm) {
        sync_compute(r);
        g_nccl.CommDestroy(r->comm);
        r->comm = nullptr;
    }
    r->rank = 0; r->n_ranks = 1;
    return RTR_OK;
}

}  // extern "C"
