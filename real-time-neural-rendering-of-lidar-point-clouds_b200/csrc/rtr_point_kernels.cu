// Point-level kernels of the hot path, hand-written for sm_100a: the per-thread LDG.128 forms.
//
//   clear_kernel   <- fillBuffer (render.cu:16-31) + cudaMemset (project_cloud.cu:316-317)
//   zmin_kernel    <- minDepthPass   (render.cu:53-83)
//   blend_kernel   <- accumulatePass (render.cu:85-130)
//
// These stream EVERY point of the cloud (frames rendered with chunk_cull = 0 — the configuration north_star's
// "16 B/point against the HBM roofline" describes: 6.4 TB/s) and, as zmin_list_kernel / blend_list_kernel, the
// visible-chunk list when option ring = 0.  Culled frames default to the TMA-fed kernels of rtr_point_ring.cu.
//
// Both point passes stream the packed 16-byte {x,y,z,bgra} record with one 128-bit no-allocate
// load per point, UNROLL independent loads in flight per thread, project in registers with the
// camera in the constant bank (kernel parameter), and touch the L2-resident frame buffers only for
// points that survive culling.  HBM-bound by design: 16 B/point/pass.
//
// Z-min atomics.  min is idempotent and monotone, so three exact optimisations are legal:
//  (1) early depth test: a plain (possibly stale) load of the z-buffer value; a point that is not
//      strictly nearer than a value the pixel has already reached can never lower it;
//  (2) warp aggregation (match.any on the pixel id, redux.min on the depth bits) — what the
//      reference does for every in-frustum point, here only for the survivors of (1);
//  (3) red.global.min (no return value).
// Which combination is fastest depends on point order; `variant` selects it at run time
// (rtr_set_option "zmin_variant"), default chosen from measurements.
#include "rtr_kernels.h"

namespace rtr {

// ---------------------------------------------------------------- clear
// zbuf[0, cov) = FLT_MAX bits ; accum[0, 4P) = 0 ; minmax = {UINT_MAX, 0}.
__global__ void __launch_bounds__(256) clear_kernel(uint32_t* __restrict__ zbuf, uint64_t cov,
                                                    uint4* __restrict__ accum, uint64_t n_px,
                                                    uint32_t* __restrict__ minmax, CullState* __restrict__ cull) {
    pdl_prologue();
    const uint64_t tid = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
    if (tid == 0) {
        minmax[0] = 0xFFFFFFFFu;
        minmax[1] = 0u;
        minmax[2] = 0u;  // float-accumulator overflow flag of this frame
        minmax[3] = 0u;  // grid-barrier counter of exact_fixup_kernel
        (void)cull;  // frames rendered without culling leave the culling state alone
    }
    if (accum) {
        for (uint64_t i = tid; i < n_px; i += stride) accum[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    const uint64_t cov4 = cov >> 2;
    uint4* z4 = reinterpret_cast<uint4*>(zbuf);
    for (uint64_t i = tid; i < cov4; i += stride)
        z4[i] = make_uint4(kEmptyDepthBits, kEmptyDepthBits, kEmptyDepthBits, kEmptyDepthBits);
    for (uint64_t i = (cov4 << 2) + tid; i < cov; i += stride) zbuf[i] = kEmptyDepthBits;
}

__global__ void __launch_bounds__(256) clear_accum_gated_kernel(uint4* __restrict__ accum, uint64_t n_px,
                                                                const uint32_t* __restrict__ gate, uint32_t* __restrict__ host_note) {
    pdl_prologue();
    if (*gate == 0u) return;
    // tell the host (mapped pinned word) that float sums overflowed in this view: the next frames start with integer sums
    if (blockIdx.x == 0 && threadIdx.x == 0 && host_note) *reinterpret_cast<volatile uint32_t*>(host_note) = 1u;
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n_px; i += uint64_t(gridDim.x) * blockDim.x)
        accum[i] = make_uint4(0u, 0u, 0u, 0u);
}

// ---------------------------------------------------------------- z-min
// variant bit 0: early depth test   bit 1: warp aggregation   bit 2: early test through L1 (ld.ca)
// bit 3: measurement only, no RED issued   bit 4: measurement only, atomicMin builtin instead of the PTX red
template <int UNROLL, int VARIANT, bool DISTORT, int KEY64>
__device__ __forceinline__ void zmin_tile(const PointRecord* __restrict__ pts, uint64_t n, uint64_t index_base,
                                          const uint64_t base, const ProjParams& pp, uint32_t* __restrict__ zbuf,
                                          unsigned long long* __restrict__ zkey) {
    // phase 1: UNROLL independent 128-bit loads in flight (tail lanes re-read the last record)
    PointRecord p[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
        const uint64_t idx = base + uint64_t(u) * kPointBlock;
        p[u] = ld_point_stream(pts + (idx < n ? idx : n - 1));
    }
    // phase 2: project, branch-free
    uint32_t pix[UNROLL], dbits[UNROLL];
    bool live[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
        float depth;
        live[u] = project<DISTORT>(pp, p[u].x, p[u].y, p[u].z, pix[u], depth) &
                  (base + uint64_t(u) * kPointBlock < n);
        dbits[u] = __float_as_uint(depth);
    }
    if constexpr (KEY64) {
        // north_star's deterministic 64-bit key: (depth bits << 32) | global point index.
        unsigned long long key[UNROLL], cur[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            key[u] = (static_cast<unsigned long long>(dbits[u]) << 32) |
                     static_cast<unsigned long long>(uint32_t(index_base + base + uint64_t(u) * kPointBlock));
            cur[u] = ~0ull;
            if ((VARIANT & 1) && live[u]) cur[u] = (VARIANT & 4) ? __ldca(zkey + pix[u]) : __ldcg(zkey + pix[u]);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
            if (live[u] && key[u] < cur[u] && !(VARIANT & 8)) {
                if constexpr (VARIANT & 16) atomicMin(zkey + pix[u], key[u]);  // measurement: the builtin (ATOMG after the fence)
                else red_min_u64(zkey + pix[u], key[u]);
            }
    } else {
        // phase 3: early depth test, UNROLL gathers in flight
        uint32_t cur[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            cur[u] = 0xFFFFFFFFu;
            if ((VARIANT & 1) && live[u]) cur[u] = (VARIANT & 4) ? __ldca(zbuf + pix[u]) : __ldcg(zbuf + pix[u]);
        }
        // phase 4: RED.MIN for the survivors
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if (live[u] && dbits[u] < cur[u]) {
                if constexpr (VARIANT & 2) {
                    const unsigned same = __match_any_sync(__activemask(), pix[u]);
                    const uint32_t mn = __reduce_min_sync(same, dbits[u]);
                    // one lane per (warp, pixel) group issues the RED
                    const unsigned winners = __ballot_sync(same, dbits[u] == mn) & same;
                    if ((threadIdx.x & 31) == (__ffs(winners) - 1)) red_min_u32(zbuf + pix[u], mn);
                } else if constexpr (VARIANT & 8) {  // measurement only: no RED issued (results are wrong)
                    if (dbits[u] == 0x12345678u && pix[u] == 0xFFFFFFFFu) zbuf[0] = 0u;
                } else if constexpr (VARIANT & 16) {  // measurement: the builtin (ATOMG after the fence)
                    atomicMin(zbuf + pix[u], dbits[u]);
                } else {
                    red_min_u32(zbuf + pix[u], dbits[u]);
                }
            }
        }
    }
}

template <int UNROLL, int VARIANT, bool DISTORT, int KEY64>
__global__ void __launch_bounds__(kPointBlock) zmin_kernel(const PointRecord* __restrict__ pts, uint64_t n,
                                                           uint64_t index_base,
                                                           const __grid_constant__ ProjParams pp,
                                                           uint32_t* __restrict__ zbuf,
                                                           unsigned long long* __restrict__ zkey) {
    pdl_prologue();
    zmin_tile<UNROLL, VARIANT, DISTORT, KEY64>(pts, n, index_base,
                                               uint64_t(blockIdx.x) * (kPointBlock * UNROLL) + threadIdx.x, pp, zbuf, zkey);
}

// Same tile body over the frame's VISIBLE chunks only (chunk-level frustum culling, see
// rtr_cull.cu): a persistent grid walks the compacted chunk list, so nothing is launched, loaded or
// scheduled for the chunks whose bounding box lies outside the frustum.
template <int VARIANT, int KEY64>
__global__ void __launch_bounds__(kPointBlock) zmin_list_kernel(const PointRecord* __restrict__ pts, uint64_t n,
                                                                uint64_t index_base,
                                                                const __grid_constant__ ProjParams pp,
                                                                const CullState* __restrict__ cull,
                                                                const uint32_t* __restrict__ vis_list,
                                                                uint32_t* __restrict__ zbuf,
                                                                unsigned long long* __restrict__ zkey) {
    pdl_prologue();
    const uint32_t n_vis = cull_count(cull);
    for (uint32_t c = blockIdx.x; c < n_vis; c += gridDim.x)
        zmin_tile<kChunkPoints / kPointBlock, VARIANT, false, KEY64>(
            pts, n, index_base, uint64_t(vis_list[c]) * kChunkPoints + threadIdx.x, pp, zbuf, zkey);
}

// ---------------------------------------------------------------- blend (2 cm depth-window colour sums)
// accum[pix] = {sum b, sum g, sum r, count} as 4 x u32 — the reference's layout.  The four u32
// atomicAdds per (warp, pixel) group of the reference become two 64-bit RED.ADDs on the same
// memory: (b | g<<32) and (r | count<<32).  Identical bits as long as no 32-bit channel sum wraps
// (> 16.8 M points in one pixel; the reference wraps silently there, this carries — documented).
// variant bit 1: warp aggregation (match.any + redux.add), as the reference does.
// variant bit 2: float accumulators, one RED.ADD.F32x4 per point (RED issue rate is per lane, not per
//                byte: measured 193 G/s for F32x4 vs 96 G/s for the u64 pair, profiles/).
// `gate`: when not null the kernel runs only if *gate != 0 (the exact re-run after a float overflow).
template <int UNROLL, int VARIANT, bool DISTORT>
__device__ __forceinline__ void blend_tile(const PointRecord* __restrict__ pts, uint64_t n, const uint64_t base,
                                           const ProjParams& pp, const uint32_t* __restrict__ zbuf,
                                           unsigned long long* __restrict__ accum2) {
    PointRecord p[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
        const uint64_t idx = base + uint64_t(u) * kPointBlock;
        p[u] = ld_point_stream(pts + (idx < n ? idx : n - 1));
    }
    uint32_t pix[UNROLL];
    float depth[UNROLL];
    bool live[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
        live[u] = project<DISTORT>(pp, p[u].x, p[u].y, p[u].z, pix[u], depth[u]) &
                  (base + uint64_t(u) * kPointBlock < n);
    uint32_t zmin[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
        zmin[u] = 0u;
        if (live[u]) zmin[u] = __ldg(zbuf + pix[u]);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
        const float lim = __fadd_rn(__uint_as_float(zmin[u]), kDepthWindow);
        if (live[u] && !(depth[u] > lim)) {  // render.cu:106 (NaN depth is accepted, as there)
            uint32_t b = p[u].bgra & 0xFFu, g = (p[u].bgra >> 8) & 0xFFu, r = (p[u].bgra >> 16) & 0xFFu, c = 1u;
            if constexpr (VARIANT & 2) {
                const unsigned same = __match_any_sync(__activemask(), pix[u]);
                b = __reduce_add_sync(same, b);
                g = __reduce_add_sync(same, g);
                r = __reduce_add_sync(same, r);
                c = __popc(same);
                if ((threadIdx.x & 31) != (__ffs(same) - 1)) continue;
            }
            unsigned long long* a = accum2 + uint64_t(pix[u]) * 2;
            if constexpr (VARIANT & 4) {
                // one 16-byte RED.ADD.F32x4 per point: the accumulator holds the same integers as floats,
                // exact while count <= kF32ExactCount (resolve detects anything beyond and the exact
                // passes below are re-run for that frame)
                red_add_f32x4(a, float(b), float(g), float(r), float(c));
            } else {
                red_add_u64(a + 0, static_cast<unsigned long long>(b) | (static_cast<unsigned long long>(g) << 32));
                red_add_u64(a + 1, static_cast<unsigned long long>(r) | (static_cast<unsigned long long>(c) << 32));
            }
        }
    }
}

template <int UNROLL, int VARIANT, bool DISTORT>
__global__ void __launch_bounds__(kPointBlock) blend_kernel(const PointRecord* __restrict__ pts, uint64_t n,
                                                            const __grid_constant__ ProjParams pp,
                                                            const uint32_t* __restrict__ zbuf,
                                                            unsigned long long* __restrict__ accum2,
                                                            const uint32_t* __restrict__ gate) {
    pdl_prologue();
    if (gate && *gate == 0u) return;
    // one tile per CTA in the normal launch; the gated exact re-run uses a small grid and strides
    const uint64_t n_tiles = (n + kPointBlock * UNROLL - 1) / (kPointBlock * UNROLL);
    for (uint64_t t = blockIdx.x; t < n_tiles; t += gridDim.x)
        blend_tile<UNROLL, VARIANT, DISTORT>(pts, n, t * (kPointBlock * UNROLL) + threadIdx.x, pp, zbuf, accum2);
}

template <int VARIANT, bool DISTORT = false>
__global__ void __launch_bounds__(kPointBlock) blend_list_kernel(const PointRecord* __restrict__ pts, uint64_t n,
                                                                 const __grid_constant__ ProjParams pp,
                                                                 const CullState* __restrict__ cull,
                                                                 const uint32_t* __restrict__ vis_list,
                                                                 const uint32_t* __restrict__ zbuf,
                                                                 unsigned long long* __restrict__ accum2,
                                                                 const uint32_t* __restrict__ gate, uint32_t need_flag) {
    pdl_prologue();
    if (gate && *gate == 0u) return;
    const uint32_t n_vis = cull_count(cull);
    for (uint32_t c = blockIdx.x; c < n_vis; c += gridDim.x) {
        const uint32_t entry = vis_list[c];  // two-camera lists (need_flag = kTileBlend): only the chunks flagged for this blend
        if ((entry & need_flag) != need_flag) continue;
        if (DISTORT) blend_tile<kChunkPoints / kPointBlock, VARIANT, true>(pts, n, uint64_t(entry & kTileIdMask) * kChunkPoints + threadIdx.x, pp, zbuf, accum2);
        else blend_tile<kChunkPoints / kPointBlock, VARIANT, false>(pts, n, uint64_t(entry & kTileIdMask) * kChunkPoints + threadIdx.x, pp, zbuf, accum2);
    }
}

// ---------------------------------------------------------------- exact re-run after a float-accumulator overflow
// One launch that almost always returns at once.  When resolve flagged a pixel beyond the exact range of
// the float sums (minmax[2] != 0) it redoes the frame's colour sums with the integer REDs and resolves
// again: clear accum -> blend (exact) -> resolve, separated by a grid-wide barrier.  The grid is 2 CTAs
// per SM and launched with the cooperative attribute, so all CTAs are co-resident and the spin barrier
// cannot deadlock (CTAs of the preceding kernel still draining do not depend on this grid; PDL schedules
// the following kernel only after every CTA here has started).
__device__ __forceinline__ void grid_barrier(uint32_t* counter, uint32_t target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        uint32_t v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        } while (v < target);
    }
    __syncthreads();
}

// The same barrier for a grid that is ONE thread-block cluster: the hardware co-schedules a cluster's CTAs, so no
// cooperative launch (which has to wait until the whole grid fits beside whatever occupies the SMs) is needed.
__device__ __forceinline__ void cluster_barrier() {
    __threadfence();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <bool LIST, bool DISTORT, bool CLUSTER = false>
__global__ void __launch_bounds__(kPointBlock) exact_fixup_kernel(const PointRecord* __restrict__ pts, uint64_t n,
                                                                  const __grid_constant__ ProjParams pp,
                                                                  const CullState* __restrict__ cull,
                                                                  const uint32_t* __restrict__ vis_list,
                                                                  const uint32_t* __restrict__ zbuf,
                                                                  uint4* __restrict__ accum, uint64_t n_px,
                                                                  uint8_t* __restrict__ image, uint64_t cov,
                                                                  uint32_t* __restrict__ minmax,
                                                                  uint32_t* __restrict__ host_note, uint32_t need_flag) {
    pdl_prologue();
    if (minmax[2] == 0u) return;
    const uint64_t tid = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x, stride = uint64_t(gridDim.x) * blockDim.x;
    // tell the host (mapped pinned word, read without any synchronisation) that float sums overflowed in this view:
    // it renders the next frames with integer sums straight away instead of paying for this re-run every frame
    if (tid == 0 && host_note) *reinterpret_cast<volatile uint32_t*>(host_note) = 1u;
    for (uint64_t i = tid; i < n_px; i += stride) accum[i] = make_uint4(0u, 0u, 0u, 0u);
    if constexpr (CLUSTER) cluster_barrier(); else grid_barrier(minmax + 3, gridDim.x);
    unsigned long long* a2 = reinterpret_cast<unsigned long long*>(accum);
    if constexpr (LIST) {
        const uint32_t n_vis = cull_count(cull);
        for (uint32_t c = blockIdx.x; c < n_vis; c += gridDim.x) {
            const uint32_t entry = vis_list[c];  // two-camera lists: only the chunks flagged for this frame's blend
            if ((entry & need_flag) != need_flag) continue;
            blend_tile<kChunkPoints / kPointBlock, 0, DISTORT>(pts, n, uint64_t(entry & kTileIdMask) * kChunkPoints + threadIdx.x, pp, zbuf, a2);
        }
    } else {
        const uint64_t n_tiles = (n + kChunkPoints - 1) / kChunkPoints;
        for (uint64_t t = blockIdx.x; t < n_tiles; t += gridDim.x)
            blend_tile<kChunkPoints / kPointBlock, 0, DISTORT>(pts, n, t * kChunkPoints + threadIdx.x, pp, zbuf, a2);
    }
    if constexpr (CLUSTER) cluster_barrier(); else grid_barrier(minmax + 3, 2u * gridDim.x);
    for (uint64_t id = tid; id < cov; id += stride) {  // resolvePass on the integer sums (render.cu:147-162)
        const uint4 a = __ldcg(accum + id);
        uint8_t b = 0, g = 0, r = 0;
        if (a.w != 0u) { b = uint8_t(a.x / a.w); g = uint8_t(a.y / a.w); r = uint8_t(a.z / a.w); }
        image[id * 3 + 0] = b; image[id * 3 + 1] = g; image[id * 3 + 2] = r;
    }
}

// ---------------------------------------------------------------- per-point projection dump (tests)
// Writes (pix or -1, depth bits) for every point: the "hybrid golden" tap of SURVEY.md §8 c.
template <bool DISTORT>
__global__ void __launch_bounds__(256) project_dump_kernel(const PointRecord* __restrict__ pts, uint64_t n,
                                                           const __grid_constant__ ProjParams pp,
                                                           int32_t* __restrict__ pix_out,
                                                           uint32_t* __restrict__ zbits_out) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const PointRecord p = ld_point_stream(pts + i);
    uint32_t pix = 0;
    float depth = 0.f;
    const bool live = project<DISTORT>(pp, p.x, p.y, p.z, pix, depth);
    pix_out[i] = live ? int32_t(pix) : -1;
    zbits_out[i] = live ? __float_as_uint(depth) : 0u;
}

// ---------------------------------------------------------------- host launchers
static inline unsigned grid_for(uint64_t n, int per_block) { return unsigned((n + per_block - 1) / per_block); }

cudaError_t launch_clear(cudaStream_t s, int sm_count, uint32_t* zbuf, uint64_t cov, uint32_t* accum, uint64_t n_px,
                         uint32_t* minmax, CullState* cull) {
    launch_pdl(clear_kernel, dim3(sm_count * 8), dim3(256), s, zbuf, cov, reinterpret_cast<uint4*>(accum), n_px, minmax, cull);
    return cudaGetLastError();
}

cudaError_t launch_zmin_list(cudaStream_t s, int sm_count, int variant, const PointRecord* pts, uint64_t n,
                             uint64_t index_base, const ProjParams& pp, const CullState* cull, const uint32_t* vis_list,
                             uint32_t* zbuf, unsigned long long* zkey) {
    if (n == 0) return cudaSuccess;
    const unsigned grid = unsigned(sm_count) * 8u;
#define RTR_ZL(V)                                                                                                  \
    do {                                                                                                           \
        if (zkey) launch_pdl((zmin_list_kernel<V, 1>), dim3(grid), dim3(kPointBlock), s, pts, n, index_base, pp, cull, vis_list, zbuf, zkey); \
        else launch_pdl((zmin_list_kernel<V, 0>), dim3(grid), dim3(kPointBlock), s, pts, n, index_base, pp, cull, vis_list, zbuf, zkey);      \
    } while (0)
    switch (variant & 31) {
        case 0: RTR_ZL(0); break;
        case 1: RTR_ZL(1); break;
        case 2: RTR_ZL(2); break;
        case 3: RTR_ZL(3); break;
        case 5: RTR_ZL(5); break;
        case 7: RTR_ZL(7); break;
#ifdef RTR_EXPERIMENTS  // measurement-only kernels (no RED / ATOMG builtin) exist in experiment builds only
        case 8: RTR_ZL(8); break;
        case 9: RTR_ZL(9); break;
        case 21: RTR_ZL(21); break;
#endif
        default: return cudaErrorInvalidValue;
    }
#undef RTR_ZL
    return cudaGetLastError();
}

cudaError_t launch_blend_list(cudaStream_t s, int sm_count, int variant, const PointRecord* pts, uint64_t n,
                              const ProjParams& pp, const CullState* cull, const uint32_t* vis_list, const uint32_t* zbuf,
                              uint32_t* accum, const uint32_t* gate, uint32_t need_flag) {
    if (n == 0) return cudaSuccess;
    const unsigned grid = unsigned(sm_count) * (gate ? 2u : 8u);  // the gated re-run almost always returns at once
    unsigned long long* a2 = reinterpret_cast<unsigned long long*>(accum);
    if (pp.distort) {  // only the gated exact re-run of a fused sequence comes here with a distorted camera: integer sums
        launch_pdl((blend_list_kernel<0, true>), dim3(grid), dim3(kPointBlock), s, pts, n, pp, cull, vis_list, zbuf, a2, gate, need_flag);
        return cudaGetLastError();
    }
    switch (variant & 6) {
        case 0: launch_pdl((blend_list_kernel<0>), dim3(grid), dim3(kPointBlock), s, pts, n, pp, cull, vis_list, zbuf, a2, gate, need_flag); break;
        case 2: launch_pdl((blend_list_kernel<2>), dim3(grid), dim3(kPointBlock), s, pts, n, pp, cull, vis_list, zbuf, a2, gate, need_flag); break;
        case 4: launch_pdl((blend_list_kernel<4>), dim3(grid), dim3(kPointBlock), s, pts, n, pp, cull, vis_list, zbuf, a2, gate, need_flag); break;
        default: launch_pdl((blend_list_kernel<6>), dim3(grid), dim3(kPointBlock), s, pts, n, pp, cull, vis_list, zbuf, a2, gate, need_flag); break;
    }
    return cudaGetLastError();
}

cudaError_t launch_exact_fixup(cudaStream_t s, int sm_count, const PointRecord* pts, uint64_t n, const ProjParams& pp,
                               const CullState* cull, const uint32_t* vis_list, const uint32_t* zbuf, uint32_t* accum,
                               uint64_t n_px, uint8_t* image, uint64_t cov, uint32_t* minmax, uint32_t* host_note,
                               uint32_t need_flag, unsigned grid_ctas) {
    if (n == 0) return cudaSuccess;
    uint4* a4 = reinterpret_cast<uint4*>(accum);
    if (grid_ctas == kFixupClusterCtas && cull) {
        // fused sequences: ONE cluster of 8 CTAs (barrier.cluster between the phases).  The launch returns at once unless a pixel
        // overflowed; when it has to work it is slow (8 SMs redo the frame's blend: of the order of a millisecond) — once, after
        // which the renderer starts the next 64 frames of that view with integer sums.
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(kFixupClusterCtas);
        cfg.blockDim = dim3(kPointBlock);
        cfg.stream = s;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = kFixupClusterCtas; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = pdl_enabled() ? 2 : 1;
        if (pp.distort) return cudaLaunchKernelEx(&cfg, exact_fixup_kernel<true, true, true>, pts, n, pp, cull, vis_list, zbuf, a4, n_px, image, cov, minmax, host_note, need_flag);
        return cudaLaunchKernelEx(&cfg, exact_fixup_kernel<true, false, true>, pts, n, pp, cull, vis_list, zbuf, a4, n_px, image, cov, minmax, host_note, need_flag);
    }
    const dim3 grid(grid_ctas ? grid_ctas : unsigned(sm_count) * 2u), block(kPointBlock);
    if (cull && pp.distort) return launch_pdl_cooperative((exact_fixup_kernel<true, true>), grid, block, s, pts, n, pp, cull, vis_list, zbuf, a4, n_px, image, cov, minmax, host_note, need_flag);
    if (cull) return launch_pdl_cooperative((exact_fixup_kernel<true, false>), grid, block, s, pts, n, pp, cull, vis_list, zbuf, a4, n_px, image, cov, minmax, host_note, need_flag);
    if (pp.distort) return launch_pdl_cooperative((exact_fixup_kernel<false, true>), grid, block, s, pts, n, pp, cull, vis_list, zbuf, a4, n_px, image, cov, minmax, host_note, need_flag);
    return launch_pdl_cooperative((exact_fixup_kernel<false, false>), grid, block, s, pts, n, pp, cull, vis_list, zbuf, a4, n_px, image, cov, minmax, host_note, need_flag);
}

cudaError_t launch_clear_accum_gated(cudaStream_t s, int sm_count, uint32_t* accum, uint64_t n_px, const uint32_t* gate, uint32_t* host_note) {
    launch_pdl(clear_accum_gated_kernel, dim3(sm_count * 2), dim3(256), s, reinterpret_cast<uint4*>(accum), n_px, gate, host_note);
    return cudaGetLastError();
}

template <int UNROLL, int VARIANT>
static cudaError_t launch_zmin_uv(cudaStream_t s, const PointRecord* pts, uint64_t n, uint64_t index_base,
                                  const ProjParams& pp, uint32_t* zbuf, unsigned long long* zkey) {
    const unsigned grid = grid_for(n, kPointBlock * UNROLL);
    if (zkey) {
        if (pp.distort) launch_pdl((zmin_kernel<UNROLL, VARIANT, true, 1>), dim3(grid), dim3(kPointBlock), s, pts, n, index_base, pp, zbuf, zkey);
        else launch_pdl((zmin_kernel<UNROLL, VARIANT, false, 1>), dim3(grid), dim3(kPointBlock), s, pts, n, index_base, pp, zbuf, zkey);
    } else {
        if (pp.distort) launch_pdl((zmin_kernel<UNROLL, VARIANT, true, 0>), dim3(grid), dim3(kPointBlock), s, pts, n, index_base, pp, zbuf, zkey);
        else launch_pdl((zmin_kernel<UNROLL, VARIANT, false, 0>), dim3(grid), dim3(kPointBlock), s, pts, n, index_base, pp, zbuf, zkey);
    }
    return cudaGetLastError();
}

template <int UNROLL>
static cudaError_t launch_zmin_u(cudaStream_t s, int variant, const PointRecord* pts, uint64_t n, uint64_t index_base,
                                 const ProjParams& pp, uint32_t* zbuf, unsigned long long* zkey) {
    switch (variant & 15) {
#ifdef RTR_EXPERIMENTS
        case 8: return launch_zmin_uv<UNROLL, 8>(s, pts, n, index_base, pp, zbuf, zkey);
#endif
        case 0: return launch_zmin_uv<UNROLL, 0>(s, pts, n, index_base, pp, zbuf, zkey);
        case 1: return launch_zmin_uv<UNROLL, 1>(s, pts, n, index_base, pp, zbuf, zkey);
        case 2: return launch_zmin_uv<UNROLL, 2>(s, pts, n, index_base, pp, zbuf, zkey);
        case 3: return launch_zmin_uv<UNROLL, 3>(s, pts, n, index_base, pp, zbuf, zkey);
        case 5: return launch_zmin_uv<UNROLL, 5>(s, pts, n, index_base, pp, zbuf, zkey);
        case 7: return launch_zmin_uv<UNROLL, 7>(s, pts, n, index_base, pp, zbuf, zkey);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_zmin(cudaStream_t s, int variant, int unroll, const PointRecord* pts, uint64_t n,
                        uint64_t index_base, const ProjParams& pp, uint32_t* zbuf, unsigned long long* zkey) {
    if (n == 0) return cudaSuccess;
    switch (unroll) {
        case 1: return launch_zmin_u<1>(s, variant, pts, n, index_base, pp, zbuf, zkey);
        case 2: return launch_zmin_u<2>(s, variant, pts, n, index_base, pp, zbuf, zkey);
        case 4: return launch_zmin_u<4>(s, variant, pts, n, index_base, pp, zbuf, zkey);
        case 8: return launch_zmin_u<8>(s, variant, pts, n, index_base, pp, zbuf, zkey);
        default: return cudaErrorInvalidValue;
    }
}

template <int UNROLL, int VARIANT>
static cudaError_t launch_blend_uv(cudaStream_t s, const PointRecord* pts, uint64_t n, const ProjParams& pp,
                                   const uint32_t* zbuf, uint32_t* accum, const uint32_t* gate) {
    const unsigned grid = gate ? 148u * 2u : grid_for(n, kPointBlock * UNROLL);
    unsigned long long* a2 = reinterpret_cast<unsigned long long*>(accum);
    if (pp.distort) launch_pdl((blend_kernel<UNROLL, VARIANT, true>), dim3(grid), dim3(kPointBlock), s, pts, n, pp, zbuf, a2, gate);
    else launch_pdl((blend_kernel<UNROLL, VARIANT, false>), dim3(grid), dim3(kPointBlock), s, pts, n, pp, zbuf, a2, gate);
    return cudaGetLastError();
}

template <int UNROLL>
static cudaError_t launch_blend_u(cudaStream_t s, int variant, const PointRecord* pts, uint64_t n, const ProjParams& pp,
                                  const uint32_t* zbuf, uint32_t* accum, const uint32_t* gate) {
    switch (variant & 6) {
        case 0: return launch_blend_uv<UNROLL, 0>(s, pts, n, pp, zbuf, accum, gate);
        case 2: return launch_blend_uv<UNROLL, 2>(s, pts, n, pp, zbuf, accum, gate);
        case 4: return launch_blend_uv<UNROLL, 4>(s, pts, n, pp, zbuf, accum, gate);
        default: return launch_blend_uv<UNROLL, 6>(s, pts, n, pp, zbuf, accum, gate);
    }
}

cudaError_t launch_blend(cudaStream_t s, int variant, int unroll, const PointRecord* pts, uint64_t n,
                         const ProjParams& pp, const uint32_t* zbuf, uint32_t* accum, const uint32_t* gate) {
    if (n == 0) return cudaSuccess;
    switch (unroll) {
        case 1: return launch_blend_u<1>(s, variant, pts, n, pp, zbuf, accum, gate);
        case 2: return launch_blend_u<2>(s, variant, pts, n, pp, zbuf, accum, gate);
        case 4: return launch_blend_u<4>(s, variant, pts, n, pp, zbuf, accum, gate);
        case 8: return launch_blend_u<8>(s, variant, pts, n, pp, zbuf, accum, gate);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_project_dump(cudaStream_t s, const PointRecord* pts, uint64_t n, const ProjParams& pp,
                                int32_t* pix_out, uint32_t* zbits_out) {
    if (n == 0) return cudaSuccess;
    const unsigned grid = grid_for(n, 256);
    if (pp.distort) project_dump_kernel<true><<<grid, 256, 0, s>>>(pts, n, pp, pix_out, zbits_out);
    else project_dump_kernel<false><<<grid, 256, 0, s>>>(pts, n, pp, pix_out, zbits_out);
    return cudaGetLastError();
}

}  // namespace rtr
