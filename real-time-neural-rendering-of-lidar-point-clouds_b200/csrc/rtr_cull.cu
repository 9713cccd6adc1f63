// Chunk-level frustum culling for the two point passes (sm_100a).
//
// The reference tests every point against the image rectangle inside minDepthPass/accumulatePass
// (render.cu:63-68, 96-101) and therefore streams the whole cloud twice per frame, although a
// camera inside a scan sees a fraction of it.  The reference loader delivers points grouped by
// 0.25 m cell (cloudreader.cpp:47-60, Octreegrid.h:162-170), so kChunkPoints consecutive records
// form a compact box.  At upload we store each chunk's axis-aligned bounds; per frame one thread
// per chunk decides whether ANY point of the box can survive the reference's per-point test and
// appends the survivors to a compact list that the point passes walk.
//
// The test must never drop a chunk holding a point the reference would keep.  For a point p of
// the box, with R0,R1,R2 the camProj rows, the reference computes in float
//     rx_f = R0.p (+-ex)   ry_f = R1.p (+-ey)   rz_f = R2.p (+-ez)
// where e* <= gamma * (sum_j |R_ij| |p_j| + |R_i3|), gamma = 2^-20 (the true bound of the
// mul/fma/fma/add chain is < 2^-22), keeps the point iff rz_f > 0 and
// u = rint(rx_f * rcp(rz_f)) in [0, W) (same for v), and rcp/mul add < 2^-21 relative error.
// A chunk is dropped only if one of these holds for the whole box (interval arithmetic in double):
//     behind : max R2.p + ez < 0                                  -> rz_f < 0 for every point
//     left   : max (R0 + 1.5 R2).p + (ex + 1.5 ez) < 0            -> rx_f/rz_f < -1.5  -> u <= -1
//     right  : min (R0 - (W+0.5) R2).p - (ex + (W+0.5) ez) > 0    -> rx_f/rz_f > W+0.5 -> u >= W
//     top / bottom likewise with R1, H.
// (one extra pixel of margin on each side).  NaN anywhere makes every comparison false, i.e. keeps
// the chunk; chunks holding a non-finite or huge (> 2^40) coordinate are always kept (the reference
// maps NaN depth to pixel 0 and __fdividef returns 0 for |z| > 2^126).  DESIGN.md "chunk culling".
#include "../../include/rtr_b200.h"
#include "rtr_kernels.h"

namespace rtr {

// ---------------------------------------------------------------- bounds of every chunk (at upload)
__global__ void __launch_bounds__(256) chunk_bounds_kernel(const PointRecord* __restrict__ pts, uint64_t n,
                                                           ChunkBounds* __restrict__ bounds) {
    __shared__ float s_lo[3][8], s_hi[3][8];
    __shared__ uint32_t s_bad[8];
    const uint64_t first = uint64_t(blockIdx.x) * kChunkPoints;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    uint32_t bad = 0;
    for (int k = 0; k < kChunkPoints / 256; ++k) {
        const uint64_t i = first + uint64_t(k) * 256 + threadIdx.x;
        if (i >= n) break;
        const PointRecord p = pts[i];
        const float c[3] = {p.x, p.y, p.z};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            if (!(fabsf(c[a]) <= 1.0995116e12f)) bad = 1;  // NaN, inf or > 2^40
            lo[a] = fminf(lo[a], c[a]);
            hi[a] = fmaxf(hi[a], c[a]);
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xFFFFFFFFu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xFFFFFFFFu, hi[a], o));
        }
    }
    bad = __any_sync(0xFFFFFFFFu, bad);
    if (lane == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) { s_lo[a][warp] = lo[a]; s_hi[a][warp] = hi[a]; }
        s_bad[warp] = bad;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        ChunkBounds b;
        b.always_visible = 0;
        b.pad = 0;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            float l = s_lo[a][0], h = s_hi[a][0];
            for (int w = 1; w < 8; ++w) { l = fminf(l, s_lo[a][w]); h = fmaxf(h, s_hi[a][w]); }
            b.lo[a] = l;
            b.hi[a] = h;
        }
        for (int w = 0; w < 8; ++w) b.always_visible |= s_bad[w];
        bounds[blockIdx.x] = b;
    }
}

// ---------------------------------------------------------------- per-frame classification
struct Interval { double lo, hi; };
// range of f.p over the box, f = (f0,f1,f2,f3)
__device__ __forceinline__ Interval box_range(const double f[4], const double lo[3], const double hi[3]) {
    Interval r{f[3], f[3]};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const double u = f[a] * lo[a], v = f[a] * hi[a];
        r.lo += fmin(u, v);
        r.hi += fmax(u, v);
    }
    return r;
}
// gamma * (sum |f_a| max|p_a| + |f3|): bound of the float evaluation error of row f over the box
__device__ __forceinline__ double row_err(const double f[4], const double mabs[3]) {
    return 9.5367431640625e-7 * (fabs(f[0]) * mabs[0] + fabs(f[1]) * mabs[1] + fabs(f[2]) * mabs[2] + fabs(f[3]));
}

// true unless the whole box provably fails the reference's per-point test (see file header)
__device__ __forceinline__ bool chunk_visible(const ChunkBounds& b, const CullParams& cp) {
    if (b.always_visible) return true;
    const double lo[3] = {b.lo[0], b.lo[1], b.lo[2]}, hi[3] = {b.hi[0], b.hi[1], b.hi[2]};
    const double mabs[3] = {fmax(fabs(lo[0]), fabs(hi[0])), fmax(fabs(lo[1]), fabs(hi[1])), fmax(fabs(lo[2]), fabs(hi[2]))};
    const double ex = row_err(cp.r0, mabs), ey = row_err(cp.r1, mabs), ez = row_err(cp.r2, mabs);
    const bool sane = (ex < 1e30) & (ey < 1e30) & (ez < 1e30);  // false for NaN / inf / absurd matrices
    if (!sane) return true;
    const double cl = 1.5, cr = cp.W + 0.5, cb = cp.H + 0.5;
    // five independent plane tests, combined without short-circuit evaluation: a chain of `||` would make each test wait
    // for the previous one's branch, and this kernel is nothing but the latency of its double-precision chain
    double fl[4], fr[4], ft[4], fb[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        fl[k] = cp.r0[k] + cl * cp.r2[k];
        fr[k] = cp.r0[k] - cr * cp.r2[k];
        ft[k] = cp.r1[k] + cl * cp.r2[k];
        fb[k] = cp.r1[k] - cb * cp.r2[k];
    }
    const bool behind = box_range(cp.r2, lo, hi).hi + ez < 0.0;
    const bool left = box_range(fl, lo, hi).hi + (ex + cl * ez) < 0.0;
    const bool right = box_range(fr, lo, hi).lo - (ex + cr * ez) > 0.0;
    const bool top = box_range(ft, lo, hi).hi + (ey + cl * ez) < 0.0;
    const bool bottom = box_range(fb, lo, hi).lo - (ey + cb * ez) > 0.0;
    const bool cut = behind | left | right | top | bottom;
    return !cut;
}

// ---------------------------------------------------------------- screen-band ordering of the list (BandSort, rtr_kernels.h)
// The band a chunk is filed under: the row band of the image its box centre projects into for camera cp.  It only
// orders the list, so float arithmetic and an approximate divide do; centres behind the camera or with a non-finite
// image land in band 0.
__device__ __forceinline__ uint32_t chunk_band(const ChunkBounds& b, const CullParams& cp, uint32_t n_bands) {
    const float cx = 0.5f * (b.lo[0] + b.hi[0]), cy = 0.5f * (b.lo[1] + b.hi[1]), cz = 0.5f * (b.lo[2] + b.hi[2]);
    const float ry = float(cp.r1[0]) * cx + float(cp.r1[1]) * cy + float(cp.r1[2]) * cz + float(cp.r1[3]);
    const float rz = float(cp.r2[0]) * cx + float(cp.r2[1]) * cy + float(cp.r2[2]) * cz + float(cp.r2[3]);
    const float t = __fdividef(ry, rz) * (float(n_bands) / float(cp.H));
    if (!(rz > 0.0f) || !(t >= 0.0f)) return 0u;
    return t >= float(n_bands) ? n_bands - 1u : uint32_t(t);
}
// Warp-wide append of the lanes with keep = true to their bands' segments (every lane of the warp must call this).
__device__ __forceinline__ void band_append(bool keep, uint32_t entry, uint32_t band, CullState* cull, const BandSort& bs) {
    const unsigned grp = __match_any_sync(0xFFFFFFFFu, keep ? band : 0xFFFFFFFFu);  // the lanes filing under the same band
    const int lane = threadIdx.x & 31, leader = __ffs(grp) - 1;
    uint32_t base = 0;
    if (keep && lane == leader) base = atomicAdd(&cull->band_count[band], uint32_t(__popc(grp)));
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    if (keep) bs.scratch[size_t(band) * bs.cap + base + __popc(grp & ((1u << lane) - 1u))] = entry;
}
// Called by every thread of every CTA once its appends are done: the CTA that arrives last copies the segments, band
// after band, into the list the point passes walk, and re-arms the counters for the next classification.
__device__ __forceinline__ void band_compact(uint32_t* __restrict__ vis_list, CullState* cull, const BandSort& bs) {
    __shared__ uint32_t s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&cull->band_ticket, 1u) == gridDim.x - 1u ? 1u : 0u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // One flat loop over the 16-byte vectors of all segments (a loop per band would pay a dependent L2 round trip or two
    // for each band, one after the other, on the next pass's critical path): all eight counts in one go, then every
    // thread finds the band of its vector from the running sums (band_layout / band_locate, rtr_kernels.h) and has four
    // loads in flight.
    const uint4 c03 = __ldcg(reinterpret_cast<const uint4*>(cull->band_count)), c47 = __ldcg(reinterpret_cast<const uint4*>(cull->band_count) + 1);
    const uint32_t cnt[kMaxBands] = {c03.x, c03.y, c03.z, c03.w, c47.x, c47.y, c47.z, c47.w};
    BandLayout L;
    band_layout(cnt, bs.n_bands, L);
    const uint32_t total_v = L.voff[kMaxBands];
    constexpr int kInFlight = 4;
    for (uint32_t v0 = threadIdx.x; v0 < total_v; v0 += blockDim.x * kInFlight) {
        uint4 val[kInFlight];
        uint32_t dst0[kInFlight], left[kInFlight];
#pragma unroll
        for (int u = 0; u < kInFlight; ++u) {
            const uint32_t v = v0 + uint32_t(u) * blockDim.x;
            uint32_t band, lv;
            band_locate(L, cnt, v, band, lv, dst0[u], left[u]);
            if (v >= total_v) left[u] = 0u;
            // cap % 4 == 0: the segments are 16-byte aligned.  Written by other SMs: through L2.
            if (v < total_v) val[u] = __ldcg(reinterpret_cast<const uint4*>(bs.scratch + size_t(band) * bs.cap) + lv);
        }
#pragma unroll
        for (int u = 0; u < kInFlight; ++u) {
            uint32_t* dst = vis_list + dst0[u];
            if (left[u] > 0u) dst[0] = val[u].x;
            if (left[u] > 1u) dst[1] = val[u].y;
            if (left[u] > 2u) dst[2] = val[u].z;
            if (left[u] > 3u) dst[3] = val[u].w;
        }
    }
    __syncthreads();
    if (threadIdx.x < uint32_t(kMaxBands)) cull->band_count[threadIdx.x] = 0u;
    if (threadIdx.x == 0) cull->band_ticket = 0u;
}

// One launch per frame: fillBuffer + cudaMemset (render.cu:16-31, project_cloud.cu:316-317) and the chunk
// classification.  The visible-chunk counter is double-buffered by frame parity so that the reset of one counter
// and the atomic appends to the other need no ordering inside the launch.
//
// MIN_BLOCKS = 4 caps the kernel at 64 registers (the interval arithmetic in double wants 80): with two frames in
// flight this kernel should start while the OTHER frame's ring kernel still holds its 2 x 512 threads x 48 registers
// per SM, which leaves room for exactly 256 threads x 64 registers.
template <int MIN_BLOCKS, bool BANDS>
__global__ void __launch_bounds__(256, MIN_BLOCKS) clear_classify_kernel(uint32_t* __restrict__ zbuf, uint64_t cov,
                                                             uint4* __restrict__ accum, uint64_t n_px,
                                                             uint32_t* __restrict__ minmax,
                                                             const ChunkBounds* __restrict__ bounds, uint32_t n_chunks,
                                                             const __grid_constant__ CullParams cp,
                                                             uint32_t* __restrict__ vis_list, CullState* __restrict__ cull,
                                                             uint32_t parity, const __grid_constant__ BandSort bs) {
    pdl_prologue();
    const uint64_t tid = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
    if (tid == 0) {
        minmax[0] = 0xFFFFFFFFu;
        minmax[1] = 0u;
        minmax[2] = 0u;  // float-accumulator overflow flag of this frame
        minmax[3] = 0u;  // grid-barrier counter of exact_fixup_kernel
        // fold the previous culled frame into the running totals, hand its counter over to the next frame
        if (cull->armed) cull_fold(cull, parity ^ 1u);
        cull->n_visible[parity ^ 1u] = 0u;
        cull->n_blend[0] = cull->n_blend[1] = cull->n_zmin[0] = cull->n_zmin[1] = 0u;  // (two-camera lists only)
        cull->parity = parity;
        cull->kind = kListWhole;
        cull->armed = 1u;
    }
    // the ring kernels' tile-claim counters: every earlier frame of this set has completed (PDL waits are transitive)
    if (tid < uint64_t(2 * kMaxTileQueues)) tile_counters(cull, 0)[tid * kTileQueueStride] = 0u;
    // ---- classification first (its appends are what the next kernel waits for), whole warps at a time
    const uint32_t n_round = (n_chunks + 31u) & ~31u;
    for (uint64_t c = tid; c < n_round; c += stride) {
        ChunkBounds b;
        if (c < n_chunks) b = bounds[c];
        const bool visible = c < n_chunks && chunk_visible(b, cp);
        const unsigned m = __ballot_sync(0xFFFFFFFFu, visible);
        if (m) {
            const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
            uint32_t base = 0;
            if (lane == leader) base = atomicAdd(&cull->n_visible[parity], uint32_t(__popc(m)));
            if constexpr (BANDS) {
                band_append(visible, uint32_t(c), visible ? chunk_band(b, cp, bs.n_bands) : 0u, cull, bs);
            } else {
                base = __shfl_sync(0xFFFFFFFFu, base, leader);
                if (visible) vis_list[base + __popc(m & ((1u << lane) - 1u))] = uint32_t(c);
            }
        }
    }
    if constexpr (BANDS) band_compact(vis_list, cull, bs);  // (the next kernel waits for this; the clear below is bulk work)
    // ---- clear
    for (uint64_t i = tid; i < n_px; i += stride) accum[i] = make_uint4(0u, 0u, 0u, 0u);
    const uint64_t cov4 = cov >> 2;
    uint4* z4 = reinterpret_cast<uint4*>(zbuf);
    for (uint64_t i = tid; i < cov4; i += stride)
        z4[i] = make_uint4(kEmptyDepthBits, kEmptyDepthBits, kEmptyDepthBits, kEmptyDepthBits);
    for (uint64_t i = (cov4 << 2) + tid; i < cov; i += stride) zbuf[i] = kEmptyDepthBits;
}

// Classification for TWO cameras, no clear (the fused point pass, rtr_point_ring.cu): consecutive trajectory poses
// see nearly the same chunks, so frame k-1's blend and frame k's z-min walk ONE list — the union of the two frames'
// visible chunks — and every chunk is streamed from HBM once per frame instead of twice.  An entry carries which of
// the two passes the chunk can contribute to (kTileBlend: camera cp_blend of the previous frame, kTileZmin: camera
// cp_zmin of this frame); each flag is the same conservative test as above, so each pass still sees a superset of
// the chunks the reference's per-point test would keep and the per-point tests inside the pass stay exact.
// 128 threads x <= 64 registers = 8 K registers: what two resident CTAs of the fused point pass (2 x 512 x 56) leave free
// on an SM, so the classification for the NEXT pass runs beside the current one (it is enqueued on the clear stream).
// LATE_WAIT (fused sequences): nothing this kernel reads or writes is touched by the point pass in front of it, so it
// lets its own dependents be scheduled at once, classifies while that pass is still draining, and only THEN waits for
// it — completion stays transitive along the stream (the next pass waits for this kernel, this kernel for the previous
// pass), but the classification no longer sits between two passes.
template <bool LATE_WAIT, bool BANDS>
__global__ void __launch_bounds__(128, 8) classify_pair_kernel(const ChunkBounds* __restrict__ bounds, uint32_t n_chunks,
                                                               const __grid_constant__ CullParams cp_blend,
                                                               const __grid_constant__ CullParams cp_zmin, uint32_t have,
                                                               uint32_t* __restrict__ vis_list, CullState* __restrict__ cull,
                                                               uint32_t parity, const __grid_constant__ BandSort bs) {
    if constexpr (LATE_WAIT) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    else pdl_prologue();
    const uint64_t tid = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
    if (tid == 0) {
        if (cull->armed) cull_fold(cull, parity ^ 1u);
        cull->n_visible[parity ^ 1u] = 0u;
        cull->n_blend[parity ^ 1u] = 0u;
        cull->n_zmin[parity ^ 1u] = 0u;
        cull->parity = parity;
        cull->kind = have;  // kListHasBlend | kListHasZmin
        cull->armed = 1u;
    }
    if (tid < uint64_t(2 * kMaxTileQueues)) tile_counters(cull, 0)[tid * kTileQueueStride] = 0u;
    const uint32_t n_round = (n_chunks + 31u) & ~31u;
    for (uint64_t c = tid; c < n_round; c += stride) {
        uint32_t flags = 0u;
        ChunkBounds b;
        if (c < n_chunks) {
            b = bounds[c];
            if ((have & 1u) && chunk_visible(b, cp_blend)) flags |= kTileBlend;
            if ((have & 2u) && chunk_visible(b, cp_zmin)) flags |= kTileZmin;
        }
        const unsigned m = __ballot_sync(0xFFFFFFFFu, flags != 0u);
        if (m) {
            const unsigned mb = __ballot_sync(0xFFFFFFFFu, (flags & kTileBlend) != 0u), mz = __ballot_sync(0xFFFFFFFFu, (flags & kTileZmin) != 0u);
            const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
            uint32_t base = 0;
            if (lane == leader) {
                base = atomicAdd(&cull->n_visible[parity], uint32_t(__popc(m)));
                if (mb) atomicAdd(&cull->n_blend[parity], uint32_t(__popc(mb)));
                if (mz) atomicAdd(&cull->n_zmin[parity], uint32_t(__popc(mz)));
            }
            if constexpr (BANDS) {
                // (both cameras of a pair are consecutive poses: the newer one's image decides the band)
                band_append(flags != 0u, uint32_t(c) | flags, flags ? chunk_band(b, (have & 2u) ? cp_zmin : cp_blend, bs.n_bands) : 0u, cull, bs);
            } else {
                base = __shfl_sync(0xFFFFFFFFu, base, leader);
                if (flags) vis_list[base + __popc(m & ((1u << lane) - 1u))] = uint32_t(c) | flags;
            }
        }
    }
    if constexpr (BANDS) band_compact(vis_list, cull, bs);
    if constexpr (LATE_WAIT) asm volatile("griddepcontrol.wait;" ::: "memory");
}

}  // namespace rtr
// Host replay of band_compact's copy loop — same layout / locate arithmetic, same thread and in-flight structure — for
// the CPU tests (tests/test_host_logic.py).  scratch: kMaxBands segments of cap entries; counts8: entries per band.
extern "C" int rtr_host_band_compact(const uint32_t* scratch, uint32_t cap, const uint32_t* counts8, uint32_t n_bands, uint32_t threads,
                                     uint32_t* list_out, uint32_t* n_out) {
    using namespace rtr;
    if (!scratch || !counts8 || !list_out || !n_out || (cap & 3u) || n_bands < 1u || n_bands > uint32_t(kMaxBands) || threads < 1u) return RTR_ERR_ARG;
    uint32_t cnt[kMaxBands];
    for (int b = 0; b < kMaxBands; ++b) {
        cnt[b] = counts8[b];
        if (uint32_t(b) < n_bands && cnt[b] > cap) return RTR_ERR_ARG;
    }
    BandLayout L;
    band_layout(cnt, n_bands, L);
    const uint32_t total_v = L.voff[kMaxBands];
    constexpr int kInFlight = 4;
    for (uint32_t t = 0; t < threads; ++t)
        for (uint32_t v0 = t; v0 < total_v; v0 += threads * kInFlight)
            for (int u = 0; u < kInFlight; ++u) {
                const uint32_t v = v0 + uint32_t(u) * threads;
                if (v >= total_v) continue;
                uint32_t band, lv, dst0, left;
                band_locate(L, cnt, v, band, lv, dst0, left);
                for (uint32_t k = 0; k < left; ++k) list_out[dst0 + k] = scratch[size_t(band) * cap + size_t(lv) * 4u + k];
            }
    *n_out = L.eoff[kMaxBands];
    return RTR_OK;
}
namespace rtr {

cudaError_t launch_classify_pair(cudaStream_t s, int sm_count, const ChunkBounds* bounds, uint32_t n_chunks,
                                 const CullParams& cp_blend, bool have_blend, const CullParams& cp_zmin, bool have_zmin,
                                 uint32_t* vis_list, CullState* cull, uint32_t parity, bool late_wait, const BandSort& bands) {
    const uint32_t have = (have_blend ? 1u : 0u) | (have_zmin ? 2u : 0u);
    // one thread per chunk, at most 8 CTAs per SM (the list is short: a few microseconds, latency-bound)
    unsigned grid = (n_chunks + 127u) / 128u;
    const unsigned cap = unsigned(sm_count) * 8u;
    if (grid > cap) grid = cap;
    if (grid < 1u) grid = 1u;
    const bool sort = bands.n_bands > 1u && bands.scratch != nullptr && bands.cap >= n_chunks;
    BandSort bs = bands;
    if (bs.n_bands > uint32_t(kMaxBands)) bs.n_bands = uint32_t(kMaxBands);
    if (sort && late_wait) launch_pdl(classify_pair_kernel<true, true>, dim3(grid), dim3(128), s, bounds, n_chunks, cp_blend, cp_zmin, have, vis_list, cull, parity, bs);
    else if (sort) launch_pdl(classify_pair_kernel<false, true>, dim3(grid), dim3(128), s, bounds, n_chunks, cp_blend, cp_zmin, have, vis_list, cull, parity, bs);
    else if (late_wait) launch_pdl(classify_pair_kernel<true, false>, dim3(grid), dim3(128), s, bounds, n_chunks, cp_blend, cp_zmin, have, vis_list, cull, parity, bs);
    else launch_pdl(classify_pair_kernel<false, false>, dim3(grid), dim3(128), s, bounds, n_chunks, cp_blend, cp_zmin, have, vis_list, cull, parity, bs);
    return cudaGetLastError();
}

cudaError_t launch_chunk_bounds(cudaStream_t s, const PointRecord* pts, uint64_t n, ChunkBounds* bounds) {
    if (n == 0) return cudaSuccess;
    chunk_bounds_kernel<<<unsigned((n + kChunkPoints - 1) / kChunkPoints), 256, 0, s>>>(pts, n, bounds);
    return cudaGetLastError();
}

cudaError_t launch_clear_classify(cudaStream_t s, int sm_count, uint32_t* zbuf, uint64_t cov, uint32_t* accum,
                                  uint64_t n_px, uint32_t* minmax, const ChunkBounds* bounds, uint32_t n_chunks,
                                  const CullParams& cp, uint32_t* vis_list, CullState* cull, uint32_t parity, bool lean,
                                  const BandSort& bands) {
    const bool sort = bands.n_bands > 1u && bands.scratch != nullptr && bands.cap >= n_chunks;
    BandSort bs = bands;
    if (bs.n_bands > uint32_t(kMaxBands)) bs.n_bands = uint32_t(kMaxBands);
    uint4* a4 = reinterpret_cast<uint4*>(accum);
    const dim3 grid(sm_count * 8), block(256);
    if (sort) launch_pdl(clear_classify_kernel<4, true>, grid, block, s, zbuf, cov, a4, n_px, minmax, bounds, n_chunks, cp, vis_list, cull, parity, bs);
    else if (lean) launch_pdl(clear_classify_kernel<4, false>, grid, block, s, zbuf, cov, a4, n_px, minmax, bounds, n_chunks, cp, vis_list, cull, parity, bs);
    else launch_pdl(clear_classify_kernel<3, false>, grid, block, s, zbuf, cov, a4, n_px, minmax, bounds, n_chunks, cp, vis_list, cull, parity, bs);
    return cudaGetLastError();
}

}  // namespace rtr
