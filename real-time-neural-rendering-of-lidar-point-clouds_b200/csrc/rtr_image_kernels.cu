// Image-space kernels of the hot path, hand-written for sm_100a.
//
// The reference runs 1 resolve + 18 filter launches with a cudaMalloc/cudaFree pair and a
// cudaDeviceSynchronize around nearly every one (project_cloud.cu:325, 331-392).  Here:
//
//   resolve_pyramid_kernel   resolvePass (render.cu:132-163)
//                          + reduce x4   (project_cloud.cu:28-53)
//                          + find_local/overall_minmax (render.cu:168-240)          -> 1 launch
//   up_level_kernel<false>   laplacianKernel + compareImgsKernel + resizeKernel
//                            (project_cloud.cu:55-161), levels 4->3, 3->2, 2->1      -> 3 launches
//   up_level_kernel<true>    laplacianKernel + compareImgsKernel + removeMask
//                            (project_cloud.cu:55-126, 163-187), level 1->0          -> 1 launch
//
// All on persistent scratch, no allocation, no host sync.  Arithmetic follows the reference's
// compiled op order (see rtr_common.cuh header and DESIGN.md); indexing follows its FLAT,
// truncated-dims semantics, so 1920x1080 behaves exactly as the reference does
// under zero-initialised buffers.  The fused resolve/pyramid kernel needs W % 16 == 0 (then every
// level width is an exact half and the flat pyramid is a true 2-D pyramid); other widths take the
// generic one-kernel-per-reference-kernel path, which is also the cross-check in the tests.
#include <cuda_fp16.h>

#include <cstdlib>

#include "rtr_kernels.h"

#ifndef RTR_L2_HINTS
#define RTR_L2_HINTS 0
#endif
// resident CTAs per SM the two per-frame image kernels are compiled for (register caps 65536 / 256 / n): build knobs of the
// A/B in profiles/r02Q_exp_image_occupancy.json
#ifdef RTR_RESOLVE_MIN_CTAS
#define RTR_RESOLVE_BOUNDS __launch_bounds__(256, RTR_RESOLVE_MIN_CTAS)
#else
#define RTR_RESOLVE_BOUNDS __launch_bounds__(256)
#endif
#ifndef RTR_UP_MIN_CTAS
#define RTR_UP_MIN_CTAS 4
#endif
// 1: the fused resolve + pyramid kernel as a persistent grid (5 CTAs per SM) whose CTAs walk the tiles and fetch the NEXT
// tile's z-buffer / colour-sum words with cp.async while they work on the current one (A/B in
// profiles/r02X_exp_resolve_persist.json)
#ifndef RTR_RESOLVE_PERSIST
#define RTR_RESOLVE_PERSIST 0
#endif

namespace rtr {

__device__ __forceinline__ float sel_min(float a, float b) { return a < b ? a : b; }  // project_cloud.cu:46-49

__device__ __forceinline__ void resolve_px(const uint4 a, uint8_t& b, uint8_t& g, uint8_t& r) {
    if (a.w == 0u) { b = g = r = 0; return; }  // render.cu:147-154
    b = uint8_t(a.x / a.w);
    g = uint8_t(a.y / a.w);
    r = uint8_t(a.z / a.w);
}

// ---------------------------------------------------------------- fused resolve + pyramid + min/max
// One CTA = 256 threads = a 64 x 16 tile of level 0 = 32 x 8 of L1 = 16 x 4 of L2 = 8 x 2 of L3 =
// 4 x 1 of L4.  Thread (tx, ty) owns the 2x2 quad of level-0 pixels under L1 pixel (tx, ty).
// Tiles cover rows [0, 16*floor(H/16)): exactly resolvePass's / the min-max's coverage when
// W % 16 == 0, and every pyramid row the up-pass can read (uh[i] = 16*floor(H/16) >> i).
// float accumulator -> the integer sums it holds; flags counts beyond the exact range
__device__ __forceinline__ uint4 accum_as_u32(const uint4 raw, bool f32acc, uint32_t* __restrict__ overflow) {
    if (!f32acc) return raw;
    const float c = __uint_as_float(raw.w);
    if (c > kF32ExactCount) *overflow = 1u;
    return make_uint4(__float2uint_rz(__uint_as_float(raw.x)), __float2uint_rz(__uint_as_float(raw.y)),
                      __float2uint_rz(__uint_as_float(raw.z)), __float2uint_rz(c));
}

// resolvePass on float accumulators: floor(sum / count) per channel without the three 32-bit integer divisions.
// sum and count are exact integers (sum < 2^24, count <= kF32ExactCount, else the frame is flagged and redone with
// integer sums), so q = floor(sum * rcp(count)) is off by at most one (|error| <= 255 * 2^-22) and the remainder
// sum - q * count, exact in one FMA, says which way.
__device__ __forceinline__ uint8_t floor_div_exact(float s, float c, float rc) {
    float q = floorf(__fmul_rn(s, rc));
    const float rem = __fmaf_rn(-q, c, s);
    if (rem < 0.0f) q -= 1.0f;
    else if (rem >= c) q += 1.0f;
    return uint8_t(__float2uint_rz(q));
}
__device__ __forceinline__ void resolve_px_f32(const uint4 raw, uint32_t* __restrict__ overflow, uint8_t& b, uint8_t& g, uint8_t& r) {
    const float c = __uint_as_float(raw.w);
    if (c > kF32ExactCount) *overflow = 1u;
    if (c == 0.0f) { b = g = r = 0; return; }  // render.cu:147-154
    const float rc = __frcp_rn(c);
    b = floor_div_exact(__uint_as_float(raw.x), c, rc);
    g = floor_div_exact(__uint_as_float(raw.y), c, rc);
    r = floor_div_exact(__uint_as_float(raw.z), c, rc);
}

template <bool PYRAMID, bool RESOLVE, bool F32ACC>
__global__ void RTR_RESOLVE_BOUNDS resolve_pyramid_kernel(const uint32_t* __restrict__ zbuf,
                                                              const uint4* __restrict__ accum,
                                                              uint8_t* __restrict__ image, float* __restrict__ l1,
                                                              float* __restrict__ l2, float* __restrict__ l3,
                                                              float* __restrict__ l4, uint32_t* __restrict__ minmax,
                                                              int W) {
    pdl_prologue();
    __shared__ float s1[8][33];
    __shared__ float s2[4][17];
    __shared__ float s3[2][9];
    __shared__ uint32_t smin[8], smax[8];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int X1 = blockIdx.x * 32 + tx, Y1 = blockIdx.y * 8 + ty;
    const int w1 = W >> 1;
    const bool in = X1 < w1;
    float v1 = __uint_as_float(kEmptyDepthBits);
    uint32_t tmin = 0xFFFFFFFFu, tmax = 0u;
    if (in) {
        const size_t p0 = size_t(2 * Y1) * W + 2 * X1, p1 = p0 + W;
        const uint2 z0 = *reinterpret_cast<const uint2*>(zbuf + p0);
        const uint2 z1 = *reinterpret_cast<const uint2*>(zbuf + p1);
        if constexpr (RESOLVE) {
#if RTR_L2_HINTS & 2
            const uint4 a00 = __ldcs(accum + p0), a01 = __ldcs(accum + p0 + 1), a10 = __ldcs(accum + p1), a11 = __ldcs(accum + p1 + 1);  // last use of the sums
#else
            const uint4 a00 = accum[p0], a01 = accum[p0 + 1], a10 = accum[p1], a11 = accum[p1 + 1];
#endif
            uint8_t c[12];
            if constexpr (F32ACC) {
                resolve_px_f32(a00, minmax + 2, c[0], c[1], c[2]);
                resolve_px_f32(a01, minmax + 2, c[3], c[4], c[5]);
                resolve_px_f32(a10, minmax + 2, c[6], c[7], c[8]);
                resolve_px_f32(a11, minmax + 2, c[9], c[10], c[11]);
            } else {
                resolve_px(a00, c[0], c[1], c[2]);
                resolve_px(a01, c[3], c[4], c[5]);
                resolve_px(a10, c[6], c[7], c[8]);
                resolve_px(a11, c[9], c[10], c[11]);
            }
            uint16_t* o0 = reinterpret_cast<uint16_t*>(image + p0 * 3);  // p0 is even -> 2-byte aligned
            uint16_t* o1 = reinterpret_cast<uint16_t*>(image + p1 * 3);
            o0[0] = uint16_t(c[0] | (c[1] << 8)); o0[1] = uint16_t(c[2] | (c[3] << 8)); o0[2] = uint16_t(c[4] | (c[5] << 8));
            o1[0] = uint16_t(c[6] | (c[7] << 8)); o1[1] = uint16_t(c[8] | (c[9] << 8)); o1[2] = uint16_t(c[10] | (c[11] << 8));
        }
        if constexpr (PYRAMID) {
            const uint32_t zz[4] = {z0.x, z0.y, z1.x, z1.y};
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (zz[k] != kEmptyDepthBits) { tmin = min(tmin, zz[k]); tmax = max(tmax, zz[k]); }  // render.cu:181-184
            v1 = sel_min(sel_min(__uint_as_float(z0.x), __uint_as_float(z0.y)),
                         sel_min(__uint_as_float(z1.x), __uint_as_float(z1.y)));
            l1[size_t(Y1) * w1 + X1] = v1;
        }
    }
    if constexpr (!PYRAMID) return;
    s1[ty][tx] = v1;
    tmin = __reduce_min_sync(0xFFFFFFFFu, tmin);
    tmax = __reduce_max_sync(0xFFFFFFFFu, tmax);
    if (tx == 0) { smin[ty] = tmin; smax[ty] = tmax; }
    __syncthreads();
    const int t = threadIdx.x;
    if (t < 64) {
        const int x = t & 15, y = t >> 4;
        const float v = sel_min(sel_min(s1[2 * y][2 * x], s1[2 * y][2 * x + 1]), sel_min(s1[2 * y + 1][2 * x], s1[2 * y + 1][2 * x + 1]));
        s2[y][x] = v;
        const int X = blockIdx.x * 16 + x, w2 = W >> 2;
        if (X < w2) l2[size_t(blockIdx.y * 4 + y) * w2 + X] = v;
    }
    if (t == 64) {
        uint32_t mn = smin[0], mx = smax[0];
#pragma unroll
        for (int k = 1; k < 8; ++k) { mn = min(mn, smin[k]); mx = max(mx, smax[k]); }
        if (mn != 0xFFFFFFFFu) {  // at least one valid pixel in the tile
            atomicMin(minmax + 0, mn);
            atomicMax(minmax + 1, mx);
        }
    }
    __syncthreads();
    if (t < 16) {
        const int x = t & 7, y = t >> 3;
        const float v = sel_min(sel_min(s2[2 * y][2 * x], s2[2 * y][2 * x + 1]), sel_min(s2[2 * y + 1][2 * x], s2[2 * y + 1][2 * x + 1]));
        s3[y][x] = v;
        const int X = blockIdx.x * 8 + x, w3 = W >> 3;
        if (X < w3) l3[size_t(blockIdx.y * 2 + y) * w3 + X] = v;
    }
    __syncthreads();
    if (t < 4) {
        const float v = sel_min(sel_min(s3[0][2 * t], s3[0][2 * t + 1]), sel_min(s3[1][2 * t], s3[1][2 * t + 1]));
        const int X = blockIdx.x * 4 + t, w4 = W >> 4;
        if (X < w4) l4[size_t(blockIdx.y) * w4 + X] = v;
    }
}

// ---------------------------------------------------------------- the same, persistent with prefetch (RTR_RESOLVE_PERSIST)
// Same tile, same thread -> pixel mapping and the same arithmetic as resolve_pyramid_kernel<true, true, F32ACC>; a CTA
// handles tiles blockIdx.x, blockIdx.x + gridDim.x, ... and every thread copies the 80 bytes it will need of the next
// tile (two z-buffer pairs, four accumulators) into its own shared-memory slots with cp.async before it starts on the
// current tile, so the load latency of all but a CTA's first tile is hidden behind arithmetic.
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(uint32_t(__cvta_generic_to_shared(dst_smem))), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst_smem, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(uint32_t(__cvta_generic_to_shared(dst_smem))), "l"(src) : "memory");
}
template <bool F32ACC>
__global__ void __launch_bounds__(256) resolve_pyramid_persist_kernel(const uint32_t* __restrict__ zbuf,
                                                                      const uint4* __restrict__ accum,
                                                                      uint8_t* __restrict__ image, float* __restrict__ l1,
                                                                      float* __restrict__ l2, float* __restrict__ l3,
                                                                      float* __restrict__ l4, uint32_t* __restrict__ minmax,
                                                                      int W, int tiles_x, int n_tiles) {
    __shared__ __align__(16) uint4 sa[2][4][256];   // [buffer][a00, a01, a10, a11][thread]
    __shared__ __align__(8) uint2 sz[2][2][256];    // [buffer][row 0, row 1][thread]
    __shared__ float s1[8][33];
    __shared__ float s2[4][17];
    __shared__ float s3[2][9];
    __shared__ uint32_t smin[8], smax[8];
    pdl_prologue();
    const int t = threadIdx.x, tx = t & 31, ty = t >> 5;
    const int w1 = W >> 1;
    auto fetch = [&](int tile, int buf) {
        const int bx = tile % tiles_x, by = tile / tiles_x;
        const int X1 = bx * 32 + tx, Y1 = by * 8 + ty;
        if (X1 < w1) {
            const size_t p0 = size_t(2 * Y1) * W + 2 * X1, p1 = p0 + W;
            cp_async8(&sz[buf][0][t], zbuf + p0);
            cp_async8(&sz[buf][1][t], zbuf + p1);
            cp_async16(&sa[buf][0][t], accum + p0);
            cp_async16(&sa[buf][1][t], accum + p0 + 1);
            cp_async16(&sa[buf][2][t], accum + p1);
            cp_async16(&sa[buf][3][t], accum + p1 + 1);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int tile = blockIdx.x, buf = 0;
    if (tile < n_tiles) fetch(tile, 0);
    for (; tile < n_tiles; tile += gridDim.x, buf ^= 1) {
        const int next = tile + int(gridDim.x);
        if (next < n_tiles) {
            fetch(next, buf ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        const int bx = tile % tiles_x, by = tile / tiles_x;
        const int X1 = bx * 32 + tx, Y1 = by * 8 + ty;
        const bool in = X1 < w1;
        float v1 = __uint_as_float(kEmptyDepthBits);
        uint32_t tmin = 0xFFFFFFFFu, tmax = 0u;
        if (in) {
            const size_t p0 = size_t(2 * Y1) * W + 2 * X1, p1 = p0 + W;
            const uint2 z0 = sz[buf][0][t], z1 = sz[buf][1][t];
            const uint4 a00 = sa[buf][0][t], a01 = sa[buf][1][t], a10 = sa[buf][2][t], a11 = sa[buf][3][t];
            uint8_t c[12];
            if constexpr (F32ACC) {
                resolve_px_f32(a00, minmax + 2, c[0], c[1], c[2]);
                resolve_px_f32(a01, minmax + 2, c[3], c[4], c[5]);
                resolve_px_f32(a10, minmax + 2, c[6], c[7], c[8]);
                resolve_px_f32(a11, minmax + 2, c[9], c[10], c[11]);
            } else {
                resolve_px(a00, c[0], c[1], c[2]);
                resolve_px(a01, c[3], c[4], c[5]);
                resolve_px(a10, c[6], c[7], c[8]);
                resolve_px(a11, c[9], c[10], c[11]);
            }
            uint16_t* o0 = reinterpret_cast<uint16_t*>(image + p0 * 3);  // p0 is even -> 2-byte aligned
            uint16_t* o1 = reinterpret_cast<uint16_t*>(image + p1 * 3);
            o0[0] = uint16_t(c[0] | (c[1] << 8)); o0[1] = uint16_t(c[2] | (c[3] << 8)); o0[2] = uint16_t(c[4] | (c[5] << 8));
            o1[0] = uint16_t(c[6] | (c[7] << 8)); o1[1] = uint16_t(c[8] | (c[9] << 8)); o1[2] = uint16_t(c[10] | (c[11] << 8));
            const uint32_t zz[4] = {z0.x, z0.y, z1.x, z1.y};
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (zz[k] != kEmptyDepthBits) { tmin = min(tmin, zz[k]); tmax = max(tmax, zz[k]); }  // render.cu:181-184
            v1 = sel_min(sel_min(__uint_as_float(z0.x), __uint_as_float(z0.y)), sel_min(__uint_as_float(z1.x), __uint_as_float(z1.y)));
            l1[size_t(Y1) * w1 + X1] = v1;
        }
        s1[ty][tx] = v1;
        tmin = __reduce_min_sync(0xFFFFFFFFu, tmin);
        tmax = __reduce_max_sync(0xFFFFFFFFu, tmax);
        if (tx == 0) { smin[ty] = tmin; smax[ty] = tmax; }
        __syncthreads();
        if (t < 64) {
            const int x = t & 15, y = t >> 4;
            const float v = sel_min(sel_min(s1[2 * y][2 * x], s1[2 * y][2 * x + 1]), sel_min(s1[2 * y + 1][2 * x], s1[2 * y + 1][2 * x + 1]));
            s2[y][x] = v;
            const int X = bx * 16 + x, w2 = W >> 2;
            if (X < w2) l2[size_t(by * 4 + y) * w2 + X] = v;
        }
        if (t == 64) {
            uint32_t mn = smin[0], mx = smax[0];
#pragma unroll
            for (int k = 1; k < 8; ++k) { mn = min(mn, smin[k]); mx = max(mx, smax[k]); }
            if (mn != 0xFFFFFFFFu) {  // at least one valid pixel in the tile
                atomicMin(minmax + 0, mn);
                atomicMax(minmax + 1, mx);
            }
        }
        __syncthreads();
        if (t < 16) {
            const int x = t & 7, y = t >> 3;
            const float v = sel_min(sel_min(s2[2 * y][2 * x], s2[2 * y][2 * x + 1]), sel_min(s2[2 * y + 1][2 * x], s2[2 * y + 1][2 * x + 1]));
            s3[y][x] = v;
            const int X = bx * 8 + x, w3 = W >> 3;
            if (X < w3) l3[size_t(by * 2 + y) * w3 + X] = v;
        }
        __syncthreads();
        if (t < 4) {
            const float v = sel_min(sel_min(s3[0][2 * t], s3[0][2 * t + 1]), sel_min(s3[1][2 * t], s3[1][2 * t + 1]));
            const int X = bx * 4 + t, w4 = W >> 4;
            if (X < w4) l4[size_t(by) * w4 + X] = v;
        }
    }
}

// ---------------------------------------------------------------- generic (any W, H) restatements
__global__ void __launch_bounds__(256) resolve_generic_kernel(const uint4* __restrict__ accum,
                                                              uint8_t* __restrict__ image, uint64_t cov, bool f32acc,
                                                              uint32_t* __restrict__ overflow, bool gated) {
    pdl_prologue();
    if (gated && *overflow == 0u) return;
    for (uint64_t id = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; id < cov; id += uint64_t(gridDim.x) * blockDim.x) {
        uint8_t b, g, r;
        resolve_px(accum_as_u32(accum[id], f32acc, overflow), b, g, r);
        image[id * 3 + 0] = b; image[id * 3 + 1] = g; image[id * 3 + 2] = r;
    }
}

__global__ void __launch_bounds__(256) minmax_generic_kernel(const uint32_t* __restrict__ zbuf, uint64_t count,
                                                             uint32_t* __restrict__ minmax) {
    pdl_prologue();
    __shared__ uint32_t smin[8], smax[8];
    uint32_t tmin = 0xFFFFFFFFu, tmax = 0u;
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += uint64_t(gridDim.x) * blockDim.x) {
        const uint32_t v = zbuf[i];
        if (v != kEmptyDepthBits) { tmin = min(tmin, v); tmax = max(tmax, v); }
    }
    tmin = __reduce_min_sync(0xFFFFFFFFu, tmin);
    tmax = __reduce_max_sync(0xFFFFFFFFu, tmax);
    if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = tmin; smax[threadIdx.x >> 5] = tmax; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; ++k) { tmin = min(tmin, smin[k]); tmax = max(tmax, smax[k]); }
        if (tmin != 0xFFFFFFFFu) { atomicMin(minmax + 0, tmin); atomicMax(minmax + 1, tmax); }
    }
}

__global__ void __launch_bounds__(256) reduce_generic_kernel(const float* __restrict__ hi, float* __restrict__ lo,
                                                             int w, int h) {
    pdl_prologue();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= w * h) return;
    const int x = idx % w, y = idx / w, wh = w * 2;
    const float* r0 = hi + size_t(2 * y) * wh + 2 * x;
    const float* r1 = r0 + wh;
    lo[idx] = sel_min(sel_min(r0[0], r0[1]), sel_min(r1[0], r1[1]));
}

// ---------------------------------------------------------------- up-pass, one level per launch
__device__ __forceinline__ float lo_or_m1(const float* __restrict__ lo, int x, int y, int w, int h) {
    return (x >= 0 && x < w && y >= 0 && y < h) ? lo[y * w + x] : -1.0f;  // getPixelValue, project_cloud.cu:81-86
}

// Accessors of a coarse level: global memory (row stride = the level's up-pass width) or a clipped rectangle of it
// staged in shared memory (up_fused_kernel).
struct LoGlobal {
    const float* __restrict__ p;
    int w;
    __device__ __forceinline__ float operator()(int x, int y) const { return p[y * w + x]; }
    __device__ __forceinline__ const float* row(int y) const { return p + y * w; }  // row(y)[x] == (*this)(x, y)
};
template <int STRIDE>  // compile-time row stride: the staged rectangle sits in the top-left corner of a fixed-size array
struct LoShared {
    const float* p;
    int x0, y0;
    __device__ __forceinline__ float operator()(int x, int y) const { return p[(y - y0) * STRIDE + (x - x0)]; }
    __device__ __forceinline__ const float* row(int y) const { return p + ((y - y0) * STRIDE - x0); }
};

// bilinear hole fill of one fine pixel (resizeKernel, project_cloud.cu:135-160; op order from SASS)
template <typename Lo>
__device__ __forceinline__ float bilinear_up_t(const Lo& lo, int lw, int lh, int x, int y) {
    const float inX = __fmaf_rn(__fadd_rn(float(x), 0.5f), 0.5f, -0.5f);
    const float inY = __fmaf_rn(__fadd_rn(float(y), 0.5f), 0.5f, -0.5f);
    int x0 = __float2int_rd(inX), y0 = __float2int_rd(inY);
    int x1 = x0 + 1, y1 = y0 + 1;
    x0 = x0 < 0 ? 0 : (x0 >= lw ? lw - 1 : x0);
    x1 = x1 < 0 ? 0 : (x1 >= lw ? lw - 1 : x1);
    y0 = y0 < 0 ? 0 : (y0 >= lh ? lh - 1 : y0);
    y1 = y1 < 0 ? 0 : (y1 >= lh ? lh - 1 : y1);
    const float wx = __fsub_rn(inX, float(x0)), wy = __fsub_rn(inY, float(y0));
    const float omx = __fsub_rn(1.0f, wx);
    const float v0 = __fmaf_rn(wx, lo(x1, y0), __fmul_rn(omx, lo(x0, y0)));
    const float v1 = __fmaf_rn(wx, lo(x1, y1), __fmul_rn(omx, lo(x0, y1)));
    return __fmaf_rn(v0, __fsub_rn(1.0f, wy), __fmul_rn(wy, v1));
}
__device__ __forceinline__ float bilinear_up(const float* __restrict__ lo, int lw, int lh, int x, int y) {
    return bilinear_up_t(LoGlobal{lo, lw}, lw, lh, x, y);
}

// laplacianKernel + compareImgsKernel for ONE fine pixel with value cur under coarse parent (lx, ly)
// (project_cloud.cu:55-126): true = the pixel keeps its value.
template <typename Lo>
__device__ __forceinline__ bool up_keep_t(const Lo& lo, int lw, int lh, int lx, int ly, float cur) {
    if (cur >= __uint_as_float(kMaxFloatThresholdBits)) return false;  // `if (currentVal >= MAX_FLOAT) mask = 0`
    const bool border = (lx == 0 || lx == lw - 1 || ly == 0 || ly == lh - 1);  // laplacianKernel border -> 0
    if (!border) {
        float nb[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) nb[k] = lo(lx + k % 3 - 1, ly + k / 3 - 1);
        float sum = 0.0f;  // nine chained FMAs, k row-major, zero-weight taps included (laplacianKernel as compiled)
        sum = __fmaf_rn(nb[0], 0.0f, sum);
        sum = __fmaf_rn(nb[1], 1.0f, sum);
        sum = __fmaf_rn(nb[2], 0.0f, sum);
        sum = __fmaf_rn(nb[3], 1.0f, sum);
        sum = __fmaf_rn(nb[4], -4.0f, sum);
        sum = __fmaf_rn(nb[5], 1.0f, sum);
        sum = __fmaf_rn(nb[6], 0.0f, sum);
        sum = __fmaf_rn(nb[7], 1.0f, sum);
        sum = __fmaf_rn(nb[8], 0.0f, sum);
        if (sum > kGradientFilter) {
            bool keep = false;
#pragma unroll
            for (int k = 0; k < 9; ++k) keep = keep || (cur <= __fmul_rn(nb[k], kFilterStrength));
            return keep;
        }
    }
    return cur <= __fmul_rn(lo(lx, ly), kFilterStrength);
}

__device__ __forceinline__ uint16_t h_bits(float f) { return __half_as_ushort(__float2half_rn(f)); }
// (c10::Half)v / d  ->  half( float(half(v)) / d )   (project_cloud.cu:181-185)
__device__ __forceinline__ uint16_t half_div(float v, float d) {
    return h_bits(__fdiv_rn(__half2float(__float2half_rn(v)), d));
}

// Each thread owns two horizontally adjacent fine pixels (2*lx, 2*lx+1) of row hy: they share the
// coarse parent (lx, hy/2), its 3x3 neighbourhood, the Laplacian and all store vectorisation.
// lo = L_i indexed with (lw, lh) = up-pass dims of level i; hi = L_{i-1} indexed with (2lw, 2lh).
template <bool FINAL>
__global__ void __launch_bounds__(256) up_level_kernel(const float* __restrict__ lo, int lw, int lh,
                                                       float* __restrict__ hi, uint8_t* __restrict__ mask_tap,
                                                       uint8_t* __restrict__ image, uint16_t* __restrict__ tensor,
                                                       const uint32_t* __restrict__ minmax) {
    pdl_prologue();
    const int pair = blockIdx.x * blockDim.x + threadIdx.x;
    const int hh = lh * 2, hw = lw * 2;
    if (pair >= lw * hh) return;
    const int lx = pair % lw, hy = pair / lw, ly = hy >> 1;
    const size_t idx = size_t(hy) * hw + 2 * lx;
    const float2 cur2 = *reinterpret_cast<const float2*>(hi + idx);
    const float cur[2] = {cur2.x, cur2.y};
    const float thr = __uint_as_float(kMaxFloatThresholdBits);
    const bool cand[2] = {!(cur[0] >= thr), !(cur[1] >= thr)};  // `if (currentVal >= MAX_FLOAT) mask = 0`
    bool keep[2] = {false, false};
    if (cand[0] || cand[1]) {
        const bool border = (lx == 0 || lx == lw - 1 || ly == 0 || ly == lh - 1);  // laplacianKernel border -> 0
        bool edge = false;
        float nb[9];
        if (!border) {
#pragma unroll
            for (int k = 0; k < 9; ++k) nb[k] = lo[(ly + k / 3 - 1) * lw + (lx + k % 3 - 1)];
            // sum += in[k] * laplaceKernel[k], k row-major, all nine taps, FMA-contracted (SASS)
            float sum = 0.0f;
            sum = __fmaf_rn(nb[0], 0.0f, sum);
            sum = __fmaf_rn(nb[1], 1.0f, sum);
            sum = __fmaf_rn(nb[2], 0.0f, sum);
            sum = __fmaf_rn(nb[3], 1.0f, sum);
            sum = __fmaf_rn(nb[4], -4.0f, sum);
            sum = __fmaf_rn(nb[5], 1.0f, sum);
            sum = __fmaf_rn(nb[6], 0.0f, sum);
            sum = __fmaf_rn(nb[7], 1.0f, sum);
            sum = __fmaf_rn(nb[8], 0.0f, sum);
            edge = sum > kGradientFilter;
        }
        if (edge) {
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const float lim = __fmul_rn(nb[k], kFilterStrength);
                keep[0] = keep[0] || (cur[0] <= lim);
                keep[1] = keep[1] || (cur[1] <= lim);
            }
        } else {
            const float lim = __fmul_rn(lo_or_m1(lo, lx, ly, lw, lh), kFilterStrength);
            keep[0] = cur[0] <= lim;
            keep[1] = cur[1] <= lim;
        }
        keep[0] = keep[0] && cand[0];
        keep[1] = keep[1] && cand[1];
    }
    if (mask_tap) *reinterpret_cast<uint16_t*>(mask_tap + idx) = uint16_t((keep[0] ? 0xFFu : 0u) | (keep[1] ? 0xFF00u : 0u));
    if constexpr (!FINAL) {
        if (keep[0] && keep[1]) return;
        float2 out = cur2;
        if (!keep[0]) out.x = bilinear_up(lo, lw, lh, 2 * lx, hy);
        if (!keep[1]) out.y = bilinear_up(lo, lw, lh, 2 * lx + 1, hy);
        *reinterpret_cast<float2*>(hi + idx) = out;
    } else {
        // removeMask (project_cloud.cu:163-187); tensor plane stride = hw*hh (the truncated dims).
        const size_t plane = size_t(hw) * hh;
        const float dmin = __uint_as_float(minmax[0]), dmax = __uint_as_float(minmax[1]);
        const float range = __fsub_rn(dmax, dmin);
        uint16_t* img2 = reinterpret_cast<uint16_t*>(image + idx * 3);  // idx even -> 2-byte aligned
        const uint16_t i0 = img2[0], i1 = img2[1], i2 = img2[2];
        uint8_t c[6] = {uint8_t(i0 & 0xFF), uint8_t(i0 >> 8), uint8_t(i1 & 0xFF), uint8_t(i1 >> 8), uint8_t(i2 & 0xFF), uint8_t(i2 >> 8)};
        uint16_t t[5][2];
        float2 dout = cur2;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            if (!keep[j]) {
                (j == 0 ? dout.x : dout.y) = -1.0f;
                c[3 * j] = c[3 * j + 1] = c[3 * j + 2] = 0;
                t[0][j] = t[1][j] = t[2][j] = t[3][j] = 0;
                t[4][j] = 0xBC00u;  // -1.0h
            } else {
                t[0][j] = half_div(float(c[3 * j + 0]), 255.0f);
                t[1][j] = half_div(float(c[3 * j + 1]), 255.0f);
                t[2][j] = half_div(float(c[3 * j + 2]), 255.0f);
                t[3][j] = 0x3C00u;  // half(float(half(255)) / 255) = 1.0h
                t[4][j] = half_div(__fsub_rn(cur[j], dmin), range);
            }
        }
        if (!(keep[0] && keep[1])) {
            *reinterpret_cast<float2*>(hi + idx) = dout;
            img2[0] = uint16_t(c[0] | (c[1] << 8)); img2[1] = uint16_t(c[2] | (c[3] << 8)); img2[2] = uint16_t(c[4] | (c[5] << 8));
        }
#pragma unroll
        for (int k = 0; k < 5; ++k)
            *reinterpret_cast<uint32_t*>(tensor + plane * k + idx) = uint32_t(t[k][0]) | (uint32_t(t[k][1]) << 16);
    }
}

// ---------------------------------------------------------------- final level, 8 fine pixels per thread
// Same arithmetic as up_level_kernel<true>; a thread owns the 8 fine pixels of row hy under 4 adjacent coarse
// pixels, so depth / tensor planes move as 16-byte vectors and the image as three 8-byte words, and the 3x3 coarse
// neighbourhoods of the four parents share one 3x6 window.  Needs lw % 4 == 0 (true whenever W % 16 == 0).
// Body shared by up_final_wide_kernel and up_fused_kernel: the 8 fine pixels of row hy under coarse pixels lx0..lx0+3.
// `c0, c1` = the 8 depths, `iw` = their 24 image bytes (loaded by the caller, possibly long before);
// `lut`: when not null, lut[v] = half(float(half(v)) / 255) for v = 0..255 (the value removeMask computes per byte).
template <typename Lo>
__device__ __forceinline__ void up_final_group(const Lo& lo, int lw, int lh, float* __restrict__ hi,
                                               uint8_t* __restrict__ mask_tap, uint8_t* __restrict__ image,
                                               uint16_t* __restrict__ tensor, float dmin, float range, int lx0, int hy,
                                               const float4 c0, const float4 c1, uint2 (&iw)[3], const uint16_t* lut) {
    const int hh = lh * 2, hw = lw * 2, ly = hy >> 1;
    const size_t idx = size_t(hy) * hw + size_t(lx0) * 2;
    const float cur[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
    const float thr = __uint_as_float(kMaxFloatThresholdBits);
    bool keep[8];
    bool any_cand = false;
#pragma unroll
    for (int j = 0; j < 8; ++j) { keep[j] = !(cur[j] >= thr); any_cand |= keep[j]; }  // candidates so far
    if (any_cand) {
        // 3 x 6 window of the coarse level around the four parents (clamped loads; out-of-range entries are never used
        // by an interior parent, and border parents do not use the window at all)
        float w[3][6];
#pragma unroll
        const int xl = max(lx0 - 1, 0), xr = min(lx0 + 4, lw - 1);  // only the window's outer columns can leave the level
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const float* row = lo.row(min(max(ly + r - 1, 0), lh - 1));
            w[r][0] = row[xl];
#pragma unroll
            for (int c = 1; c < 5; ++c) w[r][c] = row[lx0 + c - 1];
            w[r][5] = row[xr];
        }
        const bool row_border = (ly == 0 || ly == lh - 1);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int lx = lx0 + k;
            if (!(keep[2 * k] || keep[2 * k + 1])) continue;
            const bool border = row_border || lx == 0 || lx == lw - 1;
            bool edge = false;
            if (!border) {
                float sum = 0.0f;  // nine chained FMAs, k row-major, zero-weight taps included (laplacianKernel as compiled)
                sum = __fmaf_rn(w[0][k], 0.0f, sum);
                sum = __fmaf_rn(w[0][k + 1], 1.0f, sum);
                sum = __fmaf_rn(w[0][k + 2], 0.0f, sum);
                sum = __fmaf_rn(w[1][k], 1.0f, sum);
                sum = __fmaf_rn(w[1][k + 1], -4.0f, sum);
                sum = __fmaf_rn(w[1][k + 2], 1.0f, sum);
                sum = __fmaf_rn(w[2][k], 0.0f, sum);
                sum = __fmaf_rn(w[2][k + 1], 1.0f, sum);
                sum = __fmaf_rn(w[2][k + 2], 0.0f, sum);
                edge = sum > kGradientFilter;
            }
            bool k0 = false, k1 = false;
            if (edge) {
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float lim = __fmul_rn(w[r][k + c], kFilterStrength);
                        k0 = k0 || (cur[2 * k] <= lim);
                        k1 = k1 || (cur[2 * k + 1] <= lim);
                    }
            } else {
                const float lim = __fmul_rn(w[1][k + 1], kFilterStrength);
                k0 = cur[2 * k] <= lim;
                k1 = cur[2 * k + 1] <= lim;
            }
            keep[2 * k] = keep[2 * k] && k0;
            keep[2 * k + 1] = keep[2 * k + 1] && k1;
        }
    }
    if (mask_tap) {
        uint32_t m0 = 0, m1 = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) { m0 |= keep[j] ? (0xFFu << (8 * j)) : 0u; m1 |= keep[4 + j] ? (0xFFu << (8 * j)) : 0u; }
        *reinterpret_cast<uint2*>(mask_tap + idx) = make_uint2(m0, m1);
    }
    // removeMask (project_cloud.cu:163-187)
    const size_t plane = size_t(hw) * hh;
    uint2* img8 = reinterpret_cast<uint2*>(image + idx * 3);  // idx % 8 == 0 -> 24-byte group, 8-byte aligned
    uint8_t* c = reinterpret_cast<uint8_t*>(iw);
    uint16_t tp[5][8];
    float dout[8];
    bool all_keep = true;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        dout[j] = cur[j];
        if (!keep[j]) {
            all_keep = false;
            dout[j] = -1.0f;
            c[3 * j] = c[3 * j + 1] = c[3 * j + 2] = 0;
            tp[0][j] = tp[1][j] = tp[2][j] = tp[3][j] = 0;
            tp[4][j] = 0xBC00u;  // -1.0h
        } else {
            if (lut) {
                tp[0][j] = lut[c[3 * j + 0]];
                tp[1][j] = lut[c[3 * j + 1]];
                tp[2][j] = lut[c[3 * j + 2]];
            } else {
                tp[0][j] = half_div(float(c[3 * j + 0]), 255.0f);
                tp[1][j] = half_div(float(c[3 * j + 1]), 255.0f);
                tp[2][j] = half_div(float(c[3 * j + 2]), 255.0f);
            }
            tp[3][j] = 0x3C00u;  // half(float(half(255)) / 255) = 1.0h: x / x is exactly 1 in IEEE arithmetic
            tp[4][j] = half_div(__fsub_rn(cur[j], dmin), range);
        }
    }
    if (!all_keep) {
        *reinterpret_cast<float4*>(hi + idx) = make_float4(dout[0], dout[1], dout[2], dout[3]);
        *reinterpret_cast<float4*>(hi + idx + 4) = make_float4(dout[4], dout[5], dout[6], dout[7]);
        img8[0] = iw[0]; img8[1] = iw[1]; img8[2] = iw[2];
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        uint4 v;
        v.x = uint32_t(tp[k][0]) | (uint32_t(tp[k][1]) << 16);
        v.y = uint32_t(tp[k][2]) | (uint32_t(tp[k][3]) << 16);
        v.z = uint32_t(tp[k][4]) | (uint32_t(tp[k][5]) << 16);
        v.w = uint32_t(tp[k][6]) | (uint32_t(tp[k][7]) << 16);
#if RTR_L2_HINTS & 1
        __stcs(reinterpret_cast<uint4*>(tensor + plane * k + idx), v);  // nobody on this stream reads the tensor again: evict first
#else
        *reinterpret_cast<uint4*>(tensor + plane * k + idx) = v;
#endif
    }
}

__global__ void __launch_bounds__(256) up_final_wide_kernel(const float* __restrict__ lo, int lw, int lh,
                                                            float* __restrict__ hi, uint8_t* __restrict__ mask_tap,
                                                            uint8_t* __restrict__ image, uint16_t* __restrict__ tensor,
                                                            const uint32_t* __restrict__ minmax) {
    pdl_prologue();
    const int groups = lw >> 2, hh = lh * 2;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= groups * hh) return;
    const float dmin = __uint_as_float(minmax[0]), dmax = __uint_as_float(minmax[1]);
    const int lx0 = (t % groups) * 4, hy = t / groups;
    const size_t idx = size_t(hy) * (lw * 2) + size_t(lx0) * 2;
    const float4 c0 = *reinterpret_cast<const float4*>(hi + idx), c1 = *reinterpret_cast<const float4*>(hi + idx + 4);
    const uint2* img8 = reinterpret_cast<const uint2*>(image + idx * 3);
    uint2 iw[3] = {img8[0], img8[1], img8[2]};
    up_final_group(LoGlobal{lo, lw}, lw, lh, hi, mask_tap, image, tensor, dmin, __fsub_rn(dmax, dmin), lx0, hy, c0, c1, iw, nullptr);
}

// ---------------------------------------------------------------- the whole up-pass in one launch
// The four iterations of the up-pass (levels 4->3, 3->2, 2->1 with hole filling, 1->0 with removeMask) depend on
// each other only through a one-pixel ring of the coarser level: fine pixel (x, y) reads the 3x3 neighbourhood of its
// parent (x/2, y/2) (Laplacian, compare) and the bilinear taps floor((x-0.5)/2), +1.  A CTA therefore takes a
// 64 x 32 tile of level 0 (one 8-pixel group per thread), stages the regions of levels 1..4 that tile depends on in
// shared memory (34x18, 20x12, 12x8, 8x6 values), fills the holes of levels 3, 2, 1 there — the halo is recomputed by
// the neighbouring CTAs instead of being exchanged — and finishes with the level-0 pass.  Every global load (the four
// staged regions, the thread's 8 depths and 24 image bytes) is issued before the first barrier, so one memory latency
// is exposed instead of five, and the three per-byte IEEE divisions of removeMask come from a 256-entry table
// computed by the CTA with the same operations.  Levels 1..3 in global memory are only read (no cross-CTA hazard);
// they keep their down-pass values, so the per-level launches stay the path for keep_masks = 1 (taps of every
// reference kernel).
constexpr int kUpT1W = 32, kUpT1H = 16;  // level-1 core of a CTA
static_assert(kUpT1W / 4 * kUpT1H * 2 == 256, "one 8-pixel group of level 0 per thread");
struct UpRect { int x0, y0, w, h; };
// coarse rectangle a fine rectangle depends on, clipped to the coarse level's dims
__device__ __forceinline__ UpRect up_parent_rect(const UpRect f, int lw, int lh) {
    const int x0 = max((f.x0 >> 1) - 1, 0), x1 = min(((f.x0 + f.w - 1) >> 1) + 1, lw - 1);
    const int y0 = max((f.y0 >> 1) - 1, 0), y1 = min(((f.y0 + f.h - 1) >> 1) + 1, lh - 1);
    return UpRect{x0, y0, x1 - x0 + 1, y1 - y0 + 1};
}
// SW x SH = the array's fixed size (>= the clipped rectangle r); thread t handles local pixel (t % SW, t / SW)
template <int SW, int SH>
__device__ __forceinline__ void up_stage_rect(float* dst, const float* __restrict__ src, int src_w, const UpRect r) {
#pragma unroll
    for (int i = threadIdx.x; i < SW * SH; i += 256) {
        const int lx = i % SW, ly = i / SW;
        if (lx < r.w && ly < r.h) dst[i] = src[(r.y0 + ly) * src_w + r.x0 + lx];
    }
}
// hole-fill rectangle rf of the fine level in place in shared memory from the staged coarse level
template <int FW, int FH, int CW>
__device__ __forceinline__ void up_fill_rect(float* fine, const UpRect rf, const float* coarse, const UpRect rc, int lw, int lh) {
    const LoShared<CW> lo{coarse, rc.x0, rc.y0};
#pragma unroll
    for (int i = threadIdx.x; i < FW * FH; i += 256) {
        const int lx = i % FW, ly = i / FW;
        if (lx < rf.w && ly < rf.h) {
            const int x = rf.x0 + lx, y = rf.y0 + ly;
            const float cur = fine[i];
            if (!up_keep_t(lo, lw, lh, x >> 1, y >> 1, cur)) fine[i] = bilinear_up_t(lo, lw, lh, x, y);
        }
    }
}

__global__ void __launch_bounds__(256, RTR_UP_MIN_CTAS) up_fused_kernel(const float* __restrict__ l1, const float* __restrict__ l2,
                                                       const float* __restrict__ l3, const float* __restrict__ l4,
                                                       int w4, int h4, float* __restrict__ l0, uint8_t* __restrict__ image,
                                                       uint16_t* __restrict__ tensor, const uint32_t* __restrict__ minmax) {
    constexpr int S1W = kUpT1W + 2, S1H = kUpT1H + 2, S2W = kUpT1W / 2 + 4, S2H = kUpT1H / 2 + 4, S3W = kUpT1W / 4 + 4,
                  S3H = kUpT1H / 4 + 4, S4W = kUpT1W / 8 + 4, S4H = kUpT1H / 8 + 4;
    __shared__ float s1[S1W * S1H];
    __shared__ float s2[S2W * S2H];
    __shared__ float s3[S3W * S3H];
    __shared__ float s4[S4W * S4H];
    __shared__ uint16_t lut[256];
    lut[threadIdx.x] = half_div(float(threadIdx.x), 255.0f);  // removeMask's value for image byte threadIdx.x
    pdl_prologue();
    const int w3 = w4 * 2, h3 = h4 * 2, w2 = w4 * 4, h2 = h4 * 4, w1 = w4 * 8, h1 = h4 * 8;
    const UpRect core{int(blockIdx.x) * kUpT1W, int(blockIdx.y) * kUpT1H, min(kUpT1W, w1 - int(blockIdx.x) * kUpT1W),
                      min(kUpT1H, h1 - int(blockIdx.y) * kUpT1H)};
    // this thread's 8 level-0 pixels: loads issued first, consumed after the shared-memory chain
    const int groups = core.w >> 2;  // w1 % 4 == 0 and the tile width is a multiple of 4
    const bool mine = int(threadIdx.x) < groups * core.h * 2;
    const int lx0 = core.x0 + (mine ? int(threadIdx.x) % groups : 0) * 4, hy = core.y0 * 2 + (mine ? int(threadIdx.x) / groups : 0);
    const size_t idx = size_t(hy) * (w1 * 2) + size_t(lx0) * 2;
    float4 c0 = make_float4(0.f, 0.f, 0.f, 0.f), c1 = c0;
    uint2 iw[3] = {make_uint2(0u, 0u), make_uint2(0u, 0u), make_uint2(0u, 0u)};
    if (mine) {
        c0 = *reinterpret_cast<const float4*>(l0 + idx);
        c1 = *reinterpret_cast<const float4*>(l0 + idx + 4);
        const uint2* img8 = reinterpret_cast<const uint2*>(image + idx * 3);
        iw[0] = img8[0]; iw[1] = img8[1]; iw[2] = img8[2];
    }
    // level-1 pixels the level-0 tile reads: the tile's parents and one ring around them; and so on upwards
    const UpRect r1 = up_parent_rect(UpRect{core.x0 * 2, core.y0 * 2, core.w * 2, core.h * 2}, w1, h1);
    const UpRect r2 = up_parent_rect(r1, w2, h2), r3 = up_parent_rect(r2, w3, h3), r4 = up_parent_rect(r3, w4, h4);
    up_stage_rect<S4W, S4H>(s4, l4, w4, r4);
    up_stage_rect<S3W, S3H>(s3, l3, w3, r3);
    up_stage_rect<S2W, S2H>(s2, l2, w2, r2);
    up_stage_rect<S1W, S1H>(s1, l1, w1, r1);
    const float dmin = __uint_as_float(minmax[0]), dmax = __uint_as_float(minmax[1]);
    __syncthreads();
    up_fill_rect<S3W, S3H, S4W>(s3, r3, s4, r4, w4, h4);
    __syncthreads();
    up_fill_rect<S2W, S2H, S3W>(s2, r2, s3, r3, w3, h3);
    __syncthreads();
    up_fill_rect<S1W, S1H, S2W>(s1, r1, s2, r2, w2, h2);
    __syncthreads();
    if (mine)
        up_final_group(LoShared<S1W>{s1, r1.x0, r1.y0}, w1, h1, l0, nullptr, image, tensor, dmin, __fsub_rn(dmax, dmin), lx0, hy, c0, c1, iw, lut);
}

// ---------------------------------------------------------------- launchers
#ifdef RTR_EXPERIMENTS
// measurement: RTR_IMAGE_CARVEOUT = shared-memory carve-out (percent) the image kernels ask for — the same split as the
// ring kernels' would let their CTAs follow a ring CTA onto an SM without the SM having to drain and re-split first
template <typename K>
static void exp_carveout(K kernel) {
    static bool done = false;
    if (done) return;
    done = true;
    if (const char* c = std::getenv("RTR_IMAGE_CARVEOUT")) cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, std::atoi(c));
}
#else
template <typename K> static void exp_carveout(K) {}
#endif
cudaError_t launch_resolve_gated(cudaStream_t s, const FrameBuffers& fb, int W, int H) {
    const uint64_t cov = clear_coverage(W, H);
    if (cov == 0) return cudaSuccess;
    launch_pdl(resolve_generic_kernel, dim3(148 * 2), dim3(256), s, reinterpret_cast<const uint4*>(fb.accum), fb.image, cov,
               false, fb.minmax + 2, true);
    return cudaGetLastError();
}

cudaError_t launch_resolve_pyramid(cudaStream_t s, const FrameBuffers& fb, int W, int H, const PyramidDims& d,
                                   bool pyramid, bool resolve, bool force_generic, bool f32acc) {
    const uint64_t cov = clear_coverage(W, H);
    if (cov == 0 || (!pyramid && !resolve)) return cudaSuccess;
    const uint4* acc = reinterpret_cast<const uint4*>(fb.accum);
    if ((W % 16) == 0 && !force_generic) {
        dim3 grid((W + 63) / 64, H / 16);
#define RTR_RP(P_, R_, F_) launch_pdl((resolve_pyramid_kernel<P_, R_, F_>), dim3(grid), dim3(256), s, fb.zbuf, acc, fb.image, fb.level[1], fb.level[2], fb.level[3], fb.level[4], fb.minmax, W)
        exp_carveout(resolve_pyramid_kernel<true, true, true>);
#if RTR_RESOLVE_PERSIST
        if (pyramid && resolve) {
            const int tiles_x = int(grid.x), n_tiles = int(grid.x * grid.y);
            int dev = 0, sms = 148;
            if (cudaGetDevice(&dev) == cudaSuccess) (void)cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            const unsigned pgrid = unsigned(n_tiles < sms * 5 ? n_tiles : sms * 5);
            if (f32acc) launch_pdl((resolve_pyramid_persist_kernel<true>), dim3(pgrid), dim3(256), s, fb.zbuf, acc, fb.image, fb.level[1], fb.level[2], fb.level[3], fb.level[4], fb.minmax, W, tiles_x, n_tiles);
            else launch_pdl((resolve_pyramid_persist_kernel<false>), dim3(pgrid), dim3(256), s, fb.zbuf, acc, fb.image, fb.level[1], fb.level[2], fb.level[3], fb.level[4], fb.minmax, W, tiles_x, n_tiles);
            return cudaGetLastError();
        }
#endif
        if (pyramid && resolve) { if (f32acc) RTR_RP(true, true, true); else RTR_RP(true, true, false); }
        else if (pyramid) RTR_RP(true, false, false);
        else { if (f32acc) RTR_RP(false, true, true); else RTR_RP(false, true, false); }
#undef RTR_RP
        return cudaGetLastError();
    }
    if (resolve) launch_pdl(resolve_generic_kernel, dim3(unsigned((cov + 255) / 256)), dim3(256), s, acc, fb.image, cov, f32acc, fb.minmax + 2, false);
    if (pyramid) {
        const uint64_t count = uint64_t(d.uw[0]) * d.uh[0];
        const uint64_t mm_blocks = (count + 255) / 256;
        if (count) launch_pdl(minmax_generic_kernel, dim3(unsigned(mm_blocks < 148 * 8 ? mm_blocks : 148 * 8)), dim3(256), s, fb.zbuf, count, fb.minmax);
        for (int i = 1; i <= 4; ++i) {
            const int n = d.w[i] * d.h[i];
            if (n) launch_pdl(reduce_generic_kernel, dim3((n + 255) / 256), dim3(256), s, fb.level[i - 1], fb.level[i], d.w[i], d.h[i]);
        }
    }
    return cudaGetLastError();
}

// ---------------------------------------------------------------- 64-bit key mode (north_star variant)
// zkey[p] = (depth bits << 32) | point index.  Splits the key back into the reference-identical
// depth buffer and a NEAREST-point colour (not the reference's 2 cm average: opt-in, non-parity).
__global__ void __launch_bounds__(256) clear_key64_kernel(unsigned long long* __restrict__ zkey, uint64_t cov) {
    pdl_prologue();
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < cov; i += uint64_t(gridDim.x) * blockDim.x)
        zkey[i] = (static_cast<unsigned long long>(kEmptyDepthBits) << 32) | 0xFFFFFFFFull;
}
__global__ void __launch_bounds__(256) resolve_key64_kernel(const unsigned long long* __restrict__ zkey,
                                                            const PointRecord* __restrict__ pts, uint64_t index_base,
                                                            uint64_t n_local, uint32_t* __restrict__ zbuf,
                                                            uint8_t* __restrict__ image, uint64_t n_px, uint64_t cov) {
    pdl_prologue();
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n_px) return;
    const unsigned long long key = zkey[i];
    zbuf[i] = uint32_t(key >> 32);
    if (i >= cov) return;  // resolvePass coverage
    uint32_t c = 0u;
    const uint64_t idx = uint64_t(uint32_t(key)) - index_base;  // points of another shard resolve to 0 here
    if (uint32_t(key >> 32) != kEmptyDepthBits && idx < n_local) c = pts[idx].bgra;
    image[i * 3 + 0] = uint8_t(c); image[i * 3 + 1] = uint8_t(c >> 8); image[i * 3 + 2] = uint8_t(c >> 16);
}
cudaError_t launch_clear_key64(cudaStream_t s, int sm_count, unsigned long long* zkey, uint64_t cov) {
    if (cov) launch_pdl(clear_key64_kernel, dim3(sm_count * 4), dim3(256), s, zkey, cov);
    return cudaGetLastError();
}
cudaError_t launch_resolve_key64(cudaStream_t s, const unsigned long long* zkey, const PointRecord* pts,
                                 uint64_t index_base, uint64_t n_local, uint32_t* zbuf, uint8_t* image, uint64_t n_px,
                                 uint64_t cov) {
    if (n_px) launch_pdl(resolve_key64_kernel, dim3(unsigned((n_px + 255) / 256)), dim3(256), s, zkey, pts, index_base, n_local, zbuf, image, n_px, cov);
    return cudaGetLastError();
}

cudaError_t launch_up_pass(cudaStream_t s, const FrameBuffers& fb, const PyramidDims& d, bool force_generic, bool fused) {
    // tensor plane stride (uw[0]*uh[0]) must keep the 16-byte stores aligned: multiples of 8 halfs
    const bool wide_ok = !force_generic && (d.uw[1] % 4) == 0 && ((size_t(d.uw[0]) * d.uh[0]) % 8) == 0;
    const bool taps = fb.mask[0] || fb.mask[1] || fb.mask[2] || fb.mask[3];
    // one launch for all four levels: needs the flat pyramid to be a true 2-D pyramid (w[i] == uw[i], i.e. W % 16 == 0)
    if (fused && wide_ok && !taps && d.uw[4] > 0 && d.uh[4] > 0 && d.w[1] == d.uw[1] && d.w[2] == d.uw[2] && d.w[3] == d.uw[3]) {
        exp_carveout(up_fused_kernel);
        launch_pdl(up_fused_kernel, dim3((d.uw[1] + kUpT1W - 1) / kUpT1W, (d.uh[1] + kUpT1H - 1) / kUpT1H), dim3(256), s,
                   fb.level[1], fb.level[2], fb.level[3], fb.level[4], d.uw[4], d.uh[4], fb.level[0], fb.image, fb.tensor, fb.minmax);
        return cudaGetLastError();
    }
    for (int i = 4; i >= 1; --i) {
        const int pairs = d.uw[i] * d.uh[i] * 2;  // fine pixels / 2
        if (pairs == 0) continue;
        const unsigned grid = (pairs + 255) / 256;
        if (i == 1 && wide_ok)
            launch_pdl(up_final_wide_kernel, dim3((d.uw[1] / 4 * d.uh[1] * 2 + 255) / 256), dim3(256), s, fb.level[1], d.uw[1], d.uh[1], fb.level[0], fb.mask[0], fb.image, fb.tensor, fb.minmax);
        else if (i > 1)
            launch_pdl((up_level_kernel<false>), dim3(grid), dim3(256), s, fb.level[i], d.uw[i], d.uh[i], fb.level[i - 1], fb.mask[i - 1], nullptr, nullptr, fb.minmax);
        else
            launch_pdl((up_level_kernel<true>), dim3(grid), dim3(256), s, fb.level[1], d.uw[1], d.uh[1], fb.level[0], fb.mask[0], fb.image, fb.tensor, fb.minmax);
    }
    return cudaGetLastError();
}
// launches launch_up_pass issues for these dims (the renderer's launch counter)
int up_pass_launches(const FrameBuffers& fb, const PyramidDims& d, bool force_generic, bool fused) {
    const bool wide_ok = !force_generic && (d.uw[1] % 4) == 0 && ((size_t(d.uw[0]) * d.uh[0]) % 8) == 0;
    const bool taps = fb.mask[0] || fb.mask[1] || fb.mask[2] || fb.mask[3];
    if (fused && wide_ok && !taps && d.uw[4] > 0 && d.uh[4] > 0 && d.w[1] == d.uw[1] && d.w[2] == d.uw[2] && d.w[3] == d.uw[3]) return 1;
    int n = 0;
    for (int i = 4; i >= 1; --i) n += (d.uw[i] * d.uh[i] != 0);
    return n;
}

}  // namespace rtr
