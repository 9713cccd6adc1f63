// L2 atomic-throughput micro-benchmark (measurement support, not on the render path).
// north_star asks for the z-min pass as a fraction of "measured L2 atomic throughput": this issues
// RED.MIN (u32 or u64, no return value) into a frame-sized, L2-resident buffer with
//   mode 0: uniformly random addresses generated in registers (no memory reads at all),
//   mode 1: the pixel ids the current cloud + camera project to (4 B/op read, coalesced),
// so the number is the atomic units' rate, not HBM's.
#include "rtr_kernels.h"
#include "rtr_synth_common.h"

namespace rtr {

template <bool KEY64>
__global__ void __launch_bounds__(256) red_random_kernel(uint64_t n_ops, uint32_t n_px, uint32_t* __restrict__ z32,
                                                         unsigned long long* __restrict__ z64) {
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n_ops; i += uint64_t(gridDim.x) * blockDim.x) {
        const uint64_t h = rtr_splitmix64(i);
        const uint32_t px = uint32_t((uint64_t(uint32_t(h)) * n_px) >> 32);
        const uint32_t val = uint32_t(h >> 32) | 0x40000000u;
        if constexpr (KEY64) atomicMin(z64 + px, (static_cast<unsigned long long>(val) << 32) | uint32_t(i));
        else atomicMin(z32 + px, val);
    }
}

template <bool KEY64>
__global__ void __launch_bounds__(256) red_pix_kernel(const int32_t* __restrict__ pix, uint64_t n_ops,
                                                      uint32_t* __restrict__ z32, unsigned long long* __restrict__ z64) {
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n_ops; i += uint64_t(gridDim.x) * blockDim.x) {
        const int32_t px = pix[i];
        if (px < 0) continue;
        const uint32_t val = uint32_t(rtr_splitmix64(i) >> 32) | 0x40000000u;
        if constexpr (KEY64) atomicMin(z64 + px, (static_cast<unsigned long long>(val) << 32) | uint32_t(i));
        else atomicMin(z32 + px, val);
    }
}

// modes 2/3: the blend pass's update of one 16-byte {b,g,r,count} accumulator at a random pixel, as
// two RED.ADD.64 (what blend_kernel issues) or as one RED.ADD.F32x4.
template <bool VEC4F>
__global__ void __launch_bounds__(256) red_accum_kernel(uint64_t n_ops, uint32_t n_px, unsigned long long* __restrict__ acc) {
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n_ops; i += uint64_t(gridDim.x) * blockDim.x) {
        const uint64_t h = rtr_splitmix64(i);
        const uint32_t px = uint32_t((uint64_t(uint32_t(h)) * n_px) >> 32);
        const uint32_t b = uint32_t(h >> 32) & 0xFF, g = uint32_t(h >> 40) & 0xFF, r = uint32_t(h >> 48) & 0xFF;
        if constexpr (VEC4F) {
            asm volatile("red.global.v4.f32.add [%0], {%1,%2,%3,%4};" ::"l"(acc + 2 * uint64_t(px)), "f"(float(b)), "f"(float(g)),
                         "f"(float(r)), "f"(1.0f)
                         : "memory");
        } else {
            atomicAdd(acc + 2 * uint64_t(px), static_cast<unsigned long long>(b) | (static_cast<unsigned long long>(g) << 32));
            atomicAdd(acc + 2 * uint64_t(px) + 1, static_cast<unsigned long long>(r) | (1ull << 32));
        }
    }
}

// modes 4..19: what a reduction costs when the lanes of ONE warp instruction share sectors or addresses (the question
// behind the lane -> record mapping of the ring kernels, profiles/r02Q_exp_red_coalescing.json).  Groups of L = 1, 2, 4, 8
// consecutive lanes go to one random 32-byte sector, either to distinct words of it (`same` = false) or all to its first
// word (`same` = true).  VEC4F: the 16-byte RED.ADD.F32x4 (two pixels per sector) instead of RED.MIN.U32 (eight).
template <bool VEC4F>
__global__ void __launch_bounds__(256) red_group_kernel(uint64_t n_ops, uint32_t n_px, uint32_t log2_l, bool same,
                                                        uint32_t* __restrict__ z32, unsigned long long* __restrict__ acc) {
    constexpr uint32_t kPerSector = VEC4F ? 2u : 8u;
    const uint32_t n_sectors = n_px / kPerSector;
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n_ops; i += uint64_t(gridDim.x) * blockDim.x) {
        const uint64_t hg = rtr_splitmix64(i >> log2_l), h = rtr_splitmix64(i ^ 0x9E3779B97F4A7C15ull);
        const uint32_t sector = uint32_t((uint64_t(uint32_t(hg)) * n_sectors) >> 32);
        const uint32_t px = sector * kPerSector + (same ? 0u : (uint32_t(i) & ((1u << log2_l) - 1u)) % kPerSector);
        if constexpr (VEC4F) {
            asm volatile("red.global.v4.f32.add [%0], {%1,%2,%3,%4};" ::"l"(acc + 2 * uint64_t(px)), "f"(float(uint32_t(h) & 0xFF)),
                         "f"(float(uint32_t(h >> 8) & 0xFF)), "f"(float(uint32_t(h >> 16) & 0xFF)), "f"(1.0f)
                         : "memory");
        } else {
            atomicMin(z32 + px, uint32_t(h >> 32) | 0x40000000u);
        }
    }
}

// rtr_selftest_fast_divide: project4's fast path against __fdividef on raw random bit patterns.
__global__ void __launch_bounds__(256) fast_divide_selftest_kernel(uint64_t n_pairs, uint64_t seed,
                                                                   unsigned long long* __restrict__ mismatches) {
    unsigned long long bad = 0;
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n_pairs; i += uint64_t(gridDim.x) * blockDim.x) {
        const uint64_t h = rtr_splitmix64(seed + i);
        uint32_t ab = uint32_t(h), bb = uint32_t(h >> 32);
        if ((i & 7) == 1) bb = (bb & 0x807FFFFFu) | ((i >> 3) % 3 == 0 ? 0u : ((i >> 3) % 3 == 1 ? 0x00800000u : 0x7E800000u));  // denormal / smallest normal / huge b
        if ((i & 7) == 2) ab = (ab & 0x807FFFFFu);                                                                            // denormal a
        const float a = __uint_as_float(ab), b = __uint_as_float(bb);
        if (fabsf(b) < 1.17549435e-38f) continue;  // the warp takes __fdividef itself there
        const float q_ref = __fdividef(a, b), q_fast = __fmul_rn(mufu_rcp(b), a);
        const bool same_bits = __float_as_uint(q_ref) == __float_as_uint(q_fast) || (q_ref != q_ref && q_fast != q_fast);
        if (!same_bits || __float2int_rn(q_ref) != __float2int_rn(q_fast)) ++bad;
    }
    if (bad) atomicAdd(mismatches, bad);
}
cudaError_t launch_fast_divide_selftest(cudaStream_t s, int sm_count, uint64_t n_pairs, uint64_t seed, unsigned long long* mismatches) {
    fast_divide_selftest_kernel<<<unsigned(sm_count) * 8u, 256, 0, s>>>(n_pairs, seed, mismatches);
    return cudaGetLastError();
}

cudaError_t launch_red_bench(cudaStream_t s, int sm_count, int mode, bool key64, const int32_t* pix, uint64_t n_ops,
                             uint32_t n_px, uint32_t* z32, unsigned long long* z64) {
    const unsigned grid = unsigned(sm_count) * 8u;
    if (mode >= 4) {  // 4 + 4 * same + log2(L): RED.MIN.U32; 12 + 4 * same + log2(L): RED.ADD.F32x4
        const int m = (mode - 4) & 7;
        if (mode >= 12) red_group_kernel<true><<<grid, 256, 0, s>>>(n_ops, n_px, uint32_t(m & 3), (m >> 2) != 0, z32, z64);
        else red_group_kernel<false><<<grid, 256, 0, s>>>(n_ops, n_px, uint32_t(m & 3), (m >> 2) != 0, z32, z64);
    } else if (mode == 2 || mode == 3) {
        if (mode == 3) red_accum_kernel<true><<<grid, 256, 0, s>>>(n_ops, n_px, z64);
        else red_accum_kernel<false><<<grid, 256, 0, s>>>(n_ops, n_px, z64);
    } else if (mode == 0) {
        if (key64) red_random_kernel<true><<<grid, 256, 0, s>>>(n_ops, n_px, z32, z64);
        else red_random_kernel<false><<<grid, 256, 0, s>>>(n_ops, n_px, z32, z64);
    } else {
        if (key64) red_pix_kernel<true><<<grid, 256, 0, s>>>(pix, n_ops, z32, z64);
        else red_pix_kernel<false><<<grid, 256, 0, s>>>(pix, n_ops, z32, z64);
    }
    return cudaGetLastError();
}

}  // namespace rtr
