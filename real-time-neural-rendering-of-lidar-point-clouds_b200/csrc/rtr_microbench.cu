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

cudaError_t launch_red_bench(cudaStream_t s, int sm_count, int mode, bool key64, const int32_t* pix, uint64_t n_ops,
                             uint32_t n_px, uint32_t* z32, unsigned long long* z64) {
    const unsigned grid = unsigned(sm_count) * 8u;
    if (mode == 2 || mode == 3) {
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
