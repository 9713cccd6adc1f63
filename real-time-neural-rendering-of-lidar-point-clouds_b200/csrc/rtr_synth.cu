// Device-side synthetic LiDAR-hall generator (bench/test support).  Same integer generator as the
// CPU oracle uses (rtr_synth_common.h), so a 100 M / 1 B point cloud never has to cross PCIe and the
// CPU and GPU paths see bit-identical inputs.  Not a reference component: the reference loads
// PLY/E57 files (cloudreader.cpp); there is no network/dataset here (SURVEY.md §8 d).
#include "rtr_kernels.h"
#include "rtr_synth_common.h"

namespace rtr {

__global__ void __launch_bounds__(256) synth_kernel(const __grid_constant__ rtr_synth_scene scene, uint64_t first,
                                                    uint64_t count, PointRecord* __restrict__ out) {
    for (uint64_t j = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; j < count;
         j += uint64_t(gridDim.x) * blockDim.x) {
        int32_t p[3];
        uint32_t c;
        rtr_synth_point_fp(&scene, first + j, p, &c);
        uint4 rec;
        rec.x = __float_as_uint(rtr_synth_fp_to_m(p[0]));
        rec.y = __float_as_uint(rtr_synth_fp_to_m(p[1]));
        rec.z = __float_as_uint(rtr_synth_fp_to_m(p[2]));
        rec.w = c;
        reinterpret_cast<uint4*>(out)[j] = rec;
    }
}

cudaError_t launch_synth(cudaStream_t s, uint64_t seed, uint64_t n_total, uint64_t first, uint64_t count, int lx,
                         int ly, int lz, int nbox, PointRecord* out) {
    if (count == 0) return cudaSuccess;
    rtr_synth_scene scene;
    rtr_synth_build_scene(&scene, seed, n_total, lx, ly, lz, nbox);
    const uint64_t blocks = (count + 255) / 256;
    synth_kernel<<<unsigned(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, s>>>(scene, first, count, out);
    return cudaGetLastError();
}

}  // namespace rtr
