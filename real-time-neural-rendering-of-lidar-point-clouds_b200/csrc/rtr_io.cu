// The callers and data formats on either side of the hot path (SURVEY.md §8 f): C ABI in
// include/rtr_b200_io.h.  Host parsing is plain C++; the O(N) work (cell binning, output
// post-process) runs on the GPU.
//
//   f1  rtr_load_ply / rtr_bin_cells      <- loadPLY, computeGrid       cloudreader.cpp:122-177, 8-82
//   f2  rtr_load_oct / rtr_io_*_oct       <- read/writeOctreeBinary     Octreegrid.h:53-114
//   f3  rtr_io_load_calibration/_trajectory <- loadCalibration(file), parseTrajectoryLine
//                                            CameraCalibration.cpp:101-209, example/render_trajectory/main.cpp:20-65
//   f4  rtr_postprocess_unet_output       <- permute + convertTo(CV_8UC3, 255)   project_cloud.cu:475-480
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <limits>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include <cub/device/device_radix_sort.cuh>

#include "../../include/rtr_b200_io.h"
#include "rtr_internal.h"

using namespace rtr;

namespace {

#define IO_CUDA(r, call)                                                                                  \
    do {                                                                                                  \
        cudaError_t e__ = (call);                                                                         \
        if (e__ != cudaSuccess) return renderer_fail((r), RTR_ERR_CUDA, std::string(#call ": ") + cudaGetErrorString(e__)); \
    } while (0)

// ---------------------------------------------------------------- computeGrid's geometry (cloudreader.cpp:10-45)
struct GridGeom {
    float mn[3], mx[3];  // bounding box rounded outwards to whole metres
    int nb[3];           // numBlocks_x/y/z = int((max - min) / 0.25f)
};
__host__ __device__ inline int cell_coord(float p, float mn, float mx, int nb) {
    // std::floor((pt.x - bbMin.x) / (bbMax.x - bbMin.x) * numBlocks_x), all in float (cloudreader.cpp:51)
#if defined(__CUDA_ARCH__)
    return int(floorf(__fmul_rn(__fdiv_rn(__fsub_rn(p, mn), __fsub_rn(mx, mn)), float(nb))));
#else
    volatile float a = p - mn, b = mx - mn;
    volatile float q = a / b;
    volatile float s = q * float(nb);
    return int(std::floor(s));
#endif
}
__host__ __device__ inline int cell_key(const float* p, const GridGeom& g) {
    const int x = cell_coord(p[0], g.mn[0], g.mx[0], g.nb[0]);
    const int y = cell_coord(p[1], g.mn[1], g.mx[1], g.nb[1]);
    const int z = cell_coord(p[2], g.mn[2], g.mx[2], g.nb[2]);
    return x + y * g.nb[0] + z * g.nb[0] * g.nb[1];  // OctreeGrid::encodeKey, Octreegrid.h:48-50
}
GridGeom make_geom(const float lo[3], const float hi[3]) {
    GridGeom g;
    for (int a = 0; a < 3; ++a) {
        g.mx[a] = std::ceil(hi[a]);
        g.mn[a] = std::floor(lo[a]);
        g.nb[a] = int((g.mx[a] - g.mn[a]) / 0.25f);
    }
    return g;
}

// ---------------------------------------------------------------- GPU binning kernels
// min / max of the coordinates, the reference's way: `if (pt.x < bbMin.x)` / `if (pt.x > bbMax.x)` starting
// from FLT_MAX and numeric_limits<float>::min() (the smallest POSITIVE float — cloudreader.cpp:12-13), NaN never wins.
__global__ void __launch_bounds__(256) bbox_kernel(const PointRecord* __restrict__ pts, uint64_t n, float* __restrict__ out6) {
    float lo[3] = {3.402823466e38f, 3.402823466e38f, 3.402823466e38f};
    float hi[3] = {1.175494351e-38f, 1.175494351e-38f, 1.175494351e-38f};
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += uint64_t(gridDim.x) * blockDim.x) {
        const PointRecord p = pts[i];
        const float c[3] = {p.x, p.y, p.z};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            if (c[a] < lo[a]) lo[a] = c[a];
            if (c[a] > hi[a]) hi[a] = c[a];
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xFFFFFFFFu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xFFFFFFFFu, hi[a], o));
        }
        if ((threadIdx.x & 31) == 0) {  // positive-or-negative floats: order-preserving integer atomics
            const int l = __float_as_int(lo[a]), h = __float_as_int(hi[a]);
            if (l >= 0) atomicMin(reinterpret_cast<int*>(out6 + a), l); else atomicMax(reinterpret_cast<unsigned*>(out6 + a), unsigned(l));
            if (h >= 0) atomicMax(reinterpret_cast<int*>(out6 + 3 + a), h); else atomicMin(reinterpret_cast<unsigned*>(out6 + 3 + a), unsigned(h));
        }
    }
}
__global__ void __launch_bounds__(256) cell_keys_kernel(const PointRecord* __restrict__ pts, uint64_t n, GridGeom g,
                                                        uint32_t* __restrict__ keys, uint32_t* __restrict__ idx) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const PointRecord p = pts[i];
    const float c[3] = {p.x, p.y, p.z};
    keys[i] = uint32_t(cell_key(c, g)) ^ 0x80000000u;  // signed order for the unsigned radix sort
    idx[i] = uint32_t(i);
}
// 63-bit Morton code of the position quantised to 21 bits per axis over the bounding box (rtr_sort_morton).
__device__ __forceinline__ unsigned long long spread21(unsigned long long v) {
    v &= 0x1FFFFFull;
    v = (v | (v << 32)) & 0x1F00000000FFFFull;
    v = (v | (v << 16)) & 0x1F0000FF0000FFull;
    v = (v | (v << 8)) & 0x100F00F00F00F00Full;
    v = (v | (v << 4)) & 0x10C30C30C30C30C3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
}
__global__ void __launch_bounds__(256) morton_keys_kernel(const PointRecord* __restrict__ pts, uint64_t n, GridGeom g,
                                                          unsigned long long* __restrict__ keys, uint32_t* __restrict__ idx) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const PointRecord p = pts[i];
    const float c[3] = {p.x, p.y, p.z};
    unsigned long long key = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float t = (c[a] - g.mn[a]) / (g.mx[a] - g.mn[a]);
        t = t >= 0.0f ? (t <= 1.0f ? t : 1.0f) : 0.0f;  // NaN -> 0
        key |= spread21(static_cast<unsigned long long>(t * 2097151.0f)) << a;
    }
    keys[i] = key;
    idx[i] = uint32_t(i);
}

__global__ void __launch_bounds__(256) gather_kernel(const PointRecord* __restrict__ src, const uint32_t* __restrict__ idx,
                                                     uint64_t n, PointRecord* __restrict__ dst) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}

// ---------------------------------------------------------------- f4 kernel
// out[(y*W + x)*3 + c] = saturate_u8(round_half_even(float(in[c][y][x]) * 255))
__global__ void __launch_bounds__(256) unet_post_kernel(const __half* __restrict__ chw, uint64_t n_px, uint8_t* __restrict__ hwc) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n_px) return;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float v = __fmul_rn(__half2float(chw[uint64_t(c) * n_px + i]), 255.0f);
        int q = __float2int_rn(v);  // NaN -> 0, like saturate_cast<uchar>(cvRound(NaN))
        q = q < 0 ? 0 : (q > 255 ? 255 : q);
        hwc[i * 3 + c] = uint8_t(q);
    }
}

// ---------------------------------------------------------------- PLY header
struct PlyProp { std::string name; int size; bool is_float; bool is_list; };
int ply_type_size(const std::string& t, bool& is_float) {
    is_float = (t == "float" || t == "float32" || t == "double" || t == "float64");
    if (t == "char" || t == "uchar" || t == "int8" || t == "uint8") return 1;
    if (t == "short" || t == "ushort" || t == "int16" || t == "uint16") return 2;
    if (t == "int" || t == "uint" || t == "int32" || t == "uint32" || t == "float" || t == "float32") return 4;
    if (t == "double" || t == "float64") return 8;
    return 0;
}

int upload_records(rtr_renderer* r, const std::vector<PointRecord>& rec) {
    int rc = replace_cloud(r, rec.size());
    if (rc != RTR_OK || rec.empty()) return rc;
    IO_CUDA(r, cudaMemcpyAsync(r->points, rec.data(), rec.size() * sizeof(PointRecord), cudaMemcpyHostToDevice, r->stream));
    IO_CUDA(r, sync_compute(r));
    return finish_upload(r);
}

inline uint32_t pack_bgr(uint8_t b, uint8_t g, uint8_t rr) {
    return uint32_t(b) | (uint32_t(g) << 8) | (uint32_t(rr) << 16) | 0xFF000000u;  // Octreegrid.h:176
}

}  // namespace

namespace {
// Bytes between the stream's read position and its end (0 when the stream cannot seek).
uint64_t bytes_left(std::ifstream& f) {
    const std::streampos here = f.tellg();
    if (here < 0) return 0;
    f.seekg(0, std::ios::end);
    const std::streampos end = f.tellg();
    f.seekg(here);
    return end > here ? uint64_t(end - here) : 0;
}
// No exception may cross the C ABI (std::terminate in the host application): header fields of a corrupt or
// truncated file are bounded by the file size before anything is allocated, and whatever still throws
// (bad_alloc, length_error) becomes RTR_ERR_ARG.
template <typename F>
int io_guard(rtr_renderer* r, F&& body) {
    try {
        return body();
    } catch (const std::exception& e) {
        return renderer_fail(r, RTR_ERR_ARG, std::string("I/O failed: ") + e.what());
    } catch (...) {
        return renderer_fail(r, RTR_ERR_ARG, "I/O failed");
    }
}
int load_ply_impl(rtr_renderer* r, const char* path, int bin_cells);
int read_oct_impl(const char* path, float** xyz, uint8_t** bgr, uint64_t* n, int* header4, int** keys, uint64_t** counts);
int write_oct_impl(const char* path, const float* xyz, const uint8_t* bgr, uint64_t n);
}  // namespace

extern "C" {

// =============================================================== f1: PLY
int rtr_load_ply(rtr_renderer* r, const char* path, int bin_cells) {
    if (!r || !path) return RTR_ERR_ARG;
    return io_guard(r, [&] { return load_ply_impl(r, path, bin_cells); });
}
}  // extern "C"

namespace {
int load_ply_impl(rtr_renderer* r, const char* path, int bin_cells) {
    std::ifstream f(path, std::ios::binary);
    if (!f.is_open()) return renderer_fail(r, RTR_ERR_ARG, std::string("cannot open ") + path);
    std::string line, fmt;
    if (!std::getline(f, line) || line.substr(0, 3) != "ply") return renderer_fail(r, RTR_ERR_ARG, "not a PLY file");
    std::vector<PlyProp> props;
    uint64_t n_vertex = 0;
    bool in_vertex = false, seen_vertex = false, other_before = false;
    while (std::getline(f, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        std::istringstream is(line);
        std::string tok;
        is >> tok;
        if (tok == "format") is >> fmt;
        else if (tok == "element") {
            std::string name;
            uint64_t cnt = 0;
            is >> name >> cnt;
            in_vertex = (name == "vertex");
            if (in_vertex) { n_vertex = cnt; seen_vertex = true; }
            else if (!seen_vertex && cnt > 0) other_before = true;
        } else if (tok == "property" && in_vertex) {
            std::string t;
            is >> t;
            PlyProp p{"", 0, false, false};
            if (t == "list") { p.is_list = true; }
            else { p.size = ply_type_size(t, p.is_float); is >> p.name; }
            props.push_back(p);
        } else if (tok == "end_header") break;
    }
    if (!seen_vertex) return renderer_fail(r, RTR_ERR_ARG, "PLY has no vertex element");
    if (other_before) return renderer_fail(r, RTR_ERR_UNSUPPORTED, "PLY elements before 'vertex' are not supported");
    int off[6] = {-1, -1, -1, -1, -1, -1}, sz[6] = {0}, idx[6] = {-1, -1, -1, -1, -1, -1}, stride = 0;
    const char* want[6] = {"x", "y", "z", "red", "green", "blue"};
    for (size_t k = 0; k < props.size(); ++k) {
        if (props[k].is_list || props[k].size == 0) return renderer_fail(r, RTR_ERR_UNSUPPORTED, "unsupported PLY vertex property type");
        for (int w = 0; w < 6; ++w)
            if (props[k].name == want[w]) { off[w] = stride; sz[w] = props[k].size; idx[w] = int(k); }
        stride += props[k].size;
    }
    if (off[0] < 0 || off[1] < 0 || off[2] < 0) return renderer_fail(r, RTR_ERR_ARG, "PLY vertex lacks x/y/z");  // cloudreader.cpp:139-144
    for (int w = 0; w < 3; ++w)
        if (!props[idx[w]].is_float) return renderer_fail(r, RTR_ERR_UNSUPPORTED, "PLY x/y/z must be float or double");
    const bool has_rgb = off[3] >= 0 && off[4] >= 0 && off[5] >= 0 && sz[3] == 1 && sz[4] == 1 && sz[5] == 1;
    if (n_vertex > 0xFFFFFFFFull) return renderer_fail(r, RTR_ERR_UNSUPPORTED, "more than 2^32 points");
    {   // the header's vertex count against what the file can hold (binary: stride bytes, ascii: >= 2 bytes per value)
        const uint64_t left = bytes_left(f);
        const uint64_t need = fmt == "ascii" ? n_vertex * 2 * props.size() : n_vertex * uint64_t(stride);
        if (n_vertex && (stride == 0 || need > left)) return renderer_fail(r, RTR_ERR_ARG, "PLY truncated: the header promises more vertices than the file holds");
    }
    std::vector<PointRecord> rec(n_vertex);
    if (fmt == "binary_little_endian") {
        std::vector<char> buf(size_t(stride) * std::min<uint64_t>(n_vertex, 1u << 20));
        uint64_t done = 0;
        while (done < n_vertex) {
            const uint64_t m = std::min<uint64_t>(n_vertex - done, 1u << 20);
            f.read(buf.data(), std::streamsize(m * stride));
            if (uint64_t(f.gcount()) != m * stride) return renderer_fail(r, RTR_ERR_ARG, "PLY truncated");
            for (uint64_t i = 0; i < m; ++i) {
                const char* v = buf.data() + i * stride;
                float c[3];
                for (int w = 0; w < 3; ++w) {
                    if (sz[w] == 4) std::memcpy(&c[w], v + off[w], 4);
                    else { double d; std::memcpy(&d, v + off[w], 8); c[w] = float(d); }
                }
                PointRecord& p = rec[done + i];
                p.x = c[0]; p.y = c[1]; p.z = c[2];
                p.bgra = has_rgb ? pack_bgr(uint8_t(v[off[5]]), uint8_t(v[off[4]]), uint8_t(v[off[3]])) : 0xFF000000u;  // BGR, cloudreader.cpp:168
            }
            done += m;
        }
    } else if (fmt == "ascii") {
        std::vector<double> vals(props.size());
        for (uint64_t i = 0; i < n_vertex; ++i) {
            for (size_t k = 0; k < props.size(); ++k)
                if (!(f >> vals[k])) return renderer_fail(r, RTR_ERR_ARG, "PLY truncated");
            PointRecord& p = rec[i];
            p.x = float(vals[idx[0]]); p.y = float(vals[idx[1]]); p.z = float(vals[idx[2]]);
            p.bgra = has_rgb ? pack_bgr(uint8_t(vals[idx[5]]), uint8_t(vals[idx[4]]), uint8_t(vals[idx[3]])) : 0xFF000000u;
        }
    } else {
        return renderer_fail(r, RTR_ERR_UNSUPPORTED, "PLY format '" + fmt + "' not supported (ascii, binary_little_endian)");
    }
    IO_CUDA(r, cudaSetDevice(r->device));
    const int sort_opt = r->sort_on_upload;
    if (bin_cells) r->sort_on_upload = 0;  // the caller asked for the reference's grouping instead
    int rc = upload_records(r, rec);
    r->sort_on_upload = sort_opt;
    if (rc != RTR_OK || !bin_cells) return rc;
    return rtr_bin_cells(r, nullptr);
}
}  // namespace

extern "C" {

int rtr_io_write_ply(const char* path, const float* xyz, const uint8_t* bgr, uint64_t n) {
    if (!path || (n && (!xyz || !bgr))) return RTR_ERR_ARG;
    std::ofstream f(path, std::ios::binary);
    if (!f.is_open()) return renderer_fail(nullptr, RTR_ERR_ARG, std::string("cannot open ") + path);
    f << "ply\nformat binary_little_endian 1.0\nelement vertex " << n
      << "\nproperty float x\nproperty float y\nproperty float z\nproperty uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n";
    std::vector<char> buf(15 * std::min<uint64_t>(n, 1u << 20));
    for (uint64_t done = 0; done < n;) {
        const uint64_t m = std::min<uint64_t>(n - done, 1u << 20);
        for (uint64_t i = 0; i < m; ++i) {
            char* v = buf.data() + i * 15;
            std::memcpy(v, xyz + (done + i) * 3, 12);
            const uint8_t* c = bgr + (done + i) * 3;
            v[12] = char(c[2]); v[13] = char(c[1]); v[14] = char(c[0]);
        }
        f.write(buf.data(), std::streamsize(m * 15));
        done += m;
    }
    return f.good() ? RTR_OK : renderer_fail(nullptr, RTR_ERR_ARG, "write failed");
}

static int reorder_cloud(rtr_renderer* r, int* dims3, bool morton);
}  // extern "C"
namespace rtr {
int reorder_morton(rtr_renderer* r) { return reorder_cloud(r, nullptr, true); }
}
extern "C" {
int rtr_bin_cells(rtr_renderer* r, int* dims3) { return reorder_cloud(r, dims3, false); }
int rtr_sort_morton(rtr_renderer* r) { return reorder_cloud(r, nullptr, true); }

static int reorder_cloud(rtr_renderer* r, int* dims3, bool morton) {
    if (!r) return RTR_ERR_ARG;
    if (!r->points || r->n_points == 0) return renderer_fail(r, RTR_ERR_STATE, "no cloud uploaded");
    if (!r->owns_points) return renderer_fail(r, RTR_ERR_STATE, "cannot re-order an adopted device cloud");
    IO_CUDA(r, cudaSetDevice(r->device));
    const uint64_t n = r->n_points;
    cudaStream_t s = r->stream;
    float* d_box = nullptr;
    IO_CUDA(r, cudaMalloc(reinterpret_cast<void**>(&d_box), 24));
    const float init[6] = {3.402823466e38f, 3.402823466e38f, 3.402823466e38f, 1.175494351e-38f, 1.175494351e-38f, 1.175494351e-38f};
    IO_CUDA(r, cudaMemcpyAsync(d_box, init, 24, cudaMemcpyHostToDevice, s));
    bbox_kernel<<<r->sm_count * 8, 256, 0, s>>>(r->points, n, d_box);
    float box[6];
    IO_CUDA(r, cudaMemcpyAsync(box, d_box, 24, cudaMemcpyDeviceToHost, s));
    IO_CUDA(r, cudaStreamSynchronize(s));
    cudaFree(d_box);
    const GridGeom g = make_geom(box, box + 3);
    if (dims3) { dims3[0] = g.nb[0]; dims3[1] = g.nb[1]; dims3[2] = g.nb[2]; }
    uint32_t *keys = nullptr, *keys2 = nullptr, *idx = nullptr, *idx2 = nullptr;  // keys: u32, or u64 for Morton codes
    unsigned long long* k64 = nullptr;
    unsigned long long* k64b = nullptr;
    PointRecord* sorted = nullptr;
    void* tmp = nullptr;
    size_t tmp_bytes = 0;
    const size_t kb = morton ? 8 : 4;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&keys), n * kb);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&keys2), n * kb);
    k64 = reinterpret_cast<unsigned long long*>(keys);
    k64b = reinterpret_cast<unsigned long long*>(keys2);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&idx), n * 4);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&idx2), n * 4);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&sorted), n * sizeof(PointRecord));
    if (e == cudaSuccess)
        e = morton ? cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k64, k64b, idx, idx2, n, 0, 63, s)
                   : cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys2, idx, idx2, n, 0, 32, s);
    if (e == cudaSuccess) e = cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16);
    if (e == cudaSuccess) {
        const unsigned grid = unsigned((n + 255) / 256);
        if (morton) {
            morton_keys_kernel<<<grid, 256, 0, s>>>(r->points, n, g, k64, idx);
            e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k64, k64b, idx, idx2, n, 0, 63, s);
        } else {
            cell_keys_kernel<<<grid, 256, 0, s>>>(r->points, n, g, keys, idx);
            e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys2, idx, idx2, n, 0, 32, s);  // LSD radix sort: stable
        }
        if (e == cudaSuccess) {
            gather_kernel<<<grid, 256, 0, s>>>(r->points, idx2, n, sorted);
            e = cudaStreamSynchronize(s);
        }
        r->launches += 3;
    }
    cudaFree(keys); cudaFree(keys2); cudaFree(idx); cudaFree(idx2); cudaFree(tmp);
    if (e != cudaSuccess) { cudaFree(sorted); return renderer_fail(r, RTR_ERR_CUDA, std::string("rtr_bin_cells: ") + cudaGetErrorString(e)); }
    cudaFree(r->points);
    r->points = sorted;
    free_cull_storage(r);
    return build_chunk_bounds(r);
}

// =============================================================== f2: .oct cache
int rtr_io_read_oct(const char* path, float** xyz, uint8_t** bgr, uint64_t* n, int* header4, int** keys, uint64_t** counts) {
    if (!path || !xyz || !bgr || !n) return RTR_ERR_ARG;
    *xyz = nullptr; *bgr = nullptr; *n = 0;
    if (keys) *keys = nullptr;
    if (counts) *counts = nullptr;
    const int rc = io_guard(nullptr, [&] { return read_oct_impl(path, xyz, bgr, n, header4, keys, counts); });
    if (rc != RTR_OK) {  // nothing half-filled leaves the call
        std::free(*xyz); std::free(*bgr);
        *xyz = nullptr; *bgr = nullptr; *n = 0;
        if (keys) { std::free(*keys); *keys = nullptr; }
        if (counts) { std::free(*counts); *counts = nullptr; }
    }
    return rc;
}
}  // extern "C"
namespace {
int read_oct_impl(const char* path, float** xyz, uint8_t** bgr, uint64_t* n, int* header4, int** keys, uint64_t** counts) {
    std::ifstream f(path, std::ios::binary);
    if (!f.is_open()) return renderer_fail(nullptr, RTR_ERR_ARG, std::string("cannot open ") + path);
    int32_t hdr[4];
    f.read(reinterpret_cast<char*>(hdr), 16);
    if (f.gcount() != 16 || hdr[3] < 0) return renderer_fail(nullptr, RTR_ERR_ARG, "bad .oct header");
    // every block costs at least its 36-byte frame (key, count, bounds): bound the block count by the file size
    if (uint64_t(hdr[3]) * 36u > bytes_left(f)) return renderer_fail(nullptr, RTR_ERR_ARG, "bad .oct header: more blocks than the file can hold");
    if (header4) std::memcpy(header4, hdr, 16);
    std::vector<float> P;
    std::vector<uint8_t> Cc;
    std::vector<int> K(hdr[3]);
    std::vector<uint64_t> Cn(hdr[3]);
    for (int b = 0; b < hdr[3]; ++b) {
        int32_t key;
        uint64_t cnt;  // size_t in the reference (Octreegrid.h:71): 8 bytes on every platform it builds for
        f.read(reinterpret_cast<char*>(&key), 4);
        f.read(reinterpret_cast<char*>(&cnt), 8);
        if (!f.good() || cnt > (1ull << 40) || cnt * 15u + 24u > bytes_left(f)) return renderer_fail(nullptr, RTR_ERR_ARG, "bad .oct block header or truncated file");
        const size_t p0 = P.size(), c0 = Cc.size();
        P.resize(p0 + cnt * 3);
        Cc.resize(c0 + cnt * 3);
        f.read(reinterpret_cast<char*>(P.data() + p0), std::streamsize(cnt * 12));
        f.read(reinterpret_cast<char*>(Cc.data() + c0), std::streamsize(cnt * 3));
        float bb[6];
        f.read(reinterpret_cast<char*>(bb), 24);
        if (f.gcount() != 24) return renderer_fail(nullptr, RTR_ERR_ARG, ".oct truncated");
        K[b] = key;
        Cn[b] = cnt;
    }
    *n = P.size() / 3;
    *xyz = static_cast<float*>(std::malloc(P.size() * 4 + 16));
    *bgr = static_cast<uint8_t*>(std::malloc(Cc.size() + 16));
    if (!*xyz || !*bgr) return renderer_fail(nullptr, RTR_ERR_ARG, "out of memory");
    std::memcpy(*xyz, P.data(), P.size() * 4);
    std::memcpy(*bgr, Cc.data(), Cc.size());
    if (keys) {
        *keys = static_cast<int*>(std::malloc(K.size() * 4 + 16));
        if (!*keys) return renderer_fail(nullptr, RTR_ERR_ARG, "out of memory");
        std::memcpy(*keys, K.data(), K.size() * 4);
    }
    if (counts) {
        *counts = static_cast<uint64_t*>(std::malloc(Cn.size() * 8 + 16));
        if (!*counts) return renderer_fail(nullptr, RTR_ERR_ARG, "out of memory");
        std::memcpy(*counts, Cn.data(), Cn.size() * 8);
    }
    return RTR_OK;
}
}  // namespace
extern "C" {
void rtr_io_free(void* p) { std::free(p); }

int rtr_io_write_oct(const char* path, const float* xyz, const uint8_t* bgr, uint64_t n) {
    if (!path || (n && (!xyz || !bgr))) return RTR_ERR_ARG;
    return io_guard(nullptr, [&] { return write_oct_impl(path, xyz, bgr, n); });
}
}  // extern "C"
namespace {
int write_oct_impl(const char* path, const float* xyz, const uint8_t* bgr, uint64_t n) {
    // computeGrid (cloudreader.cpp:10-60)
    float lo[3] = {std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max()};
    float hi[3] = {std::numeric_limits<float>::min(), std::numeric_limits<float>::min(), std::numeric_limits<float>::min()};
    for (uint64_t i = 0; i < n; ++i)
        for (int a = 0; a < 3; ++a) {
            const float v = xyz[i * 3 + a];
            if (v < lo[a]) lo[a] = v;
            if (v > hi[a]) hi[a] = v;
        }
    const GridGeom g = make_geom(lo, hi);
    std::map<int, std::vector<uint64_t>> grid;  // block order in the file is free; ascending keys here
    for (uint64_t i = 0; i < n; ++i) grid[cell_key(xyz + i * 3, g)].push_back(i);
    std::ofstream f(path, std::ios::binary);
    if (!f.is_open()) return renderer_fail(nullptr, RTR_ERR_ARG, std::string("cannot open ") + path);
    const int32_t hdr[4] = {g.nb[0], g.nb[1], g.nb[2], int32_t(grid.size())};
    f.write(reinterpret_cast<const char*>(hdr), 16);
    std::vector<float> P;
    std::vector<uint8_t> Cc;
    for (const auto& kv : grid) {
        const int32_t key = kv.first;
        const uint64_t cnt = kv.second.size();
        f.write(reinterpret_cast<const char*>(&key), 4);
        f.write(reinterpret_cast<const char*>(&cnt), 8);
        P.resize(cnt * 3);
        Cc.resize(cnt * 3);
        for (uint64_t j = 0; j < cnt; ++j) {
            std::memcpy(P.data() + j * 3, xyz + kv.second[j] * 3, 12);
            std::memcpy(Cc.data() + j * 3, bgr + kv.second[j] * 3, 3);
        }
        f.write(reinterpret_cast<const char*>(P.data()), std::streamsize(cnt * 12));
        f.write(reinterpret_cast<const char*>(Cc.data()), std::streamsize(cnt * 3));
        // block bounds (cloudreader.cpp:62-78): decodeKey, then bbMin + idx * ((bbMax - bbMin) / numBlocks)
        int k = key;
        const int z = k / (g.nb[0] * g.nb[1]);
        k -= z * g.nb[0] * g.nb[1];
        const int y = k / g.nb[0], x = k % g.nb[0];
        const int c3[3] = {x, y, z};
        float bb[6];
        for (int a = 0; a < 3; ++a) {
            volatile float size = (g.mx[a] - g.mn[a]) / float(g.nb[a]);
            volatile float t0 = float(c3[a]) * size, t1 = float(c3[a] + 1) * size;
            bb[a] = g.mn[a] + t0;
            bb[3 + a] = g.mn[a] + t1;
        }
        f.write(reinterpret_cast<const char*>(bb), 24);
    }
    return f.good() ? RTR_OK : renderer_fail(nullptr, RTR_ERR_ARG, "write failed");
}
}  // namespace
extern "C" {

int rtr_load_oct(rtr_renderer* r, const char* path) {
    if (!r || !path) return RTR_ERR_ARG;
    float* xyz = nullptr;
    uint8_t* bgr = nullptr;
    uint64_t n = 0;
    int rc = rtr_io_read_oct(path, &xyz, &bgr, &n, nullptr, nullptr, nullptr);
    if (rc != RTR_OK) { r->err = rtr_last_error(nullptr); std::free(xyz); std::free(bgr); return rc; }
    rc = rtr_upload_cloud_xyz_bgr(r, xyz, bgr, n);  // blocks in file order == getVertexPositions' flattening
    std::free(xyz);
    std::free(bgr);
    return rc;
}

// =============================================================== f3: calibration + trajectory
int rtr_io_load_calibration(const char* path, int* width, int* height, double* K9, double* dist8, int* n_dist, int* fisheye) {
    if (!path || !width || !height || !K9 || !dist8 || !n_dist || !fisheye) return RTR_ERR_ARG;
    const std::string file(path);
    std::ifstream ifs(file);
    if (!ifs.is_open()) return renderer_fail(nullptr, RTR_ERR_ARG, "cannot open " + file);
    for (int i = 0; i < 9; ++i) K9[i] = (i % 4 == 0) ? 1.0 : 0.0;
    if (file.size() >= 11 && file.substr(file.size() - 11) == "cameras.txt") {  // CameraCalibration.cpp:103-158
        std::string line;
        while (std::getline(ifs, line)) {
            if (line.empty() || line[0] == '#') continue;
            std::istringstream iss(line);
            int id;
            std::string model;
            float fx, fy, cx, cy;  // the reference parses into float (CameraCalibration.cpp:121-122)
            iss >> id >> model >> *width >> *height;
            if (model != "OPENCV" && model != "OPENCV_FISHEYE") return renderer_fail(nullptr, RTR_ERR_UNSUPPORTED, "Unsupported camera model: " + model);
            iss >> fx >> fy >> cx >> cy;
            K9[0] = fx; K9[4] = fy; K9[2] = cx; K9[5] = cy;
            *fisheye = (model == "OPENCV_FISHEYE");
            *n_dist = *fisheye ? 4 : 5;
            for (int i = 0; i < *n_dist; ++i) { float d = 0.f; iss >> d; dist8[i] = d; }
            return RTR_OK;
        }
        return renderer_fail(nullptr, RTR_ERR_ARG, "No valid camera data found in cameras.txt");
    }
    ifs >> *width >> *height;  // CameraCalibration.cpp:167-207
    for (int i = 0; i < 9; ++i) ifs >> K9[i];
    ifs.ignore(std::numeric_limits<std::streamsize>::max(), '\n');
    std::string dl;
    std::getline(ifs, dl);
    std::replace(dl.begin(), dl.end(), ',', ' ');
    std::istringstream ds(dl);
    std::vector<double> d;
    double v;
    while (ds >> v) d.push_back(v);
    bool fe = false;
    ifs >> fe;
    *fisheye = fe ? 1 : 0;
    if (d.size() != size_t(fe ? 4 : 5))
        return renderer_fail(nullptr, RTR_ERR_ARG, std::string(fe ? "Fisheye camera expects 4" : "Pinhole camera expects 5") + " distortion parameters, got " + std::to_string(d.size()));
    *n_dist = int(d.size());
    for (size_t i = 0; i < d.size(); ++i) dist8[i] = d[i];
    return RTR_OK;
}

int rtr_io_load_trajectory(const char* path, int order, double* poses16, int max_poses, int* n_poses) {
    if (!path || !poses16 || !n_poses || max_poses < 0 || (order != 0 && order != 1)) return RTR_ERR_ARG;
    std::ifstream f(path);
    if (!f.is_open()) return renderer_fail(nullptr, RTR_ERR_ARG, std::string("cannot open ") + path);
    std::string line;
    int n = 0;
    while (std::getline(f, line)) {
        if (line.empty() || line[0] == '#') continue;  // main.cpp:58-59
        if (n >= max_poses) break;
        std::istringstream iss(line);
        double t = 0, tx = 0, ty = 0, tz = 0, qx = 0, qy = 0, qz = 0, qw = 0;
        if (order == 0) iss >> t >> tx >> ty >> tz >> qx >> qy >> qz >> qw;  // main.cpp:32
        else iss >> t >> qw >> qx >> qy >> qz >> tx >> ty >> tz;            // README.md:92 (COLMAP images.txt)
        const double nn = std::sqrt(qw * qw + qx * qx + qy * qy + qz * qz);
        if (nn > 0) { qw /= nn; qx /= nn; qy /= nn; qz /= nn; }
        double* P = poses16 + size_t(n) * 16;
        // unit quaternion -> rotation (cv::Quatd::toRotMat3x3)
        P[0] = 1 - 2 * (qy * qy + qz * qz); P[1] = 2 * (qx * qy - qw * qz);     P[2] = 2 * (qx * qz + qw * qy);     P[3] = tx;
        P[4] = 2 * (qx * qy + qw * qz);     P[5] = 1 - 2 * (qx * qx + qz * qz); P[6] = 2 * (qy * qz - qw * qx);     P[7] = ty;
        P[8] = 2 * (qx * qz - qw * qy);     P[9] = 2 * (qy * qz + qw * qx);     P[10] = 1 - 2 * (qx * qx + qy * qy); P[11] = tz;
        P[12] = 0; P[13] = 0; P[14] = 0; P[15] = 1;
        ++n;
    }
    *n_poses = n;
    return RTR_OK;
}

int rtr_io_invert_rigid(const double* p, double* o) {
    if (!p || !o) return RTR_ERR_ARG;
    double R[9] = {p[0], p[1], p[2], p[4], p[5], p[6], p[8], p[9], p[10]}, t[3] = {p[3], p[7], p[11]};
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) o[i * 4 + j] = R[j * 3 + i];
        o[i * 4 + 3] = -(R[0 * 3 + i] * t[0] + R[1 * 3 + i] * t[1] + R[2 * 3 + i] * t[2]);
    }
    o[12] = 0; o[13] = 0; o[14] = 0; o[15] = 1;
    return RTR_OK;
}

// =============================================================== f4: U-Net output post-process
int rtr_postprocess_unet_output(rtr_renderer* r, const void* device_fp16_chw, int width, int height, uint8_t* host_hwc,
                                uint8_t* device_hwc) {
    if (!r || !device_fp16_chw || width <= 0 || height <= 0 || (!host_hwc && !device_hwc)) return RTR_ERR_ARG;
    IO_CUDA(r, cudaSetDevice(r->device));
    const uint64_t n_px = uint64_t(width) * height;
    uint8_t* out = device_hwc;
    if (!out) {
        if (r->post_scratch_bytes < n_px * 3) {
            cudaFree(r->post_scratch);
            r->post_scratch = nullptr;
            r->post_scratch_bytes = 0;
            IO_CUDA(r, cudaMalloc(reinterpret_cast<void**>(&r->post_scratch), n_px * 3));
            r->post_scratch_bytes = n_px * 3;
        }
        out = r->post_scratch;
    }
    unet_post_kernel<<<unsigned((n_px + 255) / 256), 256, 0, r->stream>>>(static_cast<const __half*>(device_fp16_chw), n_px, out);
    r->launches += 1;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && host_hwc) {
        e = cudaMemcpyAsync(host_hwc, out, n_px * 3, cudaMemcpyDeviceToHost, r->stream);
        if (e == cudaSuccess) e = sync_compute(r);
    }
    if (e != cudaSuccess) return renderer_fail(r, RTR_ERR_CUDA, std::string("rtr_postprocess_unet_output: ") + cudaGetErrorString(e));
    return RTR_OK;
}

}  // extern "C"
