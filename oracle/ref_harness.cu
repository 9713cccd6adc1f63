// TEST INFRASTRUCTURE — GPU oracle.  Not part of the product; nothing under the package links this.
//
// C-ABI shim around the UNMODIFIED reference renderer.  oracle/Makefile compiles the reference's
// own translation units where they lie under /root/reference/src/RTRenderer/src
// (render.cu, project_cloud.cu, CameraCalibration.cpp) against the stand-in glm/OpenCV headers in
// oracle/stubs/, and links them with this file into oracle/_ref/libref_rtrenderer.so.
// Everything that is computed here is computed by the reference's own code:
//   ProjectCloud::ProjectCloud            project_cloud.cu:189-251
//   ProjectCloud::computeRGBD             project_cloud.cu:268-312
//   ProjectCloud::computeFilteredRGBD     project_cloud.cu:394-434
//   minDepthPass / accumulatePass / ...   render.cu:53-163   (for per-kernel timing)
// This file only (a) builds the `unordered_map<int, OctreeGrid::Block>` the constructor wants,
// (b) copies the reference object's private device buffers out for stage-by-stage comparison and
// (c) times the reference kernels with CUDA events using the reference's own launch geometry.
#include <cuda_runtime.h>
#if defined(__AVX2__) && defined(__F16C__)
#include <immintrin.h>
#endif
#include <torch/script.h>
#include <torch/cuda.h>

#include <cstdio>
#include <cstring>
#include <filesystem>
#include <sstream>
#include <unordered_map>
#include <vector>

#include "opencv2/core.hpp"

// The reference keeps its device buffers private; the oracle needs to read them.
#define private public
#include "project_cloud.h"
#include "PointCloudReader.h"
#include "Utils.h"
#include "cloudreader.h"
#undef private
#include "render.cuh"

// cv::Mat::convertTo stand-in: the one conversion computeFull performs (project_cloud.cu:480),
// CV_16FC3 -> CV_8UC3 with scale alpha, OpenCV semantics = saturate_cast<uchar>(cvRound(v*alpha)).
void cv::Mat::convertTo(cv::Mat& dst, int rtype, double alpha) const {
    if (type_ != CV_16FC3 || rtype != CV_8UC3) return;
    if (dst.rows != rows || dst.cols != cols || dst.type_ != rtype) dst = cv::Mat(cv::Size(cols, rows), rtype);
    const uint16_t* s = ptr<uint16_t>();
    uint8_t* d = dst.ptr<uint8_t>();
    const size_t n = size_t(rows) * cols * 3;
    size_t i = 0;
#if defined(__AVX2__) && defined(__F16C__)
    // OpenCV's own convertTo is a SIMD loop (float(src) * alpha, round to nearest even, saturate); keep the stand-in
    // equally cheap so that the timed reference is not slowed down by test scaffolding.
    const __m256 a = _mm256_set1_ps(float(alpha));
    for (; i + 16 <= n; i += 16) {
        const __m256 v0 = _mm256_mul_ps(_mm256_cvtph_ps(_mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i))), a);
        const __m256 v1 = _mm256_mul_ps(_mm256_cvtph_ps(_mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i + 8))), a);
        const __m256i q = _mm256_packs_epi32(_mm256_cvtps_epi32(v0), _mm256_cvtps_epi32(v1));  // lanes interleaved per 128 bit
        const __m256i q2 = _mm256_permute4x64_epi64(q, 0xD8);
        const __m128i b = _mm_packus_epi16(_mm256_castsi256_si128(q2), _mm256_extracti128_si256(q2, 1));
        _mm_storeu_si128(reinterpret_cast<__m128i*>(d + i), b);
    }
#endif
    const at::Half* sh = ptr<at::Half>();
    for (; i < n; ++i) {
        const float v = float(sh[i]) * float(alpha);
        const long r = lrintf(v);  // round-half-even, like cvRound
        d[i] = uint8_t(r < 0 ? 0 : r > 255 ? 255 : r);
    }
}

// cloudreader.cpp (compiled unmodified for its PLY path: loadPLY + computeGrid) also holds the E57 path, whose reader
// class lives in libE57Format (not in this image).  The members it links against are defined here and abort: nothing
// in the tests takes that path.
static void no_e57() { std::fprintf(stderr, "oracle/_ref: the E57 path is not available (libE57Format is not installed)\n"); std::abort(); }
PointCloudReader::PointCloudReader(const std::string&) : _reader(nullptr) { no_e57(); }
int PointCloudReader::getNumberOfClouds() { no_e57(); return 0; }
int PointCloudReader::getNumberOfImages() { no_e57(); return 0; }
cv::Mat PointCloudReader::getImage(int, cv::Matx44d&, cv::Matx33d&) { no_e57(); return cv::Mat(); }
void PointCloudReader::getScanCloud(int, cv::Matx44d&, std::vector<cv::Point3d>&, std::vector<cv::Vec3b>&, int) { no_e57(); }
std::vector<cv::Point3d> Utils::transformCloud(const std::vector<cv::Point3d>& cloud, cv::Matx44d) { no_e57(); return cloud; }

extern "C" cudaError_t rtr_ref_zero_malloc(void** p, size_t bytes) {
#undef cudaMalloc
    cudaError_t e = cudaMalloc(p, bytes);
    if (e == cudaSuccess && bytes) e = cudaMemset(*p, 0, bytes);
    return e;
}

namespace {
struct RefHandle {
    ProjectCloud* pc = nullptr;
    size_t n = 0;
};

CameraCalibration make_calib(int W, int H, const double* K9) {
    CameraCalibration c;
    cv::Matx33d K;
    for (int i = 0; i < 9; ++i) K.val[i] = K9[i];
    c.setIntrinsicsMatrix(K);
    c.setWidth(W);
    c.setHeight(H);
    return c;
}
cv::Matx44d make_E(const double* E16) {
    cv::Matx44d E;
    for (int i = 0; i < 16; ++i) E.val[i] = E16[i];
    return E;
}
}  // namespace

extern "C" {

// xyz: n*3 float32, bgr: n*3 uint8 (the order OctreeGrid::Block::colors holds, cloudreader.cpp:168).
// The whole cloud goes into one Block so the flattened order equals the input order
// (Octreegrid.h:162-180 iterates blocks, then points).
void* ref_create(const float* xyz, const uint8_t* bgr, size_t n) {
    std::unordered_map<int, OctreeGrid::Block> grid;
    OctreeGrid::Block& b = grid[0];
    b.positions.resize(n);
    b.colors.resize(n);
    std::memcpy(static_cast<void*>(b.positions.data()), xyz, n * 3 * sizeof(float));
    std::memcpy(static_cast<void*>(b.colors.data()), bgr, n * 3);
    // The constructor chats on stdout/stderr (project_cloud.cu:219-250); keep our stdout clean.
    std::ostringstream sink;
    std::streambuf* o = std::cout.rdbuf(sink.rdbuf());
    std::streambuf* e = std::cerr.rdbuf(sink.rdbuf());
    RefHandle* h = new RefHandle;
    h->pc = new ProjectCloud(grid, std::string(""));
    h->n = n;
    std::cout.rdbuf(o);
    std::cerr.rdbuf(e);
    return h;
}

// Same with a TorchScript model: the reference loads $HOME/.render_cache/<model_name> (project_cloud.cu:225-239)
// and computeFull (project_cloud.cu:437-493) runs it.  Returns NULL if the file is absent (the reference exit()s).
void* ref_create_with_model(const float* xyz, const uint8_t* bgr, size_t n, const char* model_name) {
    const char* home = getenv("HOME");
    if (!home || !std::filesystem::exists(std::filesystem::path(home) / ".render_cache" / model_name)) return nullptr;
    std::unordered_map<int, OctreeGrid::Block> grid;
    OctreeGrid::Block& b = grid[0];
    b.positions.resize(n);
    b.colors.resize(n);
    std::memcpy(static_cast<void*>(b.positions.data()), xyz, n * 3 * sizeof(float));
    std::memcpy(static_cast<void*>(b.colors.data()), bgr, n * 3);
    std::ostringstream sink;
    std::streambuf* o = std::cout.rdbuf(sink.rdbuf());
    std::streambuf* e = std::cerr.rdbuf(sink.rdbuf());
    RefHandle* h = new RefHandle;
    h->pc = new ProjectCloud(grid, std::string(model_name));
    h->n = n;
    std::cout.rdbuf(o);
    std::cerr.rdbuf(e);
    return h;
}

// computeFull: projection + prefilter + U-Net + 8-bit conversion, host outputs (color H*W*3 uint8, depth H*W float).
int ref_compute_full(void* hv, int W, int H, const double* K9, const double* E16, uint8_t* color, float* depth) {
    RefHandle* h = static_cast<RefHandle*>(hv);
    CameraCalibration c = make_calib(W, H, K9);
    cv::Mat mc(H, W, CV_8UC3, color), md(H, W, CV_32F, depth);
    std::ostringstream sink;  // RENDER_TIME print per call (project_cloud.cu:490)
    std::streambuf* o = std::cout.rdbuf(sink.rdbuf());
    const int rc = h->pc->computeFull(c, make_E(E16), color ? &mc : nullptr, depth ? &md : nullptr);
    std::cout.rdbuf(o);
    return rc;
}

void ref_destroy(void* hv) {
    RefHandle* h = static_cast<RefHandle*>(hv);
    if (!h) return;
    delete h->pc;
    delete h;
}

int ref_block_size(void* hv) { return static_cast<RefHandle*>(hv)->pc->block_size; }

// color: H*W*3 uint8 or null, depth: H*W float or null.  Returns the reference's return value.
int ref_compute_rgbd(void* hv, int W, int H, const double* K9, const double* E16, uint8_t* color, float* depth) {
    RefHandle* h = static_cast<RefHandle*>(hv);
    CameraCalibration c = make_calib(W, H, K9);
    cv::Mat mc(H, W, CV_8UC3, color), md(H, W, CV_32F, depth);
    return h->pc->computeRGBD(c, make_E(E16), color ? &mc : nullptr, depth ? &md : nullptr);
}

int ref_compute_filtered(void* hv, int W, int H, const double* K9, const double* E16, uint8_t* color, float* depth) {
    RefHandle* h = static_cast<RefHandle*>(hv);
    CameraCalibration c = make_calib(W, H, K9);
    cv::Mat mc(H, W, CV_8UC3, color), md(H, W, CV_32F, depth);
    return h->pc->computeFilteredRGBD(c, make_E(E16), color ? &mc : nullptr, depth ? &md : nullptr);
}

// Copy one of the reference object's device buffers to host.
// what: 0 zbuf/depth (P u32)  1 accum (4P u32)  2 image (3P u8)  3 tensor (5P f16)
//       4 cam_proj (16 f32)   5 final_min (u32) 6 final_max (u32)
int ref_read(void* hv, int what, void* dst, size_t bytes) {
    ProjectCloud* p = static_cast<RefHandle*>(hv)->pc;
    const void* src = nullptr;
    switch (what) {
        case 0: src = p->d_output_depth; break;
        case 1: src = p->d_output_color; break;
        case 2: src = p->d_image; break;
        case 3: src = p->tensorPtr; break;
        case 4: src = p->d_cam_proj; break;
        case 5: src = p->d_final_min; break;
        case 6: src = p->d_final_max; break;
        default: return -1;
    }
    cudaDeviceSynchronize();
    return cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost) == cudaSuccess ? 1 : -2;
}

// Override the camera matrix the NEXT direct kernel launches use (row-major float[16]); lets the
// parity tests feed both implementations the same raw matrix.
int ref_set_cam_proj_raw(void* hv, const float* m16) {
    ProjectCloud* p = static_cast<RefHandle*>(hv)->pc;
    return cudaMemcpy(p->d_cam_proj, m16, 64, cudaMemcpyHostToDevice) == cudaSuccess ? 1 : -2;
}

// Launch the reference's point kernels directly (same geometry as project_cloud.cu:316-325) on the
// buffers of an object that has already rendered one frame at W x H, `iters` times, and report
// the mean CUDA-event time of each kernel in ms:
//   ms[0] fillBuffer+memset  ms[1] minDepthPass  ms[2] accumulatePass  ms[3] resolvePass
int ref_time_point_kernels(void* hv, int W, int H, int iters, float* ms) {
    ProjectCloud* p = static_cast<RefHandle*>(hv)->pc;
    if (p->image_size.x != unsigned(W) || p->image_size.y != unsigned(H)) return -1;
    cudaEvent_t ev[5];
    for (auto& e : ev) cudaEventCreate(&e);
    double acc[4] = {0, 0, 0, 0};
    for (int it = 0; it < iters; ++it) {
        cudaEventRecord(ev[0]);
        fillBuffer<<<p->grid_dim, p->block_dim>>>(p->d_output_depth, 0x7F7FFFFF, W * H);
        cudaMemsetAsync(p->d_output_color, 0, size_t(W) * H * 4 * sizeof(uint32_t));
        cudaEventRecord(ev[1]);
        minDepthPass<<<p->numBlocksPCDLevel, p->block_size>>>(p->d_output_depth, p->d_vertices_data, (float*)p->d_cam_proj, p->image_size, p->data_size);
        cudaEventRecord(ev[2]);
        accumulatePass<<<p->numBlocksPCDLevel, p->block_size>>>(p->d_vertices_data, p->d_color_data, p->data_size, p->image_size, (float*)p->d_cam_proj, p->d_output_depth, p->d_output_color);
        cudaEventRecord(ev[3]);
        resolvePass<<<p->grid_dim, p->block_dim>>>(p->d_image, p->d_output_color, W * H);
        cudaEventRecord(ev[4]);
        if (cudaEventSynchronize(ev[4]) != cudaSuccess) return -2;
        for (int k = 0; k < 4; ++k) {
            float t = 0;
            cudaEventElapsedTime(&t, ev[k], ev[k + 1]);
            acc[k] += t;
        }
    }
    for (int k = 0; k < 4; ++k) ms[k] = float(acc[k] / iters);
    for (auto& e : ev) cudaEventDestroy(e);
    return cudaGetLastError() == cudaSuccess ? 1 : -3;
}

// ---- "next"-row pins (CPU only): the reference's own .oct writer/reader and calibration parser.
// Builds grid[key] from caller-supplied per-point keys and writes it with OctreeGrid::writeOctreeBinary
// (Octreegrid.h:53-79).
int ref_oct_write(const char* path, const float* xyz, const uint8_t* bgr, const int* keys, size_t n, int nx, int ny, int nz) {
    std::unordered_map<int, OctreeGrid::Block> grid;
    for (size_t i = 0; i < n; ++i) {
        OctreeGrid::Block& b = grid[keys[i]];
        b.positions.emplace_back(cv::Point3f(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]));
        b.colors.emplace_back(cv::Vec3b(bgr[3 * i], bgr[3 * i + 1], bgr[3 * i + 2]));
    }
    return OctreeGrid::writeOctreeBinary(path, grid, nx, ny, nz) ? 1 : -1;
}
// OctreeGrid::readOctreeBinary (Octreegrid.h:82-114) + the constructor's flattening (Octreegrid.h:162-180).
int ref_oct_read(const char* path, float* xyz4_out, uint8_t* bgra_out, size_t cap, size_t* n, int* dims3) {
    std::unordered_map<int, OctreeGrid::Block> grid;
    if (!OctreeGrid::readOctreeBinary(path, grid, dims3[0], dims3[1], dims3[2])) return -1;
    std::vector<float4> v = OctreeGrid::getVertexPositions(grid);
    std::vector<uchar4> c = OctreeGrid::getVertexColors(grid);
    *n = v.size();
    if (v.size() > cap) return -2;
    std::memcpy(xyz4_out, v.data(), v.size() * sizeof(float4));
    std::memcpy(bgra_out, c.data(), c.size() * sizeof(uchar4));
    return 1;
}
// CloudReader::loadCloud for a .ply, no cache (cloudreader.cpp:122-177 loadPLY -> :8-82 computeGrid): every point with the key
// of the 0.25 m block the reference put it in, blocks in the map's iteration order, points in arrival order inside a block.
int ref_load_ply(const char* path, float* xyz_out, uint8_t* bgr_out, int* key_out, size_t cap, size_t* n, size_t* n_blocks) {
    std::ostringstream sink;
    std::streambuf* o = std::cout.rdbuf(sink.rdbuf());
    const std::unordered_map<int, OctreeGrid::Block> grid = CloudReader::loadCloud(std::filesystem::path(path));
    std::cout.rdbuf(o);
    size_t k = 0;
    for (const auto& kv : grid) {
        const OctreeGrid::Block& b = kv.second;
        for (size_t i = 0; i < b.positions.size(); ++i, ++k) {
            if (k >= cap) return -2;
            xyz_out[3 * k] = b.positions[i].x; xyz_out[3 * k + 1] = b.positions[i].y; xyz_out[3 * k + 2] = b.positions[i].z;
            const cv::Vec3b c = i < b.colors.size() ? b.colors[i] : cv::Vec3b(0, 0, 0);
            bgr_out[3 * k] = c[0]; bgr_out[3 * k + 1] = c[1]; bgr_out[3 * k + 2] = c[2];
            key_out[k] = kv.first;
        }
    }
    *n = k;
    *n_blocks = grid.size();
    return 1;
}
// CameraCalibration::loadCalibration(file) (CameraCalibration.cpp:101-209).
int ref_load_calibration(const char* path, int* W, int* H, double* K9, double* dist8, int* n_dist, int* fisheye) {
    CameraCalibration c;
    std::ostringstream sink;
    std::streambuf* e = std::cerr.rdbuf(sink.rdbuf());
    const bool ok = c.loadCalibration(std::string(path));
    std::cerr.rdbuf(e);
    if (!ok) return -1;
    *W = c.getWidth();
    *H = c.getHeight();
    const cv::Matx33d K = c.getIntrinsicsMatrix();
    for (int i = 0; i < 9; ++i) K9[i] = K.val[i];
    const std::vector<double> d = c.getDistortionParameters();
    *n_dist = int(d.size());
    for (size_t i = 0; i < d.size() && i < 8; ++i) dist8[i] = d[i];
    *fisheye = c.isFishEye() ? 1 : 0;
    return 1;
}

}  // extern "C"
