// Header-only C++ adapter: the reference's ProjectCloud method names on top of the C ABI
// (include/rtr_b200.h).  A maintainer of the reference swaps
//     #include <RTRenderer/project_cloud.h>      ->      #include <rtr_b200/project_cloud.hpp>
// and links librtr_b200.so instead of (or next to) libRTRenderer.so; call sites such as
// example/render_trajectory/main.cpp:87-96 and cloudreader.cpp:233-246 compile unchanged when
// OpenCV's headers are present (the cv::Mat / CameraCalibration overloads below), and the
// raw-pointer overloads work without OpenCV.
//
//   reference                                              here
//   ProjectCloud(grid, modelFilename)   project_cloud.cu:189   ProjectCloud(xyz, bgr, n) / (grid)
//   computeRGBD(calib, E, &color, &depth)          :268       computeRGBD(...)
//   computeFilteredRGBD(...)                       :394       computeFilteredRGBD(...)
//   computeFull(...) up to the U-Net input         :437-471   computeTensor(...) -> device fp16*
// Return values follow the reference: 1 ok, -1 when both outputs are null; other failures are
// negative rtr_b200 codes with text in lastError() — never exit().
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../rtr_b200.h"
#include "../rtr_b200_io.h"

#if defined(__has_include)
#if __has_include(<opencv2/core.hpp>)
#include <opencv2/core.hpp>
#define RTR_B200_HAVE_OPENCV 1
#endif
#endif

namespace rtr_b200 {

struct Intrinsics {  // what the hot path reads from CameraCalibration (CameraCalibration.h:8-54)
    int width = 640, height = 480;
    double K[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};  // row-major
    double dist[5] = {0, 0, 0, 0, 0};           // k1 k2 p1 p2 k3; ignored unless applyDistortion(true)
};

class ProjectCloud {
public:
    // xyz: n*3 float32, bgr: n*3 uint8 in B,G,R order (what OctreeGrid::Block::colors holds).
    ProjectCloud(const float* xyz, const uint8_t* bgr, uint64_t n, int device = 0) {
        create(device);
        check(rtr_upload_cloud_xyz_bgr(h_, xyz, bgr, n));
    }
    explicit ProjectCloud(int device = 0) { create(device); }
    ~ProjectCloud() { rtr_destroy(h_); }
    ProjectCloud(const ProjectCloud&) = delete;
    ProjectCloud& operator=(const ProjectCloud&) = delete;

    // Any container of reference-style blocks: needs .second.positions[i].{x,y,z} and
    // .second.colors[i][0..2] (OctreeGrid::Block, Octreegrid.h:16-21).
    template <typename Grid>
    static ProjectCloud* fromGrid(const Grid& grid, int device = 0) {
        std::vector<float> xyz;
        std::vector<uint8_t> bgr;
        size_t n = 0;
        for (const auto& kv : grid) n += kv.second.positions.size();
        xyz.reserve(n * 3);
        bgr.reserve(n * 3);
        for (const auto& kv : grid)
            for (size_t i = 0; i < kv.second.positions.size(); ++i) {
                const auto& p = kv.second.positions[i];
                const auto& c = kv.second.colors[i];
                xyz.push_back(p.x); xyz.push_back(p.y); xyz.push_back(p.z);
                bgr.push_back(c[0]); bgr.push_back(c[1]); bgr.push_back(c[2]);
            }
        return new ProjectCloud(xyz.data(), bgr.data(), n, device);
    }

    void applyDistortion(bool on) { distort_ = on; }
    rtr_renderer* handle() const { return h_; }
    const char* lastError() const { return rtr_last_error(h_); }

    // extrinsics: world->camera 4x4 row-major doubles (cv::Matx44d::val).  color: H*W*3 uint8 (BGR),
    // depth: H*W float32, caller-allocated, either may be null.
    int computeRGBD(const Intrinsics& c, const double* extrinsics16, uint8_t* color, float* depth) {
        if (!color && !depth) return -1;
        int rc = setCamera(c, extrinsics16);
        return rc == RTR_OK ? rtr_render_rgbd(h_, color, depth) : rc;
    }
    int computeFilteredRGBD(const Intrinsics& c, const double* extrinsics16, uint8_t* color, float* depth) {
        if (!color && !depth) return -1;
        int rc = setCamera(c, extrinsics16);
        return rc == RTR_OK ? rtr_render_filtered(h_, color, depth) : rc;
    }
    // Projection + prefilter of computeFull; *tensor = device pointer of the 1x5xHxW fp16 U-Net input:
    //   torch::from_blob(tensor, {1, 5, H, W}, torch::TensorOptions().dtype(torch::kFloat16).device(torch::kCUDA))
    int computeTensor(const Intrinsics& c, const double* extrinsics16, void** tensor) {
        int rc = setCamera(c, extrinsics16);
        return rc == RTR_OK ? rtr_render_tensor(h_, tensor) : rc;
    }

    // computeFull (project_cloud.cu:437-493) with the neural stage supplied by the caller, so that this header needs
    // no libtorch: `unet(tensor, W, H)` receives the device pointer of the 1x5xHxW fp16 input and returns the device
    // pointer of the network's 3xHxW fp16 output (e.g. model.forward({from_blob(tensor, ...)}).toTensor()[0]
    // .contiguous().data_ptr(), kept alive by the caller until this returns).  color: H*W*3 uint8 =
    // saturate(round(v * 255)) like the reference's convertTo; depth: the filtered depth.  Either may be null.
    template <typename UNet>
    int computeFull(const Intrinsics& c, const double* extrinsics16, uint8_t* color, float* depth, UNet&& unet) {
        void* tensor = nullptr;
        int rc = computeTensor(c, extrinsics16, &tensor);
        if (rc != RTR_OK) return rc;
        const void* out = unet(tensor, c.width, c.height);
        if (color) {
            if (!out) return RTR_ERR_ARG;
            rc = rtr_postprocess_unet_output(h_, out, c.width, c.height, color, nullptr);
            if (rc != RTR_OK) return rc;
        }
        if (depth) rc = rtr_read_buffer(h_, 0, depth, size_t(c.width) * c.height * sizeof(float));
        return rc;
    }

#ifdef RTR_B200_HAVE_OPENCV
    // Drop-in signatures (Calib = the reference's CameraCalibration: getWidth/getHeight/getIntrinsicsMatrix/
    // getDistortionParameters).
    template <typename Calib>
    int computeRGBD(const Calib& calib, const cv::Matx44d& E, cv::Mat* color, cv::Mat* depth) {
        return computeRGBD(fromCalib(calib), E.val, color ? color->template ptr<uint8_t>() : nullptr, depth ? depth->template ptr<float>() : nullptr);
    }
    template <typename Calib>
    int computeFilteredRGBD(const Calib& calib, const cv::Matx44d& E, cv::Mat* color, cv::Mat* depth) {
        return computeFilteredRGBD(fromCalib(calib), E.val, color ? color->template ptr<uint8_t>() : nullptr, depth ? depth->template ptr<float>() : nullptr);
    }
    template <typename Calib>
    static Intrinsics fromCalib(const Calib& calib) {
        Intrinsics c;
        c.width = calib.getWidth();
        c.height = calib.getHeight();
        const cv::Matx33d K = calib.getIntrinsicsMatrix();
        for (int i = 0; i < 9; ++i) c.K[i] = K.val[i];
        const std::vector<double> d = calib.getDistortionParameters();
        for (size_t i = 0; i < 5 && i < d.size(); ++i) c.dist[i] = d[i];
        return c;
    }
#endif

private:
    void create(int device) {
        if (rtr_create(device, &h_) != RTR_OK) throw std::runtime_error(std::string("rtr_create: ") + rtr_last_error(nullptr));
    }
    void check(int rc) {
        if (rc != RTR_OK) throw std::runtime_error(std::string("rtr_b200: ") + rtr_last_error(h_));
    }
    int setCamera(const Intrinsics& c, const double* E16) {
        static const double zero[5] = {0, 0, 0, 0, 0};
        int rc = rtr_set_intrinsics_matrix(h_, c.width, c.height, c.K, distort_ ? c.dist : zero);
        return rc == RTR_OK ? rtr_set_pose_w2c(h_, E16) : rc;
    }
    rtr_renderer* h_ = nullptr;
    bool distort_ = false;
};

}  // namespace rtr_b200
