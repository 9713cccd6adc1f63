// Drives the header-only C++ adapter (include/rtr_b200/project_cloud.hpp) the way the reference's example drives its
// ProjectCloud (/root/reference/example/render_trajectory/main.cpp:87-96): a grid of reference-style blocks
// (Octreegrid.h:16-21) -> fromGrid -> computeRGBD / computeFilteredRGBD(calibration, world->camera Matx44d, &rgb, &depth)
// with pre-allocated cv::Mat outputs.  The look-alike types below have the members the reference's have; OpenCV's
// C++ headers are not in the image, so <opencv2/core.hpp> comes from the stand-in under oracle/stubs (test
// infrastructure).  tests/test_gpu_adapter.py builds this, feeds it a cloud file and compares the frames it writes.
//
//   adapter_main <cloud.bin> <W> <H> <fx> <fy> <cx> <cy> <pose16.bin> <out_prefix>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include <opencv2/core.hpp>

#include "rtr_b200/project_cloud.hpp"

#ifndef RTR_B200_HAVE_OPENCV
#error "the cv::Mat overloads of the adapter need <opencv2/core.hpp>"
#endif

namespace OctreeGrid {          // the members of the reference's OctreeGrid::Block (Octreegrid.h:16-21)
struct Block {
    std::vector<cv::Point3f> positions;
    std::vector<cv::Vec3b> colors;
    cv::Point3f bbMin, bbMax;
};
}  // namespace OctreeGrid

class CameraCalibration {      // the accessors the path reads (CameraCalibration.h:8-54)
public:
    bool loadCalibration(double fx, double fy, double cx, double cy, const std::vector<double>& dist, int w, int h) {
        K_ = cv::Matx33d::eye();
        K_(0, 0) = fx; K_(1, 1) = fy; K_(0, 2) = cx; K_(1, 2) = cy;
        dist_ = dist; w_ = w; h_ = h;
        return true;
    }
    int getWidth() const { return w_; }
    int getHeight() const { return h_; }
    cv::Matx33d getIntrinsicsMatrix() const { return K_; }
    std::vector<double> getDistortionParameters() const { return dist_; }
private:
    cv::Matx33d K_;
    std::vector<double> dist_;
    int w_ = 640, h_ = 480;
};

static bool write_file(const std::string& path, const void* p, size_t bytes) {
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    const bool ok = std::fwrite(p, 1, bytes, f) == bytes;
    std::fclose(f);
    return ok;
}

int main(int argc, char** argv) {
    if (argc != 10) { std::fprintf(stderr, "usage: adapter_main cloud.bin W H fx fy cx cy pose16.bin out_prefix\n"); return 64; }
    const int W = std::atoi(argv[2]), H = std::atoi(argv[3]);
    // cloud file: n x {x, y, z, bgra} 16-byte records; dealt into 0.25 m cells like CloudReader::computeGrid
    FILE* f = std::fopen(argv[1], "rb");
    if (!f) return 65;
    std::unordered_map<int, OctreeGrid::Block> grid;
    struct Rec { float x, y, z; unsigned char b, g, r, a; } rec;
    size_t n = 0;
    while (std::fread(&rec, sizeof(rec), 1, f) == 1) {
        const int cx = int(rec.x * 4.0f), cy = int(rec.y * 4.0f), cz = int(rec.z * 4.0f);
        OctreeGrid::Block& b = grid[(cz * 1000 + cy) * 1000 + cx];
        b.positions.emplace_back(rec.x, rec.y, rec.z);
        b.colors.emplace_back(rec.b, rec.g, rec.r);
        ++n;
    }
    std::fclose(f);
    cv::Matx44d pose;
    f = std::fopen(argv[8], "rb");
    if (!f || std::fread(pose.val, sizeof(double), 16, f) != 16) return 66;
    std::fclose(f);
    CameraCalibration calibration;
    calibration.loadCalibration(std::atof(argv[4]), std::atof(argv[5]), std::atof(argv[6]), std::atof(argv[7]), std::vector<double>(5, 0.0), W, H);
    try {
        std::shared_ptr<rtr_b200::ProjectCloud> projector(rtr_b200::ProjectCloud::fromGrid(grid));
        const std::string out = argv[9];
        {   // example/render_trajectory/main.cpp:92-96
            cv::Mat rgb = cv::Mat(cv::Size(calibration.getWidth(), calibration.getHeight()), CV_8UC3);
            cv::Mat depth = cv::Mat(cv::Size(calibration.getWidth(), calibration.getHeight()), CV_32F);
            if (projector->computeRGBD(calibration, pose, &rgb, &depth) != 1) { std::fprintf(stderr, "computeRGBD: %s\n", projector->lastError()); return 2; }
            if (!write_file(out + "_raw_color.bin", rgb.ptr<uint8_t>(), size_t(W) * H * 3) || !write_file(out + "_raw_depth.bin", depth.ptr<float>(), size_t(W) * H * 4)) return 67;
        }
        {
            cv::Mat rgb = cv::Mat(cv::Size(calibration.getWidth(), calibration.getHeight()), CV_8UC3);
            cv::Mat depth = cv::Mat(cv::Size(calibration.getWidth(), calibration.getHeight()), CV_32F);
            if (projector->computeFilteredRGBD(calibration, pose, &rgb, &depth) != 1) { std::fprintf(stderr, "computeFilteredRGBD: %s\n", projector->lastError()); return 3; }
            if (!write_file(out + "_flt_color.bin", rgb.ptr<uint8_t>(), size_t(W) * H * 3) || !write_file(out + "_flt_depth.bin", depth.ptr<float>(), size_t(W) * H * 4)) return 67;
            // one output only, like cloudreader.cpp:233-246 (colour without depth)
            cv::Mat only = cv::Mat(cv::Size(W, H), CV_8UC3);
            if (projector->computeFilteredRGBD(calibration, pose, &only, nullptr) != 1) return 4;
            if (std::memcmp(only.ptr<uint8_t>(), rgb.ptr<uint8_t>(), size_t(W) * H * 3) != 0) return 5;
            if (projector->computeRGBD(calibration, pose, static_cast<cv::Mat*>(nullptr), static_cast<cv::Mat*>(nullptr)) != -1) return 6;   // project_cloud.cu:270-273
        }
        {   // computeFull with a caller-supplied neural stage: identity on the first three planes of the tensor
            std::vector<uint8_t> color(size_t(W) * H * 3);
            std::vector<float> depth(size_t(W) * H);
            auto net = [](void* tensor, int, int) -> const void* { return tensor; };
            rtr_b200::Intrinsics k = rtr_b200::ProjectCloud::fromCalib(calibration);
            if (projector->computeFull(k, pose.val, color.data(), depth.data(), net) != 1) return 7;
            if (!write_file(out + "_full_color.bin", color.data(), color.size())) return 67;
        }
        std::printf("ADAPTER_OK %zu points in %zu blocks\n", n, grid.size());
    } catch (const std::exception& e) {
        std::fprintf(stderr, "exception: %s\n", e.what());
        return 1;
    }
    return 0;
}
