// Minimal stand-in for the OpenCV C++ types the reference's RTRenderer headers/sources touch.
// OpenCV's C++ headers are not installed in this image (only the Python module).  This is OUR
// code (test infrastructure for oracle/_ref): plain containers with the same member names, no
// OpenCV source.  Only what project_cloud.{h,cu}, CameraCalibration.{h,cpp} and Octreegrid.h use.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>
#include <limits>
#include <unordered_map>
#include <vector>
typedef unsigned char uchar;
#define CV_8UC3 16
#define CV_32F 5
#define CV_16FC3 23
namespace cv {
template <typename T, int M, int N> struct Matx {
    T val[M * N];
    Matx() { for (int i = 0; i < M * N; ++i) val[i] = T(0); }
    template <typename U> Matx(const Matx<U, M, N>& o) { for (int i = 0; i < M * N; ++i) val[i] = T(o.val[i]); }
    static Matx eye() { Matx m; for (int i = 0; i < (M < N ? M : N); ++i) m.val[i * N + i] = T(1); return m; }
    T& operator()(int r, int c) { return val[r * N + c]; }
    const T& operator()(int r, int c) const { return val[r * N + c]; }
};
typedef Matx<double, 3, 3> Matx33d;
typedef Matx<float, 3, 3> Matx33f;
typedef Matx<double, 4, 4> Matx44d;
struct Point3f { float x, y, z; Point3f() : x(0), y(0), z(0) {} Point3f(float a, float b, float c) : x(a), y(b), z(c) {} };
struct Point3d { double x, y, z; Point3d() : x(0), y(0), z(0) {} Point3d(double a, double b, double c) : x(a), y(b), z(c) {} };  // cloudreader.cpp's E57 path
struct Vec3b { uchar val[3]; Vec3b() { val[0] = val[1] = val[2] = 0; } Vec3b(uchar a, uchar b, uchar c) { val[0] = a; val[1] = b; val[2] = c; }
    uchar& operator[](int i) { return val[i]; } const uchar& operator[](int i) const { return val[i]; } };
struct Size { int width, height; Size() : width(0), height(0) {} Size(int w, int h) : width(w), height(h) {} };
// Dense 2-D array: owns its storage or wraps caller memory.
struct Mat {
    int rows = 0, cols = 0, type_ = 0; uint8_t* data = nullptr; std::vector<uint8_t> own;
    static size_t elemSize(int t) { return t == CV_8UC3 ? 3 : t == CV_32F ? 4 : t == CV_16FC3 ? 6 : 1; }
    Mat() {}
    Mat(Size s, int t) : rows(s.height), cols(s.width), type_(t), own(size_t(s.width) * s.height * elemSize(t)) { data = own.data(); }
    Mat(int r, int c, int t, void* p) : rows(r), cols(c), type_(t), data(static_cast<uint8_t*>(p)) {}
    Size size() const { return Size(cols, rows); }
    template <typename T> T* ptr() { return reinterpret_cast<T*>(data); }
    template <typename T> const T* ptr() const { return reinterpret_cast<const T*>(data); }
    // Only the CV_16FC3 -> CV_8UC3 scaling used by computeFull (project_cloud.cu:480).
    void convertTo(Mat& dst, int rtype, double alpha) const;
};
}  // namespace cv
