// Minimal stand-in for the parts of GLM the reference's RTRenderer sources use.
// GLM is an un-vendored dependency of the reference (find_package(glm), CMakeLists.txt:14) and is
// not installed in this image.  This is OUR code (test infrastructure for oracle/_ref), written
// from GLM's documented semantics: column-major storage, m[col][row], and the scalar
// mat4*mat4 definition  R[c] = ((A[0]*B[c][0] + A[1]*B[c][1]) + A[2]*B[c][2]) + A[3]*B[c][3].
#pragma once
#include <cstddef>
#include <limits>
#if defined(__CUDACC__)
#define RTR_GLM_HD __host__ __device__
#else
#define RTR_GLM_HD
#endif
namespace glm {
enum qualifier { packed_highp, defaultp = packed_highp };
template <int L, typename T, qualifier Q = defaultp> struct vec;
template <typename T, qualifier Q> struct vec<3, T, Q> {
    T x, y, z;
    RTR_GLM_HD vec() : x(), y(), z() {}
    RTR_GLM_HD vec(T a, T b, T c) : x(a), y(b), z(c) {}
    RTR_GLM_HD T& operator[](int i) { return (&x)[i]; }
    RTR_GLM_HD const T& operator[](int i) const { return (&x)[i]; }
};
template <typename T, qualifier Q> struct vec<4, T, Q> {
    T x, y, z, w;
    RTR_GLM_HD vec() : x(), y(), z(), w() {}
    RTR_GLM_HD vec(T a, T b, T c, T d) : x(a), y(b), z(c), w(d) {}
    RTR_GLM_HD T& operator[](int i) { return (&x)[i]; }
    RTR_GLM_HD const T& operator[](int i) const { return (&x)[i]; }
};
typedef vec<3, float, defaultp> vec3;
typedef vec<4, float, defaultp> vec4;
typedef vec<3, double, defaultp> dvec3;

template <int C, int R, typename T, qualifier Q = defaultp> struct mat;
template <typename T, qualifier Q> struct mat<3, 3, T, Q> {
    vec<3, T, Q> c[3];
    mat() {}
    vec<3, T, Q>& operator[](int i) { return c[i]; }
    const vec<3, T, Q>& operator[](int i) const { return c[i]; }
};
template <typename T, qualifier Q> struct mat<4, 4, T, Q> {
    vec<4, T, Q> c[4];
    mat() {}
    // mat4(mat3): upper-left 3x3 copied, remainder from identity.
    explicit mat(const mat<3, 3, T, Q>& m) {
        c[0] = vec<4, T, Q>(m[0].x, m[0].y, m[0].z, T(0));
        c[1] = vec<4, T, Q>(m[1].x, m[1].y, m[1].z, T(0));
        c[2] = vec<4, T, Q>(m[2].x, m[2].y, m[2].z, T(0));
        c[3] = vec<4, T, Q>(T(0), T(0), T(0), T(1));
    }
    vec<4, T, Q>& operator[](int i) { return c[i]; }
    const vec<4, T, Q>& operator[](int i) const { return c[i]; }
};
typedef mat<3, 3, float, defaultp> mat3;
typedef mat<4, 4, float, defaultp> mat4;

template <typename T, qualifier Q> inline mat<3, 3, T, Q> transpose(const mat<3, 3, T, Q>& m) {
    mat<3, 3, T, Q> r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r[i][j] = m[j][i];
    return r;
}
template <typename T, qualifier Q> inline mat<4, 4, T, Q> transpose(const mat<4, 4, T, Q>& m) {
    mat<4, 4, T, Q> r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) r[i][j] = m[j][i];
    return r;
}
template <typename T, qualifier Q>
inline mat<4, 4, T, Q> operator*(const mat<4, 4, T, Q>& a, const mat<4, 4, T, Q>& b) {
    mat<4, 4, T, Q> r;
    for (int col = 0; col < 4; ++col)
        for (int row = 0; row < 4; ++row) {
            T t = a[0][row] * b[col][0];
            t = t + a[1][row] * b[col][1];
            t = t + a[2][row] * b[col][2];
            t = t + a[3][row] * b[col][3];
            r[col][row] = t;
        }
    return r;
}
template <typename T, qualifier Q> inline const T* value_ptr(const mat<4, 4, T, Q>& m) { return &m[0].x; }
template <typename T, qualifier Q> inline T* value_ptr(mat<4, 4, T, Q>& m) { return &m[0].x; }
}  // namespace glm
