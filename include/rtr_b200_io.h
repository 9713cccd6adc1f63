/* rtr_b200_io — the callers and data formats on either side of the hot path (SURVEY.md §8 f, "next" rows).
 * Same shared library as rtr_b200.h.  Each entry point names the reference code it replaces.
 *
 *   f1  PLY loader + 0.25 m cell binning on the GPU    cloudreader.cpp:8-82 (computeGrid), 122-177 (loadPLY)
 *   f2  .oct cache reader / writer, byte-compatible      Octreegrid.h:53-114, cloudreader.cpp:182-191, 210-214
 *   f3  calibration + trajectory parsers                 CameraCalibration.cpp:101-209, example/render_trajectory/main.cpp:20-65
 *   f4  U-Net output post-process on the GPU             project_cloud.cu:475-480
 * All functions return RTR_OK (1) or a negative rtr_b200 error code; text via rtr_last_error(r) (renderer
 * calls) or rtr_last_error(NULL) (pure host calls).
 */
#ifndef RTR_B200_IO_H
#define RTR_B200_IO_H

#include "rtr_b200.h"

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

/* ---- f1.  loadPLY: ascii or binary_little_endian, vertex properties x,y,z (float/double) and optional
 * red,green,blue (uchar); colours are stored B,G,R like the reference (cloudreader.cpp:168).  With bin_cells != 0
 * the uploaded cloud is then grouped by 0.25 m cell on the GPU (rtr_bin_cells). */
int rtr_load_ply(rtr_renderer* r, const char* path, int bin_cells);
/* computeGrid's binning (cloudreader.cpp:8-60) applied to the cloud resident on the device: same bounding box
 * rounding, same float expression for the cell index, key = x + y*nx + z*nx*ny; points are re-ordered by key
 * (stable: arrival order inside a cell, as the reference's emplace_back keeps it; cells in ascending key order —
 * the reference iterates an unordered_map, i.e. an unspecified order, and no output depends on it).
 * dims3 (may be NULL) receives numBlocks_x/y/z. */
int rtr_bin_cells(rtr_renderer* r, int* dims3);
/* Re-order the resident cloud along a 63-bit Morton curve (21 bits per axis over the bounding box): consecutive
 * records are then neighbours in space, so a warp's points land on neighbouring pixels and the z-buffer gathers /
 * REDs of the point passes touch far fewer L2 sectors.  No output depends on point order. */
int rtr_sort_morton(rtr_renderer* r);
/* Fixture support: write a binary_little_endian PLY (x y z float, red green blue uchar) from B,G,R colours. */
int rtr_io_write_ply(const char* path, const float* xyz, const uint8_t* bgr, uint64_t n);

/* ---- f2.  pcd.oct: int32 nx, ny, nz, numBlocks; per block: int32 key, uint64 n, n*3 float32, n*3 uint8 (B,G,R),
 * bbMin 3 float32, bbMax 3 float32 (Octreegrid.h:53-79). */
int rtr_load_oct(rtr_renderer* r, const char* path);
/* Bins on the host exactly like computeGrid and writes the cache file the reference's readOctreeBinary accepts. */
int rtr_io_write_oct(const char* path, const float* xyz, const uint8_t* bgr, uint64_t n);
/* Reads a cache file into malloc'd arrays (release with rtr_io_free).  keys/counts (may be NULL) receive one
 * entry per block in file order; header4 = nx, ny, nz, numBlocks. */
int rtr_io_read_oct(const char* path, float** xyz, uint8_t** bgr, uint64_t* n, int* header4, int** keys,
                    uint64_t** counts);
void rtr_io_free(void* p);

/* ---- f3.  CameraCalibration::loadCalibration(file): COLMAP cameras.txt (OPENCV / OPENCV_FISHEYE, first camera
 * line) when the path ends in "cameras.txt", else the custom text format (W H, 3x3 K, distortion line, fisheye
 * flag).  dist8 receives n_dist values (5 pinhole / 4 fisheye). */
int rtr_io_load_calibration(const char* path, int* width, int* height, double* K9, double* dist8, int* n_dist,
                            int* fisheye);
/* Trajectory file -> camera->world 4x4 poses (row-major), '#' and empty lines skipped.
 * order 0: "timestamp tx ty tz qx qy qz qw" — what example/render_trajectory/main.cpp:32 parses;
 * order 1: "id qw qx qy qz tx ty tz ..."   — the COLMAP images.txt order the README documents (README.md:92).
 * Quaternions are normalised (cv::Quatd::normalize).  Returns the number of poses in *n_poses. */
int rtr_io_load_trajectory(const char* path, int order, double* poses16, int max_poses, int* n_poses);
/* world->camera from camera->world for a rigid pose (what the example's pose.inv() yields, main.cpp:96). */
int rtr_io_invert_rigid(const double* pose16, double* inv16);

/* ---- f4.  The step after the U-Net (project_cloud.cu:475-480): fp16 3xHxW network output on the device ->
 * uint8 HxWx3 = saturate(round_half_even(v * 255)), into pinned/pageable host memory (host_hwc) and/or a device
 * buffer (device_hwc); either may be NULL.  Runs on the renderer's stream; synchronises when host_hwc is given. */
int rtr_postprocess_unet_output(rtr_renderer* r, const void* device_fp16_chw, int width, int height,
                                uint8_t* host_hwc, uint8_t* device_hwc);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* RTR_B200_IO_H */
