// Stand-in for libE57Format's reader header (the library is not in this image).  OUR code, test infrastructure for
// oracle/_ref: only what PointCloudReader.h needs to DECLARE its members, so that the reference's cloudreader.cpp — whose
// PLY path (loadPLY + computeGrid, cloudreader.cpp:8-82, 122-177) pins rtr_load_ply / rtr_bin_cells — compiles unmodified.
// The E57 path itself is not available: oracle/ref_harness.cu defines PointCloudReader's members as stubs that abort.
#pragma once
namespace e57 {
class Reader;
struct RigidBodyTransform {};
struct Quaternion {};
}  // namespace e57
