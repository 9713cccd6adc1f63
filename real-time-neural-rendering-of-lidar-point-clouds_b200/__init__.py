"""rtr_b200 — B200-native point-projection path for RTRenderer (host-side mirror, Python).

The product is ``librtr_b200.so`` (hand-written sm_100a CUDA behind a C ABI, ``include/rtr_b200.h``).
This module is the thin ctypes binding plus a mirror of the reference's operator interface for
this path, so callers and parity tests read like the reference:

    reference (C++)                                   here
    ------------------------------------------------  ------------------------------------------
    CameraCalibration  (CameraCalibration.h:8-54)      CameraCalibration
    ProjectCloud(grid, model)  (project_cloud.h:13)    ProjectCloud(xyz, bgr)  /  .from_packed(...)
    computeRGBD(calib, w2c, &color, &depth)   (:16)    computeRGBD(calib, w2c, color, depth)
    computeFilteredRGBD(...)                  (:17)    computeFilteredRGBD(...)
    computeFull(...) projection+filter part   (:18)    computeTensor(calib, w2c) -> device pointer

Same argument meaning (world->camera 4x4, pre-allocated contiguous outputs, either may be None)
and the same return convention (1 ok, -1 when both outputs are None).  There is NO CPU fallback:
without the CUDA library or without a B200 the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RTR_B200_LIB") or os.path.join(_HERE, "librtr_b200.so")   # RTR_B200_LIB: an experiment build (tools/experiments)

RTR_OK = 1
RTR_ERR_ARG, RTR_ERR_CUDA, RTR_ERR_STATE, RTR_ERR_UNSUPPORTED, RTR_ERR_COMM = -1, -2, -3, -4, -5
STAGE_RGBD, STAGE_FILTERED = 0, 1

# name -> (restype, argtypes): every symbol include/rtr_b200.h declares
_vp, _i, _u64, _i64, _sz = C.c_void_p, C.c_int, C.c_uint64, C.c_int64, C.c_size_t
_dp, _fp = C.POINTER(C.c_double), C.POINTER(C.c_float)
API = {
    "rtr_create": (_i, [_i, C.POINTER(_vp)]),
    "rtr_destroy": (None, [_vp]),
    "rtr_last_error": (C.c_char_p, [_vp]),
    "rtr_upload_cloud_xyz_bgr": (_i, [_vp, _vp, _vp, _u64]),
    "rtr_upload_cloud_packed16": (_i, [_vp, _vp, _u64]),
    "rtr_adopt_device_cloud_packed16": (_i, [_vp, _vp, _u64]),
    "rtr_synth_cloud": (_i, [_vp, _u64, _u64, _u64, _u64, _i, _i, _i, _i]),
    "rtr_cloud_size": (_u64, [_vp]),
    "rtr_download_cloud_packed16": (_i, [_vp, _u64, _u64, _vp]),
    "rtr_set_intrinsics": (_i, [_vp, _i, _i, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, _dp]),
    "rtr_set_intrinsics_matrix": (_i, [_vp, _i, _i, _dp, _dp]),
    "rtr_set_pose_w2c": (_i, [_vp, _dp]),
    "rtr_set_cam_proj_raw": (_i, [_vp, _fp]),
    "rtr_get_cam_proj": (_i, [_vp, _fp]),
    "rtr_render_rgbd": (_i, [_vp, _vp, _vp]),
    "rtr_render_filtered": (_i, [_vp, _vp, _vp]),
    "rtr_render_tensor": (_i, [_vp, C.POINTER(_vp)]),
    "rtr_render_device": (_i, [_vp, _i]),
    "rtr_sync": (_i, [_vp]),
    "rtr_render_trajectory": (_i, [_vp, _i, _dp, _i, _vp, _vp]),
    "rtr_get_device_buffers": (_i, [_vp, _vp]),
    "rtr_read_buffer": (_i, [_vp, _i, _vp, _sz]),
    "rtr_project_points": (_i, [_vp, _vp, _vp]),
    "rtr_set_option": (_i, [_vp, C.c_char_p, _i64]),
    "rtr_get_option": (_i64, [_vp, C.c_char_p]),
    "rtr_get_stage_ms": (_i, [_vp, _fp]),
    "rtr_get_stage_ms_sum": (_i, [_vp, _dp, C.POINTER(_u64), _i]),
    "rtr_get_cull_stats": (_i, [_vp, C.POINTER(_u64), C.POINTER(_u64), C.POINTER(_u64), _i]),
    "rtr_get_stream_stats": (_i, [_vp, C.POINTER(_u64), C.POINTER(_u64), _i]),
    "rtr_get_smem_tile_stats": (_i, [_vp, C.POINTER(_u64)]),
    "rtr_launch_count": (_u64, [_vp]),
    "rtr_bench_red_min": (_i, [_vp, _i, _u64, _i, _i, _fp, C.POINTER(_u64)]),
    "rtr_selftest_fast_divide": (_i, [_vp, _u64, _u64, C.POINTER(_u64)]),
    "rtr_host_distortion_bounds": (_i, [_i, _i, _dp, _dp, _dp, _dp]),
    "rtr_host_ring_stride": (C.c_uint32, [_u64]),
    "rtr_host_band_compact": (C.c_int, [C.POINTER(C.c_uint32), C.c_uint32, C.POINTER(C.c_uint32), C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32),
                                        C.POINTER(C.c_uint32)]),
    "rtr_host_ring_claim": (C.c_int, [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                            C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "rtr_comm_unique_id": (_i, [_vp]),
    "rtr_comm_init": (_i, [_vp, _vp, _i, _i]),
    "rtr_comm_destroy": (_i, [_vp]),
    "rtr_peer_export": (_i, [_vp, _vp]),
    "rtr_peer_attach": (_i, [_vp, _vp, _i, _i]),
    "rtr_peer_detach": (_i, [_vp]),
    "rtr_version": (C.c_char_p, []),
}
_ip, _u8pp, _fpp = C.POINTER(_i), C.POINTER(C.POINTER(C.c_uint8)), C.POINTER(C.POINTER(C.c_float))
# include/rtr_b200_io.h ("next" rows of SURVEY.md §8 f)
API_IO = {
    "rtr_load_ply": (_i, [_vp, C.c_char_p, _i]),
    "rtr_bin_cells": (_i, [_vp, _ip]),
    "rtr_sort_morton": (_i, [_vp]),
    "rtr_io_write_ply": (_i, [C.c_char_p, _vp, _vp, _u64]),
    "rtr_load_oct": (_i, [_vp, C.c_char_p]),
    "rtr_io_write_oct": (_i, [C.c_char_p, _vp, _vp, _u64]),
    "rtr_io_read_oct": (_i, [C.c_char_p, _fpp, _u8pp, C.POINTER(_u64), _ip, C.POINTER(_ip), C.POINTER(C.POINTER(_u64))]),
    "rtr_io_free": (None, [_vp]),
    "rtr_io_load_calibration": (_i, [C.c_char_p, _ip, _ip, _dp, _dp, _ip, _ip]),
    "rtr_io_load_trajectory": (_i, [C.c_char_p, _i, _dp, _i, _ip]),
    "rtr_io_invert_rigid": (_i, [_dp, _dp]),
    "rtr_postprocess_unet_output": (_i, [_vp, _vp, _i, _i, _vp, _vp]),
}


class DeviceBuffers(C.Structure):
    """rtr_device_buffers (include/rtr_b200.h)."""
    _fields_ = [
        ("points", _vp), ("zbuf", _vp), ("accum", _vp), ("image", _vp), ("tensor", _vp), ("minmax", _vp),
        ("level", _vp * 5), ("mask", _vp * 4), ("width", _i), ("height", _i),
        ("level_w", _i * 5), ("level_h", _i * 5), ("up_w", _i * 5), ("up_h", _i * 5),
        ("tensor_plane", _u64), ("stream", _vp),
    ]


_lib = None


def load_library(path: Optional[str] = None) -> C.CDLL:
    """dlopen librtr_b200.so and bind every declared symbol.  Raises if the library is missing —
    build it with ``python __graft_entry__.py`` / ``build.build()``; nothing falls back to the CPU."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(f"{p} not found: build the CUDA library first (no CPU fallback exists)")
    lib = C.CDLL(p)
    for name, (res, args) in list(API.items()) + list(API_IO.items()):
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype, fn.argtypes = res, args
    if path is None:
        _lib = lib
    return lib


class RtrError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"rtr_b200 error {code}: {msg}")
        self.code = code


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(_vp)


def _as_f64(a, n):
    v = np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1))
    if v.size != n:
        raise ValueError(f"expected {n} values, got {v.size}")
    return v


class CameraCalibration:
    """Mirror of the reference's CameraCalibration for what the hot path reads: W, H, K and the
    (reference-unused) 5 distortion coefficients (CameraCalibration.h:8-54, .cpp:211-230)."""

    def __init__(self):
        self._K = np.eye(3, dtype=np.float64)
        self._dists = [0.0] * 5
        self._w, self._h = 640, 480  # CameraCalibration.cpp:7-8

    def loadCalibration(self, fx, fy, cx, cy, dist: Sequence[float], width: int, height: int) -> bool:
        self._K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], dtype=np.float64)
        self._dists = [float(d) for d in dist] + [0.0] * (5 - len(dist))
        self._w, self._h = int(width), int(height)
        return True

    def setIntrinsicsMatrix(self, K): self._K = np.asarray(K, dtype=np.float64).reshape(3, 3).copy()
    def setDistortionParameters(self, d): self._dists = [float(x) for x in d] + [0.0] * (5 - len(d))
    def getIntrinsicsMatrix(self): return self._K.copy()
    def getDistortionParameters(self): return list(self._dists)
    def getWidth(self): return self._w
    def getHeight(self): return self._h
    def setWidth(self, w): self._w = int(w)
    def setHeight(self, h): self._h = int(h)
    def getFocalLengthX(self): return float(self._K[0, 0])
    def getFocalLengthY(self): return float(self._K[1, 1])
    def getPrincipalPointX(self): return float(self._K[0, 2])
    def getPrincipalPointY(self): return float(self._K[1, 2])


class ProjectCloud:
    """Drop-in for the reference's ProjectCloud on the projection / prefilter path.

    ``apply_distortion=False`` (default) reproduces the reference, which parses the distortion
    coefficients but never uses them; True enables the new k1,k2,p1,p2,k3 path.
    """

    def __init__(self, xyz: Optional[np.ndarray] = None, bgr: Optional[np.ndarray] = None, device: int = 0,
                 apply_distortion: bool = False, sort: bool = True):
        self._lib = load_library()
        h = _vp()
        rc = self._lib.rtr_create(device, C.byref(h))
        if rc != RTR_OK:
            raise RtrError(rc, (self._lib.rtr_last_error(None) or b"").decode())
        self._h = h
        self.device = device
        self.apply_distortion = apply_distortion
        self._keep = None  # keeps adopted device memory alive
        if not sort:   # default: uploads are Morton-sorted on the GPU (results never depend on point order)
            self.set_option("sort_on_upload", 0)
        if xyz is not None:
            self.upload(xyz, bgr)

    # ---- construction helpers
    @classmethod
    def from_packed(cls, records: np.ndarray, device: int = 0, **kw) -> "ProjectCloud":
        pc = cls(device=device, **kw)
        rec = np.ascontiguousarray(records)
        if rec.nbytes % 16:
            raise ValueError("packed records must be 16 bytes each")
        pc._check(pc._lib.rtr_upload_cloud_packed16(pc._h, _ptr(rec), rec.nbytes // 16))
        return pc

    @classmethod
    def synthetic(cls, seed: int, n_total: int, first: int = 0, count: Optional[int] = None, hall=(48, 40, 12),
                  n_boxes: int = 12, device: int = 0, **kw) -> "ProjectCloud":
        pc = cls(device=device, **kw)
        cnt = n_total - first if count is None else count
        pc._check(pc._lib.rtr_synth_cloud(pc._h, seed, n_total, first, cnt, hall[0], hall[1], hall[2], n_boxes))
        return pc

    def upload(self, xyz: np.ndarray, bgr: np.ndarray):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        bgr = np.ascontiguousarray(bgr, dtype=np.uint8).reshape(-1, 3)
        if len(xyz) != len(bgr):
            raise ValueError("xyz and bgr differ in length")
        self._check(self._lib.rtr_upload_cloud_xyz_bgr(self._h, _ptr(xyz), _ptr(bgr), len(xyz)))

    def adopt_device_records(self, device_ptr: int, n_points: int, keepalive=None):
        self._check(self._lib.rtr_adopt_device_cloud_packed16(self._h, _vp(device_ptr), n_points))
        self._keep = keepalive

    # ---- the reference's methods
    def computeRGBD(self, calibration: CameraCalibration, extrinsics, color: Optional[np.ndarray],
                    depth: Optional[np.ndarray]) -> int:
        return self._compute(self._lib.rtr_render_rgbd, calibration, extrinsics, color, depth)

    def computeFilteredRGBD(self, calibration: CameraCalibration, extrinsics, color: Optional[np.ndarray],
                            depth: Optional[np.ndarray]) -> int:
        return self._compute(self._lib.rtr_render_filtered, calibration, extrinsics, color, depth)

    def computeTensor(self, calibration: CameraCalibration, extrinsics) -> int:
        """Projection + prefilter of computeFull; returns the device address of the 1x5xHxW fp16
        U-Net input (what the reference wraps with torch::from_blob, project_cloud.cu:471)."""
        self.set_camera(calibration, extrinsics)
        out = _vp()
        self._check(self._lib.rtr_render_tensor(self._h, C.byref(out)))
        return int(out.value)

    # ---- lower-level access used by bench/tests
    def set_camera(self, calibration: CameraCalibration, extrinsics=None):
        K = _as_f64(calibration.getIntrinsicsMatrix(), 9)
        d = _as_f64(calibration.getDistortionParameters() if self.apply_distortion else [0.0] * 5, 5)
        self._check(self._lib.rtr_set_intrinsics_matrix(self._h, calibration.getWidth(), calibration.getHeight(),
                                                        K.ctypes.data_as(_dp), d.ctypes.data_as(_dp)))
        if extrinsics is not None:
            E = _as_f64(extrinsics, 16)
            self._check(self._lib.rtr_set_pose_w2c(self._h, E.ctypes.data_as(_dp)))

    def set_cam_proj_raw(self, m16):
        m = np.ascontiguousarray(np.asarray(m16, dtype=np.float32).reshape(-1))
        self._check(self._lib.rtr_set_cam_proj_raw(self._h, m.ctypes.data_as(_fp)))

    def get_cam_proj(self) -> np.ndarray:
        m = np.zeros(16, dtype=np.float32)
        self._check(self._lib.rtr_get_cam_proj(self._h, m.ctypes.data_as(_fp)))
        return m.reshape(4, 4)

    def render_device(self, stage: int = STAGE_FILTERED):
        self._check(self._lib.rtr_render_device(self._h, stage))

    def sync(self):
        self._check(self._lib.rtr_sync(self._h))

    def render_trajectory(self, stage: int, poses_w2c: np.ndarray, color: Optional[np.ndarray] = None,
                          depth: Optional[np.ndarray] = None):
        poses = np.ascontiguousarray(poses_w2c, dtype=np.float64).reshape(-1, 16)
        self._check(self._lib.rtr_render_trajectory(self._h, stage, poses.ctypes.data_as(_dp), len(poses),
                                                    _ptr(color), _ptr(depth)))

    def render_trajectory_ptr(self, stage: int, poses_w2c: np.ndarray, color_ptr: int = 0, depth_ptr: int = 0):
        """Same, with raw host addresses (e.g. pinned torch tensors)."""
        poses = np.ascontiguousarray(poses_w2c, dtype=np.float64).reshape(-1, 16)
        self._check(self._lib.rtr_render_trajectory(self._h, stage, poses.ctypes.data_as(_dp), len(poses),
                                                    _vp(color_ptr or None), _vp(depth_ptr or None)))

    def device_buffers(self) -> DeviceBuffers:
        b = DeviceBuffers()
        self._check(self._lib.rtr_get_device_buffers(self._h, C.byref(b)))
        return b

    _WHAT = {"zbuf": 0, "accum": 1, "image": 2, "tensor": 3, "minmax": 4, "level1": 5, "level2": 6, "level3": 7,
             "level4": 8, "mask0": 9, "mask1": 10, "mask2": 11, "mask3": 12}

    def read(self, what: str, dtype, count: int) -> np.ndarray:
        out = np.empty(count, dtype=dtype)
        self._check(self._lib.rtr_read_buffer(self._h, self._WHAT[what], _ptr(out), out.nbytes))
        return out

    def project_points(self):
        n = self.cloud_size
        pix, zb = np.empty(n, dtype=np.int32), np.empty(n, dtype=np.uint32)
        self._check(self._lib.rtr_project_points(self._h, _ptr(pix), _ptr(zb)))
        return pix, zb

    def download_cloud(self, first: int = 0, count: Optional[int] = None) -> np.ndarray:
        cnt = self.cloud_size - first if count is None else count
        out = np.empty((cnt, 4), dtype=np.float32)
        self._check(self._lib.rtr_download_cloud_packed16(self._h, first, cnt, _ptr(out)))
        return out

    def set_option(self, key: str, value: int):
        self._check(self._lib.rtr_set_option(self._h, key.encode(), int(value)))

    def get_option(self, key: str) -> int:
        return int(self._lib.rtr_get_option(self._h, key.encode()))

    def stage_ms(self) -> np.ndarray:
        ms = np.zeros(6, dtype=np.float32)
        self._check(self._lib.rtr_get_stage_ms(self._h, ms.ctypes.data_as(_fp)))
        return ms

    def stage_ms_sum(self, reset: bool = True):
        """(per-stage ms summed over the frames since the last reset, number of frames); timing=2."""
        ms = np.zeros(6, dtype=np.float64)
        n = _u64(0)
        self._check(self._lib.rtr_get_stage_ms_sum(self._h, ms.ctypes.data_as(_dp), C.byref(n), int(reset)))
        return ms, int(n.value)

    def cull_stats(self, reset: bool = True):
        """(frames, visible chunks summed over those frames, chunks in the cloud)."""
        f, v, n = _u64(0), _u64(0), _u64(0)
        self._check(self._lib.rtr_get_cull_stats(self._h, C.byref(f), C.byref(v), C.byref(n), int(reset)))
        return int(f.value), int(v.value), int(n.value)

    def stream_stats(self, reset: bool = True):
        """(point passes over a visible-chunk list, chunks they streamed from HBM) since the last reset."""
        p, c = _u64(0), _u64(0)
        self._check(self._lib.rtr_get_stream_stats(self._h, C.byref(p), C.byref(c), int(reset)))
        return int(p.value), int(c.value)

    def smem_tile_stats(self):
        """(tiles through the shared-memory window, tiles direct, window pixels flushed, records entered) — zmin_variant bit 6."""
        v = (_u64 * 4)()
        self._check(self._lib.rtr_get_smem_tile_stats(self._h, v))
        return tuple(int(x) for x in v)

    def bench_red_min(self, mode: int, n_ops: int = 0, key64: bool = False, iters: int = 5):
        """(ms per launch, REDs issued per launch) of the L2 atomic micro-benchmark."""
        ms, live = C.c_float(0), _u64(0)
        self._check(self._lib.rtr_bench_red_min(self._h, mode, n_ops, int(key64), iters, C.byref(ms), C.byref(live)))
        return float(ms.value), int(live.value)

    def selftest_fast_divide(self, n_pairs: int, seed: int = 1) -> int:
        """Pairs (of n_pairs random bit patterns) for which the ring kernels' MUFU.RCP + FMUL differs from __fdividef."""
        bad = _u64(0)
        self._check(self._lib.rtr_selftest_fast_divide(self._h, n_pairs, seed, C.byref(bad)))
        return int(bad.value)

    @property
    def cloud_size(self) -> int:
        return int(self._lib.rtr_cloud_size(self._h))

    @property
    def launch_count(self) -> int:
        return int(self._lib.rtr_launch_count(self._h))

    # ---- loaders / post-process (include/rtr_b200_io.h)
    @classmethod
    def from_ply(cls, path: str, bin_cells: bool = True, device: int = 0, **kw) -> "ProjectCloud":
        """CloudReader::loadCloud for a .ply (cloudreader.cpp:122-177), binning on the GPU."""
        pc = cls(device=device, **kw)
        pc._check(pc._lib.rtr_load_ply(pc._h, os.fsencode(path), int(bin_cells)))
        return pc

    @classmethod
    def from_oct(cls, path: str, device: int = 0, **kw) -> "ProjectCloud":
        """The pcd.oct cache hit of CloudReader::loadCloud (cloudreader.cpp:182-191)."""
        pc = cls(device=device, **kw)
        pc._check(pc._lib.rtr_load_oct(pc._h, os.fsencode(path)))
        return pc

    def bin_cells(self):
        d = (_i * 3)()
        self._check(self._lib.rtr_bin_cells(self._h, d))
        return tuple(d)

    def sort_morton(self):
        self._check(self._lib.rtr_sort_morton(self._h))

    def postprocess_unet_output(self, device_fp16_chw_ptr: int, W: int, H: int) -> np.ndarray:
        out = np.empty(W * H * 3, dtype=np.uint8)
        self._check(self._lib.rtr_postprocess_unet_output(self._h, _vp(device_fp16_chw_ptr), W, H, _ptr(out), None))
        return out.reshape(H, W, 3)

    # ---- point-sharded multi-GPU plumbing (NCCL inside the library)
    def comm_init(self, unique_id: bytes, rank: int, n_ranks: int):
        buf = C.create_string_buffer(unique_id, 128)
        self._check(self._lib.rtr_comm_init(self._h, buf, rank, n_ranks))

    def peer_export(self) -> bytes:
        """512-byte blob describing this renderer's frame buffers (intrinsics must be set); all-gather them."""
        buf = C.create_string_buffer(512)
        self._check(self._lib.rtr_peer_export(self._h, buf))
        return buf.raw

    def peer_attach(self, blobs: bytes, rank: int, n_ranks: int):
        assert len(blobs) == 512 * n_ranks
        self._check(self._lib.rtr_peer_attach(self._h, C.create_string_buffer(blobs, len(blobs)), rank, n_ranks))

    def peer_detach(self):
        self._check(self._lib.rtr_peer_detach(self._h))

    @staticmethod
    def comm_unique_id() -> bytes:
        lib = load_library()
        buf = C.create_string_buffer(128)
        rc = lib.rtr_comm_unique_id(buf)
        if rc != RTR_OK:
            raise RtrError(rc, (lib.rtr_last_error(None) or b"").decode())
        return buf.raw

    # ---- internals
    def _compute(self, fn, calibration, extrinsics, color, depth) -> int:
        if color is None and depth is None:
            return -1  # project_cloud.cu:270-273
        W, H = calibration.getWidth(), calibration.getHeight()
        if color is not None and not (color.dtype == np.uint8 and color.size == W * H * 3 and color.flags.c_contiguous):
            raise ValueError("color must be a contiguous uint8 array of H*W*3 (CV_8UC3)")
        if depth is not None and not (depth.dtype == np.float32 and depth.size == W * H and depth.flags.c_contiguous):
            raise ValueError("depth must be a contiguous float32 array of H*W (CV_32F)")
        self.set_camera(calibration, extrinsics)
        rc = fn(self._h, _ptr(color), _ptr(depth))
        self._check(rc)
        return rc

    def _check(self, rc: int):
        if rc != RTR_OK:
            raise RtrError(rc, (self._lib.rtr_last_error(self._h) or b"").decode())

    def close(self):
        if getattr(self, "_h", None):
            self._lib.rtr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- file formats on either side of the path (host parsing inside the library)
def _io_check(rc):
    if rc != RTR_OK:
        raise RtrError(rc, (load_library().rtr_last_error(None) or b"").decode())


def write_ply(path, xyz, bgr):
    xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
    bgr = np.ascontiguousarray(bgr, dtype=np.uint8).reshape(-1, 3)
    _io_check(load_library().rtr_io_write_ply(os.fsencode(path), _ptr(xyz), _ptr(bgr), len(xyz)))


def write_oct(path, xyz, bgr):
    xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
    bgr = np.ascontiguousarray(bgr, dtype=np.uint8).reshape(-1, 3)
    _io_check(load_library().rtr_io_write_oct(os.fsencode(path), _ptr(xyz), _ptr(bgr), len(xyz)))


def read_oct(path):
    """-> dict(xyz (n,3) f32, bgr (n,3) u8, dims (nx,ny,nz), keys, counts) in file order."""
    lib = load_library()
    xyz, bgr, n = C.POINTER(C.c_float)(), C.POINTER(C.c_uint8)(), _u64(0)
    hdr, keys, counts = (_i * 4)(), _ip(), C.POINTER(_u64)()
    _io_check(lib.rtr_io_read_oct(os.fsencode(path), C.byref(xyz), C.byref(bgr), C.byref(n), hdr, C.byref(keys), C.byref(counts)))
    nn, nb = int(n.value), hdr[3]
    out = dict(xyz=np.ctypeslib.as_array(xyz, shape=(nn, 3)).copy() if nn else np.zeros((0, 3), np.float32),
               bgr=np.ctypeslib.as_array(bgr, shape=(nn, 3)).copy() if nn else np.zeros((0, 3), np.uint8),
               dims=(hdr[0], hdr[1], hdr[2]),
               keys=np.ctypeslib.as_array(keys, shape=(nb,)).copy() if nb else np.zeros(0, np.int32),
               counts=np.ctypeslib.as_array(counts, shape=(nb,)).copy() if nb else np.zeros(0, np.uint64))
    for p in (xyz, bgr, keys, counts):
        lib.rtr_io_free(C.cast(p, _vp))
    return out


def load_calibration(path) -> "CameraCalibration":
    """CameraCalibration::loadCalibration(file) (CameraCalibration.cpp:101-209)."""
    W, H, nd, fe = _i(0), _i(0), _i(0), _i(0)
    K, d = np.zeros(9), np.zeros(8)
    _io_check(load_library().rtr_io_load_calibration(os.fsencode(path), C.byref(W), C.byref(H), K.ctypes.data_as(_dp),
                                                     d.ctypes.data_as(_dp), C.byref(nd), C.byref(fe)))
    c = CameraCalibration()
    c.setIntrinsicsMatrix(K)
    c._dists = [float(x) for x in d[:nd.value]]
    c.setWidth(W.value)
    c.setHeight(H.value)
    c.fisheye = bool(fe.value)
    return c


def load_trajectory(path, order: int = 0, max_poses: int = 1 << 20) -> np.ndarray:
    """(n, 4, 4) camera->world poses; order 0 = 'ts tx ty tz qx qy qz qw' (what the example parses,
    main.cpp:32), 1 = COLMAP images.txt order (README.md:92)."""
    n = sum(1 for ln in open(path) if ln.strip() and not ln.startswith("#"))
    n = min(n, max_poses)
    out = np.zeros((max(n, 1), 16))
    got = _i(0)
    _io_check(load_library().rtr_io_load_trajectory(os.fsencode(path), order, out.ctypes.data_as(_dp), n, C.byref(got)))
    return out[:got.value].reshape(-1, 4, 4)


def invert_rigid(pose) -> np.ndarray:
    p = _as_f64(pose, 16)
    o = np.zeros(16)
    _io_check(load_library().rtr_io_invert_rigid(p.ctypes.data_as(_dp), o.ctypes.data_as(_dp)))
    return o.reshape(4, 4)


# ---- host-side helpers shared by bench and tests (pure host logic, no compute)
def pack_records(xyz: np.ndarray, bgr: np.ndarray) -> np.ndarray:
    """{x, y, z, b|g<<8|r<<16|255<<24} records as an (n, 4) float32 array (bit pattern in column 3)."""
    xyz = np.asarray(xyz, dtype=np.float32).reshape(-1, 3)
    bgr = np.asarray(bgr, dtype=np.uint8).reshape(-1, 3)
    rec = np.empty((len(xyz), 4), dtype=np.float32)
    rec[:, :3] = xyz
    c = bgr[:, 0].astype(np.uint32) | (bgr[:, 1].astype(np.uint32) << 8) | (bgr[:, 2].astype(np.uint32) << 16) | np.uint32(0xFF000000)
    rec[:, 3] = c.view(np.float32)
    return rec


def look_at_w2c(eye, forward, up=(0.0, 0.0, 1.0)) -> np.ndarray:
    """World->camera 4x4 (OpenCV camera axes: x right, y down, z forward)."""
    f = np.asarray(forward, dtype=np.float64); f = f / np.linalg.norm(f)
    r = np.cross(f, np.asarray(up, dtype=np.float64)); r = r / np.linalg.norm(r)
    d = np.cross(f, r)
    R = np.stack([r, d, f])
    E = np.eye(4)
    E[:3, :3] = R
    E[:3, 3] = -R @ np.asarray(eye, dtype=np.float64)
    return E


def trajectory_w2c(n_poses: int, center=(6.0, 5.0, 1.5), radius=2.0, pitch_deg=10.0) -> np.ndarray:
    """SURVEY.md §8 d trajectory: closed circle of `radius` at height center[2], yaw following the
    tangent, sinusoidal pitch.  Returns (n, 4, 4) world->camera matrices."""
    out = np.empty((n_poses, 4, 4))
    for i in range(n_poses):
        a = 2 * np.pi * i / n_poses
        eye = (center[0] + radius * np.cos(a), center[1] + radius * np.sin(a), center[2])
        pitch = np.deg2rad(pitch_deg) * np.sin(3 * a)
        fwd = (-np.sin(a) * np.cos(pitch), np.cos(a) * np.cos(pitch), np.sin(pitch))
        out[i] = look_at_w2c(eye, fwd)
    return out


def shard_frames(n_frames: int, rank: int, world: int) -> range:
    """Frame-sharded split: contiguous chunks, remainder to the low ranks."""
    base, rem = divmod(n_frames, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def shard_points(n_points: int, rank: int, world: int):
    """Point-sharded split: (first, count) of the contiguous range rank owns."""
    r = shard_frames(n_points, rank, world)
    return r.start, len(r)
