"""TEST INFRASTRUCTURE — Python bindings of the two checkers.

  cpu()  -> CpuOracle   liboracle_cpu.so, the plain-C restatement (oracle/rtr_oracle.c)
  ref()  -> RefOracle   _ref/libref_rtrenderer.so, the reference's own CUDA sources compiled
                        unmodified (oracle/Makefile `ref`); needs a GPU to run.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CPU_LIB = os.path.join(HERE, "liboracle_cpu.so")
REF_LIB = os.path.join(HERE, "_ref", "libref_rtrenderer.so")              # parity build (zero-initialising cudaMalloc)
REF_LIB_STOCK = os.path.join(HERE, "_ref", "libref_rtrenderer_stock.so")  # timing build (reference exactly as shipped)
REFERENCE_ROOT = "/root/reference"

_vp, _i, _u64 = C.c_void_p, C.c_int, C.c_uint64


def build_cpu(force: bool = False) -> str:
    src = os.path.join(HERE, "rtr_oracle.c")
    if force or not os.path.exists(CPU_LIB) or os.path.getmtime(CPU_LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-s", "-C", HERE, "cpu"], check=True)
    return CPU_LIB


def build_ref() -> str | None:
    """Compile the reference where it lies; only possible where /root/reference exists."""
    if os.path.exists(REF_LIB) and os.path.exists(REF_LIB_STOCK):
        return REF_LIB
    if not os.path.isdir(REFERENCE_ROOT):
        return None
    subprocess.run(["make", "-s", "-j4", "-C", HERE, "ref"], check=True)  # independent objects: parallel-safe
    return REF_LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(_vp)


class CpuOracle:
    def __init__(self):
        self.lib = C.CDLL(build_cpu())
        L = self.lib
        L.rtro_num_threads.restype = _i
        L.rtro_coverage.restype = _u64
        L.rtro_coverage.argtypes = [_i, _i]
        L.rtro_f32_to_f16.restype = C.c_uint16
        L.rtro_f32_to_f16.argtypes = [C.c_float]
        L.rtro_f16_to_f32.restype = C.c_float
        L.rtro_f16_to_f32.argtypes = [C.c_uint16]
        L.rtro_cam_proj.argtypes = [_vp, _vp, _vp]
        L.rtro_project.argtypes = [_vp, _i, _u64, _vp, _i, _i, _vp, _vp]
        L.rtro_project_distorted.argtypes = [_vp, _i, _u64, _vp, _vp, _vp, C.c_float, _i, _i, _vp, _vp]
        L.rtro_clear.argtypes = [_vp, _vp, _i, _i]
        L.rtro_zmin.argtypes = [_vp, _vp, _u64, _vp]
        L.rtro_accumulate.argtypes = [_vp, _vp, _vp, _u64, _vp, _vp]
        L.rtro_resolve.argtypes = [_vp, _vp, _i, _i]
        L.rtro_minmax.argtypes = [_vp, _u64, _vp, _vp]
        L.rtro_reduce.argtypes = [_vp, _vp, _i, _i]
        L.rtro_laplacian.argtypes = [_vp, _vp, _i, _i]
        L.rtro_compare.argtypes = [_vp, _vp, _vp, _vp, _i, _i]
        L.rtro_resize.argtypes = [_vp, _vp, _vp, _i, _i]
        L.rtro_remove_mask.argtypes = [_vp, _vp, _vp, _vp, C.c_uint32, C.c_uint32, _i, _i]
        L.rtro_depth_filter.argtypes = [_vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp]
        L.rtro_free.argtypes = [_vp]
        L.rtro_render.argtypes = [_vp, _vp, _vp, _u64, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]
        L.rtro_point_passes_packed.argtypes = [_vp, _u64, _vp, _i, _i, _vp, _vp]
        L.rtro_synth_packed.argtypes = [_u64, _u64, _u64, _u64, _i, _i, _i, _i, _vp]

    @property
    def threads(self) -> int:
        return int(self.lib.rtro_num_threads())

    def cam_proj(self, K, E) -> np.ndarray:
        K = np.ascontiguousarray(np.asarray(K, dtype=np.float64).reshape(9))
        E = np.ascontiguousarray(np.asarray(E, dtype=np.float64).reshape(16))
        out = np.zeros(16, dtype=np.float32)
        self.lib.rtro_cam_proj(_p(K), _p(E), _p(out))
        return out

    def project(self, records: np.ndarray, m16, W, H):
        rec = np.ascontiguousarray(records, dtype=np.float32).reshape(-1, 4)
        m = np.ascontiguousarray(np.asarray(m16, dtype=np.float32).reshape(16))
        pix = np.empty(len(rec), dtype=np.int32)
        zb = np.empty(len(rec), dtype=np.uint32)
        self.lib.rtro_project(_p(rec), 4, len(rec), _p(m), W, H, _p(pix), _p(zb))
        return pix, zb

    def project_distorted(self, records: np.ndarray, E, K, dist, r2_max: float, W, H):
        """CPU restatement of the NEW lens-distortion projection (csrc/rtr_common.cuh project_distorted), bit-exact."""
        rec = np.ascontiguousarray(records, dtype=np.float32).reshape(-1, 4)
        E, K = np.asarray(E, np.float64), np.asarray(K, np.float64).reshape(3, 3)
        e12 = np.ascontiguousarray(E.reshape(4, 4)[:3].astype(np.float32).reshape(12))
        intr = np.array([K[0, 0], K[1, 1], K[0, 2], K[1, 2], K[0, 1]], np.float32)
        d = np.ascontiguousarray(np.asarray(dist, np.float64).astype(np.float32))
        pix, zb = np.empty(len(rec), np.int32), np.empty(len(rec), np.uint32)
        self.lib.rtro_project_distorted(_p(rec), 4, len(rec), _p(e12), _p(intr), _p(d), float(r2_max), W, H, _p(pix), _p(zb))
        return pix, zb

    def new_buffers(self, W, H):
        """Zero-initialised persistent frame buffers (the parity definition of 'allocation')."""
        P = W * H
        return dict(zbuf=np.zeros(P, np.uint32), accum=np.zeros(P * 4, np.uint32), image=np.zeros(P * 3, np.uint8),
                    tensor=np.zeros(P * 5, np.uint16), minmax=np.zeros(2, np.uint32))

    def render(self, pix, zbits, bgra, W, H, filtered: bool, buf=None):
        buf = buf or self.new_buffers(W, H)
        pix = np.ascontiguousarray(pix, dtype=np.int32)
        zbits = np.ascontiguousarray(zbits, dtype=np.uint32)
        bgra = np.ascontiguousarray(bgra, dtype=np.uint32)
        mm = buf["minmax"]
        self.lib.rtro_render(_p(pix), _p(zbits), _p(bgra), len(pix), W, H, int(filtered), _p(buf["zbuf"]),
                             _p(buf["accum"]), _p(buf["image"]), _p(buf["tensor"]), _p(mm[0:1]), _p(mm[1:2]))
        return buf

    def stages(self, pix, zbits, bgra, W, H):
        """clear -> zmin -> accumulate -> resolve with a snapshot after each (stage taps 1-3)."""
        buf = self.new_buffers(W, H)
        pix = np.ascontiguousarray(pix, dtype=np.int32)
        zbits = np.ascontiguousarray(zbits, dtype=np.uint32)
        bgra = np.ascontiguousarray(bgra, dtype=np.uint32)
        self.lib.rtro_clear(_p(buf["zbuf"]), _p(buf["accum"]), W, H)
        self.lib.rtro_zmin(_p(pix), _p(zbits), len(pix), _p(buf["zbuf"]))
        self.lib.rtro_accumulate(_p(pix), _p(zbits), _p(bgra), len(pix), _p(buf["zbuf"]), _p(buf["accum"]))
        self.lib.rtro_resolve(_p(buf["accum"]), _p(buf["image"]), W, H)
        return buf

    def depth_filter(self, zbuf_u32, image, W, H, taps: bool = False):
        """applyDepthFilter on copies; returns dict(depth(u32 view), image, tensor, minmax[, levels, masks, dims])."""
        depth = np.ascontiguousarray(zbuf_u32, dtype=np.uint32).copy()
        img = np.ascontiguousarray(image, dtype=np.uint8).copy()
        tensor = np.zeros(W * H * 5, dtype=np.uint16)
        mn, mx = C.c_uint32(0), C.c_uint32(0)
        dims = (C.c_int * 10)()
        lv = (C.c_void_p * 5)()
        mk = (C.c_void_p * 4)()
        self.lib.rtro_depth_filter(_p(depth), _p(img), _p(tensor), W, H, C.byref(mn), C.byref(mx),
                                   lv if taps else None, mk if taps else None, dims)
        out = dict(depth=depth, image=img, tensor=tensor, minmax=np.array([mn.value, mx.value], np.uint32), dims=list(dims))
        if taps:
            d = list(dims)
            uw = [d[0] >> i for i in range(5)]
            uh = [d[1] >> i for i in range(5)]
            out["levels"], out["masks"] = {}, {}
            for i in range(1, 5):
                n = d[2 * i] * d[2 * i + 1]
                out["levels"][i] = np.ctypeslib.as_array(C.cast(lv[i], C.POINTER(C.c_float)), shape=(n,)).copy() if n else np.zeros(0, np.float32)
                self.lib.rtro_free(lv[i])
            for i in range(4):
                n = uw[i] * uh[i]
                out["masks"][i] = np.ctypeslib.as_array(C.cast(mk[i], C.POINTER(C.c_uint8)), shape=(n,)).copy() if n else np.zeros(0, np.uint8)
                self.lib.rtro_free(mk[i])
        return out

    def point_passes_packed(self, records, m16, W, H, zbuf, accum):
        rec = np.ascontiguousarray(records, dtype=np.float32)
        m = np.ascontiguousarray(np.asarray(m16, dtype=np.float32).reshape(16))
        self.lib.rtro_point_passes_packed(_p(rec), rec.size // 4, _p(m), W, H, _p(zbuf), _p(accum))

    def synth_packed(self, seed, n_total, first, count, hall=(48, 40, 12), n_boxes=12) -> np.ndarray:
        out = np.empty((count, 4), dtype=np.float32)
        self.lib.rtro_synth_packed(seed, n_total, first, count, hall[0], hall[1], hall[2], n_boxes, _p(out))
        return out

    def f32_to_f16(self, x: float) -> int:
        return int(self.lib.rtro_f32_to_f16(C.c_float(x)))


class RefOracle:
    """The reference's ProjectCloud, compiled unmodified (GPU required)."""

    def __init__(self, xyz: np.ndarray, bgr: np.ndarray, stock: bool = False, model_name: str | None = None):
        path = REF_LIB_STOCK if stock else REF_LIB
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `make -C oracle ref` where /root/reference exists")
        self.lib = C.CDLL(path)
        L = self.lib
        L.ref_create.restype = _vp
        L.ref_create.argtypes = [_vp, _vp, C.c_size_t]
        L.ref_destroy.argtypes = [_vp]
        L.ref_block_size.argtypes = [_vp]
        L.ref_compute_rgbd.argtypes = [_vp, _i, _i, _vp, _vp, _vp, _vp]
        L.ref_compute_filtered.argtypes = [_vp, _i, _i, _vp, _vp, _vp, _vp]
        L.ref_read.argtypes = [_vp, _i, _vp, C.c_size_t]
        L.ref_set_cam_proj_raw.argtypes = [_vp, _vp]
        L.ref_time_point_kernels.argtypes = [_vp, _i, _i, _i, _vp]
        xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        bgr = np.ascontiguousarray(bgr, dtype=np.uint8).reshape(-1, 3)
        self.n = len(xyz)
        L.ref_create_with_model.restype = _vp
        L.ref_create_with_model.argtypes = [_vp, _vp, C.c_size_t, C.c_char_p]
        L.ref_compute_full.argtypes = [_vp, _i, _i, _vp, _vp, _vp, _vp]
        if model_name:   # the reference looks the file up in $HOME/.render_cache (project_cloud.cu:227)
            self.h = L.ref_create_with_model(_p(xyz), _p(bgr), self.n, model_name.encode())
        else:
            self.h = L.ref_create(_p(xyz), _p(bgr), self.n)
        if not self.h:
            raise RuntimeError("ref_create failed" + (f" (is ~/.render_cache/{model_name} there?)" if model_name else ""))

    @property
    def block_size(self) -> int:
        return int(self.lib.ref_block_size(self.h))

    def _call(self, fn, W, H, K, E, want_color=True, want_depth=True):
        K = np.ascontiguousarray(np.asarray(K, dtype=np.float64).reshape(9))
        E = np.ascontiguousarray(np.asarray(E, dtype=np.float64).reshape(16))
        color = np.zeros(W * H * 3, np.uint8) if want_color else None
        depth = np.zeros(W * H, np.float32) if want_depth else None
        rc = fn(self.h, W, H, _p(K), _p(E), _p(color), _p(depth))
        return rc, color, depth

    def computeRGBD(self, W, H, K, E, **kw):
        return self._call(self.lib.ref_compute_rgbd, W, H, K, E, **kw)

    def computeFilteredRGBD(self, W, H, K, E, **kw):
        return self._call(self.lib.ref_compute_filtered, W, H, K, E, **kw)

    def computeFull(self, W, H, K, E, **kw):
        return self._call(self.lib.ref_compute_full, W, H, K, E, **kw)

    _WHAT = {"zbuf": (0, np.uint32), "accum": (1, np.uint32), "image": (2, np.uint8), "tensor": (3, np.uint16),
             "cam_proj": (4, np.float32), "min": (5, np.uint32), "max": (6, np.uint32)}

    def read(self, what: str, count: int) -> np.ndarray:
        code, dt = self._WHAT[what]
        out = np.empty(count, dtype=dt)
        if self.lib.ref_read(self.h, code, _p(out), out.nbytes) != 1:
            raise RuntimeError(f"ref_read({what}) failed")
        return out

    def time_point_kernels(self, W, H, iters=5) -> np.ndarray:
        ms = np.zeros(4, np.float32)
        if self.lib.ref_time_point_kernels(self.h, W, H, iters, _p(ms)) != 1:
            raise RuntimeError("ref_time_point_kernels failed")
        return ms

    def close(self):
        if self.h:
            self.lib.ref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RefHost:
    """CPU-only entry points of the compiled reference (no GPU needed): its .oct writer/reader
    (Octreegrid.h:53-114) and CameraCalibration::loadCalibration(file) (CameraCalibration.cpp:101-209)."""

    def __init__(self):
        if not os.path.exists(REF_LIB):
            raise RuntimeError(f"{REF_LIB} missing: run `make -C oracle ref` where /root/reference exists")
        self.lib = C.CDLL(REF_LIB)
        self.lib.ref_oct_write.argtypes = [C.c_char_p, _vp, _vp, _vp, C.c_size_t, _i, _i, _i]
        self.lib.ref_oct_read.argtypes = [C.c_char_p, _vp, _vp, C.c_size_t, C.POINTER(C.c_size_t), _vp]
        self.lib.ref_load_calibration.argtypes = [C.c_char_p, _vp, _vp, _vp, _vp, _vp, _vp]
        if hasattr(self.lib, "ref_load_ply"):
            self.lib.ref_load_ply.argtypes = [C.c_char_p, _vp, _vp, _vp, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]

    def oct_write(self, path, xyz, bgr, keys, dims):
        xyz = np.ascontiguousarray(xyz, np.float32)
        bgr = np.ascontiguousarray(bgr, np.uint8)
        keys = np.ascontiguousarray(keys, np.int32)
        assert self.lib.ref_oct_write(os.fsencode(path), _p(xyz), _p(bgr), _p(keys), len(keys), *[int(d) for d in dims]) == 1

    def oct_read(self, path, cap):
        xyz4 = np.zeros((cap, 4), np.float32)
        bgra = np.zeros((cap, 4), np.uint8)
        n = C.c_size_t(0)
        dims = np.zeros(3, np.int32)
        assert self.lib.ref_oct_read(os.fsencode(path), _p(xyz4), _p(bgra), cap, C.byref(n), _p(dims)) == 1
        return xyz4[:n.value], bgra[:n.value], tuple(int(d) for d in dims)

    def load_ply(self, path, cap):
        """CloudReader::loadCloud(path) for a .ply (cloudreader.cpp:122-177, 8-82): (xyz, bgr, block key per point, blocks)."""
        xyz, bgr, keys = np.zeros((cap, 3), np.float32), np.zeros((cap, 3), np.uint8), np.zeros(cap, np.int32)
        n, nb = C.c_size_t(0), C.c_size_t(0)
        assert self.lib.ref_load_ply(os.fsencode(path), _p(xyz), _p(bgr), _p(keys), cap, C.byref(n), C.byref(nb)) == 1
        return xyz[:n.value], bgr[:n.value], keys[:n.value], int(nb.value)

    def load_calibration(self, path):
        W, H, nd, fe = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
        K, d = np.zeros(9), np.zeros(8)
        rc = self.lib.ref_load_calibration(os.fsencode(path), C.byref(W), C.byref(H), _p(K), _p(d), C.byref(nd), C.byref(fe))
        if rc != 1:
            return None
        return dict(W=W.value, H=H.value, K=K.reshape(3, 3), dist=d[:nd.value].copy(), fisheye=bool(fe.value))


_cpu = None


def cpu() -> CpuOracle:
    global _cpu
    if _cpu is None:
        _cpu = CpuOracle()
    return _cpu
