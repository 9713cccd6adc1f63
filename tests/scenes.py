"""Seeded parity scenes shared by the golden generator (tests/golden/make_golden.py, runs the
REFERENCE on a B200), the CPU tests (oracle vs golden) and the GPU tests (CUDA path vs oracle /
golden / live reference).  Pure host logic."""
from __future__ import annotations

import hashlib

import numpy as np

HALL_SMALL = (32, 24, 12)   # 8 x 6 x 3 m "room"   (SURVEY.md §8 d), quarter-metre units
HALL_LARGE = (48, 40, 12)   # 12 x 10 x 3 m "hall"


def look_at_w2c(eye, forward, up=(0.0, 0.0, 1.0)) -> np.ndarray:
    """World->camera 4x4, OpenCV camera axes (x right, y down, z forward)."""
    f = np.asarray(forward, dtype=np.float64)
    f = f / np.linalg.norm(f)
    r = np.cross(f, np.asarray(up, dtype=np.float64))
    r = r / np.linalg.norm(r)
    d = np.cross(f, r)
    R = np.stack([r, d, f])
    E = np.eye(4)
    E[:3, :3] = R
    E[:3, 3] = -R @ np.asarray(eye, dtype=np.float64)
    return E


def intrinsics(W, H, f=None, cx=None, cy=None) -> np.ndarray:
    f = 0.73 * W if f is None else f
    cx = (W - 1) / 2.0 if cx is None else cx
    cy = (H - 1) / 2.0 if cy is None else cy
    return np.array([[f, 0, cx], [0, f, cy], [0, 0, 1]], dtype=np.float64)


class Case:
    def __init__(self, name, n, hall, n_boxes, seed, W, H, K, poses, full):
        self.name, self.n, self.hall, self.n_boxes, self.seed = name, n, hall, n_boxes, seed
        self.W, self.H, self.K, self.poses, self.full = W, H, K, poses, full


_P_ROOM = [look_at_w2c((4.0, 3.0, 1.5), (1.0, 0.2, 0.0)), look_at_w2c((2.0, 2.0, 1.2), (0.6, 1.0, -0.1))]
_P_HALL = [look_at_w2c((6.0, 5.0, 1.5), (1.0, 0.3, 0.05)), look_at_w2c((8.0, 3.0, 1.4), (-1.0, 0.8, -0.1))]

# `full` = True: the golden stores every output array; False: sha256 digests (+ sparse tap diff).
CASES = {c.name: c for c in [
    Case("small_160x96", 20_000, HALL_SMALL, 4, 11, 160, 96, intrinsics(160, 96), _P_ROOM[:1], True),
    Case("small_176x104", 20_000, HALL_SMALL, 4, 12, 176, 104, intrinsics(176, 104), _P_ROOM, True),   # H % 16 != 0
    Case("odd_200x120", 20_000, HALL_SMALL, 4, 13, 200, 120, intrinsics(200, 120), _P_ROOM[:1], True),  # W % 16 != 0
    Case("c1_640x480", 1_000_000, HALL_SMALL, 6, 1234, 640, 480, intrinsics(640, 480, 525.0, 319.5, 239.5), _P_ROOM[:1], False),
    Case("c2_1280x720", 2_000_000, HALL_SMALL, 6, 1234, 1280, 720, intrinsics(1280, 720, 900.0, 639.5, 359.5), _P_ROOM[:1], False),
    Case("c3_1920x1080", 2_000_000, HALL_LARGE, 12, 5678, 1920, 1080, intrinsics(1920, 1080, 1400.0, 959.5, 539.5), _P_HALL, False),
    Case("c5_3840x2160", 2_000_000, HALL_LARGE, 12, 5678, 3840, 2160, intrinsics(3840, 2160, 2800.0, 1919.5, 1079.5), _P_HALL[:1], False),
]}

OUTPUT_KEYS = ("raw_zbuf", "raw_accum", "raw_image", "raw_depth_host", "raw_color_host",
               "flt_depth_host", "flt_color_host", "flt_tensor", "flt_minmax")


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def bgra_of(records: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(records[:, 3]).view(np.uint32)


def split_records(records: np.ndarray):
    """packed (n,4) float32 records -> xyz (n,3) float32, bgr (n,3) uint8."""
    c = bgra_of(records)
    bgr = np.stack([c & 0xFF, (c >> 8) & 0xFF, (c >> 16) & 0xFF], axis=1).astype(np.uint8)
    return np.ascontiguousarray(records[:, :3]), bgr


def oracle_frames(cpu_oracle, case: Case, records: np.ndarray, taps):
    """Run the CPU oracle over the case's pose sequence on persistent zero-initialised buffers, the
    way one reference object renders them back to back: per pose computeRGBD then
    computeFilteredRGBD.  taps[i] = (pix, zbits) of pose i.  Returns a list of dicts keyed like
    OUTPUT_KEYS."""
    W, H = case.W, case.H
    bgra = bgra_of(records)
    buf = cpu_oracle.new_buffers(W, H)
    out = []
    for (pix, zb) in taps:
        o = {}
        cpu_oracle.render(pix, zb, bgra, W, H, filtered=False, buf=buf)
        o["raw_zbuf"], o["raw_accum"], o["raw_image"] = buf["zbuf"].copy(), buf["accum"].copy(), buf["image"].copy()
        o["raw_depth_host"], o["raw_color_host"] = buf["zbuf"].copy(), buf["image"].copy()
        cpu_oracle.render(pix, zb, bgra, W, H, filtered=True, buf=buf)
        o["flt_depth_host"], o["flt_color_host"] = buf["zbuf"].copy(), buf["image"].copy()
        o["flt_tensor"], o["flt_minmax"] = buf["tensor"].copy(), buf["minmax"].copy()
        out.append(o)
    return out
