#!/usr/bin/env python
"""A small band-ordered frame sequence, sized for a run under compute-sanitizer (memcheck / racecheck) where the
pool allows it (this round's pool does not: the tool is closed there, so the script was only run plainly):

    [compute-sanitizer --tool memcheck]  python tools/sanitize_bands.py

300 000 points, 640x480, 8 bands: blocking frames (clear_classify_kernel<4, true>), a fused sequence
(classify_pair_kernel<*, true>) and a two-pass sequence, each compared with the unordered list's frames."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402


def main():
    pkg = entry.load_package()
    W, H = 640, 480
    calib = pkg.CameraCalibration()
    calib.loadCalibration(525.0, 525.0, 319.5, 239.5, [0.0] * 5, W, H)
    poses = pkg.trajectory_w2c(300, center=(4.0, 3.0, 1.5), radius=2.0)[40:46]
    frames = {}
    for bands in (1, 8):
        for fuse in (2, 0):
            pc = pkg.ProjectCloud.synthetic(seed=1234, n_total=300_000, hall=(32, 24, 12), n_boxes=6)
            pc.set_option("bands", bands)
            pc.set_option("fuse", fuse)
            pc.set_camera(calib)
            color = np.zeros((len(poses), W * H * 3), np.uint8)
            depth = np.zeros((len(poses), W * H), np.float32)
            pc.render_trajectory(pkg.STAGE_FILTERED, poses, color, depth)
            c1, d1 = np.zeros(W * H * 3, np.uint8), np.zeros(W * H, np.float32)
            assert pc.computeFilteredRGBD(calib, poses[2], c1, d1) == 1
            frames[(bands, fuse)] = (color, depth.view(np.uint32).copy(), c1, d1.view(np.uint32).copy())
            pc.close()
    ref = frames[(1, 0)]
    for key, got in frames.items():
        assert all(np.array_equal(a, b) for a, b in zip(got, ref)), key
        assert np.array_equal(got[0][2], got[2]) and np.array_equal(got[1][2], got[3]), key
    print("sanitize_bands: frames identical", sorted(frames))


if __name__ == "__main__":
    main()
