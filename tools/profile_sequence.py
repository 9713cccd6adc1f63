#!/usr/bin/env python
"""A short fused frame sequence of a workload for ncu (launch list / --set full capture of the frame's kernels).

    python tools/profile_sequence.py [--workload c3] [--frames 12] [--opt key=value ...]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3", choices=list(bench.WORKLOADS))
    ap.add_argument("--frames", type=int, default=12)
    ap.add_argument("--first-pose", type=int, default=500)
    ap.add_argument("--opt", action="append", default=[])
    args = ap.parse_args()
    pkg = entry.load_package()
    n, W, H, f, cx, cy, hall, boxes, seed, n_poses = bench.WORKLOADS[args.workload]
    pc = pkg.ProjectCloud.synthetic(seed=seed, n_total=n, hall=hall, n_boxes=boxes)
    for kv in args.opt:
        k, v = kv.split("=")
        pc.set_option(k, int(v))
    calib = bench.make_calib(pkg, W, H, f, cx, cy)
    poses = bench.trajectory(pkg, hall, n_poses)
    pc.set_camera(calib)
    for i in range(args.frames):
        pc.set_camera(calib, poses[(args.first_pose + i) % n_poses])
        pc.render_device(pkg.STAGE_FILTERED)
    pc.sync()
    passes, streamed = pc.stream_stats(reset=False)
    frames, visible, n_chunks = pc.cull_stats()
    print(f"{args.workload}: {frames} frames, {passes} passes, {visible / max(frames, 1):.0f} visible chunks per frame, "
          f"{streamed / max(passes, 1):.0f} chunks per pass of {n_chunks}")
    pc.close()


if __name__ == "__main__":
    main()
