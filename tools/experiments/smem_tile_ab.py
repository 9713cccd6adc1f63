#!/usr/bin/env python
"""Shared-memory tile pre-reduction (zmin_variant bit 6) against the default z-min ring pass on two-pass frames:
per-stage CUDA-event times over a trajectory arc, the fraction of tiles whose pixels fit the 32 x 32 window, and how many
records a window absorbs per pixel it flushes.  Writes the record behind DESIGN.md's paragraph on north_star's
"shared-memory tile pre-reduction".

    python tools/experiments/smem_tile_ab.py [--workload c3] [--frames 200] --out gpurun_out/r02_exp_smem_tile.json
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--frames", type=int, default=200)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    pkg = entry.load_package()
    out = {"frames": args.frames, "workloads": {}}
    for wl_name in args.workload.split(","):
        n, W, H, f, cx, cy, hall, boxes, seed, n_poses = bench.WORKLOADS[wl_name]
        pc = pkg.ProjectCloud.synthetic(seed=seed, n_total=n, hall=hall, n_boxes=boxes)
        calib = bench.make_calib(pkg, W, H, f, cx, cy)
        poses = bench.trajectory(pkg, hall, n_poses)
        idx = bench.pose_schedule(args.frames, n_poses, 1, 0)
        pc.set_camera(calib, poses[0])
        pc.set_option("fuse", 0)
        res = {}
        for name, variant in (("default (early test + RED per surviving record)", 5), ("shared-memory tile pre-reduction", 64 | 5)):
            pc.set_option("zmin_variant", variant)
            pc.set_option("timing", 2)
            pc.cull_stats(reset=True)
            pc.stage_ms_sum(reset=True)
            for i in idx:
                pc.set_camera(calib, poses[i])
                pc.render_device(pkg.STAGE_FILTERED)
            sums, nfr = pc.stage_ms_sum(reset=True)
            stats = pc.smem_tile_stats()
            frames, visible, n_chunks = pc.cull_stats(reset=True)
            pc.set_option("timing", 0)
            r = {"zmin_us": float(sums[1]) / nfr * 1e3, "blend_us": float(sums[2]) / nfr * 1e3, "frame_us": float(sums[5]) / nfr * 1e3,
                 "visible_chunks_per_frame": visible / max(frames, 1)}
            if variant & 64:
                fit, direct, flushed, entered = stats
                r.update({"tiles_through_window": fit, "tiles_direct": direct, "window_hit_rate": fit / max(fit + direct, 1),
                          "records_entered_windows (ATOMS lanes)": entered, "window_pixels_flushed (REDG candidates)": flushed,
                          "records_per_flushed_pixel": entered / max(flushed, 1)})
            res[name] = r
        pc.close()
        out["workloads"][wl_name] = res
        print(wl_name, json.dumps(res, indent=1), flush=True)
    if args.out:
        with open(args.out, "w") as fh:
            json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
