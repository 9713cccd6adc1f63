#!/usr/bin/env python
"""What would rendering a 3840x2160 frame as N horizontal bands — each band's z-buffer and colour sums L2-resident while
its two point passes run — cost?  Estimated WITHOUT new kernels: a band of a pinhole frame is itself a pinhole frame
(same fx, fy, cx; cy shifted by the band's first row), so the existing two-pass path renders 3840 x (2160 / N) frames
whose cameras are the bands of the C5 trajectory's cameras, and the per-stage CUDA-event times are summed over the bands.

    python tools/experiments/band_pass_estimate.py [--bands 5] [--frames 200] --out gpurun_out/exp.json
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


def stage_times(pkg, pc, calib, my, frames):
    pc.set_camera(calib, my[0].reshape(4, 4))
    pc.set_option("timing", 2)
    for rep in range(2):                      # first repetition warms up
        pc.stage_ms_sum(reset=True)
        pc.cull_stats(reset=True)
        for i in range(frames):
            pc._check(pc._lib.rtr_set_pose_w2c(pc._h, my[i].ctypes.data_as(pkg._dp)))
            pc.render_device(pkg.STAGE_FILTERED)
        pc.sync()
    sums, nfr = pc.stage_ms_sum(reset=True)
    cf, cvis, nch = pc.cull_stats(reset=True)
    pc.set_option("timing", 0)
    return [float(v) / max(nfr, 1) * 1e3 for v in sums], cvis / max(cf, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bands", type=int, default=5)
    ap.add_argument("--frames", type=int, default=200)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    pkg = entry.load_package()
    n, W, H, f, cx, cy, hall, boxes, seed, n_poses = bench.WORKLOADS["c5_4k"]
    assert H % args.bands == 0 and (H // args.bands) % 16 == 0, "band height must be a multiple of 16"
    hb = H // args.bands
    pc = pkg.ProjectCloud.synthetic(seed=seed, n_total=n, hall=hall, n_boxes=boxes)
    pc.set_option("fuse", 0)
    pc.set_option("bands", 1)
    poses = bench.trajectory(pkg, hall, n_poses)
    idx = bench.pose_schedule(args.frames, n_poses, 1, 0)
    my = np.ascontiguousarray(np.stack([poses[i] for i in idx]).reshape(-1, 16))
    names = ["clear_classify", "zmin", "blend", "resolve_pyramid", "up_pass", "frame"]
    whole, vis = stage_times(pkg, pc, bench.make_calib(pkg, W, H, f, cx, cy), my, args.frames)
    out = {"bands": args.bands, "band_rows": hb, "frames": args.frames,
           "whole_frame_us": dict(zip(names, whole)), "whole_frame_visible_chunks": vis, "band_us": []}
    for b in range(args.bands):
        st, v = stage_times(pkg, pc, bench.make_calib(pkg, W, hb, f, cx, cy - hb * b), my, args.frames)
        out["band_us"].append(dict(zip(names, st), visible_chunks=v))
        print(b, out["band_us"][-1], flush=True)
    tot = {k: sum(bd[k] for bd in out["band_us"]) for k in names[:4]}
    out["sum_over_bands_us"] = tot
    out["estimate_us"] = {"point_stages_banded": tot["clear_classify"] + tot["zmin"] + tot["blend"] + tot["resolve_pyramid"],
                          "point_stages_whole": sum(whole[:4]), "up_pass_whole": whole[4]}
    print(json.dumps(out["whole_frame_us"]), json.dumps(out["sum_over_bands_us"]), json.dumps(out["estimate_us"]))
    pc.close()
    if args.out:
        json.dump(out, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
