#!/usr/bin/env python
"""Experiment: how many REDs would warp-level merging save on C3 with the Morton-ordered cloud?
Counts, over the in-frustum points of a few poses: lanes, adjacent-equal runs per 32-lane warp, distinct pixels per warp."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import __graft_entry__ as entry
import bench
pkg = entry.load_package()
n, W, H, f, cx, cy, hall, boxes, seed, n_poses = bench.WORKLOADS["c3"]
poses = bench.trajectory(pkg, hall, n_poses)
out = {}
for order in ("morton", "synthetic"):
    pc = pkg.ProjectCloud.synthetic(seed=seed, n_total=n, hall=hall, n_boxes=boxes, sort=(order == "morton"))
    calib = bench.make_calib(pkg, W, H, f, cx, cy)
    tot = dict(live=0, runs=0, distinct=0, pixels=0)
    for pi in (0, 250, 500):
        pc.set_camera(calib, poses[pi])
        pix, zb = pc.project_points()
        m = (len(pix) // 32) * 32
        p = pix[:m].reshape(-1, 32)
        live = p >= 0
        rows = live.any(axis=1)
        p, live = p[rows], live[rows]
        head = np.ones_like(p, dtype=bool)
        head[:, 1:] = p[:, 1:] != p[:, :-1]
        tot["live"] += int(live.sum())
        tot["runs"] += int((head & live).sum())
        ps = np.sort(np.where(live, p, -1), axis=1)
        d = np.ones_like(ps, dtype=bool)
        d[:, 1:] = ps[:, 1:] != ps[:, :-1]
        tot["distinct"] += int((d & (ps >= 0)).sum())
        tot["pixels"] += int(len(np.unique(pix[pix >= 0])))
    out[order] = dict(tot, runs_per_live=tot["runs"] / tot["live"], distinct_per_live=tot["distinct"] / tot["live"],
                      pixels_per_live=tot["pixels"] / tot["live"])
    pc.close()
print(json.dumps(out, indent=1))
