#!/usr/bin/env python
"""Experiment: where does the culled z-min pass spend its time?  (measurement only)"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import __graft_entry__ as entry
import bench
from sweep import stage_times

pkg = entry.load_package()
n, W, H, f, cx, cy, hall, boxes, seed, n_poses = bench.WORKLOADS["c3"]
pc = pkg.ProjectCloud.synthetic(seed=seed, n_total=n, hall=hall, n_boxes=boxes)
pc.set_camera(bench.make_calib(pkg, W, H, f, cx, cy))
poses = bench.trajectory(pkg, hall, n_poses)
poses = np.ascontiguousarray(poses[:: len(poses) // 8][:8].reshape(-1, 16))
stage_times(pc, pkg, poses)
out = {}
for cull in (1, 0):
    pc.set_option("chunk_cull", cull)
    for v in ((0, 1, 5, 8, 9, 16, 17, 24, 32, 33, 40) if cull else (0, 1, 8)):
        pc.set_option("zmin_variant", v)
        out[f"cull{cull}_v{v}"] = stage_times(pc, pkg, poses)[1]
print(json.dumps(out, indent=1))
