#!/usr/bin/env python
"""Experiment: effect of point order (synthetic patch order / cell-binned / Morton) on the point passes."""
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
poses = bench.trajectory(pkg, hall, n_poses)
poses = np.ascontiguousarray(poses[:: len(poses) // 8][:8].reshape(-1, 16))
out = {}
for order in ("synthetic", "cells", "morton"):
    pc = pkg.ProjectCloud.synthetic(seed=seed, n_total=n, hall=hall, n_boxes=boxes)
    pc.set_camera(bench.make_calib(pkg, W, H, f, cx, cy))
    if order == "cells":
        pc.bin_cells()
    elif order == "morton":
        pc.sort_morton()
    stage_times(pc, pkg, poses)
    for cull in (1, 0):
        pc.set_option("chunk_cull", cull)
        for zv in (0, 1, 5, 3):
            pc.set_option("zmin_variant", zv)
            for bv in (4, 6):
                pc.set_option("blend_variant", bv)
                pc.cull_stats()
                st = stage_times(pc, pkg, poses)
                fr, vis, nch = pc.cull_stats()
                out[f"{order}_cull{cull}_z{zv}_b{bv}"] = {"zmin": st[1], "blend": st[2], "frame": st[5], "vis": vis / max(fr, 1) / nch}
    pc.close()
print(json.dumps(out, indent=1))
