#!/usr/bin/env python
"""Experiment: distinct pixels per group of g consecutive records of the Morton-ordered C3 cloud (g = 4, 8, 16, 32,
128), and distinct 32-byte z-buffer / accumulator sectors per RED instruction under the ring kernels' lane mapping
(lane l, slot s <-> record 4l + s of a 128-record warp tile) with and without the in-register merge."""
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
pc = pkg.ProjectCloud.synthetic(seed=seed, n_total=n, hall=hall, n_boxes=boxes)
calib = bench.make_calib(pkg, W, H, f, cx, cy)
out = {}
for pi in (0, 500):
    pc.set_camera(calib, poses[pi])
    pix, zb = pc.project_points()
    m = (len(pix) // 128) * 128
    res = {"live": int((pix >= 0).sum())}
    for g in (4, 8, 16, 32, 128):
        p = pix[:m].reshape(-1, g)
        rows = (p >= 0).any(axis=1)
        ps = np.sort(p[rows], axis=1)
        d = np.ones_like(ps, dtype=bool)
        d[:, 1:] = ps[:, 1:] != ps[:, :-1]
        res[f"distinct_px_per_live_g{g}"] = float((d & (ps >= 0)).sum() / (ps >= 0).sum())
    # sectors per RED instruction: tile of 128 records, instruction s carries records 4l+s
    t = pix[:m].reshape(-1, 32, 4)
    rows = (t >= 0).any(axis=(1, 2))
    t = t[rows]
    for name, shift in (("zbuf_u32", 3), ("accum_16B", 1)):
        sec = np.where(t >= 0, t >> shift, -1)
        tot_sectors = 0
        for s in range(4):
            ss = np.sort(sec[:, :, s], axis=1)
            d = np.ones_like(ss, dtype=bool)
            d[:, 1:] = ss[:, 1:] != ss[:, :-1]
            tot_sectors += int((d & (ss >= 0)).sum())
        res[f"sectors_per_instr_{name}"] = tot_sectors / (4 * len(t))
        # union over the whole 128-record tile
        su = np.sort(sec.reshape(len(t), 128), axis=1)
        d = np.ones_like(su, dtype=bool)
        d[:, 1:] = su[:, 1:] != su[:, :-1]
        res[f"sectors_per_tile_union_{name}"] = float((d & (su >= 0)).sum() / len(t))
        # consecutive mapping: instruction s carries records 32s .. 32s+31
        c = sec.reshape(len(t), 4, 32)
        tot2 = 0
        for s in range(4):
            ss = np.sort(c[:, s, :], axis=1)
            d = np.ones_like(ss, dtype=bool)
            d[:, 1:] = ss[:, 1:] != ss[:, :-1]
            tot2 += int((d & (ss >= 0)).sum())
        res[f"sectors_per_instr_consecutive_{name}"] = tot2 / (4 * len(t))
    out[f"pose{pi}"] = res
pc.close()
print(json.dumps(out, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "merge_group_sizes.json"), "w"), indent=1)
