#!/usr/bin/env python
"""GPU sweep of the point-pass kernel variants + the L2 atomic micro-benchmark on the C3 workload
(100 M points, 1920x1080).  Writes gpurun_out/sweep.json.  Measurement support, not a test."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402
import bench  # noqa: E402


def stage_times(pc, pkg, poses, frames=8):
    pc.set_option("timing", 2)
    pc.stage_ms_sum(reset=True)
    for i in range(frames):
        pc._check(pc._lib.rtr_set_pose_w2c(pc._h, poses[i].ctypes.data_as(pkg._dp)))
        pc.render_device(pkg.STAGE_FILTERED)
    s, n = pc.stage_ms_sum(reset=True)
    pc.set_option("timing", 0)
    return (s / n).tolist()


def main():
    wl_name = sys.argv[1] if len(sys.argv) > 1 else "c3"
    pkg = entry.load_package()
    n, W, H, f, cx, cy, hall, boxes, seed, n_poses = bench.WORKLOADS[wl_name]
    if len(sys.argv) > 2:
        n = int(sys.argv[2])
    pc = pkg.ProjectCloud.synthetic(seed=seed, n_total=n, hall=hall, n_boxes=boxes)
    pc.set_camera(bench.make_calib(pkg, W, H, f, cx, cy))
    poses = bench.trajectory(pkg, hall, n_poses)
    poses = np.ascontiguousarray(poses[:: max(1, len(poses) // 8)][:8].reshape(-1, 16))
    out = {"workload": wl_name, "points": n, "zmin": {}, "blend": {}, "key64": {}, "atomics": {}, "culled": {}}
    stage_times(pc, pkg, poses)  # warm-up
    # ---- chunk culling on (default): list kernels, unroll fixed at 4
    pc.cull_stats(reset=True)
    out["culled"]["frame_stages_default"] = stage_times(pc, pkg, poses)
    fr, vis, nch = pc.cull_stats(reset=True)
    out["culled"]["visible_chunk_fraction"] = vis / fr / nch
    for v in (0, 1, 2, 3, 5, 7):
        pc.set_option("zmin_variant", v)
        out["culled"][f"zmin_v{v}"] = stage_times(pc, pkg, poses)[1]
    pc.set_option("zmin_variant", 1)
    for v in (0, 2):
        pc.set_option("blend_variant", v)
        out["culled"][f"blend_v{v}"] = stage_times(pc, pkg, poses)[2]
    pc.set_option("blend_variant", 0)
    pc.set_option("key64", 1)
    for v in (0, 1, 5):
        pc.set_option("zmin_variant", v)
        out["culled"][f"key64_zmin_v{v}"] = stage_times(pc, pkg, poses)[1]
    pc.set_option("key64", 0)
    pc.set_option("zmin_variant", 1)
    pc.set_option("chunk_cull", 0)
    # ---- chunk culling off: every point streamed
    for v in (0, 1, 2, 3, 5, 7):
        for u in (1, 2, 4, 8):
            pc.set_option("zmin_variant", v)
            pc.set_option("zmin_unroll", u)
            out["zmin"][f"v{v}_u{u}"] = stage_times(pc, pkg, poses)[1]
    pc.set_option("zmin_variant", 1)
    pc.set_option("zmin_unroll", 4)
    for v in (0, 2):
        for u in (1, 2, 4, 8):
            pc.set_option("blend_variant", v)
            pc.set_option("blend_unroll", u)
            out["blend"][f"v{v}_u{u}"] = stage_times(pc, pkg, poses)[2]
    pc.set_option("blend_variant", 0)
    pc.set_option("blend_unroll", 4)
    pc.set_option("key64", 1)
    for v in (0, 1, 5):
        pc.set_option("zmin_variant", v)
        out["key64"][f"v{v}_u4"] = stage_times(pc, pkg, poses)[1]
    pc.set_option("key64", 0)
    pc.set_option("zmin_variant", 1)
    # L2 atomics
    pc._check(pc._lib.rtr_set_pose_w2c(pc._h, poses[0].ctypes.data_as(pkg._dp)))
    for key64 in (False, True):
        ms, ops = pc.bench_red_min(0, 200_000_000, key64)
        out["atomics"][f"random_u{64 if key64 else 32}"] = {"ms": ms, "ops": ops, "Gops_per_s": ops / ms / 1e6}
        ms, ops = pc.bench_red_min(1, 0, key64)
        out["atomics"][f"projected_u{64 if key64 else 32}"] = {"ms": ms, "ops": ops, "points": n, "Gops_per_s": ops / ms / 1e6}
    pc.close()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"sweep_{wl_name}.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
