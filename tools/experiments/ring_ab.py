#!/usr/bin/env python
"""A/B of the point-pass kernels on C3 (100 M points, 1920x1080): per-stage CUDA-event times for
ring (TMA-fed persistent kernels, in-register merge) vs the per-thread LDG.128 kernels, PTX red vs
the atomicMin builtin (ATOMG after the fence), culled and stream-all.  Writes gpurun_out/exp_ring.json."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import __graft_entry__ as entry  # noqa: E402
import bench  # noqa: E402
from sweep import stage_times  # noqa: E402

NAMES = ["clear_classify", "zmin", "blend", "resolve_pyramid", "up_pass", "frame"]


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
    frames = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    pkg = entry.load_package()
    n, W, H, f, cx, cy, hall, boxes, seed, n_poses = bench.WORKLOADS[wl]
    pc = pkg.ProjectCloud.synthetic(seed=seed, n_total=n, hall=hall, n_boxes=boxes)
    pc.set_camera(bench.make_calib(pkg, W, H, f, cx, cy))
    poses = bench.trajectory(pkg, hall, n_poses)
    poses = np.ascontiguousarray(poses[:: max(1, len(poses) // frames)][:frames].reshape(-1, 16))
    out = {"workload": wl, "points": n, "frames": len(poses), "runs": {}}
    combos = [
        ("default", dict()),
        ("round_robin", dict(ring_dynamic=0)),
        ("default_again", dict()),
        ("round_robin_again", dict(ring_dynamic=0)),
        ("nomerge", dict(zmin_variant=37, blend_variant=36)),
        ("up_per_level", dict(fused_up=0)),
        ("no_early", dict(ring_early=0)),
        ("ring_nored", dict(zmin_variant=13)),
        ("ring_intblend", dict(blend_variant=0)),
        ("ldg_red", dict(ring=0)),
        ("ldg_atomg", dict(ring=0, zmin_variant=21)),
        ("all_ring", dict(chunk_cull=0, ring=2)),
        ("all_ring_noperm", dict(chunk_cull=0, ring=2, ring_perm=0)),
        ("all_ring_nored", dict(chunk_cull=0, ring=2, zmin_variant=13)),
        ("all_ldg_red", dict(ring=0, chunk_cull=0)),
    ]
    defaults = dict(ring=1, zmin_variant=5, blend_variant=4, chunk_cull=1, fused_up=1, ring_perm=1, ring_early=1, ring_dynamic=1)
    stage_times(pc, pkg, poses, len(poses))  # warm-up
    for name, opts in combos:
        for k, v in {**defaults, **opts}.items():
            pc.set_option(k, v)
        stage_times(pc, pkg, poses, 4)
        t = stage_times(pc, pkg, poses, len(poses))
        out["runs"][name] = dict(zip(NAMES, t))
        print(name, {k: round(v * 1e3, 1) for k, v in out["runs"][name].items()}, flush=True)
    pc.close()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"exp_ring_{wl}.json"), "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
