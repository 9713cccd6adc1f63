#!/usr/bin/env python
"""A/B of how the ring kernels' list passes hand out their tiles (option ring_dynamic): claimed from a counter one
iteration ahead (1) vs round-robin (0), crossed with clear_lean (clear + classify at 64 vs 80 registers).  For each setting, alternating twice: per-stage CUDA-event times (one
frame at a time) and the frame rate of `frames` back-to-back asynchronous frames (two frames in flight).  Also checks
that the two settings give byte-identical frames.  Writes gpurun_out/exp_ring_dynamic_<workload>.json.

    python tools/experiments/ring_dynamic_ab.py [workload=c3] [frames=1000] [key=value ...]   (extra renderer options)
    python tools/experiments/ring_dynamic_ab.py c3 1000 combos zmin_variant=13 zmin_variant=8,ring_dynamic=0 ...
        (after the word "combos": one comma-separated option set per argument, measured in turn, twice, against the
         defaults; variants with bit 3 set are timing-only — they issue no REDs and their frames are wrong)
"""
import json
import os
import sys
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import __graft_entry__ as entry  # noqa: E402
import bench  # noqa: E402
from sweep import stage_times  # noqa: E402

NAMES = ["clear_classify", "zmin", "blend", "resolve_pyramid", "up_pass", "frame"]


def frame_rate(pc, pkg, poses, frames):
    pc.sync()
    t0 = time.perf_counter()
    for i in range(frames):
        pc._check(pc._lib.rtr_set_pose_w2c(pc._h, poses[i % len(poses)].ctypes.data_as(pkg._dp)))
        pc.render_device(pkg.STAGE_FILTERED)
    pc.sync()
    return frames / (time.perf_counter() - t0)


def checksums(pc, pkg, poses, W, H):
    out = []
    for i in range(0, len(poses), max(1, len(poses) // 6)):
        pc._check(pc._lib.rtr_set_pose_w2c(pc._h, poses[i].ctypes.data_as(pkg._dp)))
        pc.render_device(pkg.STAGE_FILTERED)
        pc.sync()
        out.append([zlib.crc32(pc.read("zbuf", np.uint32, W * H).tobytes()), zlib.crc32(pc.read("accum", np.uint32, 4 * W * H).tobytes()),
                    zlib.crc32(pc.read("image", np.uint8, 3 * W * H).tobytes()), zlib.crc32(pc.read("tensor", np.uint16, 5 * W * H).tobytes())])
    return out


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
    frames = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    rest = sys.argv[3:]
    custom = None
    if "combos" in rest:
        i = rest.index("combos")
        custom = [dict((k, int(v)) for k, v in (kv.split("=") for kv in a.split(","))) for a in rest[i + 1:]]
        rest = rest[:i]
    extra = dict((k, int(v)) for k, v in (a.split("=") for a in rest))
    pkg = entry.load_package()
    n, W, H, f, cx, cy, hall, boxes, seed, n_poses = bench.WORKLOADS[wl]
    pc = pkg.ProjectCloud.synthetic(seed=seed, n_total=n, hall=hall, n_boxes=boxes)
    for k, v in extra.items():
        pc.set_option(k, v)
    pc.set_camera(bench.make_calib(pkg, W, H, f, cx, cy))
    poses = np.ascontiguousarray(bench.trajectory(pkg, hall, n_poses).reshape(-1, 16))
    sub = np.ascontiguousarray(poses[:: max(1, len(poses) // 40)][:40])
    out = {"workload": wl, "points": n, "frames": frames, "options": extra, "runs": []}
    sums = {}
    default_queues = pc.get_option("ring_dynamic")
    for dyn in (8, 0):
        pc.set_option("ring_dynamic", dyn)
        sums[dyn] = checksums(pc, pkg, sub, W, H)   # the float colour sums are exact integers: they compare byte for byte too
    pc.set_option("ring_dynamic", default_queues)
    out["identical_frames"] = sums[0] == sums[8]
    print("identical frames:", out["identical_frames"], flush=True)
    frame_rate(pc, pkg, poses, 200)  # warm-up
    combos = [dict(ring_dynamic=d) for d in (0, 4, 8, 16)]
    if custom is not None:
        combos = [dict()] + custom
    defaults = {k: pc.get_option(k) for c in combos for k in c}
    for combo in combos + combos:
        for k, v in {**defaults, **combo}.items():
            pc.set_option(k, v)
        stage_times(pc, pkg, sub, 4)
        st = dict(zip(NAMES, stage_times(pc, pkg, sub, len(sub))))
        fps = frame_rate(pc, pkg, poses, frames)
        out["runs"].append({"options": combo, "stage_ms": st, "frames_per_s": fps})
        print(combo, {k: round(v * 1e3, 1) for k, v in st.items()}, "frames/s", round(fps), flush=True)
    pc.close()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"exp_ring_dynamic_{wl}.json"), "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
