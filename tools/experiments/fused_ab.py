#!/usr/bin/env python
"""A/B of the fused frame sequence on one workload: frames/s of the device-resident loop for a list of
(library build, environment, options) configurations, each in its own process.

    python tools/experiments/fused_ab.py [--workload c3] [--frames 400] --out gpurun_out/exp.json  CONFIG ...
    CONFIG = name[:lib=<suffix>][:env=K=V,...][:opt=k=v,...]      e.g.  r48:lib=_r48   twopass:opt=fuse=0

Experiment builds are made on the CPU box first:  RTR_LIB_SUFFIX=_r48 RTR_NVCC_FLAGS=-DRTR_FUSED_REGS=48 python <pkg>/build.py --force
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PKG = "real-time-neural-rendering-of-lidar-point-clouds_b200"


def child(args):
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    import __graft_entry__ as entry
    import bench
    pkg = entry.load_package()
    n, W, H, f, cx, cy, hall, boxes, seed, n_poses = bench.WORKLOADS[args.workload]
    pc = pkg.ProjectCloud.synthetic(seed=seed, n_total=n, hall=hall, n_boxes=boxes)
    for kv in (args.opt.split(",") if args.opt else []):
        k, v = kv.split("=")
        pc.set_option(k, int(v))
    calib = bench.make_calib(pkg, W, H, f, cx, cy)
    if args.distort:
        calib.setDistortionParameters([-0.05, 0.01, 0.0005, -0.0005, 0.0])
        pc.apply_distortion = True
    poses = bench.trajectory(pkg, hall, n_poses)
    idx = bench.pose_schedule(args.frames + 5, n_poses, 1, 0)
    my = np.ascontiguousarray(np.stack([poses[i] for i in idx]).reshape(-1, 16))
    pc.set_camera(calib, poses[0])
    pc.render_device(pkg.STAGE_FILTERED)
    stream = torch.cuda.ExternalStream(pc.device_buffers().stream)
    best = None
    for rep in range(3):
        for i in range(5):
            pc._check(pc._lib.rtr_set_pose_w2c(pc._h, my[i].ctypes.data_as(pkg._dp)))
            pc.render_device(pkg.STAGE_FILTERED)
        pc.sync()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        t0 = time.perf_counter()
        for i in range(args.frames):
            pc._check(pc._lib.rtr_set_pose_w2c(pc._h, my[5 + i].ctypes.data_as(pkg._dp)))
            pc.render_device(pkg.STAGE_FILTERED)
        host_us = (time.perf_counter() - t0) / args.frames * 1e6   # what the enqueueing thread spends per frame (ctypes calls included)
        pc.device_buffers()
        e1.record(stream)
        pc.sync()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.frames
        if best is None or ms < best:
            best, best_host = ms, host_us
    out = {"ms_per_frame": best, "frames_per_s": 1e3 / best, "host_enqueue_us_per_frame": best_host}
    print("RESULT " + json.dumps(out), flush=True)     # (the per-stage part below can fail in measurement-only builds)
    # per-stage events of the same loop
    timing = 3 if pc.get_option("fuse_active") == 1 else 2
    pc.set_option("timing", timing)
    pc.stage_ms_sum(reset=True)
    pc.stream_stats(reset=True)
    for i in range(args.frames):
        pc._check(pc._lib.rtr_set_pose_w2c(pc._h, my[5 + i].ctypes.data_as(pkg._dp)))
        pc.render_device(pkg.STAGE_FILTERED)
    pc.sync()
    sums, nfr = pc.stage_ms_sum(reset=True)
    passes, streamed = pc.stream_stats(reset=True)
    out["timing_mode"] = timing
    out["stage_us"] = [round(float(v) / max(nfr, 1) * 1e3, 2) for v in sums]
    out["chunks_per_pass"] = streamed / max(passes, 1)
    pc.close()
    print("RESULT " + json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--frames", type=int, default=400)
    ap.add_argument("--out", default=None)
    ap.add_argument("--child", action="store_true")
    ap.add_argument("--opt", default="")
    ap.add_argument("--distort", action="store_true", help="config 2's k1,k2,p1,p2,k3")
    ap.add_argument("configs", nargs="*")
    args = ap.parse_args()
    if args.child:
        return child(args)
    results = {}
    for cfg in args.configs:
        parts = cfg.split(":")
        name, lib, env, opt = parts[0], "", {}, ""
        for p in parts[1:]:
            if p.startswith("lib="):
                lib = p[4:]
            elif p.startswith("env="):
                env = dict(kv.split("=", 1) for kv in p[4:].split(","))
            elif p.startswith("opt="):
                opt = p[4:]
        e = dict(os.environ, **env)
        if lib:
            e["RTR_B200_LIB"] = os.path.join(ROOT, PKG, f"librtr_b200{lib}.so")
        cmd = [sys.executable, os.path.abspath(__file__), "--child", "--workload", args.workload, "--frames", str(args.frames), "--opt", opt] + (["--distort"] if args.distort else [])
        res = subprocess.run(cmd, capture_output=True, text=True, env=e, timeout=600)
        line = [ln for ln in res.stdout.splitlines() if ln.startswith("RESULT ")]
        results[name] = dict(json.loads(line[-1][7:]), lib=lib or "(default)", env=env, opt=opt) if line else {"error": (res.stdout + res.stderr)[-800:]}
        print(name, json.dumps(results[name]), flush=True)
    if args.out:
        with open(args.out, "w") as fh:
            json.dump({"workload": args.workload, "frames": args.frames, "results": results}, fh, indent=1)


if __name__ == "__main__":
    main()
