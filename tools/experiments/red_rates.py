#!/usr/bin/env python
"""Measured rates of the reductions the point passes issue, to random addresses of a frame-sized buffer
(rtr_bench_red_min): RED.MIN.U32 (z-min), RED.MIN.U64 (64-bit keys), 2 x RED.ADD.U64 and RED.ADD.F32x4 (colour sums),
for a 1920x1080 frame (buffers L2-resident) and a 3840x2160 frame (133 MB of colour sums: not L2-resident).

    python tools/experiments/red_rates.py --out gpurun_out/exp.json
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ops", type=int, default=200_000_000)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    pkg = entry.load_package()
    pc = pkg.ProjectCloud.synthetic(seed=1, n_total=100_000, hall=(32, 24, 12), n_boxes=6)
    out = {"ops_per_launch": args.ops, "frames": {}}
    for name, (W, H, f, cx, cy) in {"1920x1080": (1920, 1080, 1400.0, 959.5, 539.5), "3840x2160": (3840, 2160, 2800.0, 1919.5, 1079.5)}.items():
        pc.set_camera(bench.make_calib(pkg, W, H, f, cx, cy), pkg.look_at_w2c((4.0, 3.0, 1.5), (1.0, 0.2, 0.0)))
        r = {}
        for label, mode, key64 in (("red_min_u32", 0, False), ("red_min_u64", 0, True), ("2x_red_add_u64", 2, False), ("red_add_f32x4", 3, False)):
            ms, ops = pc.bench_red_min(mode, args.ops, key64)
            r[label] = {"ms_per_launch": ms, "Gops_per_s": ops / ms / 1e6, "buffer_MB": W * H * (16 if mode >= 2 else (8 if key64 else 4)) / 1e6}
        out["frames"][name] = r
        print(name, json.dumps(r), flush=True)
    pc.close()
    if args.out:
        json.dump(out, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
