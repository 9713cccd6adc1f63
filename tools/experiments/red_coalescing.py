#!/usr/bin/env python
"""Experiment: what a reduction costs when the lanes of one warp instruction share 32-byte sectors or addresses
(rtr_bench_red_min modes 4..19): is a point pass charged per lane or per sector, and what do same-address lanes cost?

    python tools/experiments/red_coalescing.py --out gpurun_out/r02Q_exp_red_coalescing.json
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--ops", type=int, default=200_000_000)
    args = ap.parse_args()
    pkg = entry.load_package()
    pc = pkg.ProjectCloud.synthetic(seed=1, n_total=200_000, hall=(32, 24, 12), n_boxes=6)
    calib = pkg.CameraCalibration()
    calib.loadCalibration(1400.0, 1400.0, 959.5, 539.5, [0.0] * 5, 1920, 1080)
    pc.set_camera(calib, pkg.look_at_w2c((4.0, 3.0, 1.5), (1.0, 0.2, 0.0)))
    res = {}
    for base, name in ((4, "red_min_u32"), (12, "red_add_f32x4")):
        for same in (0, 1):
            for lg in range(4):
                if same and lg == 0:
                    continue
                ms, _ = pc.bench_red_min(base + 4 * same + lg, args.ops, False)
                L = 1 << lg
                per_sector = 8 if base == 4 else 2
                key = f"{name}: {L} lane(s) per sector, " + ("one address" if same else f"{min(L, per_sector)} distinct address(es)")
                res[key] = {"ms_per_launch": ms, "G_lanes_per_s": args.ops / ms / 1e6, "G_sectors_per_s": args.ops / L / ms / 1e6}
                print(key, json.dumps(res[key]), flush=True)
    pc.close()
    if args.out:
        with open(args.out, "w") as fh:
            json.dump({"ops_per_launch": args.ops, "frame": "1920x1080", "results": res}, fh, indent=1)


if __name__ == "__main__":
    main()
