#!/usr/bin/env python
"""Experiment: RED flavours for the blend accumulator (measurement only)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import __graft_entry__ as entry
import bench
pkg = entry.load_package()
pc = pkg.ProjectCloud.synthetic(seed=1, n_total=100000, hall=(32, 24, 12), n_boxes=2)
pc.set_camera(bench.make_calib(pkg, 1920, 1080, 1400.0, 959.5, 539.5), pkg.look_at_w2c((4, 3, 1.5), (1, 0, 0)))
out = {}
for mode, name in ((0, "red_min_u32"), (2, "2x_red_add_u64"), (3, "red_add_v4f32")):
    ms, ops = pc.bench_red_min(mode, 100_000_000, False, iters=5)
    out[name] = {"ms": ms, "Gupdates_per_s": ops / ms / 1e6}
ms, ops = pc.bench_red_min(0, 100_000_000, True, iters=5)
out["red_min_u64"] = {"ms": ms, "Gupdates_per_s": ops / ms / 1e6}
print(json.dumps(out, indent=1))
