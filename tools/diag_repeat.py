#!/usr/bin/env python
"""Diagnostic (one GPU): render the same frames over and over and count how often a digest changes.  The cloud is the
union of 8 unsorted 2 M-point scans (what bench.py's point-sharded digest check renders on rank 0)."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402
import bench  # noqa: E402


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    pkg = entry.load_package()
    n, W, H, f, cx, cy, hall, boxes, seed, n_poses = bench.WORKLOADS["c3"]
    P = W * H
    calib = bench.make_calib(pkg, W, H, f, cx, cy)
    poses = bench.trajectory(pkg, hall, n_poses)
    parts = []
    for r in range(8):
        s = pkg.ProjectCloud.synthetic(seed=seed + r, n_total=2_000_000, hall=hall, n_boxes=boxes, sort=False)
        parts.append(s.download_cloud())
        s.close()
    rec = np.concatenate(parts)
    base = {"blend_variant": 0}
    configs = [("unsorted int sums", dict(sort=False), dict(base)),
               ("  + ring_dynamic=0 (round-robin tiles)", dict(sort=False), dict(base, ring_dynamic=0)),
               ("  + ring_early=0", dict(sort=False), dict(base, ring_early=0)),
               ("  + zmin_variant=1 (early test through L2)", dict(sort=False), dict(base, zmin_variant=1)),
               ("  + zmin_variant=0 (no early test)", dict(sort=False), dict(base, zmin_variant=0)),
               ("  + ring=0 (per-thread list kernels)", dict(sort=False), dict(base, ring=0)),
               ("  + chunk_cull=0 (stream-all kernels)", dict(sort=False), dict(base, chunk_cull=0)),
               ("  + ring=2, chunk_cull=0 (ring, stream-all)", dict(sort=False), dict(base, ring=2, chunk_cull=0)),
               ("  + ring_ctas=1", dict(sort=False), dict(base, ring_ctas=1)),
               ("  + ring_claim_min=1000000 (never claim)", dict(sort=False), dict(base, ring_claim_min=1000000))]
    if os.environ.get("DIAG_ONLY"):
        configs = [c for c in configs if os.environ["DIAG_ONLY"] in c[0]]
    for label, kw, opts in configs:
        pc = pkg.ProjectCloud.from_packed(rec, **kw)
        for k, v in opts.items():
            pc.set_option(k, v)
        first, changed = {}, 0
        for i in range(reps):
            for pi in (0, n_poses // 3, (2 * n_poses) // 3):
                color, depth = np.zeros(P * 3, np.uint8), np.zeros(P, np.float32)
                assert pc.computeFilteredRGBD(calib, poses[pi], color, depth) == 1
                raw = pc.read("accum", np.uint32, P * 4)
                d = hashlib.sha256(color.tobytes() + depth.tobytes()).hexdigest(), hashlib.sha256(raw.tobytes()).hexdigest()
                if pi not in first:
                    first[pi] = (d, color.copy(), depth.copy(), raw.copy())
                elif d != first[pi][0]:
                    changed += 1
                    dc = int((color.reshape(P, 3) != first[pi][1].reshape(P, 3)).any(axis=1).sum())
                    dd = int((depth.view(np.uint32) != first[pi][2].view(np.uint32)).sum())
                    da = int((raw.reshape(P, 4) != first[pi][3].reshape(P, 4)).any(axis=1).sum())
                    print(f"[{label}] rep {i} pose {pi}: colour differs in {dc} px, depth in {dd} px, accum in {da} px", flush=True)
        print(f"[{label}] {reps * 3} frames, {changed} differ from the first render of their pose", flush=True)
        pc.close()


if __name__ == "__main__":
    main()
