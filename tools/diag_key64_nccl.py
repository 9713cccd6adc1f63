#!/usr/bin/env python
"""Diagnostic (torchrun, one rank per GPU): 64-bit-key frames merged by ncclAllReduce(min, uint64) against the frame one
GPU renders from the union of the shards — per buffer, how many elements differ and where."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402
import bench  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = entry.load_package()
    n, W, H, f, cx, cy, hall, boxes, seed, n_poses = bench.WORKLOADS["c3"]
    P = W * H
    calib = bench.make_calib(pkg, W, H, f, cx, cy)
    poses = bench.trajectory(pkg, hall, n_poses)
    n_small = 2_000_000
    small = pkg.ProjectCloud.synthetic(seed=seed + rank, n_total=n_small, hall=hall, n_boxes=boxes, device=local, sort=False)
    host = torch.from_numpy(small.download_cloud()).cuda()
    gathered = [torch.empty_like(host) for _ in range(world)]
    dist.all_gather(gathered, host)
    union = pkg.ProjectCloud.from_packed(torch.cat(gathered).cpu().numpy(), device=local, sort=False) if rank == 0 else None
    small.set_option("index_base", n_small * rank)
    small.set_camera(calib)
    reps = int(os.environ.get("DIAG_REPS", "1"))
    for spec in sys.argv[1:] or ["nccl:1"]:
        merge, key64 = spec.split(":")
        small.set_option("key64", int(key64))
        bench.attach_merge(pkg, torch, dist, small, merge, rank, world)
        for pi in [0, n_poses // 3, (2 * n_poses) // 3] * reps:
            E = poses[pi]
            color, depth = np.zeros(P * 3, np.uint8), np.zeros(P, np.float32)
            assert small.computeFilteredRGBD(calib, E, color, depth) == 1
            t = torch.cat([torch.from_numpy(depth.view(np.int32).copy()).cuda(), torch.from_numpy(color.astype(np.int32)).cuda()])
            if rank == 0:
                union.set_option("key64", int(key64))
                c0, d0 = np.zeros(P * 3, np.uint8), np.zeros(P, np.float32)
                assert union.computeFilteredRGBD(calib, E, c0, d0) == 1
                t0 = torch.cat([torch.from_numpy(d0.view(np.int32).copy()).cuda(), torch.from_numpy(c0.astype(np.int32)).cuda()])
            else:
                t0 = torch.empty_like(t)
            dist.broadcast(t0, 0)
            dd = (t[:P] != t0[:P])
            dc = (t[P:] != t0[P:]).view(P, 3).any(dim=1)
            msg = f"[{merge} rank {rank} pose {pi}] depth differs in {int(dd.sum())} px, colour in {int(dc.sum())} px"
            if int(dc.sum()):
                i = int(torch.nonzero(dc)[0])
                msg += f"; first colour diff px {i} (row {i // W}, col {i % W}): mine {t[P + 3 * i:P + 3 * i + 3].tolist()} union {t0[P + 3 * i:P + 3 * i + 3].tolist()} depth bits {int(t[i])} / {int(t0[i])}"
            if int(dd.sum()) or int(dc.sum()):
                print(msg, flush=True)
        if rank == 0:
            print(f"[{spec}] done", flush=True)
        bench.detach_merge(small, dist, merge)
    small.close()
    if union is not None:
        union.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
