"""torchrun worker for tests/test_gpu_multi.py: one rank per GPU, NCCL.
  point-sharded : every rank uploads its contiguous shard, the library merges z-buffers (ncclMin) and
                  colour sums (ncclSum); every rank's frame must equal the single-GPU frame of the
                  whole cloud (rendered here by rank 0 on its own GPU), bit for bit — blend mode and
                  the 64-bit key mode.
  frame-sharded : ranks render disjoint frame ranges of one trajectory from a replicated cloud; the
                  gathered per-frame digests must equal those of rank 0 rendering every frame."""
import hashlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as entry  # noqa: E402
import scenes  # noqa: E402


def frame(pc, pkg, calib, E, P, filtered=True):
    color, depth = np.zeros(P * 3, np.uint8), np.zeros(P, np.float32)
    fn = pc.computeFilteredRGBD if filtered else pc.computeRGBD
    assert fn(calib, E, color, depth) == 1
    return color, depth.view(np.uint32), pc.read("tensor", np.uint16, P * 5)


# key64's tie-break is the point index: keep the seeded order so that shard-local and whole-cloud indices agree
SORT = False


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = entry.load_package()
    case = scenes.CASES["c3_1920x1080"]
    n, W, H, P = 1_000_003, case.W, case.H, case.W * case.H   # odd size: uneven shards
    calib = pkg.CameraCalibration()
    calib.setIntrinsicsMatrix(case.K)
    calib.setWidth(W)
    calib.setHeight(H)
    # ---- point-sharded
    first, count = pkg.shard_points(n, rank, world)
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(pkg.ProjectCloud.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    shard = pkg.ProjectCloud.synthetic(seed=case.seed, n_total=n, first=first, count=count, hall=case.hall, n_boxes=case.n_boxes, device=local, sort=SORT)
    shard.comm_init(uid.cpu().numpy().tobytes(), rank, world)
    full = pkg.ProjectCloud.synthetic(seed=case.seed, n_total=n, hall=case.hall, n_boxes=case.n_boxes, device=local, sort=SORT) if rank == 0 else None
    ok = True
    for key64 in (0, 1):
        shard.set_option("key64", key64)
        for E in case.poses:
            got = frame(shard, pkg, calib, E, P)
            dig = hashlib.sha256(b"".join(np.ascontiguousarray(a).tobytes() for a in got)).digest()
            t = torch.frombuffer(bytearray(dig), dtype=torch.uint8).cuda()
            if rank == 0:
                full.set_option("key64", key64)
                want = frame(full, pkg, calib, E, P)
                wd = hashlib.sha256(b"".join(np.ascontiguousarray(a).tobytes() for a in want)).digest()
                t = torch.frombuffer(bytearray(wd), dtype=torch.uint8).cuda()
            dist.broadcast(t, 0)
            same = bytes(t.cpu().numpy().tobytes()) == dig
            if not same:
                print(f"[rank {rank}] point-sharded key64={key64}: frame differs from the single-GPU frame", flush=True)
            ok &= same
    shard.close()
    # ---- point-sharded, merged by our own two-shot all-reduce over NVLink peer memory (no library collective)
    peer = pkg.ProjectCloud.synthetic(seed=case.seed, n_total=n, first=first, count=count, hall=case.hall, n_boxes=case.n_boxes, device=local)
    peer.set_camera(calib)
    blob = torch.frombuffer(bytearray(peer.peer_export()), dtype=torch.uint8).cuda()
    blobs = [torch.zeros(512, dtype=torch.uint8, device="cuda") for _ in range(world)]
    dist.all_gather(blobs, blob)
    peer.peer_attach(b"".join(bytes(b.cpu().numpy().tobytes()) for b in blobs), rank, world)
    for rep_i in range(3):   # several frames: the epoch flags must keep working
        for E in case.poses:
            got = frame(peer, pkg, calib, E, P)
            dig = hashlib.sha256(b"".join(np.ascontiguousarray(a).tobytes() for a in got)).digest()
            t = torch.frombuffer(bytearray(dig), dtype=torch.uint8).cuda()
            if rank == 0:
                full.set_option("key64", 0)
                want = frame(full, pkg, calib, E, P)
                t = torch.frombuffer(bytearray(hashlib.sha256(b"".join(np.ascontiguousarray(a).tobytes() for a in want)).digest()), dtype=torch.uint8).cuda()
            dist.broadcast(t, 0)
            same = bytes(t.cpu().numpy().tobytes()) == dig
            if not same:
                print(f"[rank {rank}] peer merge: frame differs from the single-GPU frame", flush=True)
            ok &= same
    if peer.get_option("peer_error") != 0:
        print(f"[rank {rank}] peer merge: a cross-GPU wait timed out", flush=True)
        ok = False
    dist.barrier()

    def reattach(pc):
        """detach / export / all-gather / attach: every attach needs a fresh export on every rank."""
        pc.peer_detach()
        blob = torch.frombuffer(bytearray(pc.peer_export()), dtype=torch.uint8).cuda()
        blobs = [torch.zeros(512, dtype=torch.uint8, device="cuda") for _ in range(world)]
        dist.all_gather(blobs, blob)
        pc.peer_attach(b"".join(bytes(b.cpu().numpy().tobytes()) for b in blobs), rank, world)

    # ---- a lost peer: only rank 0 renders; its merge kernel must give up after peer_timeout_ms and the CALL must fail
    peer.set_option("peer_timeout_ms", 400)
    if rank == 0:
        try:
            frame(peer, pkg, calib, case.poses[0], P)
            print("[rank 0] lost peer: the render call returned success", flush=True)
            ok = False
        except pkg.RtrError as e:
            if e.code != pkg.RTR_ERR_COMM:
                print(f"[rank 0] lost peer: wrong error {e}", flush=True)
                ok = False
    dist.barrier()
    peer.set_option("peer_timeout_ms", 10000)
    reattach(peer)                                   # fresh flags and epochs on every rank: the next frames merge again
    for E in case.poses:
        got = frame(peer, pkg, calib, E, P)
        dig = hashlib.sha256(b"".join(np.ascontiguousarray(a).tobytes() for a in got)).digest()
        t = torch.frombuffer(bytearray(dig), dtype=torch.uint8).cuda()
        if rank == 0:
            want = frame(full, pkg, calib, E, P)
            t = torch.frombuffer(bytearray(hashlib.sha256(b"".join(np.ascontiguousarray(a).tobytes() for a in want)).digest()), dtype=torch.uint8).cuda()
        dist.broadcast(t, 0)
        same = bytes(t.cpu().numpy().tobytes()) == dig
        if not same:
            print(f"[rank {rank}] peer merge after re-attach: frame differs from the single-GPU frame", flush=True)
        ok &= same
    dist.barrier()
    peer.peer_detach()
    peer.close()
    # ---- north_star's merge through the peer kernels: min over the ranks of the 64-bit (depth bits << 32 | point index) keys
    pk = pkg.ProjectCloud.synthetic(seed=case.seed, n_total=n, first=first, count=count, hall=case.hall, n_boxes=case.n_boxes, device=local, sort=SORT)
    pk.set_option("key64", 1)                        # before the export: the key buffers are mapped by the peers too
    pk.set_camera(calib)
    reattach(pk)
    for E in case.poses:
        got = frame(pk, pkg, calib, E, P)
        dig = hashlib.sha256(b"".join(np.ascontiguousarray(a).tobytes() for a in got)).digest()
        t = torch.frombuffer(bytearray(dig), dtype=torch.uint8).cuda()
        if rank == 0:
            full.set_option("key64", 1)
            want = frame(full, pkg, calib, E, P)
            t = torch.frombuffer(bytearray(hashlib.sha256(b"".join(np.ascontiguousarray(a).tobytes() for a in want)).digest()), dtype=torch.uint8).cuda()
        dist.broadcast(t, 0)
        same = bytes(t.cpu().numpy().tobytes()) == dig
        if not same:
            print(f"[rank {rank}] peer merge of 64-bit keys: frame differs from the single-GPU frame", flush=True)
        ok &= same
    if rank == 0:
        full.set_option("key64", 0)
    dist.barrier()
    pk.peer_detach()
    pk.close()
    # ---- frame-sharded
    poses = pkg.trajectory_w2c(11, center=(6.0, 5.0, 1.5), radius=2.0)
    rep = pkg.ProjectCloud.synthetic(seed=case.seed, n_total=n, hall=case.hall, n_boxes=case.n_boxes, device=local, sort=SORT)
    mine = pkg.shard_frames(len(poses), rank, world)
    digs = torch.zeros((len(poses), 32), dtype=torch.uint8, device="cuda")
    for f in mine:
        got = frame(rep, pkg, calib, poses[f], P)
        digs[f] = torch.frombuffer(bytearray(hashlib.sha256(b"".join(np.ascontiguousarray(a).tobytes() for a in got)).digest()), dtype=torch.uint8).cuda()
    dist.all_reduce(digs.view(torch.uint8).to(torch.int32).contiguous(), op=dist.ReduceOp.SUM) if False else None
    gathered = digs.to(torch.int32)
    dist.all_reduce(gathered, op=dist.ReduceOp.SUM)   # disjoint rows: the sum is the union
    if rank == 0:
        for f in range(len(poses)):
            want = frame(full, pkg, calib, poses[f], P) if False else frame(rep, pkg, calib, poses[f], P)
            wd = hashlib.sha256(b"".join(np.ascontiguousarray(a).tobytes() for a in want)).digest()
            if bytes(gathered[f].to(torch.uint8).cpu().numpy().tobytes()) != wd:
                print(f"frame-sharded: frame {f} differs", flush=True)
                ok = False
    rep.close()
    if full is not None:
        full.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_OK" if int(flag.item()) == 1 else "MULTI_GPU_FAIL", flush=True)
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
