"""world_size-2 (gloo, CPU) test of the two multi-GPU modes' host logic (SURVEY.md §8 e):
  frame-sharded : the ranks' frame ranges tile the trajectory; no collective on the data path;
  point-sharded : all-reduce(min) of the per-rank z-buffers, local blend against the GLOBAL min,
                  all-reduce(sum) of the colour sums  ==  one rank holding every point, bit for bit.
The per-rank compute is stood in for by the CPU oracle (this is a test, the product's ranks run the
CUDA kernels and NCCL; tests/test_gpu_multi.py covers that on GPUs)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import __graft_entry__ as entry
    import oracle
    import scenes
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = entry.load_package()
    cpu = oracle.cpu()
    case = scenes.CASES["small_176x104"]
    W, H, P = case.W, case.H, case.W * case.H
    # ---- frame sharding: gather the frame ids every rank would render
    mine = torch.tensor(list(pkg.shard_frames(37, rank, world)), dtype=torch.int64)
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([len(mine)], dtype=torch.int64))
    bufs = [torch.zeros(int(s.item()), dtype=torch.int64) for s in sizes]
    pad = [torch.zeros(max(int(s.item()) for s in sizes), dtype=torch.int64) for _ in range(world)]
    me = torch.zeros(len(pad[0]), dtype=torch.int64)
    me[:len(mine)] = mine
    dist.all_gather(pad, me)
    frames = torch.cat([pad[r][:len(bufs[r])] for r in range(world)]).tolist()
    # ---- point sharding
    first, count = pkg.shard_points(case.n, rank, world)
    rec = cpu.synth_packed(case.seed, case.n, first, count, case.hall, case.n_boxes)   # this rank's shard only
    m = cpu.cam_proj(case.K, case.poses[1])
    pix, zb = cpu.project(rec, m, W, H)
    bgra = scenes.bgra_of(rec)
    buf = cpu.new_buffers(W, H)
    cpu.lib.rtro_clear(buf["zbuf"].ctypes.data, buf["accum"].ctypes.data, W, H)
    cpu.lib.rtro_zmin(pix.ctypes.data, zb.ctypes.data, len(pix), buf["zbuf"].ctypes.data)
    z = torch.from_numpy(buf["zbuf"].astype(np.int64))
    dist.all_reduce(z, op=dist.ReduceOp.MIN)                       # ncclMin on the depth bits
    buf["zbuf"][:] = z.numpy().astype(np.uint32)
    cpu.lib.rtro_accumulate(pix.ctypes.data, zb.ctypes.data, bgra.ctypes.data, len(pix), buf["zbuf"].ctypes.data, buf["accum"].ctypes.data)
    a = torch.from_numpy(buf["accum"].astype(np.int64))
    dist.all_reduce(a, op=dist.ReduceOp.SUM)                       # ncclSum on {b, g, r, count}
    buf["accum"][:] = a.numpy().astype(np.uint32)
    cpu.lib.rtro_resolve(buf["accum"].ctypes.data, buf["image"].ctypes.data, W, H)
    mm = buf["minmax"]
    cpu.lib.rtro_depth_filter(buf["zbuf"].ctypes.data, buf["image"].ctypes.data, buf["tensor"].ctypes.data, W, H,
                              mm[0:1].ctypes.data, mm[1:2].ctypes.data, None, None, None)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), frames=np.array(frames), zbuf=buf["zbuf"], accum=buf["accum"],
             image=buf["image"], tensor=buf["tensor"])
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_equal_one(tmp_path, cpu_oracle):
    import scenes
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    case = scenes.CASES["small_176x104"]
    rec = cpu_oracle.synth_packed(case.seed, case.n, 0, case.n, case.hall, case.n_boxes)
    m = cpu_oracle.cam_proj(case.K, case.poses[1])
    pix, zb = cpu_oracle.project(rec, m, case.W, case.H)
    single = cpu_oracle.render(pix, zb, scenes.bgra_of(rec), case.W, case.H, filtered=True)
    for r in range(world):
        g = np.load(tmp_path / f"rank{r}.npz")
        assert g["frames"].tolist() == list(range(37))
        for k in ("zbuf", "accum", "image", "tensor"):
            assert np.array_equal(g[k], single[k]), f"rank {r}: {k} differs from the single-rank result"
