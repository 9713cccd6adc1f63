#!/usr/bin/env python
"""bench.py — throughput of the point-projection hot path (BASELINE.json metric: Gpoints/s and
frames/s, 100 M points, 1920x1080, 1/2/4/8 B200, % of HBM roofline).

A "step" is ONE FRAME: clear + project/z-min + project/blend + resolve + depth prefilter + fp16
tensor write for one pose over the whole resident cloud (== the reference's computeFilteredRGBD /
the projection+filter half of computeFull).

  python bench.py [--gpus N] [--steps K] [--warmup W]              this repo's CUDA path
  python bench.py --impl reference ...                             the UNMODIFIED reference
        (its own CUDA sources compiled where they lie -> oracle/_ref, driven through its own
        ProjectCloud::computeFilteredRGBD with host output buffers).  The reference has no CPU
        implementation; if oracle/_ref is not on the box the OpenMP CPU port is timed instead.
  python bench.py --impl cpu ...                                   the OpenMP CPU port alone

Multi-GPU (torchrun, one rank per GPU):
  --mode frames (default)  cloud replicated, trajectory frames sharded across ranks, no collective;
                           weak scaling: every rank renders K frames.
  --mode points            points sharded (--points per GPU), z-buffers / colour sums merged by NCCL
                           min / sum all-reduces inside the library; every rank renders the same K poses.

Timing: CUDA events on the renderer's own stream, barrier + synchronize on both sides, max over
ranks.  The cloud (1.6 GB at 100 M points) is >> the 126 MB L2, so every frame streams from HBM.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (points, W, H, f, cx, cy, hall, boxes, seed, trajectory poses)
    "c1": (1_000_000, 640, 480, 525.0, 319.5, 239.5, (32, 24, 12), 6, 1234, 1),
    "c2": (20_000_000, 1280, 720, 900.0, 639.5, 359.5, (32, 24, 12), 6, 1234, 300),
    "c3": (100_000_000, 1920, 1080, 1400.0, 959.5, 539.5, (48, 40, 12), 12, 5678, 1000),
    "c5_4k": (100_000_000, 3840, 2160, 2800.0, 1919.5, 1079.5, (48, 40, 12), 12, 5678, 1000),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "cpu"])
    ap.add_argument("--workload", default="c3", choices=list(WORKLOADS))
    ap.add_argument("--points", type=int, default=None, help="override the workload's point count (per GPU in --mode points)")
    ap.add_argument("--mode", default="frames", choices=["frames", "points"])
    ap.add_argument("--stage", default="filtered", choices=["filtered", "rgbd"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-points", type=int, default=100_000_000, help="points of the CPU port's sample (default: the full c3 cloud)")
    ap.add_argument("--no-unet", action="store_true", help="skip the config-5 leg (projection + prefilter feeding the reference U-Net)")
    ap.add_argument("--no-points-mode", action="store_true", help="N > 1: skip the point-sharded (config 4) record")
    ap.add_argument("--points-mode-points", type=int, default=125_000_000, help="points per GPU of the point-sharded record")
    ap.add_argument("--distort", action="store_true", help="apply config 2's k1,k2,p1,p2,k3 = (-0.05, 0.01, 0.0005, -0.0005, 0) (new feature, no reference parity)")
    ap.add_argument("--nccl", action="store_true", help="--mode points: merge with ncclAllReduce instead of the peer-memory kernels")
    ap.add_argument("--opt", action="append", default=[], help="renderer option key=value (repeatable)")
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock + throttle reasons sampled through NVML while the timed region runs."""

    def __init__(self, index: int, interval: float = 0.01):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.interval = interval
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(self.interval)

    def __enter__(self):
        if self.nv:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t:
            self._t.join()

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def make_calib(pkg, W, H, f, cx, cy):
    c = pkg.CameraCalibration()
    c.loadCalibration(f, f, cx, cy, [0.0] * 5, W, H)
    return c


def trajectory(pkg, hall, n_poses):
    center = (hall[0] * 0.125, hall[1] * 0.125, 1.5)  # hall dims are in quarter metres
    return pkg.trajectory_w2c(max(n_poses, 1), center=center, radius=2.0)


# ------------------------------------------------------------------------------------------
def cpu_port_baseline(args, wl, sample_points, min_seconds=12.0, max_frames=400):
    """OpenMP CPU port of the same frame (oracle/rtr_oracle.c) on a bounded sample: the same scene
    and camera at `sample_points` points.  Returns the cpu_baseline object."""
    import oracle
    cpu = oracle.cpu()
    n, W, H, f, cx, cy, hall, boxes, seed, n_poses = wl
    n = min(sample_points, n)
    import importlib
    pkg = importlib.import_module("__graft_entry__").load_package()  # host helpers only (poses)
    rec = cpu.synth_packed(seed, n, 0, n, hall, boxes)
    bgra = np.ascontiguousarray(rec[:, 3]).view(np.uint32)
    K = np.array([[f, 0, cx], [0, f, cy], [0, 0, 1]], np.float64)
    poses = trajectory(pkg, hall, n_poses)
    buf = cpu.new_buffers(W, H)
    frames, t0 = 0, time.perf_counter()
    while frames < max_frames and (frames < 1 or time.perf_counter() - t0 < min_seconds):
        m = cpu.cam_proj(K, poses[(frames * 37) % len(poses)])
        cpu.lib.rtro_clear(buf["zbuf"].ctypes.data, buf["accum"].ctypes.data, W, H)
        cpu.point_passes_packed(rec, m, W, H, buf["zbuf"], buf["accum"])
        cpu.lib.rtro_resolve(buf["accum"].ctypes.data, buf["image"].ctypes.data, W, H)
        if args.stage == "filtered":
            mm = buf["minmax"]
            cpu.lib.rtro_depth_filter(buf["zbuf"].ctypes.data, buf["image"].ctypes.data, buf["tensor"].ctypes.data, W, H,
                                      mm[0:1].ctypes.data, mm[1:2].ctypes.data, None, None, None)
        frames += 1
    dt = time.perf_counter() - t0
    return {"value": n * frames / dt / 1e9, "unit": "Gpoints/s", "cores": cpu.threads, "kind": "port",
            "frames_per_s": frames / dt,
            "sample": f"{frames} frames of the same scene/camera at {n} points ({W}x{H}, stage {args.stage}), OpenMP, {dt:.1f} s"}


def run_reference(args, wl):
    """The reference arm.  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import importlib
    import oracle
    pkg = importlib.import_module("__graft_entry__").load_package()
    n, W, H, f, cx, cy, hall, boxes, seed, n_poses = wl
    if args.points:
        n = args.points
    base = {"impl": args.impl, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True,
            "metric": "Gpoints/s (points x frames / s, whole frame: projection + z-buffer + blend + resolve + prefilter)",
            "unit": "Gpoints/s", "scaling": "weak", "vs_baseline": None, "dtype": "f32+u32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {n} points, {W}x{H}, stage {args.stage}, {n_poses}-pose trajectory"}}
    use_cuda_ref = args.impl == "reference" and os.path.exists(oracle.REF_LIB)
    if use_cuda_ref:
        try:
            import ctypes
            ctypes.CDLL("libcuda.so.1")
        except OSError:
            use_cuda_ref = False
    if not use_cuda_ref:
        cb = cpu_port_baseline(args, (n,) + wl[1:], args.cpu_sample_points)
        line = dict(base, value=cb["value"], ms_per_step=1e3 / cb["frames_per_s"] if cb["frames_per_s"] else None,
                    frames_per_s=cb["frames_per_s"], cpu_baseline=cb, gpu_launches=0,
                    e2e={"value": cb["value"], "unit": "Gpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
        line["config"]["note"] = "OpenMP CPU port of the path (the reference itself is CUDA-only; oracle/_ref unavailable here)" \
            if args.impl == "reference" else "OpenMP CPU port of the path"
        print(json.dumps(line), flush=True)
        return
    cpu = oracle.cpu()
    rec = cpu.synth_packed(seed, n, 0, n, hall, boxes)
    c = np.ascontiguousarray(rec[:, 3]).view(np.uint32)
    bgr = np.stack([c & 0xFF, (c >> 8) & 0xFF, (c >> 16) & 0xFF], axis=1).astype(np.uint8)
    xyz = np.ascontiguousarray(rec[:, :3])
    del rec, c
    stock = os.path.exists(oracle.REF_LIB_STOCK)   # timing build: no zero-initialising cudaMalloc prelude
    ref = oracle.RefOracle(xyz, bgr, stock=stock)
    del xyz, bgr
    K = np.array([[f, 0, cx], [0, f, cy], [0, 0, 1]], np.float64)
    poses = trajectory(pkg, hall, n_poses)
    fn = ref.computeFilteredRGBD if args.stage == "filtered" else ref.computeRGBD
    # pre-allocated outputs, like the cv::Mat pair a reference user passes
    import ctypes as C
    color, depth = np.zeros(W * H * 3, np.uint8), np.zeros(W * H, np.float32)
    cfn = ref.lib.ref_compute_filtered if args.stage == "filtered" else ref.lib.ref_compute_rgbd
    Kc = np.ascontiguousarray(K.reshape(9))

    def step(i):
        E = np.ascontiguousarray(poses[i % len(poses)].reshape(16))
        rc = cfn(ref.h, W, H, Kc.ctypes.data_as(C.c_void_p), E.ctypes.data_as(C.c_void_p),
                 color.ctypes.data_as(C.c_void_p), depth.ctypes.data_as(C.c_void_p))
        assert rc == 1
    del fn
    for i in range(args.warmup):
        step(i)
    # (sampled four times a second only: the reference allocates and frees device memory every frame, and NVML queries
    # take driver locks those calls need — the sampler must not be what slows the reference arm down)
    with ClockSampler(0, interval=float(os.environ.get("RTR_BENCH_NVML_INTERVAL", "0.25"))) as clk:
        t0 = time.perf_counter()
        step_s = []
        for i in range(args.steps):
            ts = time.perf_counter()
            step(args.warmup + i)
            step_s.append(time.perf_counter() - ts)
        dt = time.perf_counter() - t0   # every reference call ends in blocking cudaMemcpy D2H
    kms = ref.time_point_kernels(W, H, iters=5)
    value = n * args.steps / dt / 1e9
    line = dict(base, value=value, ms_per_step=dt / args.steps * 1e3, frames_per_s=args.steps / dt,
                gpu_launches=0, clocks=clk.summary(),
                cpu_baseline={"value": value, "unit": "Gpoints/s", "cores": 1, "kind": "reference",
                              "sample": f"{args.steps} frames, full workload; the reference's path is CUDA (render.cu/project_cloud.cu "
                                        "compiled unmodified for sm_100, one host thread driving one B200), not a CPU implementation"},
                e2e={"value": value, "unit": "Gpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                # the reference's per-frame time is erratic (it allocates and frees device memory every frame: 4 ... 230 ms per
                # frame between runs, profiles/r02R_reference_nvml_ab.txt): the spread of the timed steps, for the reader
                step_ms={"min": min(step_s) * 1e3, "median": sorted(step_s)[len(step_s) // 2] * 1e3, "max": max(step_s) * 1e3},
                reference_kernels_ms={"clear": float(kms[0]), "minDepthPass": float(kms[1]), "accumulatePass": float(kms[2]),
                                      "resolvePass": float(kms[3]), "block_size": ref.block_size, "build": "stock" if stock else "zero-init parity build",
                                      "minDepthPass_GBps_at_16B_per_point": 16.0 * n / (float(kms[1]) * 1e-3) / 1e9})
    line["config"]["note"] = "unmodified reference CUDA code through ProjectCloud::compute*RGBD with host outputs (sync + pageable D2H per frame, as shipped)"
    ref.close()
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
def pose_schedule(n_steps, n_poses, world, rank, segments=4):
    """Which trajectory pose each of this rank's frames renders.  The trajectory is dealt to the ranks in contiguous runs
    of frames (SURVEY.md §8 e allows "f mod G or contiguous chunks"; contiguous keeps the frame-to-frame motion the real
    trajectory's, which is what the fused sequences' shared chunk stream depends on).  A run shorter than the loop is cut
    into `segments` arcs per rank, and the segments x world arcs are spread evenly over the loop, so any step count
    samples the easy and the hard parts of the trajectory alike; at 1000 steps on one GPU the arcs tile the whole loop."""
    seg_len = -(-n_steps // segments)
    out = []
    for i in range(n_steps):
        seg, off = divmod(i, seg_len)
        start = ((seg * world + rank) * n_poses) // (segments * world)
        out.append((start + off) % n_poses)
    return out


def d2h_ceiling(torch, dist, world, P, iters=60):
    """What the box can move device->host: every rank copies a frame's worth of output (depth 4 B/px + BGR 3 B/px) from
    device memory into pinned host memory `iters` times, all ranks at once, nothing else running.  Returns aggregate GB/s."""
    src_d, src_c = torch.empty(P, dtype=torch.float32, device="cuda"), torch.empty(P * 3, dtype=torch.uint8, device="cuda")
    n_buf = 4
    dst_d = torch.empty((n_buf, P), dtype=torch.float32, pin_memory=True)
    dst_c = torch.empty((n_buf, P * 3), dtype=torch.uint8, pin_memory=True)
    st = torch.cuda.Stream()

    def loop(n):
        with torch.cuda.stream(st):
            for i in range(n):
                dst_d[i % n_buf].copy_(src_d, non_blocking=True)
                dst_c[i % n_buf].copy_(src_c, non_blocking=True)
        st.synchronize()
    loop(5)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    loop(iters)
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    return world * iters * P * 7 / dt / 1e9


def attach_merge(pkg, torch, dist, pc, kind, rank, world):
    """kind 'peer': our two-shot all-reduce kernels over NVLink peer memory; 'nccl': ncclAllReduce inside the library."""
    if kind == "nccl":
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(pkg.ProjectCloud.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        pc.comm_init(bytes(uid.cpu().numpy().tobytes()), rank, world)
    else:
        blob = torch.frombuffer(bytearray(pc.peer_export()), dtype=torch.uint8).cuda()
        blobs = [torch.zeros(512, dtype=torch.uint8, device="cuda") for _ in range(world)]
        dist.all_gather(blobs, blob)
        pc.peer_attach(b"".join(bytes(b.cpu().numpy().tobytes()) for b in blobs), rank, world)


def detach_merge(pc, dist, kind):
    dist.barrier()
    if kind == "nccl":
        pc._check(pc._lib.rtr_comm_destroy(pc._h))
    else:
        pc.peer_detach()


def points_mode_record(args, pkg, torch, dist, wl, rank, world, local, n_per_gpu, steps):
    """BASELINE config 4 for the driver's record: points sharded over the ranks (n_per_gpu each, 1 B at 8 x 125 M), every
    rank renders the same poses, per-GPU z-buffers / colour sums merged (a) by our peer-memory kernels, (b) by
    ncclAllReduce, and in north_star's 64-bit-key mode — ncclMin on uint64 keys and the same min in the peer kernel.
    Each variant is first checked on a down-sampled cloud: every rank's merged frame must equal, bit for bit, the frame
    one GPU renders from the union of the shards."""
    import hashlib
    n, W, H, f, cx, cy, hall, boxes, seed, n_poses = wl
    P = W * H
    calib = make_calib(pkg, W, H, f, cx, cy)
    poses = trajectory(pkg, hall, n_poses)
    variants = [("peer", 0), ("nccl", 0), ("nccl", 1), ("peer", 1)]   # (merge, key64)
    out = {"points_per_gpu": n_per_gpu, "points_total": n_per_gpu * world, "resolution": f"{W}x{H}", "steps": steps, "variants": {}}

    def digest_of(pc, E):
        """sha256 of the frame plus one sha256 per buffer (colour, depth, tensor): 128 bytes."""
        color, depth = np.zeros(P * 3, np.uint8), np.zeros(P, np.float32)
        assert pc.computeFilteredRGBD(calib, E, color, depth) == 1
        tensor = pc.read("tensor", np.uint16, P * 5)
        parts = [color.tobytes(), depth.tobytes(), tensor.tobytes()]
        return hashlib.sha256(b"".join(parts)).digest() + b"".join(hashlib.sha256(p).digest() for p in parts)

    # ---- parity on a down-sampled cloud (2 M points per rank), all variants
    n_small = 2_000_000
    check_poses = [poses[0], poses[n_poses // 3], poses[(2 * n_poses) // 3]]
    small = pkg.ProjectCloud.synthetic(seed=seed + rank, n_total=n_small, hall=hall, n_boxes=boxes, device=local, sort=False)
    host = torch.empty((n_small, 4), dtype=torch.float32, device="cuda")
    host.copy_(torch.from_numpy(small.download_cloud()))
    gathered = [torch.empty_like(host) for _ in range(world)]
    dist.all_gather(gathered, host)
    union = pkg.ProjectCloud.from_packed(torch.cat(gathered).cpu().numpy(), device=local, sort=False) if rank == 0 else None
    del gathered, host
    small.set_option("index_base", n_small * rank)
    small.set_camera(calib)
    for merge, key64 in variants:
        small.set_option("key64", key64)
        attach_merge(pkg, torch, dist, small, merge, rank, world)
        equal = True
        differing = torch.zeros(3, dtype=torch.int32, device="cuda")   # poses on which colour / depth / tensor differ, summed over the ranks
        for E in check_poses:
            mine = digest_of(small, E)
            t = torch.frombuffer(bytearray(mine), dtype=torch.uint8).cuda()
            if rank == 0:
                union.set_option("key64", key64)
                t = torch.frombuffer(bytearray(digest_of(union, E)), dtype=torch.uint8).cuda()
            dist.broadcast(t, 0)
            want = bytes(t.cpu().numpy().tobytes())
            equal &= want[:32] == mine[:32]
            for b in range(3):
                differing[b] += int(want[32 * (b + 1):32 * (b + 2)] != mine[32 * (b + 1):32 * (b + 2)])
        flag = torch.tensor([1 if equal else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        dist.all_reduce(differing, op=dist.ReduceOp.SUM)
        detach_merge(small, dist, merge)
        out["variants"][f"{merge}{'_key64' if key64 else ''}"] = {"digest_equal": bool(int(flag.item()) == 1),
                                                                "differing_rank_poses": dict(zip(["colour", "depth", "tensor"], differing.tolist())),
                                                                "digest_check": f"{world} x {n_small} points, {len(check_poses)} poses: every rank's merged frame (colour, depth, tensor) "
                                                                                "== the frame one GPU renders from the union of the shards"}
    small.close()
    if union is not None:
        union.close()

    # ---- timing at full size: every rank holds an independent n_per_gpu-point scan of the whole hall (seed + rank), so the
    # shards cover the scene uniformly and the ranks' per-frame work is balanced
    pc = pkg.ProjectCloud.synthetic(seed=seed + rank, n_total=n_per_gpu, hall=hall, n_boxes=boxes, device=local)
    pc.set_camera(calib)
    pc.set_option("index_base", 0 if n_per_gpu * world > (1 << 32) else n_per_gpu * rank)
    idx = pose_schedule(steps + 3, n_poses, 1, 0)
    my = np.ascontiguousarray(np.stack([poses[i] for i in idx]).reshape(-1, 16))
    stage = pkg.STAGE_FILTERED

    def loop(timing):
        pc.set_option("timing", timing)
        for i in range(3):
            pc._check(pc._lib.rtr_set_pose_w2c(pc._h, my[i].ctypes.data_as(pkg._dp)))
            pc.render_device(stage)
        pc.sync()
        dist.barrier()
        torch.cuda.synchronize()
        if timing:
            pc.stage_ms_sum(reset=True)
        stream = torch.cuda.ExternalStream(pc.device_buffers().stream, device=torch.device("cuda", local))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(steps):
            pc._check(pc._lib.rtr_set_pose_w2c(pc._h, my[3 + i].ctypes.data_as(pkg._dp)))
            pc.render_device(stage)
        pc.device_buffers()
        e1.record(stream)
        pc.sync()
        dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        stages = None
        if timing:
            sums, nfr = pc.stage_ms_sum(reset=True)
            stages = (sums / max(nfr, 1)).tolist()
        pc.set_option("timing", 0)
        return float(t.item()) / steps, stages

    # the local passes alone (no merge attached; two passes per frame, integer sums as in the merged frames)
    pc.set_option("fuse", 0)
    pc.set_option("pipeline", 0)
    pc.set_option("blend_variant", 0)
    _, local_stages = loop(2)
    for merge, key64 in variants:
        pc.set_option("key64", key64)
        if key64:
            _, base_stages = loop(2)
        else:
            base_stages = local_stages
        attach_merge(pkg, torch, dist, pc, merge, rank, world)
        ms, _ = loop(0)
        _, st = loop(2)
        detach_merge(pc, dist, merge)
        rec = out["variants"][f"{merge}{'_key64' if key64 else ''}"]
        merge_ms = (st[1] - base_stages[1]) + (st[2] - base_stages[2]) if not key64 else (st[1] - base_stages[1]) + (st[3] - base_stages[3])
        rec.update({"ms_per_frame": ms, "frames_per_s": 1e3 / ms, "gpoints_per_s": n_per_gpu * world / (ms * 1e-3) / 1e9,
                    "merge_ms": merge_ms,
                    "merge": ("two-shot all-reduce kernels over NVLink peer memory (csrc/rtr_peer.cu)" if merge == "peer" else "ncclAllReduce") +
                             (": min of uint64 (depth bits << 32 | point index) keys + sum of the image bytes" if key64 else ": min of u32 depth bits + sum of u32 colour sums"),
                    "stage_ms_with_merge": dict(zip(["clear_classify", "zmin+merge", "blend+merge", "resolve", "up_pass", "frame"], st)),
                    "stage_ms_local_only": dict(zip(["clear_classify", "zmin", "blend", "resolve", "up_pass", "frame"], base_stages))})
    pc.close()
    return out


def unet_leg(pkg, torch, wl, local, frames=30):
    """BASELINE config 5 at 1080p for the driver's record: projection + prefilter feeding the REFERENCE's U-Net
    (TorchScript export of the reference's own model.py with seeded random weights, oracle/_ref/unet_<W>x<H>.pt — the real
    weights are a Git-LFS pointer), output converted on the GPU, image + depth to pinned host memory, per frame."""
    n, W, H, f, cx, cy, hall, boxes, seed, n_poses = wl
    model_file = os.path.join(ROOT, "oracle", "_ref", f"unet_{W}x{H}.pt")
    if not os.path.exists(model_file):
        return {"unavailable": f"{os.path.relpath(model_file, ROOT)} is not on this box (tools/export_unet.py writes it where /root/reference exists)"}
    P = W * H
    try:
        model = torch.jit.load(model_file).cuda().eval()
    except Exception as e:   # measurement support: never fail the bench line for it
        return {"unavailable": f"torch.jit.load failed: {e}"}

    class DevPtr:
        def __init__(self, ptr, shape, typestr):
            self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 2}

    calib = make_calib(pkg, W, H, f, cx, cy)
    poses = trajectory(pkg, hall, n_poses)
    idx = pose_schedule(frames + 3, n_poses, 1, 0)
    pc = pkg.ProjectCloud.synthetic(seed=seed, n_total=n, hall=hall, n_boxes=boxes, device=local)
    color = torch.empty(P * 3, dtype=torch.uint8, pin_memory=True)
    depth = torch.empty(P, dtype=torch.float32, pin_memory=True)
    pc.set_camera(calib, poses[0])
    pc.render_device(pkg.STAGE_FILTERED)
    stream = torch.cuda.ExternalStream(pc.device_buffers().stream, device=torch.device("cuda", local))

    def one(E):
        pc.set_camera(calib, E)
        pc.render_device(pkg.STAGE_FILTERED)
        bufs = pc.device_buffers()                       # completes the frame; `stream` waits for it
        tin = torch.as_tensor(DevPtr(bufs.tensor, (1, 5, H, W), "<f2"), device="cuda")   # what torch::from_blob does, project_cloud.cu:471
        with torch.no_grad(), torch.cuda.stream(stream):
            y = model(tin)[0].contiguous()
        pc._check(pc._lib.rtr_postprocess_unet_output(pc._h, y.data_ptr(), W, H, color.data_ptr(), None))
        pc._check(pc._lib.rtr_read_buffer(pc._h, 0, depth.data_ptr(), P * 4))

    for i in idx[:3]:
        one(poses[i])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in idx[3:]:
        one(poses[i])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    # split: the projection + prefilter alone, and the network alone
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    bufs = pc.device_buffers()
    tin = torch.as_tensor(DevPtr(bufs.tensor, (1, 5, H, W), "<f2"), device="cuda")
    with torch.no_grad(), torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(5):
            model(tin)
        e1.record(stream)
    torch.cuda.synchronize()
    unet_ms = e0.elapsed_time(e1) / 5
    pc.close()
    return {"frames_per_s": frames / dt, "ms_per_frame": dt / frames * 1e3, "unet_ms": unet_ms, "frames": frames,
            "model": os.path.basename(model_file), "resolution": f"{W}x{H}", "points": n,
            "d2h_bytes_per_frame": P * 7,
            "note": "per frame: rtr_render_device -> tensor wrapped zero-copy as fp16 {1,5,H,W} -> TorchScript U-Net of the reference on the renderer's "
                    "stream -> rtr_postprocess_unet_output (fp16 CHW -> uint8 HWC on the GPU) -> image + depth in pinned host memory; the network "
                    "dominates the frame"}


def run_b200(args, wl):
    import torch
    import torch.distributed as dist
    import importlib
    pkg = importlib.import_module("__graft_entry__").load_package()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU fallback (use --impl cpu for the OpenMP port)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, W, H, f, cx, cy, hall, boxes, seed, n_poses = wl
    if args.points:
        n = args.points
    P = W * H
    stage = pkg.STAGE_FILTERED if args.stage == "filtered" else pkg.STAGE_RGBD
    K_steps, Wm = args.steps, max(args.warmup, 3)

    # ---- load cloud.  The cloud is synthesised on the device (no dataset, no network), copied to
    # pinned HOST memory, and then uploaded through the public C-ABI call, so the renderer's cloud
    # really arrives from the host the way the reference's constructor receives it.
    if args.mode == "points":
        # every rank holds an independent n-point scan of the WHOLE hall (seed + rank): the union is the N*n-point cloud
        # and every shard covers the scene uniformly, so the ranks' per-frame work is balanced (a contiguous index range
        # of one cloud would give one rank the floor and another the ceiling)
        n_total, first, count, seed = n, 0, n, seed + rank
    else:
        n_total, first, count = n, 0, n
    gen = pkg.ProjectCloud.synthetic(seed=seed, n_total=n_total, first=first, count=count, hall=hall, n_boxes=boxes, device=local)
    host_cloud = torch.empty((count, 4), dtype=torch.float32, pin_memory=True)
    gen._check(gen._lib.rtr_download_cloud_packed16(gen._h, 0, count, host_cloud.data_ptr()))
    gen.close()
    pc = pkg.ProjectCloud(device=local)
    t0 = time.perf_counter()
    pc._check(pc._lib.rtr_upload_cloud_packed16(pc._h, host_cloud.data_ptr(), count))
    upload_s = time.perf_counter() - t0
    del host_cloud
    for kv in args.opt:
        k, v = kv.split("=")
        pc.set_option(k, int(v))
    calib = make_calib(pkg, W, H, f, cx, cy)
    if args.distort:
        calib.setDistortionParameters([-0.05, 0.01, 0.0005, -0.0005, 0.0])
        pc.apply_distortion = True
    pc.set_camera(calib)
    if args.mode == "points":
        pc.set_option("index_base", 0 if n * world > (1 << 32) else n * rank)
        if world > 1:
            attach_merge(pkg, torch, dist, pc, "nccl" if args.nccl else "peer", rank, world)
    poses = trajectory(pkg, hall, n_poses)
    idx = pose_schedule(K_steps + Wm, n_poses, world if args.mode == "frames" else 1, rank if args.mode == "frames" else 0)   # points mode: every rank renders the same poses
    my = np.ascontiguousarray(np.stack([poses[i] for i in idx]).reshape(-1, 16))

    def set_pose(i):
        pc._check(pc._lib.rtr_set_pose_w2c(pc._h, my[i].ctypes.data_as(pkg._dp)))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        pc.sync()

    set_pose(0)
    pc.render_device(stage)
    pc.sync()
    stream = torch.cuda.ExternalStream(pc.device_buffers().stream, device=torch.device("cuda", local))

    def timed_device_loop():
        for i in range(Wm):
            set_pose(i)
            pc.render_device(stage)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = pc.launch_count
        e0.record(stream)
        for i in range(K_steps):
            set_pose(Wm + i)
            pc.render_device(stage)
        pc.device_buffers()   # completes the last frame (a fused sequence's last blend) and makes `stream` wait for every stream of the sequence
        e1.record(stream)
        barrier()
        return e0.elapsed_time(e1), pc.launch_count - l0

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- leg 1: device-resident throughput (`value`)
    with ClockSampler(local) as clk:
        ms, launches = timed_device_loop()
    ms = max_over_ranks(ms)
    frames_total = K_steps * (world if args.mode == "frames" else 1)
    points_per_frame = n if args.mode == "frames" else n * world
    value = points_per_frame * frames_total / (ms * 1e-3) / 1e9

    # ---- leg 2: the same loop with per-stage CUDA events (roofline of the dominant kernel)
    peak, peak_src = peaks()
    culling = bool(pc.get_option("chunk_cull"))
    fused = pc.get_option("fuse_active") == 1 and args.mode == "frames"   # large clouds by default (option fuse)

    def staged_loop(timing):
        pc.set_option("timing", timing)
        pc.stage_ms_sum(reset=True)
        pc.cull_stats(reset=True)
        ms_t, _ = timed_device_loop()
        sums, nfr = pc.stage_ms_sum(reset=True)
        passes, streamed = pc.stream_stats(reset=False)
        cf, cvis, nch = pc.cull_stats(reset=True)
        pc.set_option("timing", 0)
        return {"stage_ms": (sums / max(nfr, 1)).tolist(), "frame_ms": ms_t / K_steps, "visible_chunks_per_frame": (cvis / cf if cf else None),
                "chunks_per_pass": (streamed / passes if passes else None), "passes_per_frame": (passes / cf if cf else None), "n_chunks": nch,
                "timed_passes": nfr}   # warm-up frames included (same work)

    def kernel_line(ms, pts):
        a = 16.0 * pts / (ms * 1e-3) / 1e9
        return {"launch_ms": ms, "achieved": a, "frac": a / peak, "frac_of_nominal_8TBps": a / 8000.0}

    names2 = ["clear_classify", "zmin", "blend", "resolve_pyramid_minmax", "up_pass_tensor", "frame"]
    if fused:
        # the default path: ONE point pass per frame (blend of frame k-1 + z-min of frame k over the union of the two frames'
        # visible chunks), timed on the point stream while the previous frame's image passes run on the image stream
        st = staged_loop(3)
        streamed = min(count, st["chunks_per_pass"] * 1024.0)
        names3 = ["classify_pair (point stream)", "fused_blend_zmin (point stream)", "wait_for_image_stream",
                  "resolve_pyramid_minmax_fixup_gate (image stream)", "up_pass_tensor (image stream)", "first_to_last_event"]
        fused_ms = st["stage_ms"][1]
        kl = kernel_line(fused_ms, streamed)
        roofline = {"bound": "hbm",
                    "kernel": (f"fused_ring_kernel: blend of frame k-1 + z-min of frame k over ONE stream of chunks (the union of the two frames' visible "
                               f"1024-point chunks, {streamed / count * 100:.1f}% of the cloud; 16 B/point read once, projected for both cameras)"),
                    "achieved": kl["achieved"], "peak": peak, "unit": "GB/s", "frac": kl["frac"],
                    "frac_of_nominal_8TBps": kl["frac_of_nominal_8TBps"], "peak_source": peak_src, "traffic": None,
                    "launch_ms": fused_ms, "algorithmic_bytes_per_launch": 16.0 * streamed,
                    "points_streamed_per_launch": streamed, "points_in_cloud": count,
                    "visible_chunks_per_frame": st["visible_chunks_per_frame"], "chunks_streamed_per_pass": st["chunks_per_pass"],
                    "chunks_in_cloud": st["n_chunks"], "stage_ms": dict(zip(names3, st["stage_ms"])), "timed_passes": st["timed_passes"],
                    "frame_ms_with_stage_events": st["frame_ms"],
                    "note": "the pass does the arithmetic, gathers and reductions of BOTH of the reference's point passes on every record it reads; "
                            "`two_pass` is the same trajectory with fuse=0 (each frame streams its own list twice)"}
        # SURVEY section 8(d) charges 16 B per point PER PASS: the launch does one z-min-pass unit (frame k) and one blend-pass
        # unit (frame k-1) for the points of the two visible lists while reading their union once
        units = 2.0 * min(count, (st["visible_chunks_per_frame"] or 0.0) * 1024.0)
        eq = 16.0 * units / (fused_ms * 1e-3) / 1e9
        roofline["per_reference_pass"] = {"point_pass_units_per_launch": units, "algorithmic_bytes": 16.0 * units, "achieved": eq, "frac": eq / peak,
                                          "frac_of_nominal_8TBps": eq / 8000.0,
                                          "note": "the reference's algorithm needs 16 B per point for EACH of minDepthPass and accumulatePass; `frac` above "
                                                  "counts the bytes this launch really streams (each visible point once), this entry the two passes' worth of "
                                                  "work it does on them — comparable with round 1's one-pass-per-launch 0.545 / 0.551"}
        # both currencies side by side at the top of the object: `frac` = bytes really streamed, `frac_per_reference_pass` =
        # SURVEY 8(d)'s per-unit figure (16 B per point and pass) x the pass units the launch processes
        roofline["frac_per_reference_pass"] = eq / peak
        roofline["achieved_per_reference_pass"] = eq
        # the same frames as two passes per frame (option fuse = 0): what each pass costs on its own
        fuse_opt = pc.get_option("fuse")
        pc.set_option("fuse", 0)
        ms2, _ = timed_device_loop()
        st2 = staged_loop(2)
        pc.set_option("fuse", fuse_opt)
        vis2 = min(count, st2["visible_chunks_per_frame"] * 1024.0)
        roofline["two_pass"] = {"note": "fuse=0: zmin_ring_kernel and blend_ring_kernel each walk the frame's own visible list",
                                "zmin": kernel_line(st2["stage_ms"][1], vis2), "blend": kernel_line(st2["stage_ms"][2], vis2),
                                "stage_ms": dict(zip(names2, st2["stage_ms"])), "points_streamed_per_pass": vis2,
                                "ms_per_step": max_over_ranks(ms2) / K_steps,
                                "value_gpoints_per_s": points_per_frame * frames_total / (max_over_ranks(ms2) * 1e-3) / 1e9}
        dom = "fused"
        stage_ms = st2["stage_ms"]
    else:
        st = staged_loop(2)
        stage_ms = st["stage_ms"]
        streamed = count if not (culling and st["visible_chunks_per_frame"] is not None) else min(count, st["visible_chunks_per_frame"] * 1024.0)
        dom = "zmin" if stage_ms[1] >= stage_ms[2] else "blend"
        dom_ms = max(stage_ms[1], stage_ms[2])
        kl = kernel_line(dom_ms, streamed)
        roofline = {"bound": "hbm",
                    "kernel": (f"{dom}_{'ring' if pc.get_option('ring') else 'list'}_kernel (walks the frame's visible 1024-point chunks: 16 B/point "
                               f"read for the {streamed / count * 100:.1f}% of the cloud inside or near the frustum)") if culling
                    else f"{dom}_{'ring_' if pc.get_option('ring') == 2 else ''}kernel (every point streamed, 16 B/point read)",
                    "achieved": kl["achieved"], "peak": peak, "unit": "GB/s", "frac": kl["frac"],
                    "frac_of_nominal_8TBps": kl["frac_of_nominal_8TBps"], "peak_source": peak_src, "traffic": None,
                    "launch_ms": dom_ms, "algorithmic_bytes_per_launch": 16.0 * streamed,
                    "points_streamed_per_launch": streamed, "points_in_cloud": count,
                    "zmin": kernel_line(stage_ms[1], streamed), "blend": kernel_line(stage_ms[2], streamed),
                    "stage_ms": dict(zip(names2, stage_ms)), "frame_ms_with_stage_events": st["frame_ms"]}
        if culling:
            roofline["visible_chunks_per_frame"], roofline["chunks_in_cloud"] = st["visible_chunks_per_frame"], st["n_chunks"]
    if culling and args.mode == "frames":
        # the same trajectory with chunk culling off: every record of the cloud is streamed by both passes
        # (the configuration north_star's "16 B/point against the HBM roofline" is quoted on)
        pc.set_option("chunk_cull", 0)
        sa = staged_loop(2)
        ms_all, _ = timed_device_loop()
        pc.set_option("chunk_cull", 1)
        roofline["stream_all"] = {"note": "chunk_cull=0: zmin_kernel / blend_kernel stream all points, 16 B/point/pass",
                                  "zmin": kernel_line(sa["stage_ms"][1], count), "blend": kernel_line(sa["stage_ms"][2], count),
                                  "stage_ms": dict(zip(names2, sa["stage_ms"])), "ms_per_step": max_over_ranks(ms_all) / K_steps,
                                  "value_gpoints_per_s": points_per_frame * frames_total / (max_over_ranks(ms_all) * 1e-3) / 1e9}
    if world == 1:
        # second bound named by north_star: L2 atomic (RED) throughput into a frame-sized buffer
        set_pose(Wm)
        ms_rand, ops_rand = pc.bench_red_min(0, 200_000_000, False)
        _, live = pc.bench_red_min(1, 0, False, iters=1)
        red_peak = ops_rand / ms_rand / 1e6
        roofline["l2_atomics"] = {"measured_red_min_u32_random_Gops": red_peak, "in_frustum_points_this_pose": live,
                                  "zmin_upper_bound_red_Gops": live / stage_ms[1] / 1e6,
                                  "zmin_frac_of_measured_red_peak": live / stage_ms[1] / 1e6 / red_peak,
                                  "blend_upper_bound_red64_Gops": 2 * live / stage_ms[2] / 1e6,
                                  "note": "upper bounds from the two-pass stage times: one RED per in-frustum point (z-min, before the early depth test) / two 64-bit REDs per in-frustum point (blend)"}
    tr = os.path.join(ROOT, "profiles", "traffic.json")   # dram bytes per launch from the committed ncu --set full capture
    if os.path.exists(tr):
        try:
            tj = json.load(open(tr))
            roofline["traffic"] = tj.get(f"{dom}_{args.workload}")
            roofline["traffic_source"] = (f"profiles/traffic.json ({tj.get('source', 'committed ncu --set full capture')}): dram__bytes_read.sum + dram__bytes_write.sum per "
                                          "launch of that capture's poses — not measured in this run")
            ncu_at = tj.get(f"ncu_atomics_{args.workload}")
            if ncu_at and "l2_atomics" in roofline and culling and pc.get_option("ring"):
                # REDs really issued (ncu counters of the committed capture) against the measured random-address RED rate
                # (one 32-byte sector per lane there): what fraction of the pass the reduction path alone accounts for
                for name, stt in (("zmin", stage_ms[1]), ("blend", stage_ms[2])):
                    c = ncu_at.get(f"{name}_ring_kernel")
                    if c:
                        roofline["l2_atomics"][f"{name}_ncu"] = dict(c, red_sector_time_ms_at_measured_peak=c["red_sectors"] / (red_peak * 1e6),
                                                                    frac_of_launch=c["red_sectors"] / (red_peak * 1e6) / stt)
                fd = tj.get(f"fused_{args.workload}_detail")
                if fd and fused:
                    # the fused launch of the committed capture: its scattered 32-byte-sector operations (REDs + z-buffer gathers)
                    # against the measured random-address RED rate — the pass is bound by this path, not by HBM
                    ops = fd["red_sectors"] + fd["gather_sectors"]
                    roofline["l2_atomics"]["fused_ncu"] = dict(
                        fd, scattered_sector_ops=ops, scattered_sector_Gops=ops / fd["duration_us_cold"] / 1e3,
                        frac_of_measured_red_peak=ops / fd["duration_us_cold"] / 1e3 / red_peak,
                        red_sector_time_ms_at_measured_peak=fd["red_sectors"] / (red_peak * 1e6),
                        red_frac_of_launch=fd["red_sectors"] / (red_peak * 1e6) / (fd["duration_us_cold"] * 1e-3),
                        note="REDs + gathers of one fused launch (ncu, stand-alone launch of the capture) per second, over the RED rate "
                             "rtr_bench_red_min measures in THIS run with one random 32-byte sector per lane")
        except Exception:
            pass

    # ---- leg 3: end to end through the public API with HOST buffers (pose in, images out, every step)
    color = torch.empty((2, P * 3), dtype=torch.uint8, pin_memory=True)
    depth = torch.empty((2, P), dtype=torch.float32, pin_memory=True)

    def e2e_loop(per_call):
        for i in range(Wm):
            set_pose(i)
            pc._check(pc._lib.rtr_render_filtered(pc._h, color[0].data_ptr(), depth[0].data_ptr()) if stage == pkg.STAGE_FILTERED
                      else pc._lib.rtr_render_rgbd(pc._h, color[0].data_ptr(), depth[0].data_ptr()))
        barrier()
        t0 = time.perf_counter()
        if per_call:   # the reference's call pattern: one synchronous compute*RGBD per pose
            fn = pc._lib.rtr_render_filtered if stage == pkg.STAGE_FILTERED else pc._lib.rtr_render_rgbd
            for i in range(K_steps):
                set_pose(Wm + i)
                pc._check(fn(pc._h, color[i & 1].data_ptr(), depth[i & 1].data_ptr()))
        else:          # the library's trajectory call: same frames, D2H of frame i overlaps frame i+1
            for c0 in range(0, K_steps, CHUNK):   # host ring of CHUNK frames, reused per call
                m = min(CHUNK, K_steps - c0)
                pc._check(pc._lib.rtr_render_trajectory(pc._h, stage, my[Wm + c0:].ctypes.data_as(pkg._dp), m,
                                                        e2e_bufs[0].data_ptr(), e2e_bufs[1].data_ptr()))
        barrier()
        return (time.perf_counter() - t0) * 1e3

    ms_call = max_over_ranks(e2e_loop(True))
    CHUNK = min(K_steps, 50)
    e2e_bufs = (torch.empty((CHUNK, P * 3), dtype=torch.uint8, pin_memory=True),
                torch.empty((CHUNK, P), dtype=torch.float32, pin_memory=True))
    ms_traj = max_over_ranks(e2e_loop(False))
    checksum = int(e2e_bufs[0][-1].to(torch.int64).sum().item())   # the D2H result is really read
    del e2e_bufs
    e2e_val = points_per_frame * frames_total / (ms_traj * 1e-3) / 1e9
    e2e = {"value": e2e_val, "unit": "Gpoints/s", "h2d_bytes_per_step": 128, "d2h_bytes_per_step": P * 7,
           "api": f"rtr_render_trajectory in calls of {CHUNK} poses (poses from host, BGR + depth images to pinned host memory every frame)",
           "frames_per_s": frames_total / (ms_traj * 1e-3),
           "d2h_gbs": frames_total * P * 7 / (ms_traj * 1e-3) / 1e9,
           "per_call_sync": {"api": "rtr_render_filtered per pose (== computeFilteredRGBD, blocking)",
                             "value": points_per_frame * frames_total / (ms_call * 1e-3) / 1e9,
                             "frames_per_s": frames_total / (ms_call * 1e-3)},
           "cloud_upload_s": upload_s, "last_frame_checksum": checksum}
    if args.mode == "frames":
        # what the box can copy device -> host with nothing else running: all ranks at once, the same bytes per frame
        ceil = d2h_ceiling(torch, dist, world, P)
        e2e["d2h_ceiling_gbs"] = ceil
        e2e["frac_of_d2h_ceiling"] = e2e["d2h_gbs"] / ceil
        e2e["d2h_ceiling_note"] = (f"{world} rank(s) copying {P * 7} B per frame (depth + BGR) from device to pinned host memory concurrently, no rendering; "
                                   "the end-to-end frame rate is bounded by this, not by the kernels")

    line = {"metric": "Gpoints/s (points x frames / s, whole frame: projection + z-buffer + blend + resolve + prefilter)",
            "value": value, "unit": "Gpoints/s", "n_gpus": world, "steps": K_steps, "warmup": Wm,
            "ms_per_step": ms / K_steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32+u32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {points_per_frame} points, {W}x{H}, stage {args.stage}, {n_poses}-pose trajectory",
                       "sharding": ("frame-sharded, cloud replicated" if args.mode == "frames" else
                                    ("point-sharded, ncclAllReduce min/sum" if args.nccl else "point-sharded, two-shot min/sum all-reduce kernels over NVLink peer memory")),
                       "l2": f"inputs larger than L2 ({count * 16 / 1e6:.0f} MB cloud per GPU vs 126 MB)",
                       "poses": f"{K_steps} timed frames per rank in 4 arcs of consecutive trajectory frames; the 4 x {world} arcs are spread evenly over the {n_poses}-pose loop",
                       "distortion": bool(args.distort),
                       "options": {k: pc.get_option(k) for k in ("chunk_cull", "ring", "ring_dynamic", "clear_lean", "fused_up", "pipeline", "fuse", "fuse_active", "zmin_variant", "blend_variant", "key64")}},
            "frames_per_s": frames_total / (ms * 1e-3), "gpu_launches": int(launches), "clocks": clk.summary(),
            "roofline": roofline, "e2e": e2e}
    if args.mode == "points" and world > 1 and not args.nccl:
        line["peer_error"] = pc.get_option("peer_error")
        dist.barrier()
        pc.peer_detach()
    pc.close()
    del pc
    torch.cuda.empty_cache()
    if rank == 0 and world == 1 and not args.no_unet and args.workload in ("c3", "c5_4k") and not args.points and not args.distort:
        line["unet_e2e"] = unet_leg(pkg, torch, wl, local)
    if world > 1 and args.mode == "frames" and not args.no_points_mode:
        # config 4 (point-sharded) in the same run, so that the driver's scaling record carries it
        try:
            line["points_mode"] = points_mode_record(args, pkg, torch, dist, wl, rank, world, local, args.points_mode_points,
                                                     min(K_steps, 100))
        except Exception as e:
            line["points_mode"] = {"error": repr(e)}
            raise
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_port_baseline(args, (n,) + wl[1:], args.cpu_sample_points)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl in ("reference", "cpu"):
        run_reference(args, wl)
    else:
        run_b200(args, wl)


if __name__ == "__main__":
    main()
