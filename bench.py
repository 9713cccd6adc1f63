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
    ap.add_argument("--cpu-sample-points", type=int, default=20_000_000)
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

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
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
            self._stop.wait(0.01)

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
    with ClockSampler(0) as clk:
        t0 = time.perf_counter()
        for i in range(args.steps):
            step(args.warmup + i)
        dt = time.perf_counter() - t0   # every reference call ends in blocking cudaMemcpy D2H
    kms = ref.time_point_kernels(W, H, iters=5)
    value = n * args.steps / dt / 1e9
    line = dict(base, value=value, ms_per_step=dt / args.steps * 1e3, frames_per_s=args.steps / dt,
                gpu_launches=0, clocks=clk.summary(),
                cpu_baseline={"value": value, "unit": "Gpoints/s", "cores": 1, "kind": "reference",
                              "sample": f"{args.steps} frames, full workload; the reference's path is CUDA (render.cu/project_cloud.cu "
                                        "compiled unmodified for sm_100, one host thread driving one B200), not a CPU implementation"},
                e2e={"value": value, "unit": "Gpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                reference_kernels_ms={"clear": float(kms[0]), "minDepthPass": float(kms[1]), "accumulatePass": float(kms[2]),
                                      "resolvePass": float(kms[3]), "block_size": ref.block_size, "build": "stock" if stock else "zero-init parity build",
                                      "minDepthPass_GBps_at_16B_per_point": 16.0 * n / (float(kms[1]) * 1e-3) / 1e9})
    line["config"]["note"] = "unmodified reference CUDA code through ProjectCloud::compute*RGBD with host outputs (sync + pageable D2H per frame, as shipped)"
    ref.close()
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
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
        pc.set_option("index_base", n * rank)
        if world > 1 and args.nccl:
            uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                uid.copy_(torch.frombuffer(bytearray(pkg.ProjectCloud.comm_unique_id()), dtype=torch.uint8))
            dist.broadcast(uid, 0)
            pc.comm_init(bytes(uid.cpu().numpy().tobytes()), rank, world)
        elif world > 1:   # our own two-shot all-reduce over NVLink peer memory
            blob = torch.frombuffer(bytearray(pc.peer_export()), dtype=torch.uint8).cuda()
            blobs = [torch.zeros(512, dtype=torch.uint8, device="cuda") for _ in range(world)]
            dist.all_gather(blobs, blob)
            pc.peer_attach(b"".join(bytes(b.cpu().numpy().tobytes()) for b in blobs), rank, world)
    poses = trajectory(pkg, hall, n_poses)
    if args.mode == "frames":   # frame f of the (looping) trajectory goes to rank f mod N (SURVEY.md §8 e)
        my = [poses[(i * world + rank) % len(poses)] for i in range(K_steps + Wm)]
    else:
        my = [poses[i % len(poses)] for i in range(K_steps + Wm)]
    my = np.ascontiguousarray(np.stack(my).reshape(-1, 16))

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
        pc.device_buffers()   # consecutive frames alternate between two streams: make `stream` wait for both before the end event
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

    # ---- leg 2: same loop with per-stage CUDA events (roofline of the dominant kernel)
    peak, peak_src = peaks()
    names = ["clear_classify", "zmin", "blend", "resolve_pyramid_minmax", "up_pass_tensor", "frame"]

    def staged_loop():
        pc.set_option("timing", 2)
        pc.stage_ms_sum(reset=True)
        pc.cull_stats(reset=True)
        ms_t, _ = timed_device_loop()
        sums, nfr = pc.stage_ms_sum(reset=True)
        cf, cvis, nch = pc.cull_stats(reset=True)
        pc.set_option("timing", 0)
        return (sums / max(nfr, 1)).tolist(), ms_t / K_steps, (cvis / cf if cf else None), nch   # warm-up frames included (same work)

    culling = bool(pc.get_option("chunk_cull"))
    stage_ms, frame_ms_ev, vis_chunks, n_chunks = staged_loop()
    streamed = count if not (culling and vis_chunks is not None) else min(count, vis_chunks * 1024.0)

    def kernel_line(ms, pts):
        a = 16.0 * pts / (ms * 1e-3) / 1e9
        return {"launch_ms": ms, "achieved": a, "frac": a / peak, "frac_of_nominal_8TBps": a / 8000.0}
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
                "stage_ms": dict(zip(names, stage_ms)), "frame_ms_with_stage_events": frame_ms_ev}
    if culling:
        # the same trajectory with chunk culling off: every record of the cloud is streamed by both passes
        # (the configuration north_star's "16 B/point against the HBM roofline" is quoted on)
        roofline["visible_chunks_per_frame"], roofline["chunks_in_cloud"] = vis_chunks, n_chunks
        pc.set_option("chunk_cull", 0)
        sm_all, fm_all, _, _ = staged_loop()
        ms_all, _ = timed_device_loop()
        pc.set_option("chunk_cull", 1)
        roofline["stream_all"] = {"note": "chunk_cull=0: zmin_kernel / blend_kernel stream all points, 16 B/point/pass",
                                  "zmin": kernel_line(sm_all[1], count), "blend": kernel_line(sm_all[2], count),
                                  "stage_ms": dict(zip(names, sm_all)), "ms_per_step": max_over_ranks(ms_all) / K_steps,
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
                                  "note": "upper bounds: one RED per in-frustum point (z-min, before the early depth test) / two 64-bit REDs per in-frustum point (blend)"}
    tr = os.path.join(ROOT, "profiles", "traffic.json")   # dram bytes per launch from the committed ncu --set full capture
    if os.path.exists(tr):
        try:
            tj = json.load(open(tr))
            roofline["traffic"] = tj.get(f"{dom}_{args.workload}")
            ncu_at = tj.get(f"ncu_atomics_{args.workload}")
            if ncu_at and "l2_atomics" in roofline and culling and pc.get_option("ring"):
                # REDs really issued (ncu counters of the committed capture) against the measured random-address RED rate
                # (one 32-byte sector per lane there): what fraction of the pass the reduction path alone accounts for
                for name, st in (("zmin", stage_ms[1]), ("blend", stage_ms[2])):
                    c = ncu_at.get(f"{name}_ring_kernel")
                    if c:
                        roofline["l2_atomics"][f"{name}_ncu"] = dict(c, red_sector_time_ms_at_measured_peak=c["red_sectors"] / (red_peak * 1e6),
                                                                    frac_of_launch=c["red_sectors"] / (red_peak * 1e6) / st)
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
    e2e_val = points_per_frame * frames_total / (ms_traj * 1e-3) / 1e9
    e2e = {"value": e2e_val, "unit": "Gpoints/s", "h2d_bytes_per_step": 128, "d2h_bytes_per_step": P * 7,
           "api": f"rtr_render_trajectory in calls of {CHUNK} poses (poses from host, BGR + depth images to pinned host memory every frame)",
           "frames_per_s": frames_total / (ms_traj * 1e-3),
           "per_call_sync": {"api": "rtr_render_filtered per pose (== computeFilteredRGBD, blocking)",
                             "value": points_per_frame * frames_total / (ms_call * 1e-3) / 1e9,
                             "frames_per_s": frames_total / (ms_call * 1e-3)},
           "cloud_upload_s": upload_s, "last_frame_checksum": checksum}

    line = {"metric": "Gpoints/s (points x frames / s, whole frame: projection + z-buffer + blend + resolve + prefilter)",
            "value": value, "unit": "Gpoints/s", "n_gpus": world, "steps": K_steps, "warmup": Wm,
            "ms_per_step": ms / K_steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32+u32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {points_per_frame} points, {W}x{H}, stage {args.stage}, {n_poses}-pose trajectory",
                       "sharding": ("frame-sharded, cloud replicated" if args.mode == "frames" else
                                    ("point-sharded, ncclAllReduce min/sum" if args.nccl else "point-sharded, two-shot min/sum all-reduce kernels over NVLink peer memory")),
                       "l2": f"inputs larger than L2 ({count * 16 / 1e6:.0f} MB cloud per GPU vs 126 MB)",
                       "distortion": bool(args.distort),
                       "options": {k: pc.get_option(k) for k in ("chunk_cull", "ring", "ring_dynamic", "clear_lean", "fused_up", "pipeline", "zmin_variant", "blend_variant", "key64")}},
            "frames_per_s": frames_total / (ms * 1e-3), "gpu_launches": int(launches), "clocks": clk.summary(),
            "roofline": roofline, "e2e": e2e}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_port_baseline(args, (n,) + wl[1:], args.cpu_sample_points)
    if args.mode == "points" and world > 1 and not args.nccl:
        line["peer_error"] = pc.get_option("peer_error")
        dist.barrier()
        pc.peer_detach()
    pc.close()
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
