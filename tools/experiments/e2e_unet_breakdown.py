#!/usr/bin/env python
"""Where do the e2e U-Net milliseconds go? (measurement only)"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import __graft_entry__ as entry, bench
from e2e_unet import DevPtr
pkg = entry.load_package()
W, H = 1920, 1080
n, _, _, f, cx, cy, hall, boxes, seed, n_poses = bench.WORKLOADS["c3"]
n = 20_000_000
model = torch.jit.load(os.path.join(ROOT, "oracle", "_ref", "unet_1920x1080.pt")).cuda().eval()
pc = pkg.ProjectCloud.synthetic(seed=seed, n_total=n, hall=hall, n_boxes=boxes)
calib = bench.make_calib(pkg, W, H, f, cx, cy)
poses = bench.trajectory(pkg, hall, n_poses)
pc.set_camera(calib, poses[0]); pc.render_device(pkg.STAGE_FILTERED); pc.sync()
bufs = pc.device_buffers()
tin = torch.as_tensor(DevPtr(bufs.tensor, (1, 5, H, W), "<f2"), device="cuda")
color = torch.empty(W * H * 3, dtype=torch.uint8, pin_memory=True)
depth = torch.empty(W * H, dtype=torch.float32, pin_memory=True)
for mode in ("default_stream", "external_stream", "default_stream_benchmark"):
    if mode.endswith("benchmark"):
        torch.backends.cudnn.benchmark = True
    stream = torch.cuda.ExternalStream(bufs.stream) if mode == "external_stream" else torch.cuda.current_stream()
    ts = []
    for i in range(12):
        t0 = time.perf_counter()
        pc.set_camera(calib, poses[i * 7]); pc.render_device(pkg.STAGE_FILTERED)
        if mode != "external_stream":
            pc.sync()
        t1 = time.perf_counter()
        with torch.no_grad(), torch.cuda.stream(stream):
            y = model(tin)[0].contiguous()
        t2 = time.perf_counter()
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        pc._check(pc._lib.rtr_postprocess_unet_output(pc._h, y.data_ptr(), W, H, color.data_ptr(), None))
        t4 = time.perf_counter()
        pc._check(pc._lib.rtr_read_buffer(pc._h, 0, depth.data_ptr(), W * H * 4))
        t5 = time.perf_counter()
        ts.append((t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4))
    print(mode, ["%.1f/%.1f/%.1f/%.1f/%.1f" % tuple(1e3 * v for v in t) for t in ts[2:8]])
