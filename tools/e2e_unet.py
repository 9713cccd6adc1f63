#!/usr/bin/env python
"""Config 5 (BASELINE.json configs[4]): end-to-end frames/s of the new projection + prefilter feeding the
REFERENCE's libtorch U-Net (TorchScript exported by tools/export_unet.py from the reference's own model.py, seeded
random weights), next to the reference's own computeFull on the same B200, same cloud, same poses, same model file.

    python tools/e2e_unet.py --res 1920x1080 --points 100000000 --frames 30

ours      : rtr_render_device (async) -> tensor wrapped zero-copy as torch fp16 {1,5,H,W} (what torch::from_blob does
            at project_cloud.cu:471) -> model.forward on the renderer's stream -> rtr_postprocess_unet_output
            (fp16 CHW -> uint8 HWC on the GPU) -> pinned host image + depth.
reference : ProjectCloud::computeFull (project_cloud.cu:437-493) through oracle/_ref (stock build), model loaded by
            the reference itself from ~/.render_cache.
Measurement support; writes gpurun_out/e2e_unet_<res>.json."""
import argparse
import json
import os
import shutil
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402
import bench  # noqa: E402


class DevPtr:
    """Zero-copy view of a device buffer for torch.as_tensor (the Python spelling of torch::from_blob)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 2}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", default="1920x1080")
    ap.add_argument("--points", type=int, default=100_000_000)
    ap.add_argument("--frames", type=int, default=30)
    ap.add_argument("--no-reference", action="store_true")
    args = ap.parse_args()
    W, H = (int(v) for v in args.res.split("x"))
    model_file = os.path.join(ROOT, "oracle", "_ref", f"unet_{W}x{H}.pt")
    if not os.path.exists(model_file):
        raise SystemExit(f"{model_file} missing: run tools/export_unet.py {W}x{H} where /root/reference exists")
    pkg = entry.load_package()
    wl = bench.WORKLOADS["c3" if W == 1920 else "c5_4k"]
    n, _, _, f, cx, cy, hall, boxes, seed, n_poses = wl
    n = args.points
    calib = bench.make_calib(pkg, W, H, f, cx, cy)
    poses = bench.trajectory(pkg, hall, n_poses)
    poses = [poses[(i * 7) % len(poses)] for i in range(args.frames + 3)]
    P = W * H
    out = {"resolution": args.res, "points": n, "frames": args.frames, "model": os.path.basename(model_file)}

    # ---------------- ours
    torch.cuda.set_device(0)
    model = torch.jit.load(model_file).cuda().eval()
    pc = pkg.ProjectCloud.synthetic(seed=seed, n_total=n, hall=hall, n_boxes=boxes)
    pc.set_camera(calib, poses[0])
    pc.render_device(pkg.STAGE_FILTERED)
    pc.sync()
    bufs = pc.device_buffers()
    stream = torch.cuda.ExternalStream(bufs.stream)
    tin = torch.as_tensor(DevPtr(bufs.tensor, (1, 5, H, W), "<f2"), device="cuda")
    color = torch.empty(P * 3, dtype=torch.uint8, pin_memory=True)
    depth = torch.empty(P, dtype=torch.float32, pin_memory=True)
    frames_ours = []

    def ours(E):
        pc.set_camera(calib, E)
        pc.render_device(pkg.STAGE_FILTERED)
        assert pc.device_buffers().tensor == bufs.tensor
        with torch.no_grad(), torch.cuda.stream(stream):
            y = model(tin)[0].contiguous()                       # 3 x H x W fp16
        pc._check(pc._lib.rtr_postprocess_unet_output(pc._h, y.data_ptr(), W, H, color.data_ptr(), None))
        pc._check(pc._lib.rtr_read_buffer(pc._h, 0, depth.data_ptr(), P * 4))
        return y

    for E in poses[:3]:
        ours(E)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for E in poses[3:]:
        ours(E)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out["ours"] = {"frames_per_s": args.frames / dt, "ms_per_frame": dt / args.frames * 1e3}
    # stage split of the last frame
    pc.set_option("timing", 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pc.set_camera(calib, poses[-1])
    pc.render_device(pkg.STAGE_FILTERED)
    with torch.no_grad(), torch.cuda.stream(stream):
        e0.record(stream)
        y = model(tin)[0].contiguous()
        e1.record(stream)
    torch.cuda.synchronize()
    out["ours"]["projection_prefilter_ms"] = float(pc.stage_ms()[5])
    out["ours"]["unet_ms"] = float(e0.elapsed_time(e1))
    pc.set_option("timing", 0)
    ours(poses[-1])
    last_ours = color.numpy().copy()
    last_depth = depth.numpy().copy()
    pc.close()
    del model, tin
    torch.cuda.empty_cache()

    # ---------------- reference
    if not args.no_reference:
        import oracle
        cache = os.path.join(os.environ.get("HOME", "/root"), ".render_cache")
        os.makedirs(cache, exist_ok=True)
        shutil.copy(model_file, os.path.join(cache, os.path.basename(model_file)))
        cpu = oracle.cpu()
        rec = cpu.synth_packed(seed, n, 0, n, hall, boxes)
        c = np.ascontiguousarray(rec[:, 3]).view(np.uint32)
        bgr = np.stack([c & 0xFF, (c >> 8) & 0xFF, (c >> 16) & 0xFF], axis=1).astype(np.uint8)
        xyz = np.ascontiguousarray(rec[:, :3])
        del rec, c
        ref = oracle.RefOracle(xyz, bgr, stock=os.path.exists(oracle.REF_LIB_STOCK), model_name=os.path.basename(model_file))
        del xyz, bgr
        K = calib.getIntrinsicsMatrix()
        for E in poses[:3]:
            ref.computeFull(W, H, K, E)
        t0 = time.perf_counter()
        for E in poses[3:]:
            rc, rcol, rdep = ref.computeFull(W, H, K, E)
        dt = time.perf_counter() - t0
        out["reference"] = {"frames_per_s": args.frames / dt, "ms_per_frame": dt / args.frames * 1e3}
        out["speedup_frames_per_s"] = out["ours"]["frames_per_s"] / out["reference"]["frames_per_s"]
        d = np.abs(rcol.astype(np.int32) - last_ours.astype(np.int32))
        out["last_frame_vs_reference"] = {"depth_identical": bool(np.array_equal(rdep.view(np.uint32), last_depth.view(np.uint32))),
                                          "colour_max_abs_diff": int(d.max()), "colour_differing_bytes": int((d > 0).sum()),
                                          "colour_bytes": int(d.size)}
        ref.close()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"e2e_unet_{args.res}.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
