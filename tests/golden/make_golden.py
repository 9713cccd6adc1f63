"""Generate tests/golden/*.npz by running the UNMODIFIED REFERENCE (oracle/_ref/libref_rtrenderer.so:
the reference's render.cu / project_cloud.cu / CameraCalibration.cpp compiled where they lie, see
oracle/Makefile) on a B200.  Run on the GPU box:

    gpurun -- 'python tests/golden/make_golden.py'     # writes gpurun_out/golden/*.npz
    cp gpurun_out/golden/*.npz tests/golden/           # then commit

Every array in a golden file is an output of the reference's own code, except
  tap_*      per-point (pixel id, depth bits) of the projection, dumped by rtr_project_points —
             needed because the perspective divide uses MUFU.RCP, which no CPU reproduces; the
             golden stores it SPARSELY as the points where it differs from the CPU oracle's
             correctly-rounded projection (these are the documented pixel-boundary ties);
  cam_proj   read back from the reference object's d_cam_proj (so glm's K*E is pinned too).
The CPU tests re-create the cloud from its seed, apply the tap, run the CPU oracle and compare with
the reference outputs stored here (full arrays for the small cases, sha256 for the large ones).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import __graft_entry__ as entry  # noqa: E402
import oracle  # noqa: E402
import scenes  # noqa: E402


def run_case(pkg, cpu, case, out_dir):
    rec = cpu.synth_packed(case.seed, case.n, 0, case.n, case.hall, case.n_boxes)
    xyz, bgr = scenes.split_records(rec)
    ref = oracle.RefOracle(xyz, bgr)
    mine = pkg.ProjectCloud.from_packed(rec, sort=False)   # taps must follow the seeded cloud's own order
    W, H, P = case.W, case.H, case.W * case.H
    data = {"block_size": np.array([ref.block_size], np.int32), "n_frames": np.array([len(case.poses)], np.int32)}
    for fi, E in enumerate(case.poses):
        o = {}
        rc, color, depth = ref.computeRGBD(W, H, case.K, E)
        assert rc == 1
        o["raw_depth_host"], o["raw_color_host"] = depth.view(np.uint32), color
        o["raw_zbuf"], o["raw_accum"], o["raw_image"] = ref.read("zbuf", P), ref.read("accum", P * 4), ref.read("image", P * 3)
        cam = ref.read("cam_proj", 16)
        rc, color, depth = ref.computeFilteredRGBD(W, H, case.K, E)
        assert rc == 1
        o["flt_depth_host"], o["flt_color_host"] = depth.view(np.uint32), color
        o["flt_tensor"] = ref.read("tensor", P * 5)
        o["flt_minmax"] = np.array([ref.read("min", 1)[0], ref.read("max", 1)[0]], np.uint32)
        # projection tap from the CUDA path fed the reference's own matrix
        mine.set_camera(_calib(pkg, case))
        mine.set_cam_proj_raw(cam)
        pix, zb = mine.project_points()
        cpix, czb = cpu.project(rec, cam, W, H)
        diff = np.nonzero((pix != cpix) | (zb != czb))[0].astype(np.int64)
        data[f"f{fi}_cam_proj"] = cam
        data[f"f{fi}_tap_idx"], data[f"f{fi}_tap_pix"], data[f"f{fi}_tap_zb"] = diff, pix[diff], zb[diff]
        for k, v in o.items():
            if case.full:
                data[f"f{fi}_{k}"] = v
            data[f"f{fi}_{k}_sha"] = np.frombuffer(scenes.sha(v).encode(), dtype=np.uint8)
        print(f"[golden] {case.name} frame {fi}: {int((pix >= 0).sum())} in-frustum points, "
              f"{len(diff)} CPU/GPU projection ties, block_size {ref.block_size}")
    ref.close()
    mine.close()
    np.savez_compressed(os.path.join(out_dir, case.name + ".npz"), **data)


def _calib(pkg, case):
    c = pkg.CameraCalibration()
    c.setIntrinsicsMatrix(case.K)
    c.setWidth(case.W)
    c.setHeight(case.H)
    return c


def main():
    out_dir = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out_dir, exist_ok=True)
    pkg = entry.load_package()
    cpu = oracle.cpu()
    for name in (sys.argv[1:] or list(scenes.CASES)):
        run_case(pkg, cpu, scenes.CASES[name], out_dir)


if __name__ == "__main__":
    main()
