"""GPU side of the "next" rows: PLY / .oct loading into the renderer, 0.25 m cell
binning on the device, U-Net output post-process."""
import numpy as np
import pytest

import scenes
from test_io_formats import grid_keys

pytestmark = pytest.mark.gpu


def _frame(pkg, pc, case):
    calib = pkg.CameraCalibration()
    calib.setIntrinsicsMatrix(case.K)
    calib.setWidth(case.W)
    calib.setHeight(case.H)
    P = case.W * case.H
    color, depth = np.zeros(P * 3, np.uint8), np.zeros(P, np.float32)
    assert pc.computeFilteredRGBD(calib, case.poses[0], color, depth) == 1
    return color, depth.view(np.uint32), pc.read("tensor", np.uint16, P * 5)


@pytest.fixture(scope="module")
def shuffled(cpu_oracle):
    case = scenes.CASES["c1_640x480"]
    rec = cpu_oracle.synth_packed(case.seed, 300_000, 0, 300_000, case.hall, case.n_boxes)
    perm = np.random.default_rng(3).permutation(len(rec))
    return case, np.ascontiguousarray(rec[perm])


def test_ply_loads_like_the_reference_loader(gpu, shuffled, tmp_path):
    case, rec = shuffled
    xyz, bgr = scenes.split_records(rec)
    path = str(tmp_path / "cloud.ply")
    gpu.write_ply(path, xyz, bgr)
    raw = gpu.ProjectCloud.from_ply(path, bin_cells=False, sort=False)
    assert np.array_equal(raw.download_cloud().view(np.uint32), rec.view(np.uint32))       # file order, B,G,R packing
    direct = gpu.ProjectCloud.from_packed(rec)
    want = _frame(gpu, direct, case)
    binned = gpu.ProjectCloud.from_ply(path, bin_cells=True)
    keys, dims, _, _ = grid_keys(xyz)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(binned.download_cloud().view(np.uint32), rec[order].view(np.uint32))  # computeGrid grouping, arrival order kept
    for pc in (raw, binned):
        got = _frame(gpu, pc, case)
        assert all(np.array_equal(a, b) for a, b in zip(got, want))
    # grouping by cell is what makes chunk culling bite on an unordered file
    f1, v1, n1 = raw.cull_stats()
    f2, v2, n2 = binned.cull_stats()
    assert n1 == n2 and v1 / f1 > 0.95 * n1 and v2 / f2 < 0.8 * n2
    # the default upload (Morton order) culls at least as well as the reference's cell grouping
    f3, v3, n3 = direct.cull_stats()
    assert v3 / f3 <= v2 / f2 * 1.1
    for pc in (raw, binned, direct):
        pc.close()
    # ascii flavour with double coordinates and extra properties
    apath = tmp_path / "ascii.ply"
    n = 500
    with open(apath, "w") as f:
        f.write(f"ply\nformat ascii 1.0\ncomment test\nelement vertex {n}\nproperty double x\nproperty double y\nproperty double z\n"
                "property float intensity\nproperty uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n")
        for i in range(n):
            f.write(f"{float(xyz[i, 0])!r} {float(xyz[i, 1])!r} {float(xyz[i, 2])!r} 0.5 {bgr[i, 2]} {bgr[i, 1]} {bgr[i, 0]}\n")
    a = gpu.ProjectCloud.from_ply(str(apath), bin_cells=False, sort=False)
    assert np.array_equal(a.download_cloud().view(np.uint32), rec[:n].view(np.uint32))
    a.close()
    with pytest.raises(gpu.RtrError):
        gpu.ProjectCloud.from_ply(str(tmp_path / "nope.ply"))


def test_bin_cells_on_resident_cloud(gpu, shuffled):
    case, rec = shuffled
    pc = gpu.ProjectCloud.from_packed(rec, sort=False)
    before = _frame(gpu, pc, case)
    keys, dims, _, _ = grid_keys(rec[:, :3])
    assert pc.bin_cells() == dims
    assert np.array_equal(pc.download_cloud().view(np.uint32), rec[np.argsort(keys, kind="stable")].view(np.uint32))
    after = _frame(gpu, pc, case)
    assert all(np.array_equal(a, b) for a, b in zip(before, after))
    pc.close()


def test_oct_cache_loads(gpu, shuffled, tmp_path):
    case, rec = shuffled
    xyz, bgr = scenes.split_records(rec)
    path = str(tmp_path / "pcd.oct")
    gpu.write_oct(path, xyz, bgr)
    pc = gpu.ProjectCloud.from_oct(path, sort=False)
    keys, _, _, _ = grid_keys(xyz)
    assert np.array_equal(pc.download_cloud().view(np.uint32), rec[np.argsort(keys, kind="stable")].view(np.uint32))
    direct = gpu.ProjectCloud.from_packed(rec)
    assert all(np.array_equal(a, b) for a, b in zip(_frame(gpu, pc, case), _frame(gpu, direct, case)))
    pc.close()
    direct.close()


def test_unet_output_postprocess(gpu):
    """fp16 3xHxW -> uint8 HxWx3 = saturate(round_half_even(v * 255)) (permute + convertTo, project_cloud.cu:475-480)."""
    import torch
    W, H = 96, 40
    g = torch.Generator().manual_seed(0)
    x = (torch.rand(3, H, W, generator=g) * 1.4 - 0.2).half()
    x[0, 0, :8] = torch.tensor([0.5 / 255, 1.5 / 255, 2.5 / 255, float("nan"), float("inf"), -float("inf"), 1.0, 0.0]).half()
    xd = x.cuda().contiguous()
    pc = gpu.ProjectCloud()
    got = pc.postprocess_unet_output(xd.data_ptr(), W, H)
    pc.close()
    v = x.float().numpy().astype(np.float64) * 255.0
    with np.errstate(invalid="ignore"):
        want = np.clip(np.nan_to_num(np.rint(v), nan=0.0, posinf=255, neginf=0), 0, 255).astype(np.uint8)
    assert np.array_equal(got, want.transpose(1, 2, 0))


def test_config1_ply_file_to_reference_frame(gpu, cpu_oracle, tmp_path):
    """BASELINE config 1 end to end: a 1 M-point coloured .ply, one 640x480 pinhole pose, projection + z-buffer +
    prefilter — loaded through the PLY loader (binned like the reference's loader), the frame must equal what the
    unmodified reference produced for that cloud and pose (golden c1_640x480)."""
    import os
    from conftest import ROOT
    case = scenes.CASES["c1_640x480"]
    g = np.load(os.path.join(ROOT, "tests", "golden", case.name + ".npz"))
    rec = cpu_oracle.synth_packed(case.seed, case.n, 0, case.n, case.hall, case.n_boxes)
    xyz, bgr = scenes.split_records(rec)
    path = str(tmp_path / "c1.ply")
    gpu.write_ply(path, xyz, bgr)
    assert os.path.getsize(path) > 15 * case.n
    for kw in (dict(bin_cells=True), dict(bin_cells=False)):
        pc = gpu.ProjectCloud.from_ply(path, **kw)
        color, depth, tensor = _frame(gpu, pc, case)
        pc.close()
        assert scenes.sha(depth) == bytes(g["f0_flt_depth_host_sha"]).decode()
        assert scenes.sha(color) == bytes(g["f0_flt_color_host_sha"]).decode()
        assert scenes.sha(tensor) == bytes(g["f0_flt_tensor_sha"]).decode()


def test_ply_header_promising_more_vertices_than_the_file_holds(gpu, tmp_path):
    """A corrupt vertex count (2^32 - 1 here: 64 GB of records if trusted) is checked against the file size before
    anything is allocated; the call returns RTR_ERR_ARG instead of throwing bad_alloc through the C ABI."""
    p = tmp_path / "lying.ply"
    header = ("ply\nformat binary_little_endian 1.0\nelement vertex 4294967295\nproperty float x\nproperty float y\n"
              "property float z\nproperty uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n").encode()
    p.write_bytes(header + b"\0" * 150)
    with pytest.raises(gpu.RtrError) as e:
        gpu.ProjectCloud.from_ply(str(p))
    assert e.value.code == gpu.RTR_ERR_ARG and "truncated" in str(e.value)
    q = tmp_path / "lying_ascii.ply"
    q.write_bytes(b"ply\nformat ascii 1.0\nelement vertex 100000000\nproperty float x\nproperty float y\nproperty float z\nend_header\n1 2 3\n")
    with pytest.raises(gpu.RtrError) as e:
        gpu.ProjectCloud.from_ply(str(q))
    assert e.value.code == gpu.RTR_ERR_ARG
