"""File formats on either side of the hot path (SURVEY.md §8 f2, f3 + the PLY fixture writer of f1): host parsing
inside librtr_b200.so, checked against a numpy restatement and — where oracle/_ref travelled — against the
reference's OWN code running on the CPU (Octreegrid.h read/writeOctreeBinary, CameraCalibration::loadCalibration)."""
import os

import numpy as np
import pytest

import scenes


def have_ref():
    import oracle
    return os.path.exists(oracle.REF_LIB)


def grid_keys(xyz):
    """computeGrid (cloudreader.cpp:10-55) in numpy float32."""
    lo = np.minimum(xyz.min(axis=0), np.float32(3.4028235e38))
    hi = np.maximum(xyz.max(axis=0), np.float32(1.1754944e-38))
    mn, mx = np.floor(lo).astype(np.float32), np.ceil(hi).astype(np.float32)
    nb = ((mx - mn) / np.float32(0.25)).astype(np.int32)
    c = np.floor(((xyz - mn) / (mx - mn)).astype(np.float32) * nb.astype(np.float32)).astype(np.int32)
    return c[:, 0] + c[:, 1] * nb[0] + c[:, 2] * nb[0] * nb[1], tuple(int(v) for v in nb), mn, mx


@pytest.fixture(scope="module")
def small_cloud(cpu_oracle):
    rec = cpu_oracle.synth_packed(21, 30_000, 0, 30_000, scenes.HALL_SMALL, 4)
    perm = np.random.default_rng(1).permutation(len(rec))          # file order != cell order
    return scenes.split_records(np.ascontiguousarray(rec[perm]))


def test_oct_write_read_round_trip_and_binning(pkg, small_cloud, tmp_path):
    xyz, bgr = small_cloud
    path = str(tmp_path / "pcd.oct")
    pkg.write_oct(path, xyz, bgr)
    got = pkg.read_oct(path)
    keys, dims, mn, mx = grid_keys(xyz)
    assert got["dims"] == dims == (40, 32, 20)                      # 8 x 6 x 3 m +- noise, rounded OUTWARDS to whole metres, 0.25 m cells
    assert np.array_equal(got["keys"], np.unique(keys)) and int(got["counts"].sum()) == len(xyz)
    order = np.argsort(keys, kind="stable")                        # blocks ascending, arrival order inside a block
    assert np.array_equal(got["xyz"], xyz[order]) and np.array_equal(got["bgr"], bgr[order])
    assert np.array_equal(got["counts"], np.unique(keys, return_counts=True)[1].astype(np.uint64))
    # file size: 16 B header + per block 4 + 8 + 24 B + 15 B per point (Octreegrid.h:62-76)
    assert os.path.getsize(path) == 16 + len(got["keys"]) * 36 + len(xyz) * 15


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
def test_oct_is_byte_compatible_with_the_reference(pkg, small_cloud, tmp_path):
    import oracle
    ref = oracle.RefHost()
    xyz, bgr = small_cloud
    keys, dims, _, _ = grid_keys(xyz)
    # (1) the reference's readOctreeBinary reads OUR file
    mine = str(tmp_path / "mine.oct")
    pkg.write_oct(mine, xyz, bgr)
    rx, rc, rdims = ref.oct_read(mine, len(xyz))
    assert rdims == dims and len(rx) == len(xyz)
    assert (rx[:, 3] == 1.0).all() and (rc[:, 3] == 255).all()
    a = np.concatenate([rx[:, :3].view(np.uint32), rc[:, :3].astype(np.uint32)], axis=1)
    b = np.concatenate([xyz.view(np.uint32), bgr.astype(np.uint32)], axis=1)
    assert np.array_equal(a[np.lexsort(a.T)], b[np.lexsort(b.T)])   # same multiset of points (block order is unspecified)
    # (2) OUR reader reads a file written by the reference's writeOctreeBinary
    theirs = str(tmp_path / "theirs.oct")
    ref.oct_write(theirs, xyz, bgr, keys, dims)
    got = pkg.read_oct(theirs)
    assert got["dims"] == dims and sorted(got["keys"].tolist()) == np.unique(keys).tolist()
    off = 0
    for k, c in zip(got["keys"], got["counts"]):                   # every block: the points of that key, arrival order
        sel = keys == k
        assert np.array_equal(got["xyz"][off:off + int(c)], xyz[sel]) and np.array_equal(got["bgr"][off:off + int(c)], bgr[sel])
        off += int(c)
    # (3) identical bytes apart from the block order: compare per-block payloads incl. the block bounds
    def blocks(path):
        raw = open(path, "rb").read()
        out, p = {}, 16
        for _ in range(np.frombuffer(raw, np.int32, 1, 12)[0]):
            key = int(np.frombuffer(raw, np.int32, 1, p)[0])
            n = int(np.frombuffer(raw, np.uint64, 1, p + 4)[0])
            size = 12 + n * 15 + 24
            out[key] = raw[p:p + size]
            p += size
        assert p == len(raw)
        return raw[:12], out
    h1, b1 = blocks(mine)
    h2, b2 = blocks(theirs)
    assert h1 == h2 and b1.keys() == b2.keys()
    for k in b1:
        n = int(np.frombuffer(b1[k], np.uint64, 1, 4)[0])
        assert b1[k][:12 + n * 15] == b2[k][:12 + n * 15]
    # block bounds: the reference writes whatever its grid holds (zeros here: ref_oct_write does not run computeGrid);
    # ours follow cloudreader.cpp:62-78
    k0 = sorted(b1)[0]
    bb = np.frombuffer(b1[k0][-24:], np.float32)
    assert np.allclose(bb[3:] - bb[:3], 0.25)


CAMERAS_TXT = """# Camera list with one line of data per camera:
#   CAMERA_ID, MODEL, WIDTH, HEIGHT, PARAMS[]
1 OPENCV 1752 1168 1211.45 1210.98 877.3 580.41 -0.05 0.01 0.0005 -0.0005 0.002
2 OPENCV 640 480 1 1 1 1 0 0 0 0 0
"""
FISHEYE_TXT = "7 OPENCV_FISHEYE 1440 1440 600.5 601.25 719.5 720.5 0.03 -0.004 0.0007 -0.0001\n"
CUSTOM_TXT = "1280 720\n900.0 0.0 639.5\n0.0 901.5 359.5\n0 0 1\n-0.05, 0.01, 0.0005, -0.0005, 0.0\n0\n"


@pytest.mark.parametrize("name,text,exp", [
    ("cameras.txt", CAMERAS_TXT, (1752, 1168, 5, False)),
    ("fish/cameras.txt", FISHEYE_TXT, (1440, 1440, 4, True)),
    ("calib.txt", CUSTOM_TXT, (1280, 720, 5, False)),
])
def test_calibration_parsers(pkg, tmp_path, name, text, exp):
    path = tmp_path / name
    path.parent.mkdir(parents=True, exist_ok=True)
    path.write_text(text)
    c = pkg.load_calibration(str(path))
    assert (c.getWidth(), c.getHeight(), len(c.getDistortionParameters()), c.fisheye) == exp
    if have_ref():
        import oracle
        want = oracle.RefHost().load_calibration(str(path))
        assert want is not None and (want["W"], want["H"], want["fisheye"]) == (exp[0], exp[1], exp[3])
        assert np.array_equal(c.getIntrinsicsMatrix(), want["K"])           # cameras.txt values pass through float
        assert np.array_equal(np.array(c.getDistortionParameters()), want["dist"])
    if name == "calib.txt":
        assert c.getIntrinsicsMatrix().tolist() == [[900.0, 0.0, 639.5], [0.0, 901.5, 359.5], [0.0, 0.0, 1.0]]
        assert c.getDistortionParameters() == [-0.05, 0.01, 0.0005, -0.0005, 0.0]
    elif name == "cameras.txt":   # first camera line wins; values pass through float like the reference's locals
        assert c.getFocalLengthX() == float(np.float32(1211.45)) and c.getPrincipalPointY() == float(np.float32(580.41))
        assert c.getDistortionParameters() == [float(np.float32(v)) for v in (-0.05, 0.01, 0.0005, -0.0005, 0.002)]


def test_calibration_errors(pkg, tmp_path):
    bad = tmp_path / "cameras.txt"
    bad.write_text("1 SIMPLE_RADIAL 640 480 500 320 240 0.1\n")
    with pytest.raises(pkg.RtrError):
        pkg.load_calibration(str(bad))
    with pytest.raises(pkg.RtrError):
        pkg.load_calibration(str(tmp_path / "missing.txt"))
    short = tmp_path / "calib.txt"
    short.write_text("640 480\n1 0 0\n0 1 0\n0 0 1\n0.1 0.2\n0\n")
    with pytest.raises(pkg.RtrError):
        pkg.load_calibration(str(short))     # "Pinhole camera expects 5 distortion parameters"


def test_trajectory_parser_and_pose_inverse(pkg, tmp_path):
    rng = np.random.default_rng(2)
    n = 17
    q = rng.standard_normal((n, 4)) * rng.uniform(0.5, 2.0, (n, 1))          # not normalised on purpose (qx qy qz qw)
    t = rng.uniform(-5, 5, (n, 3))
    lines = ["# timestamp tx ty tz qx qy qz qw", ""]
    for i in range(n):
        lines.append(" ".join(repr(float(v)) for v in [i * 0.1, *t[i], *q[i]]))
    p = tmp_path / "traj.txt"
    p.write_text("\n".join(lines) + "\n")
    poses = pkg.load_trajectory(str(p), order=0)
    assert poses.shape == (n, 4, 4)
    for i in range(n):
        x, y, z, w = q[i] / np.linalg.norm(q[i])
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                      [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                      [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
        assert np.allclose(poses[i, :3, :3], R, atol=1e-14) and np.allclose(poses[i, :3, 3], t[i])
        assert poses[i, 3].tolist() == [0, 0, 0, 1]
        assert np.allclose(pkg.invert_rigid(poses[i]), np.linalg.inv(poses[i]), atol=1e-12)   # pose.inv(), main.cpp:96
    # COLMAP images.txt order (README.md:92): id qw qx qy qz tx ty tz
    lines = [" ".join(repr(float(v)) for v in [i + 1, q[i][3], q[i][0], q[i][1], q[i][2], *t[i]]) for i in range(n)]
    p2 = tmp_path / "images.txt"
    p2.write_text("\n".join(lines) + "\n")
    assert np.allclose(pkg.load_trajectory(str(p2), order=1), poses, atol=1e-14)
    assert len(pkg.load_trajectory(str(p), order=0, max_poses=5)) == 5


def test_ply_writer_layout(pkg, small_cloud, tmp_path):
    xyz, bgr = small_cloud
    path = tmp_path / "c.ply"
    pkg.write_ply(str(path), xyz[:100], bgr[:100])
    raw = path.read_bytes()
    head, body = raw.split(b"end_header\n", 1)
    assert head.startswith(b"ply\nformat binary_little_endian 1.0\nelement vertex 100\n") and len(body) == 100 * 15
    rec = np.frombuffer(body, dtype=np.dtype([("p", "<f4", 3), ("rgb", "u1", 3)]))
    assert np.array_equal(rec["p"], xyz[:100]) and np.array_equal(rec["rgb"][:, ::-1], bgr[:100])   # file is R,G,B


def test_corrupt_files_return_errors_instead_of_throwing(pkg, tmp_path):
    """Header fields of a corrupt / truncated file are bounded by the file size before anything is allocated, and no
    C++ exception crosses the C ABI: the calls return RTR_ERR_ARG (a 2^32-vertex PLY header, an .oct header announcing
    2^31 - 1 blocks, a block announcing 2^40 points, truncated payloads)."""
    import ctypes as C
    import struct
    lib = pkg.load_library()
    # .oct: absurd block count / absurd block size / truncated block
    for name, payload in (("blocks.oct", struct.pack("<4i", 4, 4, 4, 2**31 - 1)),
                          ("count.oct", struct.pack("<4i", 4, 4, 4, 1) + struct.pack("<iQ", 7, 2**40)),
                          ("trunc.oct", struct.pack("<4i", 4, 4, 4, 1) + struct.pack("<iQ", 7, 1000) + b"\0" * 100),
                          ("neg.oct", struct.pack("<4i", 4, 4, 4, -5))):
        p = tmp_path / name
        p.write_bytes(payload)
        xyz, bgr, n = C.POINTER(C.c_float)(), C.POINTER(C.c_uint8)(), C.c_uint64(0)
        hdr = (C.c_int * 4)()
        rc = lib.rtr_io_read_oct(os.fsencode(str(p)), C.byref(xyz), C.byref(bgr), C.byref(n), hdr, None, None)
        assert rc == pkg.RTR_ERR_ARG, name
        assert not xyz and not bgr and n.value == 0
    # (the PLY loader needs a renderer, i.e. a GPU: its header bound is checked in tests/test_gpu_io.py)


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
def test_cell_binning_is_pinned_to_the_reference_loader(pkg, small_cloud, tmp_path):
    """f1 against the reference's OWN loader: its cloudreader.cpp compiled unmodified (tinyply is vendored) reads a .ply and
    bins it into 0.25 m blocks (loadPLY + computeGrid, cloudreader.cpp:122-177, 8-82).  Every point must come back with the
    B,G,R colour the file implies and in the block our numpy restatement of computeGrid (grid_keys — what rtr_bin_cells /
    rtr_io_write_oct are tested against on the GPU and above) assigns it, points in arrival order inside a block."""
    import oracle
    xyz, bgr = small_cloud
    ref = oracle.RefHost()
    if not hasattr(ref.lib, "ref_load_ply"):
        pytest.skip("oracle/_ref predates ref_load_ply: rebuild it (make -C oracle ref)")
    path = str(tmp_path / "cloud.ply")
    pkg.write_ply(path, xyz, bgr)
    rx, rc, rk, n_blocks = ref.load_ply(path, len(xyz) + 16)
    assert len(rx) == len(xyz)
    keys, dims, _, _ = grid_keys(xyz)
    assert n_blocks == len(np.unique(keys))
    # per block: the same points in the same (arrival) order
    order_ref = np.argsort(rk, kind="stable")
    order_mine = np.argsort(keys, kind="stable")
    assert np.array_equal(rk[order_ref], keys[order_mine])
    assert np.array_equal(rx[order_ref].view(np.uint32), xyz[order_mine].view(np.uint32))
    assert np.array_equal(rc[order_ref], bgr[order_mine])          # red/green/blue of the file -> B,G,R (cloudreader.cpp:168)
