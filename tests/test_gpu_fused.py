"""Fused frame sequences (csrc/rtr_renderer.cu enqueue_fused, csrc/rtr_point_ring.cu fused_ring_kernel): frame k-1's
blend and frame k's z-min walk ONE list — the union of the two frames' visible chunks — so every chunk is read once per
frame.  Every frame must stay byte-identical to the blocking per-pose call (two passes, one frame at a time), whatever
the sequence of API calls around it."""
import numpy as np
import pytest

import scenes
from test_gpu_parity import calib_of, cloud_of

pytestmark = pytest.mark.gpu


def _blocking(gpu, rec, calib, poses, stage_filtered=True, options=None):
    pc = gpu.ProjectCloud.from_packed(rec)
    for k, v in (options or {}).items():
        pc.set_option(k, v)
    P = calib.getWidth() * calib.getHeight()
    out = []
    for E in poses:
        color, depth = np.zeros(P * 3, np.uint8), np.zeros(P, np.float32)
        fn = pc.computeFilteredRGBD if stage_filtered else pc.computeRGBD
        assert fn(calib, E, color, depth) == 1
        out.append((color, depth.view(np.uint32).copy(), pc.read("tensor", np.uint16, P * 5) if stage_filtered else None))
    pc.close()
    return out


def _trajectory(gpu, n, hall=scenes.HALL_LARGE):
    return gpu.trajectory_w2c(n, center=(hall[0] * 0.125, hall[1] * 0.125, 1.5), radius=2.0)


def test_fused_trajectory_equals_frame_by_frame_64_poses(gpu, cpu_oracle):
    case = scenes.CASES["c3_1920x1080"]
    rec = cloud_of(cpu_oracle, case)
    calib = calib_of(gpu, case)
    P = case.W * case.H
    poses = _trajectory(gpu, 1000)[100:164]           # 64 consecutive poses of the 1000-pose loop
    want = _blocking(gpu, rec, calib, poses)
    for opts in ({}, {"fuse": 0}, {"ring_dynamic": 0}, {"ring_ctas": 1}, {"ring_claim_min": 0, "ring_dynamic": 3}, {"fused_tiles_per_cta": 3}):
        pc = gpu.ProjectCloud.from_packed(rec)
        assert pc.get_option("fuse") == 1                       # default: fused sequences for large clouds only ...
        pc.set_option("fuse", 2)                                # ... this 2 M-point cloud takes them when told to
        for k, v in opts.items():
            pc.set_option(k, v)
        pc.set_camera(calib)
        color = np.zeros((len(poses), P * 3), np.uint8)
        depth = np.zeros((len(poses), P), np.float32)
        pc.stream_stats(reset=True)
        pc.render_trajectory(gpu.STAGE_FILTERED, poses, color, depth)
        tensor = pc.read("tensor", np.uint16, P * 5)
        passes, streamed = pc.stream_stats(reset=False)
        frames, visible, n_chunks = pc.cull_stats(reset=True)
        pc.close()
        for i in range(len(poses)):
            assert np.array_equal(color[i], want[i][0]) and np.array_equal(depth[i].view(np.uint32), want[i][1]), f"{opts}: frame {i}"
        assert np.array_equal(tensor, want[-1][2]), f"{opts}: tensor of the last frame"
        assert frames == len(poses) and 0 < visible <= frames * n_chunks
        if opts.get("fuse", 2):
            # one pass per frame plus the last frame's blend; consecutive poses share nearly all their chunks
            assert passes == len(poses) + 1
            assert visible <= streamed < 1.35 * visible, (visible, streamed)
        else:
            assert passes == 2 * len(poses) and streamed == 2 * visible


def test_fused_sequence_with_api_calls_in_between(gpu, cpu_oracle):
    """render_device back to back with reads, option changes, stage changes, a resolution change, a distorted camera and
    pose jumps in between: every read sees exactly the blocking call's frame."""
    a, b = scenes.CASES["c1_640x480"], scenes.CASES["small_176x104"]
    rec = cloud_of(cpu_oracle, a)
    rng = np.random.default_rng(5)
    traj = _trajectory(gpu, 300, scenes.HALL_SMALL)
    jump = gpu.look_at_w2c((4.0, 3.0, 1.5), (-1.0, -0.2, 0.0))
    steps = []
    for i in range(40):
        case = b if 14 <= i < 20 else a
        E = jump if i % 9 == 4 else traj[(3 * i) % 300]
        filtered = (i % 5) != 3
        dist = [-0.05, 0.01, 0.0005, -0.0005, 0.0] if 26 <= i < 31 else None
        steps.append((case, E, filtered, dist, rng.integers(0, 4)))
    pc = gpu.ProjectCloud.from_packed(rec, apply_distortion=True)
    ref = gpu.ProjectCloud.from_packed(rec, apply_distortion=True)
    pc.set_option("fuse", 2)
    assert pc.get_option("pipeline") == 1
    saw_pending = False
    for i, (case, E, filtered, dist, action) in enumerate(steps):
        P = case.W * case.H
        calib = calib_of(gpu, case)
        if dist:
            calib.setDistortionParameters(dist)
        pc.set_camera(calib, E)
        pc.render_device(gpu.STAGE_FILTERED if filtered else gpu.STAGE_RGBD)
        saw_pending |= pc.get_option("pending") == 1
        color, depth = np.zeros(P * 3, np.uint8), np.zeros(P, np.float32)
        fn = ref.computeFilteredRGBD if filtered else ref.computeRGBD
        assert fn(calib, E, color, depth) == 1
        if action == 0:
            continue                                   # no read: the next frame's z-min shares a pass with this frame's blend
        if action == 1:
            pc.sync()
        elif action == 2:
            pc.set_option("ring_dynamic", int(rng.choice([0, 1, 8])))
        got_image, got_depth = pc.read("image", np.uint8, P * 3), pc.read("zbuf", np.uint32, P)
        assert pc.get_option("pending") == 0
        assert np.array_equal(got_image, color) and np.array_equal(got_depth, depth.view(np.uint32)), f"step {i}"
        if filtered:
            assert np.array_equal(pc.read("tensor", np.uint16, P * 5), ref.read("tensor", np.uint16, P * 5)), f"step {i} tensor"
        assert np.array_equal(pc.read("accum", np.uint32, P * 4), ref.read("accum", np.uint32, P * 4)), f"step {i} accum"
    assert saw_pending
    pc.close()
    ref.close()


def test_fused_sequence_of_disjoint_views(gpu, cpu_oracle):
    """Consecutive poses that share no chunk (opposite directions, outside views): the union list is the two lists side
    by side, tiles carry one flag each."""
    case = scenes.CASES["c1_640x480"]
    rec = cloud_of(cpu_oracle, case)
    calib = calib_of(gpu, case)
    P = case.W * case.H
    poses = []
    for i in range(12):
        d = (1.0, 0.1 * i, 0.0) if i % 2 == 0 else (-1.0, -0.1 * i, 0.05)
        poses.append(gpu.look_at_w2c((4.0, 3.0, 1.5), d))
    poses.append(gpu.look_at_w2c((-30.0, 3.0, 1.5), (-1.0, 0.0, 0.0)))      # sees nothing at all
    poses.append(gpu.look_at_w2c((4.0, 3.0, 1.5), (0.0, 1.0, 0.0)))
    poses = np.stack(poses)
    want = _blocking(gpu, rec, calib, poses)
    pc = gpu.ProjectCloud.from_packed(rec)
    pc.set_option("fuse", 2)
    pc.set_camera(calib)
    color = np.zeros((len(poses), P * 3), np.uint8)
    depth = np.zeros((len(poses), P), np.float32)
    pc.render_trajectory(gpu.STAGE_FILTERED, poses, color, depth)
    passes, streamed = pc.stream_stats(reset=False)
    frames, visible, _ = pc.cull_stats(reset=True)
    pc.close()
    for i in range(len(poses)):
        assert np.array_equal(color[i], want[i][0]) and np.array_equal(depth[i].view(np.uint32), want[i][1]), f"frame {i}"
    assert frames == len(poses) and passes == len(poses) + 1
    assert streamed > 1.6 * visible          # hardly anything is shared between these views


@pytest.mark.parametrize("fixup_launches", [1, 2, 3])
def test_fused_sequence_float_sum_overflow(gpu, cpu_oracle, fixup_launches):
    """A pixel with > 65 793 accepted points inside a fused sequence: the gated exact re-run on the image stream (one
    thread-block cluster, a small cooperative grid, or three gated launches) redoes that frame's colour sums from the
    two-camera list; later frames start with integer sums."""
    W, H, heavy = 64, 48, 70_000
    m = np.array([32, 0, 31.5, 0, 0, 32, 23.5, 0, 0, 0, 1, 0, 0, 0, 0, 1], np.float32)
    rng = np.random.default_rng(9)
    n = heavy + 5000
    xyz = rng.uniform(-1.0, 1.0, (n, 3)).astype(np.float32)
    xyz[:, 2] = rng.uniform(1.5, 3.0, n).astype(np.float32)
    xyz[:heavy] = np.array([0.013, 0.009, 1.0], np.float32)
    bgr = rng.integers(0, 256, (n, 3), dtype=np.uint8)
    pc = gpu.ProjectCloud.from_packed(gpu.pack_records(xyz, bgr))
    pc.set_option("fuse", 2)
    pc.set_option("fixup_launches", fixup_launches)
    c = gpu.CameraCalibration()
    c.setWidth(W)
    c.setHeight(H)
    pc.set_camera(c)
    pc.set_cam_proj_raw(m)
    frames = []
    for i in range(6):
        pc.render_device(gpu.STAGE_FILTERED)
        if i in (2, 5):
            frames.append((pc.read("image", np.uint8, W * H * 3), pc.read("zbuf", np.uint32, W * H), pc.read("tensor", np.uint16, W * H * 5),
                           pc.read("accum", np.uint32, W * H * 4)))
    assert pc.get_option("int_sum_frames") > 0        # the note of the exact re-run reached the host
    tap, resident = pc.project_points(), pc.download_cloud()
    pc.close()
    gold = cpu_oracle.render(tap[0], tap[1], scenes.bgra_of(resident), W, H, filtered=True)
    for image, zbuf, tensor, accum in frames:
        assert np.array_equal(image, gold["image"]) and np.array_equal(zbuf, gold["zbuf"]) and np.array_equal(tensor, gold["tensor"])
        assert np.array_equal(accum, gold["accum"]) and accum.reshape(-1, 4)[:, 3].max() == heavy


def test_fused_sequence_replaced_cloud_and_reuse(gpu, cpu_oracle):
    """Uploading another cloud while a frame's blend is outstanding drops that frame; the renderer keeps working."""
    case = scenes.CASES["small_160x96"]
    rec = cloud_of(cpu_oracle, case)
    calib = calib_of(gpu, case)
    P = case.W * case.H
    pc = gpu.ProjectCloud.from_packed(rec)
    pc.set_option("fuse", 2)
    pc.set_camera(calib, case.poses[0])
    assert pc.get_option("fuse_active") == 1
    pc.render_device(gpu.STAGE_FILTERED)
    assert pc.get_option("pending") == 1
    other = cpu_oracle.synth_packed(99, 30_000, 0, 30_000, case.hall, case.n_boxes)
    pc._check(pc._lib.rtr_upload_cloud_packed16(pc._h, other.ctypes.data, len(other)))
    assert pc.get_option("pending") == 0
    pc.render_device(gpu.STAGE_FILTERED)
    pc.render_device(gpu.STAGE_FILTERED)
    got = pc.read("image", np.uint8, P * 3), pc.read("zbuf", np.uint32, P)
    pc.close()
    want = _blocking(gpu, other, calib, [case.poses[0]])[0]
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])


@pytest.mark.parametrize("bands", [2, 3, 8])
def test_band_ordered_lists_give_identical_frames(gpu, cpu_oracle, bands):
    """Option bands (csrc/rtr_cull.cu band_append / band_compact): the classification files the visible chunks under
    screen bands and its last CTA copies the segments, band after band, into the list.  The list must stay a permutation
    of the unordered one — a lost chunk would drop points, a doubled one would double colour sums — so every frame is
    byte-identical and the streamed-chunk counts are the same, through the fused pass, the two-pass sequence, the
    blocking calls (clear_classify_kernel) and the 64-bit-key frame."""
    case = scenes.CASES["c3_1920x1080"]
    rec = cloud_of(cpu_oracle, case)
    calib = calib_of(gpu, case)
    P = case.W * case.H
    poses = _trajectory(gpu, 1000)[300:316]
    want = _blocking(gpu, rec, calib, poses)
    got_blocking = _blocking(gpu, rec, calib, poses[:4], options={"bands": bands})
    for i in range(4):
        assert all(np.array_equal(a, b) for a, b in zip(got_blocking[i], want[i])), f"blocking call, frame {i}"
    counts = {}
    for fuse in (2, 0):
        for b in (1, bands):
            pc = gpu.ProjectCloud.from_packed(rec)
            pc.set_option("fuse", fuse)
            pc.set_option("bands", b)
            pc.set_camera(calib)
            assert pc.get_option("bands_active") == b
            color = np.zeros((len(poses), P * 3), np.uint8)
            depth = np.zeros((len(poses), P), np.float32)
            pc.stream_stats(reset=True)
            pc.render_trajectory(gpu.STAGE_FILTERED, poses, color, depth)
            tensor, accum = pc.read("tensor", np.uint16, P * 5), pc.read("accum", np.uint32, P * 4)
            counts[(fuse, b)] = (pc.stream_stats(reset=False), pc.cull_stats(reset=True))
            pc.close()
            for i in range(len(poses)):
                assert np.array_equal(color[i], want[i][0]) and np.array_equal(depth[i].view(np.uint32), want[i][1]), f"fuse {fuse}, bands {b}: frame {i}"
            assert np.array_equal(tensor, want[-1][2])
            if b == 1:
                accum_want = accum
            else:
                assert np.array_equal(accum, accum_want)
        assert counts[(fuse, 1)] == counts[(fuse, bands)]
    # 64-bit keys (depth bits << 32 | point index): clear_classify without the colour-sum clear
    frames = []
    for b in (1, bands):
        pc = gpu.ProjectCloud.from_packed(rec)
        pc.set_option("key64", 1)
        pc.set_option("bands", b)
        color, depth = np.zeros(P * 3, np.uint8), np.zeros(P, np.float32)
        assert pc.computeFilteredRGBD(calib, poses[5], color, depth) == 1
        frames.append((color, depth.view(np.uint32).copy()))
        pc.close()
    assert np.array_equal(frames[0][0], frames[1][0]) and np.array_equal(frames[0][1], frames[1][1])


def test_bands_follow_the_frame_size(gpu, cpu_oracle):
    """bands = 0 (default): list order while z-buffer + colour sums of a frame fit the 126 MB L2 (1080p: 50 MB), screen bands
    beyond (4K: 199 MB -> 8 bands); and the 4K frame through band-ordered lists equals the golden of the reference."""
    small, big = scenes.CASES["c3_1920x1080"], scenes.CASES["c5_3840x2160"]
    pc = gpu.ProjectCloud.from_packed(cloud_of(cpu_oracle, big))
    assert pc.get_option("bands") == 0
    pc.set_camera(calib_of(gpu, small))
    assert pc.get_option("bands_active") == 1
    calib = calib_of(gpu, big)
    pc.set_camera(calib)
    assert pc.get_option("bands_active") == 8
    P = big.W * big.H
    out = []
    for b in (0, 1):
        pc.set_option("bands", b)
        color, depth = np.zeros(P * 3, np.uint8), np.zeros(P, np.float32)
        assert pc.computeFilteredRGBD(calib, big.poses[0], color, depth) == 1
        out.append((color, depth.view(np.uint32).copy(), pc.read("tensor", np.uint16, P * 5)))
    pc.close()
    assert all(np.array_equal(a, b) for a, b in zip(out[0], out[1]))
    with pytest.raises(Exception):
        gpu.ProjectCloud.from_packed(cloud_of(cpu_oracle, small)).set_option("bands", 9)


def test_fused_sequence_at_3840x2160_with_band_ordered_lists(gpu, cpu_oracle):
    """Config 5's 4K frames as a fused sequence (band-ordered lists by default there) against the blocking calls with
    the list left in cloud order."""
    case = scenes.CASES["c5_3840x2160"]
    rec = cloud_of(cpu_oracle, case)
    calib = calib_of(gpu, case)
    P = case.W * case.H
    poses = _trajectory(gpu, 1000)[700:706]
    want = _blocking(gpu, rec, calib, poses, options={"bands": 1})
    pc = gpu.ProjectCloud.from_packed(rec)
    pc.set_option("fuse", 2)
    pc.set_camera(calib, poses[0])
    assert pc.get_option("bands_active") == 8 and pc.get_option("fuse_active") == 1
    color = np.zeros((len(poses), P * 3), np.uint8)
    depth = np.zeros((len(poses), P), np.float32)
    pc.render_trajectory(gpu.STAGE_FILTERED, poses, color, depth)
    tensor = pc.read("tensor", np.uint16, P * 5)
    pc.close()
    for i in range(len(poses)):
        assert np.array_equal(color[i], want[i][0]) and np.array_equal(depth[i].view(np.uint32), want[i][1]), f"frame {i}"
    assert np.array_equal(tensor, want[-1][2])


def test_two_pass_sequence_with_three_frames_in_flight(gpu, cpu_oracle):
    """Option pipeline_depth = 3 (default): whole frames of a two-pass (non-fused) sequence rotate through three frame
    sets on three streams (2: two sets / streams, round 1's pipeline).  The trajectory call (images to host every frame) and render_device back to back with reads in between must
    give the blocking call's frames; switching the depth mid-sequence must too."""
    case = scenes.CASES["c2_1280x720"]
    rec = cloud_of(cpu_oracle, case)
    calib = calib_of(gpu, case)
    P = case.W * case.H
    poses = _trajectory(gpu, 300)[40:61]
    want = _blocking(gpu, rec, calib, poses)
    pc = gpu.ProjectCloud.from_packed(rec)
    pc.set_option("fuse", 0)
    assert pc.get_option("pipeline_depth") == 3 and pc.get_option("fuse_active") == 0     # the default
    pc.set_camera(calib)
    color = np.zeros((len(poses), P * 3), np.uint8)
    depth = np.zeros((len(poses), P), np.float32)
    pc.render_trajectory(gpu.STAGE_FILTERED, poses, color, depth)
    for i in range(len(poses)):
        assert np.array_equal(color[i], want[i][0]) and np.array_equal(depth[i].view(np.uint32), want[i][1]), f"trajectory frame {i}"
    assert np.array_equal(pc.read("tensor", np.uint16, P * 5), want[-1][2])
    for i, E in enumerate(poses):
        if i == 13:
            pc.set_option("pipeline_depth", 2)        # back to two sets while frames of the third are in flight
        pc.set_camera(calib, E)
        pc.render_device(gpu.STAGE_FILTERED)
        if i % 4 == 3 or i == len(poses) - 1:
            assert np.array_equal(pc.read("image", np.uint8, P * 3), want[i][0]), f"device frame {i}"
            assert np.array_equal(pc.read("zbuf", np.uint32, P), want[i][1]), f"device frame {i}"
            assert np.array_equal(pc.read("tensor", np.uint16, P * 5), want[i][2]), f"device frame {i}"
    pc.close()
