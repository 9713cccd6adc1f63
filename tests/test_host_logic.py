"""Host-side logic of the package (no GPU, no compute through the library)."""
import numpy as np
import pytest


@pytest.mark.parametrize("n,world", [(1000, 1), (1000, 8), (1003, 8), (5, 8), (0, 4), (300, 7)])
def test_shards_partition_the_range(pkg, n, world):
    seen = []
    for r in range(world):
        fr = pkg.shard_frames(n, r, world)
        first, count = pkg.shard_points(n, r, world)
        assert (fr.start, len(fr)) == (first, count)
        seen += list(fr)
    assert seen == list(range(n))
    sizes = [len(pkg.shard_frames(n, r, world)) for r in range(world)]
    assert max(sizes) - min(sizes) <= 1


def test_look_at_and_trajectory_are_rigid_world_to_camera(pkg):
    T = pkg.trajectory_w2c(12, center=(6.0, 5.0, 1.5), radius=2.0)
    assert T.shape == (12, 4, 4)
    for i, E in enumerate(T):
        R = E[:3, :3]
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-12) and np.isclose(np.linalg.det(R), 1.0)
        a = 2 * np.pi * i / 12
        eye = np.array([6.0 + 2 * np.cos(a), 5.0 + 2 * np.sin(a), 1.5, 1.0])
        assert np.allclose(E @ eye, [0, 0, 0, 1], atol=1e-12)          # the camera centre maps to the origin
        ahead = eye[:3] + np.linalg.inv(E)[:3, 2]                       # one metre along the optical axis
        assert np.allclose(E @ np.append(ahead, 1.0), [0, 0, 1, 1], atol=1e-12)
    assert np.allclose(T[0], pkg.trajectory_w2c(12, center=(6.0, 5.0, 1.5), radius=2.0)[0])


def test_pack_records_layout(pkg):
    xyz = np.array([[1.5, -2.0, 3.25], [0, 0, 0]], np.float32)
    bgr = np.array([[1, 2, 3], [255, 128, 0]], np.uint8)
    rec = pkg.pack_records(xyz, bgr)
    assert rec.dtype == np.float32 and rec.shape == (2, 4) and rec.nbytes == 32
    assert np.array_equal(rec[:, :3], xyz)
    assert rec[:, 3].view(np.uint32).tolist() == [0xFF030201, 0xFF0080FF]   # b | g<<8 | r<<16 | 255<<24 (Octreegrid.h:176)


def test_camera_calibration_mirror(pkg):
    c = pkg.CameraCalibration()
    assert (c.getWidth(), c.getHeight()) == (640, 480)                        # CameraCalibration.cpp:7-8
    assert c.loadCalibration(525.0, 526.0, 319.5, 239.5, [0.1, 0.2], 1280, 720)
    assert c.getIntrinsicsMatrix().tolist() == [[525.0, 0, 319.5], [0, 526.0, 239.5], [0, 0, 1]]
    assert c.getDistortionParameters() == [0.1, 0.2, 0.0, 0.0, 0.0]
    assert (c.getFocalLengthX(), c.getFocalLengthY(), c.getPrincipalPointX(), c.getPrincipalPointY()) == (525.0, 526.0, 319.5, 239.5)
    assert (c.getWidth(), c.getHeight()) == (1280, 720)


def test_device_buffers_struct_matches_header(pkg):
    import ctypes as C
    # rtr_device_buffers: 6 pointers + 5 + 4 pointers, 2 ints, 4 x 5 ints, u64, pointer
    assert C.sizeof(pkg.DeviceBuffers) == 15 * 8 + 2 * 4 + 20 * 4 + 8 + 8


def test_distortion_cull_radius_bounds_every_visible_point(pkg):
    """rtr_host_distortion_bounds (host code, what chunk culling under lens distortion rests on): no point whose
    distorted projection (OpenCV model, float64) rounds to a pixel of the image may lie beyond r*."""
    import ctypes as C
    lib = pkg.load_library()
    rng = np.random.default_rng(5)
    dists = [[-0.05, 0.01, 0.0005, -0.0005, 0.0], [0.2, 0.05, 0.0, 0.0, 0.01], [-0.3, 0.1, 0.0, 0.0, -0.01],
             [0.0, 0.0, 0.02, -0.03, 0.0], [-0.2, 0.03, 0.01, 0.01, 0.001], [1e-9, 0.0, 0.0, 0.0, 0.0], [0.5, 0.0, 0.05, 0.05, 0.0]]
    effective = 0
    for it in range(40):
        W, H = int(rng.choice([320, 640, 1280, 1920])), int(rng.choice([208, 480, 720, 1080]))
        f = float(rng.uniform(0.3, 1.5) * W)
        K = np.array([f, float(rng.uniform(-2, 2)), rng.uniform(0.3 * W, 0.7 * W), 0, f * rng.uniform(0.9, 1.1), rng.uniform(0.3 * H, 0.7 * H), 0, 0, 1], np.float64)
        d = np.array(dists[it % len(dists)], np.float64)
        r2max, rstar = C.c_double(0), C.c_double(0)
        assert lib.rtr_host_distortion_bounds(W, H, K.ctypes.data_as(pkg._dp), d.ctypes.data_as(pkg._dp), C.byref(r2max), C.byref(rstar)) == pkg.RTR_OK
        rmax = np.sqrt(r2max.value)
        rr = np.linspace(0, rmax, 1500)[:, None]
        th = np.linspace(0, 2 * np.pi, 720, endpoint=False)[None, :]
        x, y = rr * np.cos(th), rr * np.sin(th)
        r2 = x * x + y * y
        radial = 1 + d[0] * r2 + d[1] * r2 ** 2 + d[4] * r2 ** 3
        xd = x * radial + 2 * d[2] * x * y + d[3] * (r2 + 2 * x * x)
        yd = y * radial + d[2] * (r2 + 2 * y * y) + 2 * d[3] * x * y
        u, v = np.rint(K[0] * xd + K[1] * yd + K[2]), np.rint(K[4] * yd + K[5])
        vis = (u >= 0) & (u < W) & (v >= 0) & (v < H)
        assert vis.any()
        far = np.sqrt(r2[vis]).max()
        assert rstar.value == 0 or far <= rstar.value, (it, far, rstar.value)
        effective += 0 < rstar.value < 0.8 * rmax
    assert effective >= 30      # the bound is tight enough to cull for most cameras


def test_ring_stream_all_order_is_a_permutation(pkg):
    """Tile t of a stream-all ring pass reads chunk (t * m) mod n_chunks (rtr_point_ring.cu): m must be coprime with
    n_chunks for every cloud size, and spread consecutive tiles over the cloud."""
    import math
    lib = pkg.load_library()
    for n_points in [1, 1023, 1024, 1025, 2048, 3 * 1024, 5000, 100_000, 1_000_003, 20_000_000, 100_000_000, 2 ** 32 - 1]:
        n_chunks = (n_points + 1023) // 1024
        m = lib.rtr_host_ring_stride(n_points)
        if n_chunks == 1:
            continue
        assert 0 < m < n_chunks and math.gcd(m, n_chunks) == 1, (n_points, n_chunks, m)
        if n_chunks <= 4096:
            assert sorted((t * m) % n_chunks for t in range(n_chunks)) == list(range(n_chunks))
        if n_chunks >= 1000:   # golden-ratio stride: any 64 consecutive tiles land in at least 32 different 64ths of the cloud
            assert len({((t * m) % n_chunks) * 64 // n_chunks for t in range(64)}) >= 32


def test_ring_tile_claims_cover_every_tile_once(pkg):
    """rtr_host_ring_claim is the arithmetic the ring kernels' list passes use to hand out tiles (csrc/rtr_kernels.h:
    ring_queue_of / ring_claimed_tile).  Replay the protocol of csrc/rtr_point_ring.cu:ring_walk on the CPU with consumer
    groups advancing in random order: a CTA's first `stages` tiles are fixed; every refill streams the tile claimed from
    the group's queue one iteration earlier; a group stops at the first stage that got no tile.  Every tile of the launch
    must be consumed exactly once and every group must stop, for any grid / queue count / tile count."""
    import ctypes as C
    import random
    lib = pkg.load_library()
    u32 = C.c_uint32

    def claim_of(grid, nq, block, group, claim):
        q, t, st, gr = u32(), u32(), u32(), u32()
        assert lib.rtr_host_ring_claim(grid, nq, block, group, claim, C.byref(q), C.byref(t), C.byref(st), C.byref(gr)) == 1
        return q.value, t.value, st.value, gr.value

    _, _, stages, groups = claim_of(1, 1, 0, 0, 0)
    assert stages % groups == 0
    assert lib.rtr_host_ring_claim(4, 0, 0, 0, 0, C.byref(u32()), C.byref(u32()), None, None) < 0        # no queue
    assert lib.rtr_host_ring_claim(4, 1, 0, groups, 0, C.byref(u32()), C.byref(u32()), None, None) < 0   # no such group
    rng = random.Random(5)
    for grid, nq, n_tiles in [(4, 1, 0), (4, 1, 3), (4, 3, 24), (4, 3, 25), (6, 8, 500), (296, 8, 12535), (296, 64, 1777), (148, 5, 4000), (3, 64, 200)]:
        counters = [0] * nq
        seen = [0] * n_tiles
        NO = None
        state = []   # per group: ring (stage -> tile or NO), pending claim, next k
        for block in range(grid):
            for g in range(groups):
                ring = {}
                for k in range(g, stages, groups):          # the fixed first ring-full
                    t = block + k * grid
                    ring[k % stages] = t if t < n_tiles else NO
                q = claim_of(grid, nq, block, g, 0)[0]
                c = counters[q]; counters[q] += 1            # the claim made before the loop
                state.append(dict(block=block, g=g, q=q, ring=ring, claim=c, k=g, done=False))
        live = list(range(len(state)))
        steps = 0
        while live:
            i = rng.choice(live)
            s = state[i]
            stage = s["k"] % stages
            # the refill this iteration will issue: the tile claimed one iteration ago
            t_refill = NO
            if s["claim"] is not NO and s["claim"] < n_tiles:
                t = claim_of(grid, nq, s["block"], s["g"], s["claim"])[1]
                if t < n_tiles:
                    t_refill = t
            if t_refill is not NO:
                s["claim"] = counters[s["q"]]; counters[s["q"]] += 1
            else:
                s["claim"] = NO
            tile = s["ring"][stage]
            if tile is NO:                                   # end mark: the group stops
                s["done"] = True
                live.remove(i)
                continue
            seen[tile] += 1
            s["ring"][stage] = t_refill
            s["k"] += groups
            steps += 1
            assert steps <= n_tiles + 1
        assert all(v == 1 for v in seen), (grid, nq, n_tiles, [j for j, v in enumerate(seen) if v != 1][:5])
        assert all(s["done"] for s in state)


def test_bench_pose_schedule_covers_the_loop_in_contiguous_arcs():
    """bench.py deals the trajectory in arcs of consecutive frames (what fused sequences rely on), spreads the arcs of all
    ranks evenly over the loop, and tiles the whole loop when the run is as long as the trajectory."""
    import bench
    n_poses = 1000
    # a 25-frame run on one GPU: four arcs of consecutive poses, a quarter of the loop apart
    s = bench.pose_schedule(25, n_poses, 1, 0)
    assert len(s) == 25
    jumps = [i for i in range(1, 25) if s[i] != (s[i - 1] + 1) % n_poses]
    assert len(jumps) == 3 and [s[j] for j in jumps] == [250, 500, 750]
    # the full loop: every pose exactly once
    assert sorted(bench.pose_schedule(1000, n_poses, 1, 0)) == list(range(1000))
    # 8 ranks: arcs of different ranks never overlap, and together they sample every eighth of the loop
    seen = set()
    for r in range(8):
        sr = bench.pose_schedule(25, n_poses, 8, r)
        assert not (seen & set(sr))
        seen |= set(sr)
    assert {p * 8 // n_poses for p in seen} == set(range(8))
    # the full trajectory dealt to 8 ranks in contiguous runs: a partition of the loop
    every = sorted(p for r in range(8) for p in bench.pose_schedule(125, n_poses, 8, r))
    assert len(set(every)) >= 970   # (arcs are whole frames: neighbouring arcs can overlap by a pose at the seams)


def test_band_ordered_list_copy_is_the_concatenation_of_the_segments(pkg):
    """rtr_host_band_compact replays, with the kernel's own layout / locate arithmetic (csrc/rtr_kernels.h band_layout,
    band_locate; csrc/rtr_cull.cu band_compact), the flat copy loop that turns the per-band segments of a band-ordered
    classification into the list the point passes walk: for any counts (zero, not multiples of 4, a single band) and
    any thread count the list must be the segments' valid entries, band after band, and nothing else is written."""
    import ctypes as C
    lib = pkg.load_library()
    rng = np.random.default_rng(11)
    u32p = C.POINTER(C.c_uint32)
    for it in range(200):
        n_bands = int(rng.integers(1, 9))
        cap = int(rng.integers(1, 300)) * 4
        counts = np.zeros(8, np.uint32)
        mode = it % 4
        for b in range(n_bands):
            counts[b] = 0 if (mode == 1 and rng.random() < 0.5) else (cap if mode == 2 else int(rng.integers(0, cap + 1)))
        if mode == 3:
            counts[n_bands:] = rng.integers(1, cap + 1, 8 - n_bands)      # stale counters of unused bands must be ignored
        scratch = rng.integers(0, 2 ** 32, 8 * cap, dtype=np.uint64).astype(np.uint32)
        want = np.concatenate([scratch[b * cap: b * cap + int(counts[b])] for b in range(n_bands)] + [np.zeros(0, np.uint32)])
        guard = np.uint32(0xDEADBEEF)
        out = np.full(len(want) + 16, guard, np.uint32)
        n_out = C.c_uint32(0)
        threads = int(rng.choice([1, 3, 32, 128, 256]))
        rc = lib.rtr_host_band_compact(scratch.ctypes.data_as(u32p), cap, counts.ctypes.data_as(u32p), n_bands, threads,
                                       out.ctypes.data_as(u32p), C.byref(n_out))
        assert rc == 1 and n_out.value == len(want), (it, rc, n_out.value, len(want))
        assert np.array_equal(out[:len(want)], want), it
        assert (out[len(want):] == guard).all(), it
    bad = np.zeros(8, np.uint32)
    assert lib.rtr_host_band_compact(bad.ctypes.data_as(u32p), 6, bad.ctypes.data_as(u32p), 2, 128, bad.ctypes.data_as(u32p), C.byref(n_out)) < 0
    assert lib.rtr_host_band_compact(bad.ctypes.data_as(u32p), 8, bad.ctypes.data_as(u32p), 9, 128, bad.ctypes.data_as(u32p), C.byref(n_out)) < 0
