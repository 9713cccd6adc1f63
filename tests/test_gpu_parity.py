"""Parity of the sm_100a CUDA path (through the C ABI) with
  (1) the CPU oracle fed the GPU's own per-point projection (hybrid golden, SURVEY.md §8 c) — every
      stage tap, bit for bit;
  (2) the committed golden vectors = outputs of the unmodified reference on a B200 (sha256);
  (3) the live reference (oracle/_ref), when its library travelled to this box — 0 differing
      pixels allowed (bar from north_star: bit-exact framebuffer, depth, mask/tensor).
"""
import os

import numpy as np
import pytest

import scenes
from conftest import ROOT

pytestmark = pytest.mark.gpu

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
_clouds = {}


def cloud_of(cpu_oracle, case):
    key = (case.seed, case.n, case.hall, case.n_boxes)
    if key not in _clouds:
        _clouds[key] = cpu_oracle.synth_packed(case.seed, case.n, 0, case.n, case.hall, case.n_boxes)
    return _clouds[key]


def calib_of(pkg, case):
    c = pkg.CameraCalibration()
    c.setIntrinsicsMatrix(case.K)
    c.setWidth(case.W)
    c.setHeight(case.H)
    return c


def render_mine(pkg, case, rec, keep_masks=False, options=None, with_taps=True):
    """Per pose: computeRGBD then computeFilteredRGBD on one renderer (like one reference object).
    Returns (frames, taps, extras)."""
    pc = pkg.ProjectCloud.from_packed(rec)
    resident = pc.download_cloud() if with_taps else None   # uploads are Morton-sorted: taps follow THIS order
    for k, v in (options or {}).items():
        pc.set_option(k, v)
    if keep_masks:
        pc.set_option("keep_masks", 1)
    W, H, P = case.W, case.H, case.W * case.H
    calib = calib_of(pkg, case)
    frames, taps, extras = [], [], []
    for E in case.poses:
        o = {}
        color, depth = np.zeros(P * 3, np.uint8), np.zeros(P, np.float32)
        assert pc.computeRGBD(calib, E, color, depth) == 1
        o["raw_depth_host"], o["raw_color_host"] = depth.view(np.uint32).copy(), color.copy()
        o["raw_zbuf"], o["raw_accum"], o["raw_image"] = pc.read("zbuf", np.uint32, P), pc.read("accum", np.uint32, P * 4), pc.read("image", np.uint8, P * 3)
        color, depth = np.zeros(P * 3, np.uint8), np.zeros(P, np.float32)
        assert pc.computeFilteredRGBD(calib, E, color, depth) == 1
        o["flt_depth_host"], o["flt_color_host"] = depth.view(np.uint32).copy(), color.copy()
        o["flt_tensor"], o["flt_minmax"] = pc.read("tensor", np.uint16, P * 5), pc.read("minmax", np.uint32, 2)
        frames.append(o)
        ex = {"cam_proj": pc.get_cam_proj().reshape(16), "records": resident}
        if keep_masks:
            b = pc.device_buffers()
            ex["levels"] = {i: pc.read(f"level{i}", np.float32, b.level_w[i] * b.level_h[i]) for i in range(1, 5)}
            ex["masks"] = {i: pc.read(f"mask{i}", np.uint8, b.up_w[i] * b.up_h[i]) for i in range(4)}
            ex["up"] = [(b.up_w[i], b.up_h[i]) for i in range(5)]
        extras.append(ex)
        if with_taps:
            taps.append(pc.project_points())
    pc.close()
    return frames, taps, extras


def assert_frames_equal(a, b, what):
    for fi, (fa, fb) in enumerate(zip(a, b)):
        for k in scenes.OUTPUT_KEYS:
            if not np.array_equal(fa[k], fb[k]):
                n = int((np.asarray(fa[k]) != np.asarray(fb[k])).sum())
                raise AssertionError(f"{what}: frame {fi} '{k}' differs in {n} of {fa[k].size} elements")


@pytest.mark.parametrize("name", list(scenes.CASES))
def test_cuda_vs_cpu_oracle_all_stages(gpu, cpu_oracle, name):
    case = scenes.CASES[name]
    rec = cloud_of(cpu_oracle, case)
    mine, taps, extras = render_mine(gpu, case, rec, keep_masks=True)
    assert (taps[0][0] >= 0).sum() > 1000
    # host matrix: K4*E as the reference's glm expression evaluates it
    for E, ex in zip(case.poses, extras):
        assert np.array_equal(ex["cam_proj"], cpu_oracle.cam_proj(case.K, E))
    resident = extras[0]["records"]
    assert np.array_equal(np.sort(resident.view(np.uint32).view([("", np.uint32)] * 4).ravel()),
                          np.sort(rec.view(np.uint32).view([("", np.uint32)] * 4).ravel())), "upload re-ordering lost or changed a record"
    gold = scenes.oracle_frames(cpu_oracle, case, resident, taps)
    assert_frames_equal(mine, gold, f"{name} CUDA vs CPU oracle")
    # pyramid levels and masks of the last frame (stage taps 4-5)
    f = gold[-1]
    flt = cpu_oracle.depth_filter(f["raw_zbuf"], f["raw_image"], case.W, case.H, taps=True)
    ex = extras[-1]
    for i in range(1, 5):
        n = ex["up"][i][0] * ex["up"][i][1] if case.W % 16 == 0 else flt["levels"][i].size
        a, b = ex["levels"][i][:n], flt["levels"][i][:n]
        # hole-filled levels can hold NaN (bilinear of FLT_MAX overflows to inf - inf); its sign/payload is
        # not defined by IEEE (x86 gives 0xFFC00000, the GPU 0x7FFFFFFF) and no comparison ever sees it
        same = (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))
        assert same.all(), f"level {i}"
    for i in range(4):
        assert np.array_equal(ex["masks"][i], flt["masks"][i]), f"mask {i}"


def load_golden(name):
    p = os.path.join(GOLDEN_DIR, name + ".npz")
    return np.load(p) if os.path.exists(p) else None


@pytest.mark.parametrize("name", list(scenes.CASES))
def test_cuda_vs_golden_reference_outputs(gpu, cpu_oracle, name):
    g = load_golden(name)
    if g is None:
        pytest.skip("golden not generated yet (tests/golden/make_golden.py)")
    case = scenes.CASES[name]
    mine, _, extras = render_mine(gpu, case, cloud_of(cpu_oracle, case), with_taps=False)
    for fi, f in enumerate(mine):
        assert np.array_equal(extras[fi]["cam_proj"], g[f"f{fi}_cam_proj"]), "camProj differs from the reference's glm product"
        for k in scenes.OUTPUT_KEYS:
            assert scenes.sha(f[k]) == bytes(g[f"f{fi}_{k}_sha"]).decode(), f"{name} frame {fi} {k}: differs from the reference"


def have_ref():
    import oracle
    return os.path.exists(oracle.REF_LIB)


@pytest.mark.parametrize("name", list(scenes.CASES))
def test_cuda_vs_live_reference(gpu, cpu_oracle, name):
    if not have_ref():
        pytest.skip("oracle/_ref/libref_rtrenderer.so not on this box")
    import oracle
    case = scenes.CASES[name]
    rec = cloud_of(cpu_oracle, case)
    xyz, bgr = scenes.split_records(rec)
    ref = oracle.RefOracle(xyz, bgr)
    W, H, P = case.W, case.H, case.W * case.H
    ref_frames = []
    for E in case.poses:
        o = {}
        rc, color, depth = ref.computeRGBD(W, H, case.K, E)
        assert rc == 1
        o["raw_depth_host"], o["raw_color_host"] = depth.view(np.uint32), color
        o["raw_zbuf"], o["raw_accum"], o["raw_image"] = ref.read("zbuf", P), ref.read("accum", P * 4), ref.read("image", P * 3)
        rc, color, depth = ref.computeFilteredRGBD(W, H, case.K, E)
        assert rc == 1
        o["flt_depth_host"], o["flt_color_host"] = depth.view(np.uint32), color
        o["flt_tensor"] = ref.read("tensor", P * 5)
        o["flt_minmax"] = np.array([ref.read("min", 1)[0], ref.read("max", 1)[0]], np.uint32)
        ref_frames.append(o)
    ref.close()
    mine, _, _ = render_mine(gpu, case, rec, with_taps=False)
    assert_frames_equal(mine, ref_frames, f"{name} CUDA vs live reference")


# ------------------------------------------------------------------ invariances / variants
# ring = 0: the per-thread LDG.128 kernels (rtr_point_kernels.cu); ring = 1 (default): the TMA-fed persistent kernels
# (rtr_point_ring.cu) over the visible-chunk list when chunk_cull = 1; ring = 2: also over every chunk (permuted order)
VARIANTS = [dict(zmin_variant=v, zmin_unroll=u, blend_variant=b, blend_unroll=u, chunk_cull=c, ring=r)
            for v, u, b, c, r in [(0, 1, 0, 0, 0), (1, 2, 2, 0, 0), (2, 4, 0, 0, 0), (3, 8, 2, 0, 0), (5, 4, 0, 0, 0), (7, 4, 2, 0, 0),
                                  (0, 4, 2, 1, 0), (1, 4, 0, 1, 0), (3, 4, 0, 1, 0), (7, 4, 2, 1, 0), (5, 4, 4, 1, 0), (5, 4, 6, 0, 0),
                                  (5, 2, 4, 0, 0),
                                  (0, 4, 0, 0, 2), (1, 4, 4, 0, 2), (5, 4, 4, 0, 2), (0, 4, 4, 1, 1), (1, 4, 0, 1, 1), (5, 4, 0, 1, 2)]]
# zmin_variant bit 6: shared-memory tile pre-reduction in the z-min ring pass (two-pass frames; list and stream-all)
VARIANTS += [dict(zmin_variant=64 | 5, ring=1, chunk_cull=1, fuse=0), dict(zmin_variant=64 | 5, ring=2, chunk_cull=0),
             dict(zmin_variant=64 | 5, ring=1, chunk_cull=1, fuse=0, ring_claim_min=0, ring_ctas=1)]
# ring_dynamic: how the list passes of the ring kernels hand out tiles — 0 round-robin, q > 0 claimed from q counters
VARIANTS += [dict(ring=1, chunk_cull=1, ring_dynamic=q, ring_claim_min=0) for q in (0, 1, 3, 8, 64)] + [dict(ring=1, chunk_cull=1, clear_lean=0)]
# ring_ctas = 1: one ring-kernel CTA per SM (half the grid)
VARIANTS += [dict(ring=1, chunk_cull=1, ring_ctas=1, ring_claim_min=0), dict(ring=2, chunk_cull=0, ring_ctas=1), dict(ring=1, chunk_cull=1, ring_ctas=1, ring_dynamic=0)]


@pytest.mark.parametrize("opts", VARIANTS)
def test_kernel_variants_give_identical_frames(gpu, cpu_oracle, opts):
    case = scenes.CASES["c1_640x480"]
    rec = cloud_of(cpu_oracle, case)
    base, _, _ = render_mine(gpu, case, rec, with_taps=False)
    other, _, _ = render_mine(gpu, case, rec, options=opts, with_taps=False)
    assert_frames_equal(other, base, f"variant {opts}")


def test_measurement_only_kernels_are_not_selectable(gpu, cpu_oracle):
    """Variant bits 8 / 16 / 32 (no RED issued, ATOMG builtin, no in-register merge) select kernels whose frames are wrong
    or slower by design; they exist only in -DRTR_EXPERIMENTS builds and a caller of the shipped library cannot reach them."""
    pc = gpu.ProjectCloud.from_packed(cloud_of(cpu_oracle, scenes.CASES["small_160x96"]))
    if pc.get_option("experiments") == 1:
        pc.close()
        pytest.skip("this is an RTR_EXPERIMENTS build")
    for key, value in (("zmin_variant", 8 | 5), ("zmin_variant", 16 | 5), ("zmin_variant", 32 | 5), ("zmin_variant", 128 | 5),
                       ("blend_variant", 32 | 4), ("blend_variant", 8), ("zmin_variant", -1)):
        with pytest.raises(gpu.RtrError) as e:
            pc.set_option(key, value)
        assert e.value.code == gpu.RTR_ERR_ARG
        assert pc.get_option(key) in (5, 4)                     # unchanged defaults
    with pytest.raises(gpu.RtrError):
        pc.set_option("index_base", (1 << 32) - 5)              # index_base + cloud size must fit the key's 32-bit index
    pc.close()


def test_generic_path_equals_fused_path(gpu, cpu_oracle):
    for name in ("small_176x104", "c3_1920x1080"):
        case = scenes.CASES[name]
        rec = cloud_of(cpu_oracle, case)
        base, _, _ = render_mine(gpu, case, rec, with_taps=False)
        other, _, _ = render_mine(gpu, case, rec, options={"force_generic": 1}, with_taps=False)
        assert_frames_equal(other, base, f"{name} generic vs fused")


def test_one_launch_up_pass_equals_per_level_launches(gpu, cpu_oracle):
    """up_fused_kernel (default when W % 16 == 0) recomputes halos in shared memory instead of running the four
    levels as four launches; every output must stay identical, incl. partial tiles (1080 rows -> 1072) and 4K."""
    for name in ("small_160x96", "small_176x104", "c1_640x480", "c2_1280x720", "c3_1920x1080", "c5_3840x2160"):
        case = scenes.CASES[name]
        rec = cloud_of(cpu_oracle, case)
        base, _, _ = render_mine(gpu, case, rec, with_taps=False)
        other, _, _ = render_mine(gpu, case, rec, options={"fused_up": 0}, with_taps=False)
        assert_frames_equal(other, base, f"{name} per-level vs one-launch up-pass")


def test_fast_divide_path_is_the_reference_divide(gpu, cpu_oracle):
    """The ring kernels issue MUFU.RCP + FMUL directly when no depth of the warp is denormal (rtr_common.cuh project4):
    on 200 M random bit patterns (NaN, inf, denormal, huge included) the quotient bits and the rounded pixel must equal
    __fdividef's, which is what the reference compiles."""
    pc = gpu.ProjectCloud.from_packed(cloud_of(cpu_oracle, scenes.CASES["small_160x96"]))
    assert pc.selftest_fast_divide(200_000_000, seed=3) == 0
    pc.close()


def test_point_order_does_not_matter(gpu, cpu_oracle):
    case = scenes.CASES["c1_640x480"]
    rec = cloud_of(cpu_oracle, case)
    base, _, _ = render_mine(gpu, case, rec, with_taps=False)
    perm = np.random.default_rng(7).permutation(len(rec))
    other, _, _ = render_mine(gpu, case, np.ascontiguousarray(rec[perm]), with_taps=False)
    assert_frames_equal(other, base, "shuffled cloud")


def test_upload_xyz_bgr_equals_packed(gpu, cpu_oracle):
    case = scenes.CASES["small_160x96"]
    rec = cloud_of(cpu_oracle, case)
    xyz, bgr = scenes.split_records(rec)
    pc = gpu.ProjectCloud(xyz, bgr, sort=False)
    assert pc.cloud_size == len(rec)
    assert np.array_equal(pc.download_cloud().view(np.uint32), rec.view(np.uint32))
    pc.close()


def test_device_synth_equals_oracle_synth(gpu, cpu_oracle):
    n = 300_000
    pc = gpu.ProjectCloud.synthetic(seed=99, n_total=n, hall=scenes.HALL_LARGE, n_boxes=12, sort=False)
    rec = cpu_oracle.synth_packed(99, n, 0, n, scenes.HALL_LARGE, 12)
    assert np.array_equal(pc.download_cloud().view(np.uint32), rec.view(np.uint32))
    pc.close()
    pc = gpu.ProjectCloud.synthetic(seed=99, n_total=n, first=1000, count=5000, hall=scenes.HALL_LARGE, n_boxes=12, sort=False)
    assert np.array_equal(pc.download_cloud().view(np.uint32), rec[1000:6000].view(np.uint32))
    pc.close()


def test_key64_depth_is_reference_depth_and_colour_is_nearest_point(gpu, cpu_oracle):
    case = scenes.CASES["c1_640x480"]
    rec = cloud_of(cpu_oracle, case)
    W, H, P = case.W, case.H, case.W * case.H
    base, taps, extras = render_mine(gpu, case, rec)
    rec = extras[0]["records"]                       # the key's point index is the index in the RESIDENT order
    pc = gpu.ProjectCloud.from_packed(rec, sort=False)
    pc.set_option("key64", 1)
    color, depth = np.zeros(P * 3, np.uint8), np.zeros(P, np.float32)
    assert pc.computeRGBD(calib_of(gpu, case), case.poses[0], color, depth) == 1
    pc.close()
    assert np.array_equal(depth.view(np.uint32), base[0]["raw_depth_host"])
    # nearest point with the lowest index wins (deterministic 64-bit key)
    pix, zb = taps[0]
    live = np.nonzero(pix >= 0)[0]
    key = (zb[live].astype(np.uint64) << np.uint64(32)) | live.astype(np.uint64)
    best = np.full(P, np.iinfo(np.uint64).max, np.uint64)
    np.minimum.at(best, pix[live], key)
    hit = best != np.iinfo(np.uint64).max
    idx = (best[hit] & np.uint64(0xFFFFFFFF)).astype(np.int64)
    exp = np.zeros((P, 3), np.uint8)
    c = scenes.bgra_of(rec)[idx]
    exp[hit] = np.stack([c & 0xFF, (c >> 8) & 0xFF, (c >> 16) & 0xFF], axis=1).astype(np.uint8)
    assert np.array_equal(color.reshape(P, 3), exp)


def test_trajectory_equals_frame_by_frame(gpu, cpu_oracle):
    case = scenes.CASES["c3_1920x1080"]
    rec = cloud_of(cpu_oracle, case)
    P = case.W * case.H
    poses = np.stack([case.poses[i % 2] for i in range(5)])
    pc = gpu.ProjectCloud.from_packed(rec)
    calib = calib_of(gpu, case)
    pc.set_camera(calib)
    color = np.zeros((5, P * 3), np.uint8)
    depth = np.zeros((5, P), np.float32)
    pc.render_trajectory(gpu.STAGE_FILTERED, poses, color, depth)
    pc.close()
    # each frame set is persistent; with only filtered frames the stale 1080p tail is all-zero
    one = gpu.ProjectCloud.from_packed(rec)
    for i in range(5):
        c1, d1 = np.zeros(P * 3, np.uint8), np.zeros(P, np.float32)
        one.computeFilteredRGBD(calib, poses[i], c1, d1)
        assert np.array_equal(c1, color[i]) and np.array_equal(d1.view(np.uint32), depth[i].view(np.uint32)), f"frame {i}"
    one.close()


# ------------------------------------------------------------------ edge cases
def _render_records(gpu, rec, W, H, m16, filtered=True):
    pc = gpu.ProjectCloud.from_packed(rec)
    c = gpu.CameraCalibration()
    c.setWidth(W)
    c.setHeight(H)
    pc.set_camera(c)
    pc.set_cam_proj_raw(m16)
    color, depth = np.zeros(W * H * 3, np.uint8), np.zeros(W * H, np.float32)
    fn = pc._lib.rtr_render_filtered if filtered else pc._lib.rtr_render_rgbd
    pc._check(fn(pc._h, color.ctypes.data, depth.ctypes.data))
    out = dict(color=color, depth=depth.view(np.uint32), tensor=pc.read("tensor", np.uint16, W * H * 5),
               accum=pc.read("accum", np.uint32, W * H * 4), minmax=pc.read("minmax", np.uint32, 2),
               records=pc.download_cloud())
    tap = pc.project_points()
    pc.close()
    return out, tap


def _oracle_records(cpu_oracle, rec, W, H, tap, filtered=True):
    return cpu_oracle.render(tap[0], tap[1], scenes.bgra_of(rec), W, H, filtered=filtered)


def test_adversarial_points(gpu, cpu_oracle):
    """NaN / inf / denormal / behind-camera / exactly-on-boundary / huge coordinates, many points per
    pixel, colour sums near the byte limits."""
    W, H = 64, 48
    m = np.array([50, 0, 31.5, 0, 0, 50, 23.5, 0, 0, 0, 1, 0, 0, 0, 0, 1], np.float32)
    rng = np.random.default_rng(3)
    n = 40_000
    xyz = rng.uniform(-1.5, 1.5, (n, 3)).astype(np.float32)
    xyz[:, 2] = rng.uniform(0.5, 3.0, n).astype(np.float32)
    special = np.array([[np.nan, 0, 1], [0, np.nan, 1], [0, 0, np.nan], [np.inf, 0, 1], [0, 0, np.inf], [-np.inf, 1, 1],
                        [0, 0, 0], [0, 0, -0.0], [0, 0, -1], [1e-41, 1e-41, 1e-41], [1e-39, 0, 1e-39], [3e38, 3e38, 1],
                        [3e38, 0, 3e38], [0, 0, 1e-45], [-0.63, -0.47, 1.0], [0.65, 0.49, 1.0], [-0.64, 0.0, 1.0],
                        [0.01, 0.01, 1.0], [0.01, 0.01, 1.0199], [0.01, 0.01, 1.02], [0.01, 0.01, 1.0201]], np.float32)
    xyz[:len(special)] = special
    xyz[1000:3000, :2] = 0.0          # 2000 points in one pixel, depths spread around the 2 cm window
    xyz[1000:3000, 2] = (1.0 + rng.uniform(0, 0.04, 2000)).astype(np.float32)
    bgr = rng.integers(0, 256, (n, 3), dtype=np.uint8)
    bgr[1000:2000] = 255
    rec = gpu.pack_records(xyz, bgr)
    for filtered in (False, True):
        out, tap = _render_records(gpu, rec, W, H, m, filtered)
        gold = _oracle_records(cpu_oracle, out["records"], W, H, tap, filtered)
        assert np.array_equal(out["depth"], gold["zbuf"])
        assert np.array_equal(out["accum"], gold["accum"])
        assert np.array_equal(out["color"], gold["image"])
        if filtered:
            assert np.array_equal(out["tensor"], gold["tensor"])
            assert np.array_equal(out["minmax"], gold["minmax"])


def test_nothing_in_frustum_and_single_point(gpu, cpu_oracle):
    W, H = 32, 16
    m = np.array([20, 0, 15.5, 0, 0, 20, 7.5, 0, 0, 0, 1, 0, 0, 0, 0, 1], np.float32)
    behind = gpu.pack_records(np.array([[0, 0, -1.0], [0.1, 0.1, -2.0]], np.float32), np.full((2, 3), 200, np.uint8))
    out, tap = _render_records(gpu, behind, W, H, m)
    assert (tap[0] == -1).all()
    assert (out["depth"] == np.float32(-1.0).view(np.uint32)).all() and not out["color"].any()
    assert out["minmax"].tolist() == [0xFFFFFFFF, 0]
    gold = _oracle_records(cpu_oracle, out["records"], W, H, tap)
    assert np.array_equal(out["tensor"], gold["tensor"]) and np.array_equal(out["depth"], gold["zbuf"])
    one = gpu.pack_records(np.array([[0.0, 0.0, 2.0]], np.float32), np.array([[10, 20, 30]], np.uint8))
    out, tap = _render_records(gpu, one, W, H, m, filtered=False)
    pix = 8 * W + 16  # rint(7.5) = 8, rint(15.5) = 16 (ties to even)
    assert tap[0][0] == pix
    assert out["depth"][pix] == np.float32(2.0).view(np.uint32) and out["color"].reshape(-1, 3)[pix].tolist() == [10, 20, 30]
    assert (np.delete(out["depth"], pix) == 0x7F7FFFFF).all()


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 255, 1023, 1024, 1025, 2047, 4099, 6 * 1024 * 3 + 1])
def test_cloud_sizes_around_chunk_and_ring_boundaries(gpu, cpu_oracle, n):
    """Clouds of 1 .. a few chunks (partial last chunk, fewer chunks than ring stages, fewer tiles than CTAs): the
    TMA-fed ring kernels (culled list and stream-all), the per-thread kernels and the CPU oracle must agree, in blend
    and in 64-bit-key mode."""
    case = scenes.CASES["small_160x96"]
    rec = cpu_oracle.synth_packed(77, 20_000, 0, 20_000, case.hall, case.n_boxes)[:n].copy()
    W, H, P = case.W, case.H, case.W * case.H
    calib, E = calib_of(gpu, case), case.poses[0]
    outs = {}
    for name, opts in (("ring_list", dict(ring=1)), ("ring_all", dict(ring=2, chunk_cull=0)), ("ldg_list", dict(ring=0)),
                       ("ldg_all", dict(ring=0, chunk_cull=0)), ("ring_key64", dict(ring=2, key64=1)), ("ldg_key64", dict(ring=0, key64=1)),
                       ("ring_list_rr", dict(ring=1, ring_dynamic=0)), ("ring_list_q1", dict(ring=1, ring_dynamic=1, ring_claim_min=0)),
                       ("ring_list_q8", dict(ring=1, ring_claim_min=0))):
        pc = gpu.ProjectCloud.from_packed(rec, sort=False)
        for k, v in opts.items():
            pc.set_option(k, v)
        color, depth = np.zeros(P * 3, np.uint8), np.zeros(P, np.float32)
        assert pc.computeFilteredRGBD(calib, E, color, depth) == 1
        outs[name] = (color, depth.view(np.uint32).copy(), pc.read("tensor", np.uint16, P * 5))
        if name == "ring_list":
            tap = pc.project_points()
        pc.close()
    gold = cpu_oracle.render(tap[0], tap[1], scenes.bgra_of(rec), W, H, filtered=True)
    for name in ("ring_list", "ring_all", "ldg_list", "ldg_all", "ring_list_rr", "ring_list_q1", "ring_list_q8"):
        assert np.array_equal(outs[name][0], gold["image"]) and np.array_equal(outs[name][1], gold["zbuf"]), name
        assert np.array_equal(outs[name][2], gold["tensor"]), name
    for a, b in zip(outs["ring_key64"], outs["ldg_key64"]):
        assert np.array_equal(a, b)
    assert np.array_equal(outs["ring_key64"][1], gold["zbuf"])      # key64's depth is the reference depth


def test_large_cloud_ring_equals_per_thread_kernels(gpu):
    """30 M points seen from a corner of the hall (most chunks visible: ~100 tiles per CTA, i.e. the ring's refill path
    runs deep) and from inside: the TMA-fed kernels, the per-thread kernels over the list and the stream-all kernels
    must produce identical frames, also with two frames in flight."""
    n, W, H = 30_000_000, 1280, 720
    P = W * H
    calib = gpu.CameraCalibration()
    calib.loadCalibration(500.0, 500.0, 639.5, 359.5, [0.0] * 5, W, H)
    poses = [gpu.look_at_w2c((0.3, 0.3, 2.7), (1.0, 0.8, -0.2)), gpu.look_at_w2c((6.0, 5.0, 1.5), (1.0, 0.3, 0.0))]
    pc = gpu.ProjectCloud.synthetic(seed=4242, n_total=n, hall=scenes.HALL_LARGE, n_boxes=12)
    digests = {}
    dyn = pc.get_option("ring_dynamic")
    assert dyn > 0   # the default claims tiles from counters
    for name, opts in (("ring", dict(ring=1, chunk_cull=1)), ("ring_rr", dict(ring_dynamic=0)), ("ring_q1", dict(ring_dynamic=1)),
                       ("ring_q64", dict(ring_dynamic=64)), ("ring_1cta", dict(ring_dynamic=dyn, ring_ctas=1)),
                       ("ldg_list", dict(ring=0, chunk_cull=1, ring_ctas=2)),
                       ("ldg_all", dict(ring=0, chunk_cull=0)), ("ring_all", dict(ring=2, chunk_cull=0))):
        for k, v in opts.items():
            pc.set_option(k, v)
        out = []
        for E in poses:
            color, depth = np.zeros(P * 3, np.uint8), np.zeros(P, np.float32)
            assert pc.computeFilteredRGBD(calib, E, color, depth) == 1
            out.append(scenes.sha(color) + scenes.sha(depth) + scenes.sha(pc.read("tensor", np.uint16, P * 5)))
        digests[name] = out
    fr, vis, nch = 0, 0, 0
    pc.set_option("ring", 1)
    pc.set_option("chunk_cull", 1)
    pc.cull_stats(reset=True)
    # the same two poses through the asynchronous trajectory call (frames alternate between two streams)
    color = np.zeros((4, P * 3), np.uint8)
    depth = np.zeros((4, P), np.float32)
    traj = np.ascontiguousarray(np.stack([poses[0], poses[1], poses[0], poses[1]]).reshape(-1, 16))
    pc._check(pc._lib.rtr_render_trajectory(pc._h, gpu.STAGE_FILTERED, traj.ctypes.data_as(gpu._dp), 4, color.ctypes.data, depth.ctypes.data))
    fr, vis, nch = pc.cull_stats(reset=True)
    pc.close()
    assert fr == 4 and vis / fr > 0.3 * nch          # the corner view really sees a large part of the cloud
    for name in ("ring_rr", "ring_q1", "ring_q64", "ring_1cta", "ldg_list", "ldg_all", "ring_all"):
        assert digests[name] == digests["ring"], name
    for i in range(4):
        assert scenes.sha(color[i]) + scenes.sha(depth[i]) == digests["ring"][i % 2][:128], f"trajectory frame {i}"


def test_error_behaviour(gpu):
    pc = gpu.ProjectCloud()
    calib = gpu.CameraCalibration()
    assert pc.computeRGBD(calib, np.eye(4), None, None) == -1          # project_cloud.cu:270-273
    with pytest.raises(gpu.RtrError) as e:                             # no cloud uploaded
        pc.computeRGBD(calib, np.eye(4), np.zeros(640 * 480 * 3, np.uint8), None)
    assert e.value.code == gpu.RTR_ERR_STATE
    with pytest.raises(gpu.RtrError):
        pc.set_option("no_such_option", 1)
    pc.close()


def test_resolution_change_reallocates(gpu, cpu_oracle):
    a, b = scenes.CASES["small_160x96"], scenes.CASES["small_176x104"]
    rec = cloud_of(cpu_oracle, a)
    pc = gpu.ProjectCloud.from_packed(rec)
    outs = []
    for case in (a, b, a):
        P = case.W * case.H
        color, depth = np.zeros(P * 3, np.uint8), np.zeros(P, np.float32)
        pc.computeFilteredRGBD(calib_of(gpu, case), case.poses[0], color, depth)
        outs.append((color, depth.view(np.uint32)))
    pc.close()
    assert np.array_equal(outs[0][0], outs[2][0]) and np.array_equal(outs[0][1], outs[2][1])
    fresh = gpu.ProjectCloud.from_packed(rec)
    P = b.W * b.H
    color, depth = np.zeros(P * 3, np.uint8), np.zeros(P, np.float32)
    fresh.computeFilteredRGBD(calib_of(gpu, b), b.poses[0], color, depth)
    fresh.close()
    assert np.array_equal(outs[1][0], color) and np.array_equal(outs[1][1], depth.view(np.uint32))


def test_distortion_matches_opencv_model(gpu, cpu_oracle):
    """New feature (the reference never applies distortion): parity unpinned, checked against
    cv2.projectPoints; tolerance: the rounded pixel may differ by at most 1 in u or v, and for
    >= 99.9 % of the points it is identical."""
    cv2 = pytest.importorskip("cv2")
    case = scenes.CASES["c2_1280x720"]
    rec = cloud_of(cpu_oracle, case)[:500_000]
    dist = [-0.05, 0.01, 0.0005, -0.0005, 0.0]
    pc = gpu.ProjectCloud.from_packed(rec, apply_distortion=True)
    calib = calib_of(gpu, case)
    calib.setDistortionParameters(dist)
    E = case.poses[0]
    pc.set_camera(calib, E)
    pix, zb = pc.project_points()
    rec = pc.download_cloud()
    pc.close()
    xyz = rec[:, :3].astype(np.float64)
    cam = xyz @ E[:3, :3].T + E[:3, 3]
    front = cam[:, 2] > 0.05
    uv, _ = cv2.projectPoints(cam[front].reshape(-1, 1, 3), np.zeros(3), np.zeros(3), case.K, np.array(dist))
    uv = uv.reshape(-1, 2)
    u, v = np.rint(uv[:, 0]).astype(np.int64), np.rint(uv[:, 1]).astype(np.int64)
    inside = (u >= 0) & (u < case.W) & (v >= 0) & (v < case.H)
    got = pix[front]
    both = inside & (got >= 0)
    assert both.sum() > 5_000
    gu, gv = got[both] % case.W, got[both] // case.W
    du, dv = np.abs(gu - u[both]), np.abs(gv - v[both])
    assert du.max() <= 1 and dv.max() <= 1
    assert ((du == 0) & (dv == 0)).mean() >= 0.999
    # in/out disagreement only at the image border
    assert (inside != (got >= 0)).mean() < 2e-3
    # depth is the camera-space z
    z = zb[front][both].view(np.float32)
    assert np.allclose(z, cam[front][both, 2], rtol=1e-5)


def test_distorted_projection_equals_its_cpu_restatement(gpu, cpu_oracle):
    """The distorted projection uses IEEE-rounded operations only, so — unlike the pinhole path's MUFU.RCP — it is
    reproducible on a CPU: the GPU's per-point (pixel, depth bits) must equal oracle.project_distorted bit for bit."""
    import ctypes as C
    case = scenes.CASES["c2_1280x720"]
    rec = cloud_of(cpu_oracle, case)[:600_000]
    for dist in ([-0.05, 0.01, 0.0005, -0.0005, 0.0], [0.2, 0.05, 0.0, 0.0, 0.01], [-0.3, 0.1, 0.004, -0.003, -0.01]):
        pc = gpu.ProjectCloud.from_packed(rec, apply_distortion=True)
        calib = calib_of(gpu, case)
        calib.setDistortionParameters(dist)
        for E in (case.poses[0], gpu.look_at_w2c((1.0, 1.0, 1.0), (1.0, 0.7, 0.1))):
            pc.set_camera(calib, E)
            pix, zb = pc.project_points()
            r2max, rstar = C.c_double(0), C.c_double(0)
            K9 = np.ascontiguousarray(case.K.reshape(9))
            d5 = np.asarray(dist, np.float64)
            assert gpu.load_library().rtr_host_distortion_bounds(case.W, case.H, K9.ctypes.data_as(gpu._dp), d5.ctypes.data_as(gpu._dp), C.byref(r2max), C.byref(rstar)) == 1
            want_pix, want_zb = cpu_oracle.project_distorted(pc.download_cloud(), E, case.K, dist, r2max.value, case.W, case.H)
            assert (want_pix >= 0).sum() > 1_000
            assert np.array_equal(pix, want_pix) and np.array_equal(zb, want_zb)
            # ... and with it the whole distorted frame is reproducible on the CPU without any GPU tap
            P = case.W * case.H
            color, depth = np.zeros(P * 3, np.uint8), np.zeros(P, np.float32)
            assert pc.computeFilteredRGBD(calib, E, color, depth) == 1
            gold = cpu_oracle.render(want_pix, want_zb, scenes.bgra_of(pc.download_cloud()), case.W, case.H, filtered=True)
            assert np.array_equal(color, gold["image"]) and np.array_equal(depth.view(np.uint32), gold["zbuf"])
            assert np.array_equal(pc.read("tensor", np.uint16, P * 5), gold["tensor"])
        pc.close()


# ------------------------------------------------------------------ chunk-level frustum culling
def test_views_that_overflow_float_sums_switch_to_integer_sums(gpu, cpu_oracle):
    """A view in which a pixel collects > 65 793 points pays for the exact re-run once: the re-run leaves a note for the
    host, the following frames start with integer colour sums (no overflow flag, no re-run) and every frame is identical."""
    W, H, heavy = 64, 48, 70_000
    m = np.array([32, 0, 31.5, 0, 0, 32, 23.5, 0, 0, 0, 1, 0, 0, 0, 0, 1], np.float32)
    rng = np.random.default_rng(9)
    n = heavy + 5000
    xyz = rng.uniform(-1.0, 1.0, (n, 3)).astype(np.float32)
    xyz[:, 2] = rng.uniform(1.5, 3.0, n).astype(np.float32)
    xyz[:heavy] = np.array([0.013, 0.009, 1.0], np.float32)
    bgr = rng.integers(0, 256, (n, 3), dtype=np.uint8)
    pc = gpu.ProjectCloud.from_packed(gpu.pack_records(xyz, bgr))
    c = gpu.CameraCalibration()
    c.setWidth(W)
    c.setHeight(H)
    pc.set_camera(c)
    pc.set_cam_proj_raw(m)
    frames, flags, left = [], [], []
    for _ in range(4):
        color, depth = np.zeros(W * H * 3, np.uint8), np.zeros(W * H, np.float32)
        pc._check(pc._lib.rtr_render_filtered(pc._h, color.ctypes.data, depth.ctypes.data))
        frames.append((color, depth.view(np.uint32).copy(), pc.read("tensor", np.uint16, W * H * 5)))
        flags.append(int(pc.read("minmax", np.uint32, 3)[2]))
        left.append(pc.get_option("int_sum_frames"))
    tap, resident = pc.project_points(), pc.download_cloud()
    pc.close()
    assert flags == [1, 0, 0, 0] and left[0] == 0 and left[1] == 63 and left[3] == 61
    gold = cpu_oracle.render(tap[0], tap[1], scenes.bgra_of(resident), W, H, filtered=True)
    for color, depth, tensor in frames:
        assert np.array_equal(color, gold["image"]) and np.array_equal(depth, gold["zbuf"]) and np.array_equal(tensor, gold["tensor"])


def test_chunk_culling_never_changes_a_frame(gpu, cpu_oracle):
    """Culling on (default) vs off over many random cameras, including cameras outside the cloud, grazing
    views, huge focal lengths and chunks that hold NaN / inf / huge points: every buffer identical."""
    case = scenes.CASES["c1_640x480"]
    rec = cloud_of(cpu_oracle, case)[:400_000].copy()
    rng = np.random.default_rng(11)
    rec[5000:5003, 0] = [np.nan, np.inf, -np.inf]     # chunk 4
    rec[123456, 2] = np.nan
    rec[200000, 1] = 3e38
    rec[200001, :3] = [1e20, -1e20, 1e20]
    rec[300000, :3] = [2e12, 0, 0]
    W, H, P = 320, 208, 320 * 208                       # H % 16 == 0, small for speed
    on = gpu.ProjectCloud.from_packed(rec)
    off = gpu.ProjectCloud.from_packed(rec)
    off.set_option("chunk_cull", 0)
    assert on.get_option("chunk_cull") == 1
    culled_any = False
    for it in range(60):
        f = float(rng.choice([80.0, 230.0, 1000.0, 20000.0]))
        calib = gpu.CameraCalibration()
        calib.loadCalibration(f, f * rng.uniform(0.8, 1.2), rng.uniform(0, W), rng.uniform(0, H), [0.0] * 5, W, H)
        eye = rng.uniform([-3, -3, -1], [11, 9, 4])
        fwd = rng.standard_normal(3)
        E = gpu.look_at_w2c(eye, fwd, up=(0.0, 0.3, 1.0))
        outs = []
        for pc in (on, off):
            pc.set_camera(calib, E)
            pc.render_device(gpu.STAGE_FILTERED)
            outs.append((pc.read("zbuf", np.uint32, P), pc.read("accum", np.uint32, P * 4), pc.read("image", np.uint8, P * 3),
                         pc.read("tensor", np.uint16, P * 5)))
        for a, b, what in zip(outs[0], outs[1], ("zbuf", "accum", "image", "tensor")):
            assert np.array_equal(a, b), f"camera {it}: {what} changed by chunk culling"
        fr, vis, nch = on.cull_stats(reset=True)
        assert fr == 1 and 0 < vis <= nch           # the always-visible chunks are never dropped
        culled_any |= vis < nch
    assert culled_any
    # raw matrices with NaN / inf / absurd entries must simply disable the culling
    for bad in (np.nan, np.inf, 1e38):
        m = np.array([300, 0, 160, 0, 0, 300, 104, 0, 0, 0, 1, 0, 0, 0, 0, 1], np.float32)
        m[3] = bad
        res = []
        for pc in (on, off):
            pc.set_cam_proj_raw(m)
            pc.render_device(gpu.STAGE_RGBD)
            res.append((pc.read("zbuf", np.uint32, P), pc.read("accum", np.uint32, P * 4)))
        assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    on.close()
    off.close()


def test_chunk_culling_under_distortion_never_changes_a_frame(gpu, cpu_oracle):
    """Distorted cameras cull chunks against the square of normalised coordinates that bounds every point able to
    reach the image (rtr_renderer.cu make_params): culling on vs off must give identical buffers for barrel,
    pincushion, tangential and strong mixed distortion, cameras inside and outside the cloud."""
    case = scenes.CASES["c1_640x480"]
    rec = cloud_of(cpu_oracle, case)[:400_000].copy()
    rec[5000, 0] = np.nan
    rec[200001, :3] = [1e20, -1e20, 1e20]
    rng = np.random.default_rng(23)
    W, H, P = 320, 208, 320 * 208
    on = gpu.ProjectCloud.from_packed(rec, apply_distortion=True)
    off = gpu.ProjectCloud.from_packed(rec, apply_distortion=True)
    off.set_option("chunk_cull", 0)
    dists = [[-0.05, 0.01, 0.0005, -0.0005, 0.0], [0.2, 0.05, 0.0, 0.0, 0.01], [-0.3, 0.1, 0.0, 0.0, -0.01],
             [0.0, 0.0, 0.02, -0.03, 0.0], [-0.2, 0.03, 0.01, 0.01, 0.001], [1e-9, 0.0, 0.0, 0.0, 0.0]]
    culled_any = False
    for it in range(48):
        f = float(rng.choice([120.0, 230.0, 600.0]))
        calib = gpu.CameraCalibration()
        calib.loadCalibration(f, f * rng.uniform(0.9, 1.1), rng.uniform(0.3 * W, 0.7 * W), rng.uniform(0.3 * H, 0.7 * H),
                              dists[it % len(dists)], W, H)
        eye = rng.uniform([-2, -2, 0], [10, 8, 3])
        E = gpu.look_at_w2c(eye, rng.standard_normal(3), up=(0.0, 0.3, 1.0))
        outs = []
        for pc in (on, off):
            pc.set_camera(calib, E)
            pc.render_device(gpu.STAGE_FILTERED)
            outs.append((pc.read("zbuf", np.uint32, P), pc.read("accum", np.uint32, P * 4), pc.read("image", np.uint8, P * 3),
                         pc.read("tensor", np.uint16, P * 5)))
        for a, b, what in zip(outs[0], outs[1], ("zbuf", "accum", "image", "tensor")):
            assert np.array_equal(a, b), f"camera {it} dist {dists[it % len(dists)]}: {what} changed by chunk culling"
        fr, vis, nch = on.cull_stats(reset=True)
        assert fr == 1 and 0 < vis <= nch
        culled_any |= vis < nch
    assert culled_any
    on.close()
    off.close()


def test_chunk_culling_pixel_boundary_points(gpu):
    """Points placed exactly on and one ulp around the four image borders and the z = 0 plane, one chunk
    each, with a camera whose rows make u and v land on k + 0.5 ties: cull on == cull off."""
    W, H = 64, 48
    P = W * H
    m = np.array([32, 0, 31.5, 0, 0, 32, 23.5, 0, 0, 0, 1, 0, 0, 0, 0, 1], np.float32)
    chunks = []
    for ux in (-0.5, -0.5000001, -0.4999999, 63.5, 63.49999, 63.50001, 31.0):
        for vy in (-0.5, -0.5000001, 47.5, 47.50001, 23.0):
            z = np.float32(2.0)
            x = np.float32((ux - 31.5) * 2.0 / 32.0)
            y = np.float32((vy - 23.5) * 2.0 / 32.0)
            pts = np.tile(np.array([x, y, z], np.float32), (1024, 1))
            pts[:, 0] = np.nextafter(pts[:, 0], np.float32(np.inf) * np.where(np.arange(1024) % 2, 1, -1)).astype(np.float32)
            chunks.append(pts)
    for z in (0.0, -0.0, 1e-45, -1e-45, 1e-38, 1e-30):
        chunks.append(np.tile(np.array([0.0, 0.0, z], np.float32), (1024, 1)))
    xyz = np.concatenate(chunks)
    rec = gpu.pack_records(xyz, np.full((len(xyz), 3), 77, np.uint8))
    res = []
    for cull in (1, 0):
        pc = gpu.ProjectCloud.from_packed(rec, sort=False)    # keep the one-chunk-per-border-case layout
        pc.set_option("chunk_cull", cull)
        c = gpu.CameraCalibration()
        c.setWidth(W)
        c.setHeight(H)
        pc.set_camera(c)
        pc.set_cam_proj_raw(m)
        pc.render_device(gpu.STAGE_RGBD)
        res.append((pc.read("zbuf", np.uint32, P), pc.read("accum", np.uint32, P * 4)))
        pc.close()
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    assert (res[0][0] != 0x7F7FFFFF).sum() > 4


def test_float_accumulator_overflow_falls_back_to_exact_sums(gpu, cpu_oracle):
    """The default blend keeps {b,g,r,count} as floats (one 16-byte RED per point), exact up to 65793 points per
    pixel.  70 000 white points in one pixel must trigger the in-stream exact re-run; 65 793 must not need it; both
    equal the oracle (the reference's u32 sums)."""
    W, H = 64, 48
    m = np.array([50, 0, 31.5, 0, 0, 50, 23.5, 0, 0, 0, 1, 0, 0, 0, 0, 1], np.float32)
    rng = np.random.default_rng(5)
    for heavy in (70_000, 65_793, 65_794):
        n = heavy + 3000
        xyz = rng.uniform(-1.0, 1.0, (n, 3)).astype(np.float32)
        xyz[:, 2] = rng.uniform(1.5, 3.0, n).astype(np.float32)
        xyz[:heavy] = np.array([0.013, 0.009, 1.0], np.float32)
        bgr = rng.integers(0, 256, (n, 3), dtype=np.uint8)
        bgr[:heavy] = 255
        rec = gpu.pack_records(xyz, bgr)
        for filtered in (False, True):
            for opts in ({}, {"chunk_cull": 0}, {"blend_variant": 0}):
                pc = gpu.ProjectCloud.from_packed(rec)
                for k, v in opts.items():
                    pc.set_option(k, v)
                c = gpu.CameraCalibration()
                c.setWidth(W)
                c.setHeight(H)
                pc.set_camera(c)
                pc.set_cam_proj_raw(m)
                color, depth = np.zeros(W * H * 3, np.uint8), np.zeros(W * H, np.float32)
                fn = pc._lib.rtr_render_filtered if filtered else pc._lib.rtr_render_rgbd
                pc._check(fn(pc._h, color.ctypes.data, depth.ctypes.data))
                accum = pc.read("accum", np.uint32, W * H * 4)
                flag = pc.read("minmax", np.uint32, 3)[2]
                tap = pc.project_points()
                resident = pc.download_cloud()
                pc.close()
                gold = cpu_oracle.render(tap[0], tap[1], scenes.bgra_of(resident), W, H, filtered=filtered)
                assert accum.reshape(-1, 4)[:, 3].max() == heavy
                assert np.array_equal(accum, gold["accum"]) and np.array_equal(color, gold["image"])
                assert np.array_equal(depth.view(np.uint32), gold["zbuf"])
                assert flag == (1 if (heavy > 65_793 and opts.get("blend_variant", 4) == 4) else 0)


def test_ring_kernels_hand_a_stage_back_only_after_its_records_landed(gpu):
    """Regression (round 2): the ring kernels released a shared-memory stage as soon as the four LDS of a warp were ISSUED.
    With the SM's load/store queue backed up by scattered reductions — an unsorted cloud, every chunk streamed, no early
    depth test — the mbarrier arrive overtook the loads, the next tile's bulk copy landed under them, and a few records
    were processed twice / never: frames that differed from run to run in a handful of pixels (92 % of the frames in
    this configuration).  Every render must now give the same frame, equal to the per-thread kernels'."""
    n, W, H = 16_000_000, 1920, 1080
    P = W * H
    calib = gpu.CameraCalibration()
    calib.loadCalibration(1400.0, 1400.0, 959.5, 539.5, [0.0] * 5, W, H)
    E = gpu.trajectory_w2c(1000, center=(6.0, 5.0, 1.5), radius=2.0)[333]
    pc = gpu.ProjectCloud.synthetic(seed=5678, n_total=n, hall=scenes.HALL_LARGE, n_boxes=12, sort=False)   # scan order: a warp's records are all over the image

    def digest():
        color, depth = np.zeros(P * 3, np.uint8), np.zeros(P, np.float32)
        assert pc.computeRGBD(calib, E, color, depth) == 1
        return scenes.sha(color) + scenes.sha(depth) + scenes.sha(pc.read("accum", np.uint32, P * 4))

    pc.set_option("ring", 0)
    pc.set_option("chunk_cull", 0)
    want = digest()                                          # per-thread LDG kernels
    for opts in (dict(ring=2, chunk_cull=0, zmin_variant=0, blend_variant=0), dict(ring=2, chunk_cull=0), dict(ring=1, chunk_cull=1, blend_variant=0),
                 dict(ring=1, chunk_cull=1, ring_ctas=1)):
        for k, v in {**dict(zmin_variant=5, blend_variant=4, ring_ctas=2), **opts}.items():
            pc.set_option(k, v)
        got = [digest() for _ in range(25)]
        assert all(g == want for g in got), f"{opts}: {sum(g != want for g in got)} of {len(got)} renders differ from the per-thread kernels' frame"
    # ... and as a fused sequence (blend of frame k-1 + z-min of frame k through the same ring): 25 frames of the same pose
    for k, v in dict(ring=1, chunk_cull=1, zmin_variant=5, blend_variant=4, ring_ctas=2, fuse=2).items():
        pc.set_option(k, v)
    pc.set_camera(calib, E)
    bad = 0
    for i in range(25):
        pc.render_device(gpu.STAGE_RGBD)
        if i % 4 == 3:   # every fourth frame is read back (which completes it); the others are completed by their successor
            got = scenes.sha(pc.read("image", np.uint8, P * 3)) + scenes.sha(pc.read("zbuf", np.uint32, P)) + scenes.sha(pc.read("accum", np.uint32, P * 4))
            bad += got != want
    assert bad == 0, f"fused sequence: {bad} frames differ from the per-thread kernels' frame"
    pc.close()
