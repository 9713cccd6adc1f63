"""Known-answer tests of the CPU oracle (oracle/rtr_oracle.c) against an independent numpy
restatement of the same reference lines, on small seeded inputs and hand-made edge cases."""
import ctypes as C

import numpy as np
import pytest


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def fma32(a, b, c):
    """float32 fma through float64: the product of two float32 is exact in float64."""
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(np.float32)


def test_f16_conversion_matches_ieee(cpu_oracle):
    rng = np.random.default_rng(0)
    xs = np.concatenate([
        rng.standard_normal(20000).astype(np.float32) * np.float32(10.0) ** rng.integers(-9, 6, 20000).astype(np.float32),
        np.array([0.0, -0.0, 1.0, -1.0, 65504.0, 65519.9, 65520.0, 1e9, -1e9, 5.96e-8, 2.98e-8, 2.9802322e-8, 3e-8, 6.1e-5,
                  6.0975552e-5, 1.0 / 255.0, 254.0 / 255.0, np.inf, -np.inf, 0.333251953125, 0.3332519531250001], np.float32)])
    with np.errstate(over="ignore"):
        want = xs.astype(np.float16).view(np.uint16)
    got = np.array([cpu_oracle.f32_to_f16(float(x)) for x in xs], np.uint16)
    assert np.array_equal(got, want)
    assert cpu_oracle.f32_to_f16(float("nan")) == 0x7FFF           # cvt.rn.f16.f32 canonical NaN
    back = np.array([cpu_oracle.lib.rtro_f16_to_f32(int(h)) for h in range(0, 0x7C01, 7)], np.float32)
    assert np.array_equal(back, np.arange(0, 0x7C01, 7, dtype=np.uint16).view(np.float16).astype(np.float32))


def test_cam_proj_is_float_K4_times_E_left_to_right(cpu_oracle):
    rng = np.random.default_rng(1)
    for _ in range(20):
        K = np.array([[rng.uniform(300, 3000), rng.uniform(-1, 1), rng.uniform(100, 2000)],
                      [0, rng.uniform(300, 3000), rng.uniform(100, 2000)], [0, 0, 1]])
        E = np.eye(4)
        E[:3, :4] = rng.standard_normal((3, 4))
        K4 = np.zeros((4, 4), np.float32)
        K4[:3, :3] = K.astype(np.float32)
        K4[3, 3] = 1
        Ef = E.astype(np.float32)
        want = np.zeros((4, 4), np.float32)
        for r in range(4):
            for c in range(4):
                t = np.float32(K4[r, 0] * Ef[0, c])
                for k in (1, 2, 3):
                    t = np.float32(t + np.float32(K4[r, k] * Ef[k, c]))
                want[r, c] = t
        assert np.array_equal(cpu_oracle.cam_proj(K, E).reshape(4, 4), want)


def test_projection_op_order_and_culling(cpu_oracle):
    rng = np.random.default_rng(2)
    n, W, H = 50000, 320, 200
    pts = rng.uniform(-4, 4, (n, 4)).astype(np.float32)
    m = np.array([250, 0.3, 159.5, 1.5, 0, 250, 99.5, -2.0, 0.01, -0.02, 1, 0.25, 0, 0, 0, 1], np.float32)
    pix, zb = cpu_oracle.project(pts, m, W, H)
    x, y, z = pts[:, 0], pts[:, 1], pts[:, 2]

    def row(i):   # render.cu:33-40 as compiled: y*m1, fma(x,m0,.), fma(z,m2,.), + m3
        t = (y * m[4 * i + 1]).astype(np.float32)
        t = fma32(x, m[4 * i + 0], t)
        t = fma32(z, m[4 * i + 2], t)
        return (t + m[4 * i + 3]).astype(np.float32)
    rx, ry, rz = row(0), row(1), row(2)
    with np.errstate(divide="ignore", invalid="ignore"):
        rcp = (np.float32(1.0) / rz).astype(np.float32)
        u = np.rint((rcp * rx).astype(np.float32)).astype(np.int64)
        v = np.rint((rcp * ry).astype(np.float32)).astype(np.int64)
    live = (rz > 0) & (u >= 0) & (u < W) & (v >= 0) & (v < H)
    assert live.sum() > 2000
    assert np.array_equal(pix >= 0, live)
    assert np.array_equal(pix[live], (v * W + u)[live])
    assert np.array_equal(zb[live], rz.view(np.uint32)[live])
    assert (zb[~live] == 0).all()


def test_point_stages_against_numpy(cpu_oracle):
    rng = np.random.default_rng(3)
    W, H, n = 48, 40, 30000       # H' = 32: rows 32..39 are outside the reference's clear/resolve coverage
    P = W * H
    pix = rng.integers(-1, P, n).astype(np.int32)
    z = rng.uniform(0.5, 0.6, n).astype(np.float32)
    z[::7] = rng.uniform(0.0, 0.03, len(z[::7])).astype(np.float32)      # can land inside 0 + 0.02 in the stale tail
    bgra = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
    buf = cpu_oracle.stages(pix, z.view(np.uint32), bgra, W, H)
    cov = (W // 16) * (H // 16) * 256
    assert cpu_oracle.lib.rtro_coverage(W, H) == cov == 1536
    live = pix >= 0
    zb = np.zeros(P, np.uint32)
    zb[:cov] = 0x7F7FFFFF                                  # fillBuffer coverage; the tail keeps its zero-init
    np.minimum.at(zb, pix[live], z.view(np.uint32)[live])
    assert np.array_equal(buf["zbuf"], zb)
    lim = (zb.view(np.float32)[pix[live]] + np.float32(0.02)).astype(np.float32)
    ok = ~(z[live] > lim)
    acc = np.zeros((P, 4), np.uint32)
    pl, cl = pix[live][ok], bgra[live][ok]
    for k in range(3):
        np.add.at(acc[:, k], pl, (cl >> (8 * k)) & 0xFF)
    np.add.at(acc[:, 3], pl, 1)
    assert np.array_equal(buf["accum"].reshape(P, 4), acc)
    img = np.zeros((P, 3), np.uint8)
    has = acc[:cov, 3] > 0
    img[:cov][has] = (acc[:cov, :3][has] // acc[:cov, 3:4][has]).astype(np.uint8)
    assert np.array_equal(buf["image"].reshape(P, 3), img)
    assert acc[cov:, 3].sum() > 0 and not buf["image"][cov * 3:].any()   # tail accumulates but is never resolved


def test_minmax_skips_empty(cpu_oracle):
    z = np.array([0x7F7FFFFF, 0x3F800000, 0x40000000, 0x7F7FFFFF, 0x3F000000], np.uint32)
    mn, mx = C.c_uint32(), C.c_uint32()
    cpu_oracle.lib.rtro_minmax(_p(z), 5, C.byref(mn), C.byref(mx))
    assert (mn.value, mx.value) == (0x3F000000, 0x40000000)
    e = np.full(4, 0x7F7FFFFF, np.uint32)
    cpu_oracle.lib.rtro_minmax(_p(e), 4, C.byref(mn), C.byref(mx))
    assert (mn.value, mx.value) == (0xFFFFFFFF, 0)


def test_reduce_and_laplacian_against_numpy(cpu_oracle):
    rng = np.random.default_rng(4)
    w, h = 24, 14
    hi = rng.uniform(1, 5, (2 * h, 2 * w)).astype(np.float32)
    hi[rng.random(hi.shape) < 0.3] = np.float32(3.4028234663852886e38)
    lo = np.zeros((h, w), np.float32)
    cpu_oracle.lib.rtro_reduce(_p(hi), _p(lo), w, h)
    want = np.minimum(np.minimum(hi[0::2, 0::2], hi[0::2, 1::2]), np.minimum(hi[1::2, 0::2], hi[1::2, 1::2]))
    assert np.array_equal(lo, want)
    out = np.zeros((h, w), np.uint8)
    cpu_oracle.lib.rtro_laplacian(_p(lo), _p(out), w, h)
    kern = np.array([0, 1, 0, 1, -4, 1, 0, 1, 0], np.float32)
    want = np.zeros((h, w), np.uint8)
    with np.errstate(over="ignore", invalid="ignore"):
        for y in range(1, h - 1):
            for x in range(1, w - 1):
                s = np.float32(0)
                for k in range(9):
                    s = fma32(lo[y + k // 3 - 1, x + k % 3 - 1], kern[k], s)
                want[y, x] = 255 if s > np.float32(0.03) else 0
    assert np.array_equal(out, want)
    assert want.any() and not want.all()
    # FLT_MAX neighbourhoods (empty regions): the running sum overflows to +inf at the second unit tap and the
    # FUSED -4*FLT_MAX (exact, finite product) leaves it +inf -> flagged as an edge; an unfused product would
    # have given inf - inf = NaN -> 0.  This is why the contraction order matters for parity.
    flat = np.full((5, 5), 3.4028234663852886e38, np.float32)
    out = np.ones((5, 5), np.uint8)
    cpu_oracle.lib.rtro_laplacian(_p(flat), _p(out), 5, 5)
    assert (out[1:4, 1:4] == 255).all() and out.sum() == 9 * 255


@pytest.mark.parametrize("W,H,exp", [(640, 480, (640, 480)), (1920, 1080, (1920, 1072)), (1280, 720, (1280, 720)),
                                     (3840, 2160, (3840, 2160)), (200, 120, (192, 112)), (1752, 1168, (1744, 1168))])
def test_filter_dims_follow_the_reference_truncation(cpu_oracle, W, H, exp):
    z = np.full(W * H, 0x7F7FFFFF, np.uint32)
    out = cpu_oracle.depth_filter(z, np.zeros(W * H * 3, np.uint8), W, H)
    assert tuple(out["dims"][:2]) == exp
    assert out["minmax"].tolist() == [0xFFFFFFFF, 0]
    n = exp[0] * exp[1]
    assert (out["depth"][:n] == np.float32(-1).view(np.uint32)).all() and (out["depth"][n:] == 0x7F7FFFFF).all()
    t = out["tensor"]
    assert (t[4 * n:5 * n] == 0xBC00).all() and not t[:4 * n].any() and not t[5 * n:].any()


def test_remove_mask_tensor_values(cpu_oracle):
    """removeMask: plane k = half(float(half(c)) / 255), plane 3 = 1, plane 4 = half(float(half(d - min)) / (max - min))."""
    W = H = 64
    P = W * H
    z = np.full(P, 0x7F7FFFFF, np.uint32)
    d = np.linspace(2.0, 2.03, P).astype(np.float32)
    z[:] = d.view(np.uint32)
    img = (np.arange(P * 3) % 256).astype(np.uint8)
    out = cpu_oracle.depth_filter(z, img, W, H)
    keep = out["depth"] != np.float32(-1).view(np.uint32)
    assert keep.sum() > P // 2
    t = out["tensor"].reshape(5, P).view(np.float16)
    c = img.reshape(P, 3)
    for k in range(3):
        want = (c[:, k].astype(np.float16).astype(np.float32) / np.float32(255)).astype(np.float16)
        assert np.array_equal(t[k][keep].view(np.uint16), want[keep].view(np.uint16))
    assert (t[3][keep] == 1).all() and not t[3][~keep].any()
    want = ((d - d.min()).astype(np.float16).astype(np.float32) / np.float32(d.max() - d.min())).astype(np.float16)
    assert np.array_equal(t[4][keep].view(np.uint16), want[keep].view(np.uint16))
    assert (t[4][~keep] == -1).all()


def test_oracle_distorted_projection_follows_the_opencv_model(cpu_oracle):
    """rtro_project_distorted (the new lens-distortion feature's CPU restatement, float32 with the CUDA path's op order)
    against a float64 evaluation of the OpenCV model: same pixel for >= 99.9 % of the points, never more than 1 px off."""
    rng = np.random.default_rng(3)
    n, W, H = 200_000, 1280, 720
    rec = np.zeros((n, 4), np.float32)
    rec[:, :3] = rng.uniform([-4, -3, 0.5], [4, 3, 9], (n, 3)).astype(np.float32)
    K = np.array([[900.0, 0.3, 639.5], [0, 905.0, 359.5], [0, 0, 1]])
    dist = [-0.05, 0.01, 0.0005, -0.0005, 0.001]
    E = np.eye(4)
    E[:3, 3] = [0.1, -0.2, 0.3]
    pix, zb = cpu_oracle.project_distorted(rec, E, K, dist, 4.0, W, H)
    cam = rec[:, :3].astype(np.float64) + E[:3, 3]
    x, y = cam[:, 0] / cam[:, 2], cam[:, 1] / cam[:, 2]
    r2 = x * x + y * y
    rad = 1 + dist[0] * r2 + dist[1] * r2 ** 2 + dist[4] * r2 ** 3
    xd = x * rad + 2 * dist[2] * x * y + dist[3] * (r2 + 2 * x * x)
    yd = y * rad + dist[2] * (r2 + 2 * y * y) + 2 * dist[3] * x * y
    u, v = np.rint(K[0, 0] * xd + K[0, 1] * yd + K[0, 2]), np.rint(K[1, 1] * yd + K[1, 2])
    inside = (cam[:, 2] > 0) & (r2 <= 4.0) & (u >= 0) & (u < W) & (v >= 0) & (v < H)
    both = inside & (pix >= 0)
    assert both.sum() > 50_000 and (inside != (pix >= 0)).mean() < 2e-3
    du, dv = np.abs(pix[both] % W - u[both]), np.abs(pix[both] // W - v[both])
    assert du.max() <= 1 and dv.max() <= 1 and ((du == 0) & (dv == 0)).mean() >= 0.999
    assert np.allclose(zb[both].view(np.float32), cam[both, 2], rtol=1e-6)
