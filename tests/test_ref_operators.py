"""Operator by operator on RANDOM inputs: the oracle against the reference's own code (oracle/_ref/librd_ref.so).

tests/test_ref_device.py pins the oracle on the planes a real frame produces; this file feeds every order-independent
operator of the path with random planes instead - saturating / negative / out-of-range values, degenerate gradient vectors,
dense noise masks, odd sizes, all blur radii - through the reference's L2 entry points (oclimgutil.h:74-100) or, for the
kernels of oclrect.cl that have no wrapper, through the kernel itself (refcl_rect_<name>, work-items in raster order).
Everything here is bit-exact, floats included.
"""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol
import ref_lib as rl

pytestmark = pytest.mark.skipif(not rl.available(), reason="oracle/_ref/librd_ref.so neither built nor buildable here")
vp, ci, cf = C.c_void_p, C.c_int, C.c_float
SIZES = [(64, 48), (130, 97), (33, 21)]


def P(a):
    return a.ctypes.data


@pytest.fixture(scope="module")
def ctx():
    rl.set_threads(1)
    return rl.RefContext()


def bits(a):
    return np.ascontiguousarray(a).view(np.uint8)


def rand_plab(rng, n):
    return (rng.integers(0, 4096, n, dtype=np.uint32) | (rng.integers(0, 1024, n, dtype=np.uint32) << 12) | (rng.integers(0, 1024, n, dtype=np.uint32) << 22)).astype(np.uint32)


def smooth_plab(rng, iw, ih):
    """piecewise-constant colours with a little noise: realistic input for the blur / quantise / despeckle kernels"""
    yy, xx = np.mgrid[0:ih, 0:iw]
    cell = (yy // 9) * 7 + (xx // 11)
    base = rng.integers(0, 1 << 30, int(cell.max()) + 1, dtype=np.uint32)
    lab = base[cell]
    l = (lab & 0xfff).astype(np.int64) + rng.integers(-3, 4, (ih, iw))
    a = ((lab >> 12) & 0x3ff).astype(np.int64) + rng.integers(-2, 3, (ih, iw))
    b = ((lab >> 22) & 0x3ff).astype(np.int64) + rng.integers(-2, 3, (ih, iw))
    return (np.clip(l, 0, 4095) | (np.clip(a, 0, 1023) << 12) | (np.clip(b, 0, 1023) << 22)).astype(np.uint32).ravel()


@pytest.mark.parametrize("iw,ih", SIZES)
def test_elementwise_and_colour_operators(ctx, iw, ih):
    L, LO, u, q = ctx.L, ol.oracle(), ctx.imgutil, ctx.queue
    rng = np.random.default_rng(iw * 1000 + ih)
    n = iw * ih
    # bgr2plab with a row stride wider than the image (oclimgutil_convert_plab_bgr: the name is swapped in the reference, Q9)
    ws = 3 * iw + 5
    img = rng.integers(0, 256, ws * ih, dtype=np.uint8)
    m_out, m_in = ctx.mem(4 * n), ctx.mem(4 * n + ws * ih, img)
    L.oclimgutil_convert_plab_bgr(u, m_out, m_in, iw, ih, ws, q, None)
    out = np.zeros(n, np.uint32)
    LO.ora_convert_plab_bgr(P(out), P(img), iw, ih, ws)
    assert np.array_equal(ctx.view(m_out, n, np.uint32), out)
    # unpack -> three float planes; pack back, and pack of wild floats (negative, huge, NaN-free): floor + clamp (Q14)
    plab = rand_plab(rng, n)
    m = [ctx.mem(4 * n) for _ in range(3)]
    m_p = ctx.mem(4 * n, plab)
    L.oclimgutil_unpack_f_f_f_plab(u, m[0], m[1], m[2], m_p, iw, ih, q, None)
    f = [np.zeros(n, np.float32) for _ in range(3)]
    LO.ora_unpack_f_f_f_plab(P(f[0]), P(f[1]), P(f[2]), P(plab), iw, ih)
    for k in range(3):
        assert np.array_equal(bits(ctx.view(m[k], n, np.float32)), bits(f[k]))
    wild = [(rng.random(n, dtype=np.float32) * 3 - 1).astype(np.float32) for _ in range(3)]
    wild[0][::7] = 1e9
    wild[1][::5] = -1e9
    mw = [ctx.mem(4 * n, w) for w in wild]
    L.oclimgutil_pack_plab_f_f_f(u, m_p, mw[0], mw[1], mw[2], iw, ih, q, None)
    packed = np.zeros(n, np.uint32)
    LO.ora_pack_plab_f_f_f(P(packed), P(wild[0]), P(wild[1]), P(wild[2]), iw, ih)
    assert np.array_equal(ctx.view(m_p, n, np.uint32), packed)
    # threshold_f_f, cast_i_f, threshold_i_i, cast_c_i, clear (size argument in BYTES, oclimgutil.cl:197)
    src = (rng.random(n, dtype=np.float32) * 4 - 2).astype(np.float32)
    a, b = ctx.mem(4 * n), ctx.mem(4 * n, src)
    L.oclimgutil_threshold_f_f(u, a, b, -1.5, 0.25, 7.0, n, q, None)
    o = np.zeros(n, np.float32)
    LO.ora_threshold_f_f(P(o), P(src), -1.5, 0.25, 7.0, n)
    assert np.array_equal(bits(ctx.view(a, n, np.float32)), bits(o))
    L.oclimgutil_cast_i_f(u, a, b, 1000.0, n, q, None)
    oi = np.zeros(n, np.int32)
    LO.ora_cast_i_f(P(oi), P(src), 1000.0, n)
    assert np.array_equal(ctx.view(a, n), oi)
    isrc = rng.integers(-500, 500, n, dtype=np.int32)
    c = ctx.mem(4 * n, isrc)
    L.oclimgutil_threshold_i_i(u, a, c, -3, 17, 9, n, q, None)
    LO.ora_threshold_i_i(P(oi), P(isrc), -3, 17, 9, n)
    assert np.array_equal(ctx.view(a, n), oi)
    L.oclimgutil_cast_c_i(u, a, c, n, q, None)
    oc = np.zeros(n, np.int8)
    LO.ora_cast_c_i(P(oc), P(isrc), n)
    assert np.array_equal(ctx.view(a, n, np.int8), oc)
    L.oclimgutil_clear(u, c, 4 * n, q, None)
    assert not ctx.view(c, n).any()
    ctx.release()


@pytest.mark.parametrize("iw,ih", SIZES)
def test_recursive_gaussian_all_radii(ctx, iw, ih):
    """oclimgutil_iirblur_f_f (oclimgutil.c:243-273: 2 clears + 6 passes) for every coefficient row the table holds that the
    frame is large enough for: the anti-causal recurrences start at index iw + r + 9 (oclimgutil.cl:569, :618) and fold that
    lead-in back with mirror1 (oclimgutil.cl:47), which stays inside the plane only while r <= min(iw, ih) - 11 - beyond
    that the reference itself reads out of bounds (the path uses r = 2 on frames of at least 13 x 13)"""
    L, LO = ctx.L, ol.oracle()
    rng = np.random.default_rng(7)
    n = iw * ih
    src = rng.random(n, dtype=np.float32)
    for r in range(0, min(32, min(iw, ih) - 10)):
        mo, mi, t0, t1 = ctx.mem(4 * n), ctx.mem(4 * n, src), ctx.mem(4 * n), ctx.mem(4 * n)
        L.oclimgutil_iirblur_f_f(ctx.imgutil, mo, mi, t0, t1, r, iw, ih, ctx.queue, None)
        out, a, b = np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros(n, np.float32)
        LO.ora_iirblur_f_f(P(out), P(src), P(a), P(b), r, iw, ih)
        assert np.array_equal(bits(ctx.view(mo, n, np.float32)), bits(out)), r
        ctx.release()


@pytest.mark.parametrize("iw,ih", SIZES)
def test_edge_operators_on_noise(ctx, iw, ih):
    """edgevec_f2_f (incl. flat areas -> the degenerate (sqrt 1/2, sqrt 1/2) vector), edge_f_plab, thinthres_f_f_f2 (bicubic
    samples with mirror borders) on noise, on a flat plane and on vectors pointing everywhere"""
    L, LO, u, q = ctx.L, ol.oracle(), ctx.imgutil, ctx.queue
    rng = np.random.default_rng(11)
    n = iw * ih
    lum = rng.random(n, dtype=np.float32)
    lum.reshape(ih, iw)[: ih // 3] = 0.5                                          # flat third: zero gradient
    m_v, m_l = ctx.mem(8 * n), ctx.mem(4 * n, lum)
    L.oclimgutil_edgevec_f2_f(u, m_v, m_l, iw, ih, q, None)
    vec = np.zeros(2 * n, np.float32)
    LO.ora_edgevec_f2_f(P(vec), P(lum), iw, ih)
    assert np.array_equal(bits(ctx.view(m_v, 2 * n, np.float32)), bits(vec))
    plab = rand_plab(rng, n)
    m_e, m_p = ctx.mem(4 * n), ctx.mem(4 * n, plab)
    L.oclimgutil_edge_f_plab(u, m_e, m_p, iw, ih, q, None)
    mag = np.zeros(n, np.float32)
    LO.ora_edge_f_plab(P(mag), P(plab), iw, ih)
    assert np.array_equal(bits(ctx.view(m_e, n, np.float32)), bits(mag))
    ang = rng.random(n) * 2 * np.pi
    anyvec = np.stack([np.cos(ang), np.sin(ang)], 1).astype(np.float32).ravel()
    anyvec[: 2 * iw] = 0                                                          # zero vectors too
    m_t, m_av = ctx.mem(4 * n), ctx.mem(8 * n, anyvec)
    L.oclimgutil_thinthres_f_f_f2(u, m_t, m_e, m_av, iw, ih, q, None)
    thin = np.zeros(n, np.float32)
    LO.ora_thinthres_f_f_f2(P(thin), P(mag), P(anyvec), iw, ih)
    assert np.array_equal(bits(ctx.view(m_t, n, np.float32)), bits(thin))
    assert (thin > 0).sum() > n // 50
    ctx.release()


@pytest.mark.parametrize("iw,ih", SIZES)
def test_strength_operators(ctx, iw, ih):
    """calcStrength / filterStrength (oclimgutil.cl:641-659 and their oclrect.cl copies :137-153) on random labels"""
    L, LO, u, q = ctx.L, ol.oracle(), ctx.imgutil, ctx.queue
    rng = np.random.default_rng(13)
    n = iw * ih
    label = rng.integers(0, 40, n, dtype=np.int32)
    label[rng.random(n) < 0.5] = -1
    edge = (rng.random(n, dtype=np.float32) * 0.3).astype(np.float32)
    m_s, m_e, m_l = ctx.mem(4 * n), ctx.mem(4 * n, edge), ctx.mem(4 * n, label)
    L.oclimgutil_calcStrength(u, m_s, m_e, m_l, iw, ih, q, None)
    s = np.zeros(n, np.int32)
    LO.ora_calcStrength(P(s), P(edge), P(label), iw, ih)
    assert np.array_equal(ctx.view(m_s, n), s)
    L.oclimgutil_filterStrength(u, m_l, m_s, 500, iw, ih, q, None)
    lab2 = label.copy()
    LO.ora_filterStrength(P(lab2), P(s), 500, iw, ih)
    assert np.array_equal(ctx.view(m_l, n), lab2)
    k1, k2 = rl.kernel_direct("rect", "calcStrength"), rl.kernel_direct("rect", "filterStrength")
    k1.argtypes, k2.argtypes = [ci, ci, vp, vp, vp, ci, ci], [ci, ci, vp, vp, ci, ci, ci]
    s2, lab3 = np.zeros(n, np.int32), label.copy()
    k1(iw, ih, P(s2), P(edge), P(label), iw, ih)
    k2(iw, ih, P(lab3), P(s2), 500, iw, ih)
    assert np.array_equal(s2, s) and np.array_equal(lab3, lab2)
    ctx.release()


@pytest.mark.parametrize("iw,ih", SIZES)
def test_string_cleanup_kernels_on_noise(ctx, iw, ih):
    """simpleJunction / simpleConnect / stringify of oclrect.cl (:74-135) on sparse and dense random bitmaps"""
    LO = ol.oracle()
    rng = np.random.default_rng(17)
    n = iw * ih
    for density in (0.05, 0.3, 0.7):
        src = (rng.random(n) < density).astype(np.int32)
        for name, fn, extra in (("simpleJunction", LO.ora_rect_simpleJunction, ()), ("simpleConnect", LO.ora_rect_simpleConnect, ()),
                                ("stringify", LO.ora_rect_stringify, (0,)), ("stringify", LO.ora_rect_stringify, (1,))):
            k = rl.kernel_direct("rect", name)
            k.argtypes = [ci, ci, vp, vp] + [ci] * (2 + len(extra))
            inp = src if name != "simpleConnect" else _junction(LO, src, iw, ih)
            a, b = np.full(n, -7, np.int32), np.full(n, -7, np.int32)
            k(iw, ih, P(a), P(inp), *extra, iw, ih)
            fn(P(b), P(inp), *extra, iw, ih)
            assert np.array_equal(a, b), (name, density, extra)


def _junction(LO, src, iw, ih):
    out = np.zeros(iw * ih, np.int32)
    LO.ora_rect_simpleJunction(P(out), P(src), iw, ih)
    return out


@pytest.mark.parametrize("iw,ih", SIZES)
def test_colour_smoothing_kernels(ctx, iw, ih):
    """blblur0 / blblur1 (edge-stopped box blur), quantize, despeckle (oclrect.cl:155-244) on piecewise-constant colours with
    noise and a random edge mask, and on pure noise"""
    LO = ol.oracle()
    rng = np.random.default_rng(19)
    n = iw * ih
    k0, k1, kq, kd = (rl.kernel_direct("rect", k) for k in ("blblur0", "blblur1", "quantize", "despeckle"))
    k0.argtypes = k1.argtypes = [ci, ci, vp, vp, vp, ci, ci]
    kq.argtypes = [ci, ci, vp, vp, ci, ci, ci, ci, ci]
    kd.argtypes = [ci, ci, vp, vp, vp, ci, ci]
    for plab in (smooth_plab(rng, iw, ih), rand_plab(rng, n)):
        edge8 = (rng.random(n) < 0.08).astype(np.int8)
        cur_r, cur_o = plab.copy(), plab.copy()
        for it in range(3):
            for k, fn in ((k0, LO.ora_rect_blblur0), (k1, LO.ora_rect_blblur1)):
                a, b = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
                k(iw, ih, P(a), P(edge8), P(cur_r), iw, ih)
                fn(P(b), P(edge8), P(cur_o), iw, ih)
                assert np.array_equal(a, b), it
                cur_r, cur_o = a, b
        a, b = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
        kq(iw, ih, P(a), P(cur_r), 24, 24, 24, iw, ih)
        LO.ora_rect_quantize(P(b), P(cur_o), 24, 24, 24, iw, ih)
        assert np.array_equal(a, b)
        thin = np.where(rng.random(n) < 0.1, rng.random(n), 0).astype(np.float32)
        c, d = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
        kd(iw, ih, P(c), P(a), P(thin), iw, ih)
        LO.ora_rect_despeckle(P(d), P(b), P(thin), iw, ih)
        assert np.array_equal(c, d)


@pytest.mark.parametrize("iw,ih", SIZES)
def test_merge_mask_boundary_and_size_kernels(ctx, iw, ih):
    """mkMergeMask0 / mkMergeMask1 (scatter of rings and discs), markBoundary, calcSize (oclrect.cl:246-287, 336-390)"""
    LO = ol.oracle()
    rng = np.random.default_rng(23)
    n = iw * ih
    junction = np.zeros(n, np.int32)
    pts = rng.integers(0, n, max(4, n // 400))
    junction[pts] = rng.integers(1, 5, pts.size)                                  # 2 = end pixels, others = junctions
    m0, m1 = rl.kernel_direct("rect", "mkMergeMask0"), rl.kernel_direct("rect", "mkMergeMask1")
    m0.argtypes = m1.argtypes = [ci, ci, vp, vp, ci, ci]
    a, b = np.zeros(n, np.int32), np.zeros(n, np.int32)
    m0(iw, ih, P(a), P(junction), iw, ih)
    LO.ora_rect_mkMergeMask0(P(b), P(junction), iw, ih)
    assert np.array_equal(a, b) and a.any()
    m1(iw, ih, P(a), P(junction), iw, ih)
    LO.ora_rect_mkMergeMask1(P(b), P(junction), iw, ih)
    assert np.array_equal(a, b)
    yy, xx = np.mgrid[0:ih, 0:iw]
    label = ((yy // 6) * 50 + xx // 8).astype(np.int32).ravel()
    mb, cs = rl.kernel_direct("rect", "markBoundary"), rl.kernel_direct("rect", "calcSize")
    mb.argtypes, cs.argtypes = [ci, ci, vp, vp, vp, ci, ci], [ci, ci, vp, vp, ci, ci]
    a, b = np.zeros(n, np.int32), np.zeros(n, np.int32)
    mb(iw, ih, P(a), P(label), P(junction), iw, ih)
    LO.ora_rect_markBoundary(P(b), P(label), iw, ih)
    assert np.array_equal(a, b) and (a >= 0).any()
    a, b = junction.copy(), junction.copy()                                      # accumulates on top of what is there (Q2)
    cs(iw, ih, P(a), P(label), iw, ih)
    LO.ora_rect_calcSize(P(b), P(label), iw, ih)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("iw,ih", SIZES)
def test_label8x_on_blobs(ctx, iw, ih):
    """oclimgutil_label8x_int_int (1 + 10 launches, MAXPASS 10) on random blobs, both background conventions (bgc 0 / -1)"""
    L, LO = ctx.L, ol.oracle()
    rng = np.random.default_rng(29)
    n = iw * ih
    yy, xx = np.mgrid[0:ih, 0:iw]
    pix = np.zeros((ih, iw), np.int32)
    for _ in range(12):
        cy, cx, r = rng.integers(0, ih), rng.integers(0, iw), rng.integers(2, 9)
        pix[(yy - cy) ** 2 + (xx - cx) ** 2 <= r * r] = rng.integers(1, 4)
    for bgc, src in ((0, pix.ravel()), (-1, np.where(pix.ravel() == 0, -1, pix.ravel()).astype(np.int32))):
        mo, mi, mt = ctx.mem(4 * n), ctx.mem(4 * n, src), ctx.mem(4 * n)
        L.oclimgutil_label8x_int_int(ctx.imgutil, mo, mi, mt, bgc, iw, ih, ctx.queue, None)
        out, tmp = np.zeros(n, np.int32), np.zeros(n, np.int32)
        LO.ora_label8x_int_int(P(out), P(np.ascontiguousarray(src)), P(tmp), bgc, iw, ih)
        assert np.array_equal(ctx.view(mo, n), out), bgc
        ctx.release()


@pytest.mark.parametrize("iw,ih", SIZES)
def test_operators_off_the_configured_paths(ctx, iw, ih):
    """the L2 operators no program of the reference enqueues (oclimgutil.h:85-94): packed Lab -> BGR (every 10th of the 2^32 codes would
    be too many: random codes + the code-space corners), float / label planes -> BGR, edge_f_f, thincubic, edgevec_f2_plab -
    the oracle against the reference's kernels, bit-exact"""
    L, LO, u, q = ctx.L, ol.oracle(), ctx.imgutil, ctx.queue
    rng = np.random.default_rng(iw * 77 + ih)
    n = iw * ih
    ws = 3 * iw + 7
    plab = rand_plab(rng, n)
    plab[:8] = [0, 0xffffffff, 4095, 1023 << 12, 1023 << 22, 2048 | (512 << 12) | (512 << 22), 1, 0x80000000]
    m_out, m_in = ctx.mem(ws * ih + 16), ctx.mem(4 * n, plab)
    L.oclimgutil_convert_bgr_plab(u, m_out, m_in, iw, ih, ws, q, None)            # plab2bgr (Q9)
    out = np.zeros(ws * ih, np.uint8)
    LO.ora_convert_bgr_plab(P(out), P(plab), iw, ih, ws)
    got = ctx.view(m_out, ws * ih, np.uint8).reshape(ih, ws)[:, :3 * iw]
    assert np.array_equal(got, out.reshape(ih, ws)[:, :3 * iw]) and got.max() > 200 and got.min() == 0
    # round trip through the forward conversion stays close in Lab (a sanity check of the restated formula, not of rounding)
    fsrc = (rng.random(n, dtype=np.float32) * 1.6 - 0.3).astype(np.float32)
    m_f = ctx.mem(4 * n, fsrc)
    for f in (1.0, 0.37, 3.0):
        L.oclimgutil_convert_bgr_lumaf(u, m_out, m_f, f, iw, ih, ws, q, None)
        LO.ora_convert_bgr_lumaf(P(out), P(fsrc), f, iw, ih, ws)
        assert np.array_equal(ctx.view(m_out, ws * ih, np.uint8).reshape(ih, ws)[:, :3 * iw], out.reshape(ih, ws)[:, :3 * iw]), f
    lab = rng.integers(-3, 1 << 20, n, dtype=np.int32)
    lab[::9] = 5
    m_l = ctx.mem(4 * n, lab)
    L.oclimgutil_convert_bgr_labeli(u, m_out, m_l, 5, iw, ih, ws, q, None)
    LO.ora_convert_bgr_labeli(P(out), P(lab), 5, iw, ih, ws)
    assert np.array_equal(ctx.view(m_out, ws * ih, np.uint8).reshape(ih, ws)[:, :3 * iw], out.reshape(ih, ws)[:, :3 * iw])
    # edge_f_f on noise (negative sums -> 0), thincubic with vectors pointing everywhere, edgevec_f2_plab on noise and on a flat plane
    src = (rng.random(n, dtype=np.float32) * 2 - 0.5).astype(np.float32)
    a, b = ctx.mem(4 * n), ctx.mem(4 * n, src)
    L.oclimgutil_edge_f_f(u, a, b, iw, ih, q, None)
    o = np.zeros(n, np.float32)
    LO.ora_edge_f_f(P(o), P(src), iw, ih)
    assert np.array_equal(bits(ctx.view(a, n, np.float32)), bits(o)) and (o == 0).any() and (o > 0).any()
    ang = rng.random(n) * 2 * np.pi
    vec = np.stack([np.cos(ang), np.sin(ang)], 1).astype(np.float32) * rng.choice([0.0, 0.5, 1.0, 1.7], n)[:, None].astype(np.float32)
    mv = ctx.mem(8 * n, vec)
    L.oclimgutil_thincubic_f_f_f2(u, a, b, mv, iw, ih, q, None)
    LO.ora_thincubic_f_f_f2(P(o), P(src), P(vec), iw, ih)
    assert np.array_equal(bits(ctx.view(a, n, np.float32)), bits(o)) and (o != 0).any()
    for pl in (plab, smooth_plab(rng, iw, ih), np.full(n, 2048 | (512 << 12) | (512 << 22), np.uint32)):
        mp, mo = ctx.mem(4 * n, pl), ctx.mem(8 * n)
        L.oclimgutil_edgevec_f2_plab(u, mo, mp, iw, ih, q, None)
        ov = np.zeros(2 * n, np.float32)
        LO.ora_edgevec_f2_plab(P(ov), P(pl), iw, ih)
        assert np.array_equal(bits(ctx.view(mo, 2 * n, np.float32)), bits(ov))
    ctx.release()
