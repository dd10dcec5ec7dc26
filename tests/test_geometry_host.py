"""CPU check of the per-candidate geometry code the CUDA kernels execute
(pytorchocr_b200/csrc/geometry.cuh compiled for the host by tests/host_shim) against the oracle
and against cv2 / the reference's Clipper."""
import ctypes as C
import os
import subprocess

import cv2
import numpy as np
import pytest

from conftest import ROOT
from oracle import db_oracle, geometry_oracle as G
from oracle.pse_oracle import order_points_clockwise


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("shim") / "libgeomshim.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off",
                           os.path.join(ROOT, "tests", "host_shim", "geom_shim.cpp"), "-o", out])
    L = C.CDLL(out)
    L.shim_min_area_rect.restype = C.c_int
    L.shim_do_offset.restype = C.c_int
    L.shim_do_offset.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_int]
    L.shim_unclip_distance.restype = C.c_float
    L.shim_unclip_distance.argtypes = [C.c_void_p, C.c_float]
    L.shim_roundf.restype = C.c_float
    L.shim_roundf.argtypes = [C.c_float]
    L.shim_round_half_even.restype = C.c_double
    L.shim_round_half_even.argtypes = [C.c_double]
    return L


def _sorted_pts(p):
    p = np.asarray(p, np.float64).reshape(-1, 2)
    return p[np.lexsort((p[:, 1], p[:, 0]))]


def _set_dist(a, b):
    """symmetric max over points of the distance to the nearest point of the other set"""
    a = np.asarray(a, np.float64).reshape(-1, 2)
    b = np.asarray(b, np.float64).reshape(-1, 2)
    d = np.abs(a[:, None, :] - b[None, :, :]).max(-1)
    return max(d.min(1).max(), d.min(0).max())


def _rect(shim, pts):
    pts = np.ascontiguousarray(pts, np.int32)
    corners = np.zeros(8, np.float64)
    wh = np.zeros(2, np.float64)
    hn = shim.shim_min_area_rect(pts.ctypes.data_as(C.c_void_p), len(pts), corners.ctypes.data_as(C.c_void_p),
                                 wh.ctypes.data_as(C.c_void_p))
    return corners.reshape(4, 2), wh, hn


def _raster_blob(rng, ragged):
    c = rng.uniform(60, 300, 2)
    box = cv2.boxPoints(((float(c[0]), float(c[1])), (float(rng.uniform(6, 90)), float(rng.uniform(3, 25))),
                         float(rng.uniform(-90, 90))))
    box -= box.min(0) - 3
    m = np.zeros((int(box[:, 1].max()) + 6, int(box[:, 0].max()) + 6), np.uint8)
    cv2.fillPoly(m, [np.round(box).astype(np.int32)], 1)
    if ragged:
        m &= (rng.random(m.shape) > 0.05).astype(np.uint8)
    ys, xs = np.nonzero(m)
    return np.stack([xs, ys], 1)


def test_min_area_rect_vs_cv2_and_oracle(shim):
    rng = np.random.default_rng(3)
    worst = 0.0
    for i in range(800):
        pts = _raster_blob(rng, i % 2 == 0)
        if len(pts) < 3:
            continue
        got, wh, hn = _rect(shim, pts)
        ref = cv2.boxPoints(cv2.minAreaRect(pts.astype(np.int32)))
        err = _set_dist(ref, got)
        worst = max(worst, err)
        assert err < 1e-3, (i, err)
        o_c, (ow, oh) = G.min_area_rect(pts)
        assert _set_dist(o_c, got) < 1e-9
        assert hn == len(G.convex_hull(pts))
    assert worst < 1e-3


def test_min_area_rect_degenerate(shim):
    got, wh, hn = _rect(shim, [[5, 7]])
    assert hn == 1 and wh.tolist() == [0.0, 0.0] and np.all(got == [5, 7])
    got, wh, hn = _rect(shim, [[1, 1], [4, 4], [2, 2], [3, 3]])
    assert hn == 2 and abs(wh[0] - 3 * 2 ** 0.5) < 1e-12 and wh[1] == 0.0
    got, wh, hn = _rect(shim, [[x, y] for x in range(2, 11) for y in range(3, 7)])
    assert hn == 4 and sorted(wh.tolist()) == [3.0, 8.0]
    assert np.array_equal(_sorted_pts(got), _sorted_pts([[2, 3], [10, 3], [10, 6], [2, 6]]))  # exact integers


def test_do_offset_vs_oracle_and_clipper(shim):
    rng = np.random.default_rng(5)
    for i in range(1500):
        c = rng.uniform(50, 1200, 2)
        box = cv2.boxPoints(((float(c[0]), float(c[1])), (float(rng.uniform(2, 400)), float(rng.uniform(1, 60))),
                             float(rng.uniform(-90, 90))))
        mini, _ = db_oracle.get_mini_boxes(cv2.minAreaRect(box))
        d_ref = db_oracle.get_contour_area(mini, 1.7)
        d_got = shim.shim_unclip_distance(np.ascontiguousarray(mini, np.float32).ctypes.data_as(C.c_void_p), 1.7)
        assert np.float32(d_got) == np.float32(d_ref)
        quad = np.array([(int(mini[k][0]), int(mini[k][1])) for k in range(4)], np.int32)
        out = np.zeros((512, 2), np.int32)
        m = shim.shim_do_offset(quad.ctypes.data_as(C.c_void_p), float(d_ref), out.ctypes.data_as(C.c_void_p), 512)
        want = G.do_offset(quad.tolist(), float(d_ref))
        assert m == len(want)
        assert out[:m].tolist() == [list(p) for p in want]
    # degenerate: duplicate corners -> rejected
    quad = np.array([[3, 3], [3, 3], [9, 3], [9, 3]], np.int32)
    out = np.zeros((64, 2), np.int32)
    assert shim.shim_do_offset(quad.ctypes.data_as(C.c_void_p), 2.0, out.ctypes.data_as(C.c_void_p), 64) == 0
    # capacity overflow is reported
    quad = np.array([[0, 0], [900, 0], [900, 700], [0, 700]], np.int32)
    assert shim.shim_do_offset(quad.ctypes.data_as(C.c_void_p), 300.0, out.ctypes.data_as(C.c_void_p), 8) == -1


def test_box_orderings_and_rounding(shim):
    rng = np.random.default_rng(9)
    for _ in range(500):
        rect = ((float(rng.uniform(20, 500)), float(rng.uniform(20, 500))),
                (float(rng.uniform(3, 100)), float(rng.uniform(3, 40))), float(rng.uniform(-90, 90)))
        pts = cv2.boxPoints(rect)
        o = np.zeros((4, 2), np.float32)
        shim.shim_mini_box(np.ascontiguousarray(pts).ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p))
        want, _ = db_oracle.get_mini_boxes(rect)
        assert np.array_equal(o, want)
        shim.shim_order_clockwise(np.ascontiguousarray(pts).ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p))
        assert np.array_equal(o, order_points_clockwise(pts))
    for v in (0.5, 1.5, 2.5, -0.5, -1.5, 0.49999997, 3.4999998, 1e6 + 0.5):
        assert shim.shim_roundf(v) == db_oracle.roundf(np.float32(v))
        assert shim.shim_round_half_even(v) == np.round(v)


def test_box_points_order_matches_cv2(shim):
    """cv_box_order reproduces the corner ORDER of cv2.boxPoints(cv2.minAreaRect(.)), including exact
    45-degree diamonds where order_points_clockwise (utility.py:21-29) meets ties and repeats a corner."""
    rng = np.random.default_rng(11)
    n_diamond = 0
    for i in range(1500):
        if i % 3 == 0:   # lattice diamond: all edges at +-45 degrees
            c = rng.integers(10, 60, 2)
            a, b = int(rng.integers(1, 15)), int(rng.integers(1, 15))
            corners = np.array([c, c + [a, a], c + [a - b, a + b], c + [-b, b]])
            m = np.zeros((120, 120), np.uint8)
            cv2.fillPoly(m, [corners.astype(np.int32)], 1)
            ys, xs = np.nonzero(m)
            pts = np.stack([xs, ys], 1)
        elif i % 3 == 1:  # axis-aligned block
            x0, y0 = rng.integers(0, 50, 2)
            w, h = rng.integers(1, 40, 2)
            pts = np.array([[x, y] for x in range(x0, x0 + w + 1) for y in range(y0, y0 + h + 1)])
        else:
            pts = _raster_blob(rng, i % 2 == 0)
        if len(pts) < 3:
            continue
        pts = np.ascontiguousarray(pts, np.int32)
        ref_cv = cv2.boxPoints(cv2.minAreaRect(pts))
        if min(cv2.minAreaRect(pts)[1]) < 1e-6:
            continue   # degenerate (collinear) sets: cv2's angle is arbitrary
        box_cv = np.zeros((4, 2), np.float32)
        box_o = np.zeros((4, 2), np.float32)
        shim.shim_generate_box(pts.ctypes.data_as(C.c_void_p), len(pts), box_cv.ctypes.data_as(C.c_void_p),
                               box_o.ctypes.data_as(C.c_void_p))
        if _set_dist(ref_cv, box_cv) > 1e-3:
            continue   # equal-area tie: a different rectangle (counted elsewhere)
        assert np.abs(ref_cv - box_cv).max() < 1e-3, (i, ref_cv, box_cv)
        want = order_points_clockwise(ref_cv)
        s = ref_cv.sum(1)
        d = ref_cv[:, 1] - ref_cv[:, 0]
        noisy_tie = any(0 < abs(v[i] - v[j]) < 1e-4 for v in (s, d) for i in range(4) for j in range(i))
        if not noisy_tie:   # cv2's float32 noise decides near-ties; exact ties and clear cases must match
            assert np.abs(want - box_o).max() < 1e-3, (i, want, box_o)
        if i % 3 == 0:
            n_diamond += 1
    assert n_diamond > 300


def _cross(o, a, b):
    return (a[0] - o[0]) * (b[1] - o[1]) - (a[1] - o[1]) * (b[0] - o[0])


def _hull_sorted(pts):
    """hull_sorted32 (dev_geom.cuh) / geom::hull_sorted: monotone chain over points sorted by (y, x)."""
    out = []
    for i, q in enumerate(pts):
        if i > 0 and q == pts[i - 1]:
            continue
        while len(out) >= 2 and _cross(out[-2], out[-1], q) <= 0:
            out.pop()
        out.append(q)
    if len(out) == 1:
        return out
    lo = len(out) + 1
    for i in range(len(pts) - 2, -1, -1):
        q = pts[i]
        if q == pts[i + 1]:
            continue
        while len(out) >= lo and _cross(out[-2], out[-1], q) <= 0:
            out.pop()
        out.append(q)
    return out[:-1]


def _hull_row_extents(ext):
    """hull_row_extents32 (dev_geom.cuh) / the loops of db_hull_kernel: the first pass visits right extents only,
    the second left extents only."""
    out = [ext[0][0]]
    prev = ext[0][0]
    for l, r in ext:
        if r == prev:
            continue
        prev = r
        while len(out) >= 2 and _cross(out[-2], out[-1], r) <= 0:
            out.pop()
        out.append(r)
    if len(out) == 1:
        return out
    lo = len(out) + 1
    for l, r in reversed(ext):
        if l == prev:
            continue
        prev = l
        while len(out) >= lo and _cross(out[-2], out[-1], l) <= 0:
            out.pop()
        out.append(l)
    return out[:-1]


def test_row_extent_hull_equals_generic_monotone_chain():
    """The CUDA hull kernels visit each row extent once per chain (right extents going down, left extents coming
    back). Same vertices in the same order as the generic chain over all 2*rows points, for any row-extent set:
    ragged blobs, single-pixel rows, straight edges, one-row and one-column sets."""
    rng = np.random.default_rng(0)
    for it in range(6000):
        nrows = int(rng.integers(1, 40))
        y0 = int(rng.integers(0, 50))
        mode = it % 4
        ext = []
        c = int(rng.integers(20, 60))
        for i in range(nrows):
            if mode == 0:
                l = int(rng.integers(0, 80)); r = l + int(rng.integers(0, 30))
            elif mode == 1:      # smooth blob
                c += int(rng.integers(-2, 3)); w = int(rng.integers(0, 12)); l, r = c - w, c + w
            elif mode == 2:      # many single-pixel rows / straight diagonals
                l = r = 10 + i * int(rng.integers(0, 3))
            else:                # straight left edge, ragged right edge
                l = 5; r = 5 + int(rng.integers(0, 20))
            ext.append(((l, y0 + i), (r, y0 + i)))
        flat = [p for pair in ext for p in pair]
        assert _hull_row_extents(ext) == _hull_sorted(flat), ext


def test_scan_kernel_fixed_point_trick():
    """db_scan_kernel turns a probability f in [0,1] into round(f * 2^23) with `bits(f + 1.0f) - 0x3f800000` (one FADD,
    one IADD) and flags everything else through the same number: restated in numpy - the trick is exact
    round-to-nearest-even on the 2^-23 grid for every float32 in [0,1], and any value outside [0,1] (negative, > 1,
    NaN, +-Inf) yields a result > 2^23, which is what sets OCRPP_IMG_VALUE_OUT_OF_RANGE."""
    rng = np.random.default_rng(0)
    f = np.concatenate([rng.random(1 << 20, dtype=np.float32),
                        np.float32(2.0) ** -np.arange(1, 60, dtype=np.float32),          # tiny values and ties
                        (np.arange(0, 4096, dtype=np.float32) + np.float32(0.5)) * np.float32(2.0 ** -23),
                        np.array([0.0, 1.0, np.nextafter(np.float32(1), np.float32(0)), 1e-45], np.float32)])
    with np.errstate(invalid="ignore"):
        u = (f + np.float32(1.0)).astype(np.float32).view(np.uint32) - np.uint32(0x3f800000)
    assert np.array_equal(u.astype(np.int64), np.rint(f.astype(np.float64) * 2.0 ** 23).astype(np.int64))
    one_2ulp = np.nextafter(np.nextafter(np.float32(1), np.float32(2)), np.float32(2))
    bad = np.array([-1e-45, -1e-8, -0.5, -3.0, one_2ulp, 1.5, 2.0, 1e30, np.inf, -np.inf, np.nan], np.float32)
    with np.errstate(invalid="ignore", over="ignore"):
        ub = ((bad + np.float32(1.0)).astype(np.float32).view(np.uint32) - np.uint32(0x3f800000)).astype(np.uint32)
    # -1e-45 and -1e-8 round to +1.0f exactly (they are below half an ulp of 1.0): indistinguishable from 0 and harmless,
    # as is 1 + 1 ulp, which rounds to 2.0f like 1.0 does; everything further out is flagged
    assert (ub[2:] > (1 << 23)).all() and (ub[:2] == 0).all()
    one_1ulp = np.array([np.nextafter(np.float32(1), np.float32(2))], np.float32)
    assert int(((one_1ulp + np.float32(1.0)).view(np.uint32) - np.uint32(0x3f800000))[0]) == 1 << 23


def test_dilated_scan_word_formula_equals_cv2_dilate():
    """use_dilation: db_scan_kernel<kDilate> builds the dilated mask from the raw threshold bits in the lane-major
    layout (word k of a 128-pixel group, bit `lane` = pixel 4*lane + k): pixel x-1 is word k-1 at the same bit, or
    word 3 shifted up by one lane (with the last bit of the previous group carried in) for k = 0; rows y and y-1 are
    ORed. Restated word for word in numpy and compared with cv2.dilate(seg, [[1,1],[1,1]])."""
    rng = np.random.default_rng(3)
    epl = 4
    for W in (128, 256, 200, 131, 1280):
        H = 9
        seg = (rng.random((H, W)) > 0.7).astype(np.uint8)
        ng = (W + 32 * epl - 1) // (32 * epl)

        def pack(row):
            words = np.zeros((ng, epl), np.uint64)
            for x in np.nonzero(row)[0]:
                g, r = divmod(int(x), 32 * epl)
                words[g, r % epl] |= np.uint64(1) << np.uint64(r // epl)
            return words
        raw = [pack(seg[y]) for y in range(H)]
        got = np.zeros_like(seg)
        M32 = np.uint64(0xffffffff)
        for y in range(H):
            cy = cu = np.uint64(0)
            for g in range(ng):
                ry = raw[y][g]
                ru = raw[y - 1][g] if y > 0 else np.zeros(epl, np.uint64)
                for k in range(epl):
                    py = (((ry[epl - 1] << np.uint64(1)) | cy) & M32) if k == 0 else ry[k - 1]
                    pu = (((ru[epl - 1] << np.uint64(1)) | cu) & M32) if k == 0 else ru[k - 1]
                    m = int(ry[k] | py | ru[k] | pu)
                    for lane in range(32):
                        x = g * 32 * epl + lane * epl + k
                        if x < W and (m >> lane) & 1:
                            got[y, x] = 1
                cy, cu = ry[epl - 1] >> np.uint64(31), ru[epl - 1] >> np.uint64(31)
        want = cv2.dilate(seg, np.array([[1, 1], [1, 1]], np.uint8))
        assert np.array_equal(got, want), W


def test_db_rescale_padding_resize_matches_the_reference_affine_map(shim):
    """use_padding_resize (db_postprocess.cpp:111-145,293-302): the closed form in geom::db_rescale equals the
    reference's get_affine_transform (cv::getAffineTransform on three float32 point pairs, inverse direction) +
    transform_preds (float64 product, cast to float32), for landscape, portrait and square sources."""
    shim.shim_db_rescale.argtypes = [C.c_float, C.c_float, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_void_p]
    rng = np.random.default_rng(2)
    f = np.float32
    for src_w, src_h, side in ((640, 480, 256), (600, 900, 256), (512, 512, 128), (1279, 733, 736), (31, 977, 320)):
        center = np.array([f(src_w / 2.0), f(src_h / 2.0)], np.float32)
        img_max = f(src_w if src_w > src_h else src_h)
        s_tri, d_tri = np.zeros((3, 2), np.float32), np.zeros((3, 2), np.float32)
        s_tri[0] = center
        s_tri[1] = center + np.array([0, img_max / 2.0], np.float32)
        d_tri[0] = (f(side) / 2.0, f(side) / 2.0)
        d_tri[1] = d_tri[0] + np.array([0, f(side) / 2.0], np.float32)
        d_tri[2] = (0, 0)
        s_tri[2] = (0, center[1] - center[0]) if center[0] >= center[1] else (center[0] - center[1], 0)
        warp = cv2.getAffineTransform(d_tri, s_tri).T          # map -> source, as the reference builds it (inv=1)
        out = np.zeros(2, np.float32)
        for _ in range(200):
            mx, my = f(rng.uniform(0, side)), f(rng.uniform(0, side))
            want = np.array([[float(mx), float(my), 1.0]], np.float64) @ warp
            shim.shim_db_rescale(mx, my, side, side, f(src_w), f(src_h), 1, out.ctypes.data_as(C.c_void_p))
            assert abs(float(out[0]) - float(f(want[0, 0]))) <= 1e-3 and abs(float(out[1]) - float(f(want[0, 1]))) <= 1e-3
        # plain scaling branch (:303-311), float32 in source order
        shim.shim_db_rescale(f(100.25), f(7.5), side, side, f(src_w), f(src_h), 0, out.ctypes.data_as(C.c_void_p))
        assert out[0] == f(f(f(100.25) / f(side)) * f(src_w)) and out[1] == f(f(f(7.5) / f(side)) * f(src_h))


def test_fill_quad_rows_equals_cv2_fillpoly(shim):
    """score_mode "box": geom::fill_quad_rows restates cv2.fillPoly(mask, quad, 1) (LINE_8) for the quads box_score
    builds. Exact whenever the integer quad lies inside the mask (every mini box that does not leave the map); for
    quads that leave it the mask depends on how the OpenCV build at hand rebuilds edges from clipLine's end points
    (that code changed between the reference's pinned 4.1.2 and the 4.13 of this image): a few masks in a thousand
    differ there, which the GPU parity tests classify (box on the frame) instead of hiding."""
    rng = np.random.default_rng(9)
    W, H = 200, 120
    stats = {"in": [0, 0], "out": [0, 0]}
    for t in range(6000):
        cx, cy = rng.uniform(-2, W + 2), rng.uniform(-2, H + 2)
        wd, ht = (rng.uniform(0.5, 80), rng.uniform(0.5, 30)) if t % 4 else (rng.uniform(0.5, 3), rng.uniform(0.5, 60))
        pts = cv2.boxPoints(((cx, cy), (wd, ht), rng.uniform(0, 180))).astype(np.float32)
        # db_postprocess.py:183-192
        xmin = int(np.clip(np.floor(pts[:, 0].min()), 0, W - 1))
        xmax = int(np.clip(np.ceil(pts[:, 0].max()), 0, W - 1))
        ymin = int(np.clip(np.floor(pts[:, 1].min()), 0, H - 1))
        ymax = int(np.clip(np.ceil(pts[:, 1].max()), 0, H - 1))
        q = pts.copy()
        q[:, 0] -= xmin
        q[:, 1] -= ymin
        qi = q.astype(np.int32)
        w, h = xmax - xmin + 1, ymax - ymin + 1
        mask = np.zeros((h, w), np.uint8)
        cv2.fillPoly(mask, [qi], 1)
        rect = np.zeros(4, np.int32)
        Lr, Rr = np.zeros(h + 4, np.int32), np.zeros(h + 4, np.int32)
        shim.shim_box_score_rows(np.ascontiguousarray(pts).ctypes.data_as(C.c_void_p), W, H,
                                 rect.ctypes.data_as(C.c_void_p), Lr.ctypes.data_as(C.c_void_p), Rr.ctypes.data_as(C.c_void_p))
        assert rect.tolist() == [xmin, ymin, w, h]
        got = np.zeros((h, w), np.uint8)
        for y in range(h):
            if Rr[y] >= Lr[y]:
                got[y, Lr[y]:Rr[y] + 1] = 1
        inside = bool(((qi[:, 0] >= 0) & (qi[:, 0] < w) & (qi[:, 1] >= 0) & (qi[:, 1] < h)).all())
        k = "in" if inside else "out"
        stats[k][0] += 1
        if not np.array_equal(mask, got):
            stats[k][1] += 1
    print(stats)
    assert stats["in"][0] > 2000 and stats["in"][1] == 0, stats
    assert stats["out"][0] > 1500 and stats["out"][1] <= 0.01 * stats["out"][0], stats


def _ref_rec_preprocess(part_img, mode, image_shape):
    from oracle.rec_prep_oracle import rec_preprocess
    return rec_preprocess(part_img, mode, image_shape)


def test_rec_preprocess_equals_cv2(shim):
    """pytorchocr_b200/csrc/prep.cuh (the code the rec_preprocess kernel runs per output pixel) against the
    reference's cvtColor + cv2.resize + normalise + pad on random crops: float32 tensors bit for bit."""
    rng = np.random.default_rng(4)
    shim.shim_rec_preprocess.restype = C.c_int
    n = 0
    for t in range(400):
        if t % 10 == 0:
            h, w = 64, 2 * int(rng.integers(3, 150))          # exact 2x2 decimation
        elif t % 10 == 1:
            h, w = 32, int(rng.integers(2, 400))               # no vertical resize
        else:
            h, w = int(rng.integers(2, 90)), int(rng.integers(2, 500))
        crop = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        if t % 7 == 0:
            crop = cv2.GaussianBlur(crop, (0, 0), 2.0)
        for mode, name, shape in ((0, "GRAY", (1, 32, 320)), (1, "RGB", (3, 32, 320)), (2, "BGR", (3, 32, 100))):
            want = _ref_rec_preprocess(crop, name, shape)
            got = np.empty(shape, np.float32)
            shim.shim_rec_preprocess(crop.ctypes.data_as(C.c_void_p), h, w, 3, mode, shape[1], shape[2], got.ctypes.data_as(C.c_void_p))
            assert np.array_equal(got, want), (t, h, w, name, np.abs(got - want).max(), int((got != want).sum()))
            n += 1
    assert n == 1200
