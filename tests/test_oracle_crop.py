"""CPU: the crop oracle (oracle/crop_oracle.py) pinned against cv2 and against the reference's own
sort_boxes / get_part_img outputs (tests/golden/reference_crops.npz, made by make_golden.py crops)."""
import os

import cv2
import numpy as np
import pytest

from conftest import ROOT
from oracle import crop_oracle as co
from pytorchocr_b200 import synth

G = np.load(os.path.join(ROOT, "tests", "golden", "reference_crops.npz"))


def _golden_crops():
    out, o = [], 0
    for rows, cols in G["dims"]:
        out.append(G["pixels"][o:o + rows * cols * 3].reshape(rows, cols, 3))
        o += rows * cols * 3
    return out


def test_sort_boxes_matches_reference():
    got = co.sort_boxes(G["boxes"])
    assert np.array_equal(np.asarray(got, np.int16), G["sorted_boxes"])
    assert np.array_equal(G["boxes"][co.sort_order(G["boxes"])], G["sorted_boxes"])


def test_crops_match_reference():
    for box, want in zip(G["sorted_boxes"], _golden_crops()):
        assert np.array_equal(co.crop_for_rec(G["img"], box), want)
        part = co.get_part_img_restated(G["img"], box)
        if part.shape[0] >= 1.5 * part.shape[1]:
            part = np.rot90(part, 1)
        assert np.array_equal(part, want)


def test_transform_matrices_bitwise():
    boxes = synth.page_boxes(5, n=300)
    for b in boxes:
        pts = b.astype(np.float32)
        l, t, r, bt = co.crop_rect(pts)
        pts = pts - np.array([l, t], np.float32)
        w, h = r - l, bt - t
        dst = np.array([[0, 0], [w - 1, 0], [w - 1, h - 1], [0, h - 1]], np.float32)
        M = cv2.getPerspectiveTransform(pts, dst)
        M2 = co.perspective_transform(pts, dst)
        assert np.array_equal(M, M2)
        assert np.array_equal(cv2.invert(M)[1], co.invert3(M2))


@pytest.mark.parametrize("seed", [0, 1])
def test_restated_warp_equals_cv2(seed):
    """every pixel of every crop: the first-principles restatement == cv2.warpPerspective"""
    img = synth.page_image(seed, 300, 400)
    boxes = synth.page_boxes(seed + 10, n=120, H=300, W=400, tall_frac=0.3, skew=3.0)
    boxes[::9] = _axis_aligned(boxes[::9])          # identity-like transforms: all fractions zero
    for b in boxes:
        assert np.array_equal(co.get_part_img(img, b), co.get_part_img_restated(img, b))


def _axis_aligned(b):
    out = b.copy()
    l, r = b[:, :, 0].min(1), b[:, :, 0].max(1)
    t, bt = b[:, :, 1].min(1), b[:, :, 1].max(1)
    out[:, 0], out[:, 1], out[:, 2], out[:, 3] = np.stack([l, t], 1), np.stack([r, t], 1), np.stack([r, bt], 1), np.stack([l, bt], 1)
    return out


def test_box_touching_page_border():
    """right == W / bottom == H: the slice is clipped, the taps replicate the crop's own border"""
    img = synth.page_image(3, 120, 160)
    b = np.array([[100, 90], [160, 95], [158, 120], [98, 116]], np.int16)
    assert np.array_equal(co.get_part_img(img, b), co.get_part_img_restated(img, b))


def test_collinear_box_is_rejected_by_restatement():
    with pytest.raises(ValueError):
        co.get_part_img_restated(synth.page_image(0, 64, 64), np.array([[0, 0], [10, 10], [20, 20], [30, 30]], np.int16))


def test_tied_min_area_rects_contains_cv2_choice():
    """oracle/geometry_oracle.tied_min_area_rects (used by the GPU comparators to verify equal-area ties): cv2's
    rectangle is always one of the enumerated minimal rectangles, and the known tied contour yields both."""
    from oracle import geometry_oracle as G
    tied = np.array([[132, 178], [133, 177], [134, 178], [134, 179], [135, 180], [134, 181], [131, 181], [130, 180],
                     [130, 179], [131, 178], [132, 179]])
    alts = G.tied_min_area_rects(tied)
    assert len(alts) == 2 and all(abs(w * h - 18.0) < 1e-9 for _, (w, h), _ in alts)
    rng = np.random.default_rng(0)
    for _ in range(200):
        pts = rng.integers(0, 40, (int(rng.integers(3, 30)), 2))
        if len(cv2.convexHull(pts.astype(np.int32))) < 3:
            continue
        want = np.sort(cv2.boxPoints(cv2.minAreaRect(pts.astype(np.float32))), axis=0)
        cands = G.tied_min_area_rects(pts, rel=2e-6, corners=True)
        assert any(np.abs(np.sort(c, axis=0) - want).max() < 1e-2 for c in cands), pts.tolist()


def test_sort_order_is_sort_boxes_as_a_permutation():
    """`sort_order` (the form the CUDA plan kernel implements: stable rank by (y, x, index) of the first corner, then
    ONE carried-element swap pass) reproduces `sort_boxes` on clustered rows, exact ties and long swap chains."""
    rng = np.random.default_rng(1)
    for it in range(400):
        n = int(rng.integers(1, 60))
        boxes = np.zeros((n, 4, 2), np.int16)
        boxes[:, 0, 1] = rng.choice([10, 14, 19, 21, 30, 39, 40, 41], n) if it % 2 else rng.integers(0, 100, n)
        boxes[:, 0, 0] = rng.integers(0, 12 if it % 3 == 0 else 300, n)
        boxes[:, 1:] = rng.integers(0, 300, (n, 3, 2))
        want = np.asarray(co.sort_boxes(boxes), np.int16).reshape(-1, 4, 2)
        assert np.array_equal(boxes[co.sort_order(boxes)], want)


def test_integer_bilinear_weights_equal_the_float32_table():
    """crop_warp_kernel computes cv2's int16 bilinear table entries as 32*(32-ay)*(32-ax) etc. (saturated at 32767)
    instead of saturate_cast<short>((1-fy)(1-fx) * 32768) in float32: identical for all 32 x 32 fractions."""
    f32 = np.float32
    for ay in range(32):
        for ax in range(32):
            fy, fx = f32(ay) / f32(32), f32(ax) / f32(32)
            tab = [(f32(1) - fy) * (f32(1) - fx), (f32(1) - fy) * fx, fy * (f32(1) - fx), fy * fx]
            tab = [int(np.clip(np.rint(t * f32(32768)), -32768, 32767)) for t in tab]
            ints = [min(32767, 32 * (32 - ay) * (32 - ax)), 32 * (32 - ay) * ax, 32 * ay * (32 - ax), 32 * ay * ax]
            assert tab == ints, (ay, ax)
            assert sum(ints) in (32768, 32767)
