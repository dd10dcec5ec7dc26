"""Shared comparison logic for the PSE / PAN GPU parity tests.

Gates (BASELINE.json north_star): label maps bit-exact (ids are cv2's, no renumbering needed),
box scores 1e-5 relative, vertices 1e-3 px on the pre-rounding floats, integer boxes exact outside
a +-2e-3 band around .5. Boxes come out in label order on both sides, so they are compared
pairwise; a box whose rectangle differs but has the same area (an exact equal-area tie between two
edge-aligned rectangles) or whose corner order is ambiguous (order_points_clockwise
ties on x+y / y-x) is counted, not failed, and bounded by the caller."""
import cv2
import numpy as np

from oracle import geometry_oracle as G


def _set_dist(a, b):
    d = np.abs(np.asarray(a)[:, None, :] - np.asarray(b)[None, :, :]).max(-1)
    return max(d.min(1).max(), d.min(0).max())


def _subset_dist(a, b):
    """max over points of `a` of the distance to the nearest point of `b`"""
    d = np.abs(np.asarray(a)[:, None, :] - np.asarray(b)[None, :, :]).max(-1)
    return d.min(1).max()


def compare_image(boxes, boxes_f, scores, want, shape, tol_px=1e-3, tol_score=1e-5):
    det = want["details"]
    ratio = np.array([shape[3], shape[2]], np.float64)   # (ratio_w, ratio_h)
    stats = {"n": len(det), "exact": 0, "tie": 0, "ordering": 0}
    assert len(boxes) == len(det), "box count differs: gpu %d oracle %d" % (len(boxes), len(det))
    for j, d in enumerate(det):
        assert abs(float(scores[j]) - d["score"]) <= tol_score * abs(d["score"]) + 1e-7, (scores[j], d["score"])
        of = np.asarray(d["box_scaled"], np.float64)
        gf = np.asarray(boxes_f[j], np.float64)
        if np.abs(of - gf).max() < tol_px:
            stable = np.abs(of - np.floor(of) - 0.5) > 2e-3
            ob = np.asarray(want["points"][j], np.int64)
            assert np.array_equal(ob[stable], np.asarray(boxes[j], np.int64)[stable]), (ob, boxes[j], of)
            stats["exact"] += 1
        elif _set_dist(of, gf) < tol_px:
            stats["ordering"] += 1
        elif _subset_dist(gf, cv2.boxPoints(d["rect"]).astype(np.float64) / ratio) < tol_px:
            # same rectangle; order_points_clockwise met a tie (45-degree diamond) that cv2's float32
            # noise resolved differently, so the two boxes repeat different corners of it
            stats["ordering"] += 1
        elif any(_subset_dist(gf, cv2.boxPoints(r).astype(np.float64) / ratio) < 5e-3
                 for r in G.tied_min_area_rects(d["hull"], rel=2e-6)):
            # verified equal-area tie: the GPU box is made of corners of another minimal rectangle of the label
            stats["tie"] += 1
        elif min(d["rect"][1]) < 3.0 or d["area"] < 64:
            # tiny / degenerate label (a few pixels, 1-px lines, small triangles): several edge-aligned
            # rectangles have exactly the same area and order_points_clockwise may repeat a corner of a
            # diamond, so only closeness is required
            assert _set_dist(of, gf) <= 0.75 * max(d["rect"][1]) + 1.5, ("tiny box far off", of, gf, d["rect"])
            stats["tiny"] = stats.get("tiny", 0) + 1
        else:
            # a different rectangle: acceptable only as a near-tie below cv2's float32 resolution, and then it must be a
            # valid answer - enclose the label's hull and have (nearly) its minimal area
            from db_compare import encloses_minimally
            assert encloses_minimally(gf * ratio, d["hull"]), ("box mismatch", of, gf)
            stats["tie"] += 1
    return stats


def merge(total, s):
    for k, v in s.items():
        total[k] = total.get(k, 0) + v
    return total
