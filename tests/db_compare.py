"""Shared comparison logic for DB parity tests (GPU result vs the cv2-based oracle).

Gates (BASELINE.json north_star): box scores 1e-5 relative, vertices 1e-3 px on the pre-rounding
floats, integer boxes exact outside a +-2e-3 band around .5.

The reference algorithm itself is discontinuous in three places, and cv2's float32 rotating
calipers differ from our exact-integer/fp64 rectangle by up to ~1e-4 px BEFORE those
discontinuities, so a box is classified instead of failed when the oracle's own data shows it sits
on one of them:
  * truncation - UnClip casts the 4 mini-box corners to int (db_postprocess.cpp:42-45): a corner
                 coordinate within 2e-3 of an integer may truncate differently
  * rounding   - Clipper rounds every offset vertex to an integer (clipper.cpp:136-140): the offset
                 polygon changes when `distance` moves by 1e-4 relative
  * ordering   - GetMiniBoxes sorts corners by x only (std::sort, unstable): for a "diamond" (two
                 corners with equal x) the TL/TR/BR/BL order is not defined
  * tie        - two edge-aligned rectangles of exactly equal area (tiny symmetric blobs): which one
                 cv::minAreaRect returns depends on float32 noise
Each class is counted and bounded by the caller; classified boxes must still be within 2.5 px.
"""
import numpy as np

from oracle import geometry_oracle as G


def canonical_labels(lab):
    """Renumber labels by the raster order of each label's first pixel (0 stays 0)."""
    flat = lab.ravel()
    idx = np.nonzero(flat)[0]
    if idx.size == 0:
        return lab.copy()
    labs = flat[idx]
    first = np.full(int(flat.max()) + 1, -1, np.int64)
    first[labs[::-1]] = idx[::-1]
    present = np.nonzero(first >= 0)[0]
    order = present[np.argsort(first[present])]
    remap = np.zeros(int(flat.max()) + 1, np.int64)
    remap[order] = np.arange(1, len(order) + 1)
    return remap[lab]


def _set_dist(a, b):
    d = np.abs(np.asarray(a)[:, None, :] - np.asarray(b)[None, :, :]).max(-1)
    return max(d.min(1).max(), d.min(0).max())


def classify(d):
    """Which discontinuities of the reference algorithm the oracle box `d` sits on."""
    cls = set()
    mini = np.asarray(d["mini"], np.float64)
    if np.abs(mini - np.round(mini)).min() < 2e-3:
        cls.add("truncation")
    for quad in (mini, np.asarray(d["clip"], np.float64)):
        if np.min(np.diff(np.sort(quad[:, 0]))) < 1e-3:
            cls.add("ordering")
    dist = float(d["distance"])
    base = G.do_offset(d["quad"], dist)
    if any(G.do_offset(d["quad"], dist * (1 + e)) != base for e in (-1e-4, 1e-4)):
        cls.add("rounding")
    return cls


def compare_image(boxes, boxes_f, scores, details, tol_px=1e-3, tol_score=1e-5):
    """boxes int16 [K,4,2], boxes_f float32 [K,4,2], scores [K] from the GPU; details from
    oracle.db_oracle.boxes_from_bitmap(return_details=True).
    Returns a dict of counters: exact, classified (by class), unmatched_gpu, unmatched_oracle."""
    ok = [d for d in details if d["status"] == "ok"]
    stats = {"n_oracle": len(ok), "n_gpu": len(boxes), "exact": 0, "truncation": 0, "rounding": 0,
             "ordering": 0, "tie": 0, "unmatched_gpu": 0, "unmatched_oracle": 0}
    if not ok or not len(boxes):
        stats["unmatched_gpu"], stats["unmatched_oracle"] = len(boxes), len(ok)
        return stats
    of = np.array([d["out_f"] for d in ok], np.float64)           # [K,4,2]
    gf = np.asarray(boxes_f, np.float64)
    # corner-order independent distance between boxes
    dist = np.array([[_set_dist(of[i], gf[j]) for j in range(len(gf))] for i in range(len(of))])
    used = set()
    for i in np.argsort(dist.min(1)):
        cand = [(dist[i, j], j) for j in range(len(gf)) if j not in used]
        if not cand or min(cand)[0] > 2.5:
            stats["unmatched_oracle"] += 1
            continue
        dd, j = min(cand)
        used.add(j)
        d = ok[i]
        assert abs(scores[j] - d["score"]) <= tol_score * abs(d["score"]) + 1e-7, (scores[j], d["score"])
        ordered = np.abs(of[i] - gf[j]).max()
        if ordered < tol_px:
            stable = np.abs(of[i] - np.floor(of[i]) - 0.5) > 2e-3
            assert np.array_equal(np.asarray(d["out"])[stable], np.asarray(boxes[j], np.int64)[stable]), (d["out"], boxes[j])
            stats["exact"] += 1
            continue
        cls = classify(d)
        if dd < tol_px:            # same rectangle, corners in a rotated order
            assert "ordering" in cls, ("corner order differs on a non-diamond", d["out_f"], gf[j].tolist())
            stats["ordering"] += 1
        elif cls & {"truncation", "rounding"}:
            stats["truncation" if "truncation" in cls else "rounding"] += 1
        else:
            # different rectangle of the same area around the same points: equal-area tie
            def area(b):
                return np.linalg.norm(b[1] - b[0]) * np.linalg.norm(b[2] - b[1])
            assert abs(area(of[i]) - area(gf[j])) <= 2e-3 * area(of[i]) + 0.6, ("box mismatch", dd, d["out_f"], gf[j].tolist(), d["mini"])
            stats["tie"] += 1
    stats["unmatched_gpu"] = len(gf) - len(used)
    return stats


def merge(total, s):
    for k, v in s.items():
        total[k] = total.get(k, 0) + v
    return total
