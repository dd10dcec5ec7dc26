"""TEST INFRASTRUCTURE ONLY (oracle). DB / DB++ post-processing: CPU restatement of the CONFIGURED
reference path (cpp_speedup=True):

    R/pytocr/postprocess/db_postprocess.py:40-74            DBPostProcess.__call__
    R/pytocr/postprocess/db_postprocess_fast/__init__.py:10-22   cpp_boxes_from_bitmap (uint8 cast)
    R/pytocr/postprocess/db_postprocess_fast/src/db_postprocess.cpp:231-317  BoxesFromBitmap
        :16-32 GetContourArea (float32), :34-64 UnClip, :159-192 GetMiniBoxes, :194-229 BoxScore

The C++ module itself cannot be compiled in this image (needs OpenCV C++ headers/libs,
include/db_postprocess.h:3-5), so it is restated line by line with the SAME OpenCV algorithms via
cv2-python 4.13 (findContours/minAreaRect/boxPoints/fillPoly/mean) and the reference's OWN Clipper
(oracle/_ref/libclipper_ref.so, built from R/.../src/clipper.cpp by oracle/build_ref.py). When that
.so is absent (it always travels with gpurun), the DoOffset restatement in geometry_oracle.py is
used instead and `CLIPPER_KIND` says so.

Details that matter and are easy to miss:
  * BoxScore calls cv::fillPoly(mask, pts, Scalar(1), 1): the 4th argument is lineType = 1, which
    OpenCV's Line() maps to a 4-CONNECTED boundary (db_postprocess.cpp:222). The mask is therefore
    the filled contour plus one extra "stair" pixel per diagonal contour step. We pass lineType=1
    to cv2.fillPoly to get exactly that (SURVEY.md A.2 described the lineType=8 mask).
  * GetMiniBoxes uses max(w,h) as `ssid` (:161), std::sort on x only (:147-151,168; unstable - we
    use a stable sort, generators avoid exact ties).
  * the Python wrapper discards scores (all 1.0) and returns int16 (db_postprocess.py:64-69).
"""
import ctypes as C
import math
import os

import cv2
import numpy as np

from . import geometry_oracle as G

_HERE = os.path.dirname(os.path.abspath(__file__))
_CLIPPER = None
CLIPPER_KIND = None


def _clipper():
    global _CLIPPER, CLIPPER_KIND
    if CLIPPER_KIND is None:
        path = os.path.join(_HERE, "_ref", "libclipper_ref.so")
        if os.path.exists(path):
            L = C.CDLL(path)
            L.clipper_ref_offset.argtypes = [C.POINTER(C.c_int64), C.c_int, C.c_double,
                                             C.POINTER(C.c_int64), C.c_int, C.POINTER(C.c_int), C.c_int]
            L.clipper_ref_offset.restype = C.c_int
            _CLIPPER, CLIPPER_KIND = L, "reference"
        else:
            CLIPPER_KIND = "port"
    return _CLIPPER


def clipper_offset(quad_int, delta):
    """ClipperOffset.AddPath(jtRound, etClosedPolygon) + Execute (db_postprocess.cpp:41-51).
    Returns a list of paths (each int64 [n,2])."""
    L = _clipper()
    if L is None:
        pts = G.do_offset(quad_int, float(delta))
        return [np.array(pts, np.int64).reshape(-1, 2)] if len(pts) else []
    xy = np.ascontiguousarray(np.asarray(quad_int, np.int64).reshape(-1))
    cap = 4096
    out = np.zeros(2 * cap, np.int64)
    plen = np.zeros(16, np.int32)
    n = L.clipper_ref_offset(xy.ctypes.data_as(C.POINTER(C.c_int64)), len(xy) // 2, float(delta),
                             out.ctypes.data_as(C.POINTER(C.c_int64)), cap,
                             plen.ctypes.data_as(C.POINTER(C.c_int)), 16)
    assert n >= 0, "clipper shim capacity exceeded"
    paths, o = [], 0
    for j in range(n):
        paths.append(out[2 * o:2 * (o + plen[j])].reshape(-1, 2).copy())
        o += plen[j]
    return paths


def roundf(v):
    """C roundf on a float32 value: half away from zero."""
    v = float(v)
    return math.copysign(math.floor(abs(v) + 0.5), v)


# ----------------------------------------------------------------------------------------------------------
# Stages shared by the reference's two branches. The C++ branch (cpp_speedup=True, db_postprocess.cpp) and the
# pure-Python branch (cpp_speedup=False, db_postprocess.py:76-194) run the same sequence of OpenCV / Clipper calls
# and differ in the details below, each an explicit switch so that ONE body of code is pinned by both
# tests/golden/reference_db_python.npz (the unmodified Python branch, run by make_golden.py) and the C++ reading:
#
#   switch          C++ branch (semantics="cpp")                      Python branch (semantics="python")
#   side            ssid = max(w, h)        (cpp:161)                 sside = min(w, h)            (py:176)
#   line_type       fillPoly(..., lineType=1) 4-connected (cpp:222)   fillPoly default LINE_8      (py:193)
#   score points    always the contour      (cpp:260)                 contour | mini box (score_mode) (py:109-112)
#   <=2-point skip  yes                     (cpp:252)                 no
#   distance        float32 shoelace/perimeter (cpp:16-32)            shapely area*ratio/length, float64 (py:145-146)
#   offset result   all paths' points, quirky loop (cpp:52-57)        exactly one path or skip     (py:116-118)
#   <1.001 skip     yes                     (cpp:273)                 no
#   2nd rectangle   minAreaRect(float points)                         minAreaRect(int points)      (py:119-123)
#   rounding        roundf, half away       (cpp:298-309)             np.round, half to even       (py:135-141)
#   max_candidates  constant 1000           (cpp:239)                 the constructor's value      (py:92)
#   scores          discarded by the wrapper (py:64-67 -> 1.0)        returned (float64 of cv2.mean)
# ----------------------------------------------------------------------------------------------------------
def get_mini_boxes(rect, side="max"):
    """db_postprocess.cpp:159-192 / db_postprocess.py:152-176. rect = ((cx,cy),(w,h),angle) as cv2 returns it.
    Corners sorted by x only (stable: Python's sorted, and libstdc++'s std::sort is an insertion sort below 16
    elements), then TL, TR, BR, BL."""
    w, h = np.float32(rect[1][0]), np.float32(rect[1][1])
    ssid = max(w, h) if side == "max" else min(w, h)
    pts = cv2.boxPoints(rect)  # float32 [4,2]
    order = sorted(range(4), key=lambda i: pts[i][0])
    a = [pts[i] for i in order]
    if a[3][1] <= a[2][1]:
        idx2, idx3 = a[3], a[2]
    else:
        idx2, idx3 = a[2], a[3]
    if a[1][1] <= a[0][1]:
        idx1, idx4 = a[1], a[0]
    else:
        idx1, idx4 = a[0], a[1]
    return np.array([idx1, idx2, idx3, idx4], np.float32), ssid


def get_contour_area(box, unclip_ratio):
    """db_postprocess.cpp:16-32, float32 arithmetic in source order."""
    f = np.float32
    area, dist = f(0), f(0)
    for i in range(4):
        j = (i + 1) % 4
        area = f(area + f(f(box[i][0] * box[j][1]) - f(box[i][1] * box[j][0])))
        dx, dy = f(box[i][0] - box[j][0]), f(box[i][1] - box[j][1])
        dist = f(dist + f(np.sqrt(f(f(dx * dx) + f(dy * dy)))))
    area = f(abs(f(float(area) / 2.0)))
    return f(f(area * f(unclip_ratio)) / dist)


def unclip_distance_py(box, unclip_ratio):
    """db_postprocess.py:144-146: shapely Polygon(box).area * unclip_ratio / .length in float64 (GEOS'
    Area::ofRingSigned and Length::ofLine, restated in oracle/ref_shims.py)."""
    from . import ref_shims
    return ref_shims.ring_area(box) * unclip_ratio / ref_shims.ring_length(box)


def unclip(box, unclip_ratio):
    """db_postprocess.cpp:34-64 -> RotatedRect tuple."""
    distance = get_contour_area(box, unclip_ratio)
    quad = [(int(box[i][0]), int(box[i][1])) for i in range(4)]  # C int() truncation toward zero
    soln = clipper_offset(quad, float(distance))
    points = []
    if len(soln):
        last_n = len(soln[-1])
        for j in range(len(soln)):
            # reference loop bound is soln[soln.size()-1].size() for every j (:53-55)
            for i in range(min(last_n, len(soln[j]))):
                points.append((np.float32(soln[j][i][0]), np.float32(soln[j][i][1])))
    if len(points) <= 0:
        return ((0.0, 0.0), (1.0, 1.0), 0.0), distance, quad, soln
    return cv2.minAreaRect(np.array(points, np.float32)), distance, quad, soln


def unclip_py(box, unclip_ratio):
    """db_postprocess.py:143-150 (+ :116-118): returns (int32 [n,1,2] offset polygon or None, distance, quad, soln).
    pyclipper truncates the float32 corners toward zero (see oracle/ref_shims.py)."""
    distance = unclip_distance_py(box, unclip_ratio)
    quad = [(int(box[i][0]), int(box[i][1])) for i in range(4)]
    soln = clipper_offset(quad, float(distance))
    if len(soln) != 1:
        return None, distance, quad, soln
    return np.asarray(soln[0], np.int32).reshape(-1, 1, 2), distance, quad, soln


def box_score(points, pred, line_type=1):
    """db_postprocess.cpp:194-229 (integer contour, lineType 1) and db_postprocess.py:178-194 (`box_score`: the
    contour, or - score_mode "box" - the float32 mini box; default lineType 8). points [n,2] (x,y)."""
    h, w = pred.shape
    pts = np.array(points).copy()
    if np.issubdtype(pts.dtype, np.integer):
        xs, ys = pts[:, 0], pts[:, 1]
        xmin, xmax, ymin, ymax = int(xs.min()), int(xs.max()), int(ys.min()), int(ys.max())
        xmax = min(max(xmax, 0), w - 1)
        xmin = max(min(xmin, w - 1), 0)
        ymax = min(max(ymax, 0), h - 1)
        ymin = max(min(ymin, h - 1), 0)
        poly = pts.astype(np.int32)
        poly[:, 0] -= xmin
        poly[:, 1] -= ymin
    else:
        # py:183-191: floor/ceil of the float corners, clip into the map, subtract in float32, truncate to int32
        xmin = int(np.clip(np.floor(pts[:, 0].min()).astype(int), 0, w - 1))
        xmax = int(np.clip(np.ceil(pts[:, 0].max()).astype(int), 0, w - 1))
        ymin = int(np.clip(np.floor(pts[:, 1].min()).astype(int), 0, h - 1))
        ymax = int(np.clip(np.ceil(pts[:, 1].max()).astype(int), 0, h - 1))
        pts[:, 0] = pts[:, 0] - xmin
        pts[:, 1] = pts[:, 1] - ymin
        poly = pts.astype(np.int32)
    mask = np.zeros((ymax - ymin + 1, xmax - xmin + 1), np.uint8)
    cv2.fillPoly(mask, [poly.reshape(-1, 2)], 1, line_type)  # lineType 1 -> 4-connected boundary, 8 -> 8-connected
    cropped = np.ascontiguousarray(pred[ymin:ymax + 1, xmin:xmax + 1])
    score = cv2.mean(cropped, mask)[0]
    return score, int(mask.sum())


def _padding_resize_warp(src_w, src_h, height):
    """db_postprocess.cpp:111-135,293-296 == utility.py:81-109 get_affine_transform(inv=1): float32 triangle points,
    cv2.getAffineTransform in double. Returns the 2x3 float64 matrix (map -> source image)."""
    f = np.float32
    center = np.array([f(src_w / 2.0), f(src_h / 2.0)], np.float32)
    img_maxsize = f(src_w if src_w > src_h else src_h)
    square = f(height)
    s_tri = np.zeros((3, 2), np.float32)
    d_tri = np.zeros((3, 2), np.float32)
    s_tri[0] = center
    s_tri[1] = center + np.array([0, img_maxsize / 2.0], np.float32)
    d_tri[0] = (square / 2.0, square / 2.0)
    d_tri[1] = d_tri[0] + np.array([0, square / 2.0], np.float32)
    d_tri[2] = (0, 0)
    s_tri[2] = (0, center[1] - center[0]) if center[0] >= center[1] else (center[0] - center[1], 0)
    return cv2.getAffineTransform(d_tri, s_tri)


def boxes_from_bitmap(pred, bitmap, box_thresh, unclip_ratio, src_w, src_h,
                      use_padding_resize=False, return_details=False,
                      semantics="cpp", score_mode="poly", max_candidates=1000):
    """BoxesFromBitmap, db_postprocess.cpp:231-317 (semantics="cpp") or DBPostProcess.boxes_from_bitmap,
    db_postprocess.py:76-141 (semantics="python", out_polygon=False). pred f32 [H,W]; bitmap uint8 [H,W].
    Returns list of 4x2 int lists (and, optionally, per-candidate details for parity tests); in Python semantics
    the details carry the returned score of every box."""
    assert semantics in ("cpp", "python")
    py = semantics == "python"
    min_size = 3
    if not py:
        max_candidates = 1000
    pred = np.ascontiguousarray(pred, np.float32)
    bitmap = np.ascontiguousarray(bitmap, np.uint8)
    height, width = bitmap.shape
    contours, _ = cv2.findContours(bitmap * 255 if py else bitmap, cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)
    num_contours = min(len(contours), max_candidates)
    boxes, details = [], []
    f = np.float32
    for ci in range(num_contours):
        contour = contours[ci].reshape(-1, 2)
        d = {"contour_first": (int(contour[0][0]), int(contour[0][1])), "npts": len(contour), "status": "ok"}
        if return_details:   # what the parity tests need to enumerate exact equal-area ties (tests/db_compare.py)
            d["contour"] = contour.copy()
            d["unclip_ratio"], d["scale"] = unclip_ratio, (width, height, src_w, src_h)
            d["semantics"], d["score_mode"] = semantics, score_mode
        details.append(d)
        if not py and len(contour) <= 2:
            d["status"] = "le2pts"
            continue
        rect = cv2.minAreaRect(contours[ci])
        array, ssid = get_mini_boxes(rect, side="min" if py else "max")
        d["rect"], d["mini"], d["ssid"] = rect, array, float(ssid)
        if ssid < min_size:
            d["status"] = "small"
            continue
        if py:
            score, cnt = box_score(array if score_mode == "box" else contour, pred, line_type=8)
        else:
            score, cnt = box_score(contour, pred, line_type=1)
        d["score"], d["fill_count"] = score, cnt
        if score < box_thresh:
            d["status"] = "lowscore"
            continue
        if py:
            poly2, distance, quad, soln = unclip_py(array, unclip_ratio)
            d["distance"], d["quad"], d["offset"] = float(distance), quad, soln
            if poly2 is None:
                d["status"] = "unclip_empty"
                continue
            rect2 = cv2.minAreaRect(poly2)
            d["rect2"] = rect2
        else:
            rect2, distance, quad, soln = unclip(array, unclip_ratio)
            d["distance"], d["quad"], d["rect2"], d["offset"] = float(distance), quad, rect2, soln
            if rect2[1][1] < 1.001 and rect2[1][0] < 1.001:
                d["status"] = "unclip_empty"
                continue
        cliparray, ssid2 = get_mini_boxes(rect2, side="min" if py else "max")
        d["clip"], d["ssid2"] = cliparray, float(ssid2)
        if ssid2 < min_size + 2:
            d["status"] = "small2"
            continue
        out, out_f = [], []
        if use_padding_resize:
            warp = _padding_resize_warp(src_w, src_h, height)          # 2x3 float64
        for j in range(4):
            if use_padding_resize and py:
                # utility.py:111-121: float32 point, np.dot with the float64 matrix, stored into a float64 array
                new_pt = np.dot(warp, np.array([cliparray[j][0], cliparray[j][1], 1.0], dtype=np.float32).T)
                fx, fy = float(new_pt[0]), float(new_pt[1])
            elif use_padding_resize:
                # db_postprocess.cpp:137-145,297-302: double product, result narrowed to Point2f
                new_pt = np.array([[float(cliparray[j][0]), float(cliparray[j][1]), 1.0]], np.float64) @ warp.T
                fx, fy = f(new_pt[0, 0]), f(new_pt[0, 1])
            else:
                fx = f(f(cliparray[j][0] / f(width)) * f(src_w))
                fy = f(f(cliparray[j][1] / f(height)) * f(src_h))
            out_f.append((float(fx), float(fy)))
            if py:      # np.round: half to even; np.clip; astype(int16)
                out.append([int(np.int16(np.clip(np.round(fx), 0, src_w))),
                            int(np.int16(np.clip(np.round(fy), 0, src_h)))])
            else:
                out.append([int(min(max(roundf(fx), 0.0), float(src_w))),
                            int(min(max(roundf(fy), 0.0), float(src_h)))])
        d["out_f"], d["out"] = out_f, out
        boxes.append(out)
    if return_details:
        return boxes, details
    return boxes


class DBPostProcessOracle(object):
    """db_postprocess.py:10-74. `cpp_speedup=True` (the configured path, and this class's default) runs the C++
    branch's semantics, `cpp_speedup=False` the pure-Python branch's (scores returned, score_mode and max_candidates
    honoured)."""

    def __init__(self, thresh=0.3, box_thresh=0.5, max_candidates=1000, unclip_ratio=1.5,
                 use_dilation=False, score_mode="poly", cpp_speedup=True, out_polygon=False, **kwargs):
        assert score_mode in ["box", "poly"]
        assert not out_polygon, "out_polygon: the reference's branch raises on ragged polygons (DESIGN.md)"
        self.thresh, self.box_thresh, self.max_candidates = thresh, box_thresh, max_candidates
        self.unclip_ratio, self.score_mode, self.cpp_speedup = unclip_ratio, score_mode, cpp_speedup
        self.dilation_kernel = None if not use_dilation else np.array([[1, 1], [1, 1]])

    def __call__(self, outs_dict, shape_list, use_padding_resize=False, return_details=False):
        pred = outs_dict["maps"]
        if hasattr(pred, "detach"):
            pred = pred.detach().cpu().numpy()
        pred = pred[:, 0, :, :]
        segmentation = pred > self.thresh
        res = []
        for b in range(pred.shape[0]):
            src_h, src_w, ratio_h, ratio_w = shape_list[b]
            src_h, src_w = int(src_h), int(src_w)
            if self.dilation_kernel is not None:
                mask = cv2.dilate(np.array(segmentation[b]).astype(np.uint8), self.dilation_kernel)
            else:
                mask = segmentation[b]
            out = boxes_from_bitmap(pred[b].astype(np.float32), mask.astype(np.uint8), self.box_thresh,
                                    self.unclip_ratio, src_w, src_h, use_padding_resize, True,
                                    semantics="cpp" if self.cpp_speedup else "python",
                                    score_mode=self.score_mode, max_candidates=self.max_candidates)
            tmp = out[0]
            scores = [1.0] * len(tmp) if self.cpp_speedup else [x["score"] for x in out[1] if x["status"] == "ok"]
            d = {"points": np.array(tmp, dtype=np.int16), "scores": scores}
            if return_details:
                d["details"] = out[1]
            res.append(d)
        return res
