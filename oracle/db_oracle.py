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


def get_mini_boxes(rect):
    """db_postprocess.cpp:159-192. rect = ((cx,cy),(w,h),angle) float32 semantics of cv::RotatedRect."""
    ssid = max(np.float32(rect[1][0]), np.float32(rect[1][1]))
    pts = cv2.boxPoints(rect)  # float32 [4,2]
    order = sorted(range(4), key=lambda i: pts[i][0])  # XsortFp32: by x only (stable here)
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


def box_score(contour, pred):
    """db_postprocess.cpp:194-229. contour int [n,2] (x,y)."""
    h, w = pred.shape
    xs, ys = contour[:, 0], contour[:, 1]
    xmin, xmax, ymin, ymax = int(xs.min()), int(xs.max()), int(ys.min()), int(ys.max())
    xmax = min(max(xmax, 0), w - 1)
    xmin = max(min(xmin, w - 1), 0)
    ymax = min(max(ymax, 0), h - 1)
    ymin = max(min(ymin, h - 1), 0)
    poly = contour.astype(np.int32).copy()
    poly[:, 0] -= xmin
    poly[:, 1] -= ymin
    mask = np.zeros((ymax - ymin + 1, xmax - xmin + 1), np.uint8)
    cv2.fillPoly(mask, [poly], 1, 1)  # lineType = 1 -> 4-connected boundary (see module docstring)
    cropped = np.ascontiguousarray(pred[ymin:ymax + 1, xmin:xmax + 1])
    score = cv2.mean(cropped, mask)[0]
    return score, int(mask.sum())


def boxes_from_bitmap(pred, bitmap, box_thresh, unclip_ratio, src_w, src_h,
                      use_padding_resize=False, return_details=False):
    """BoxesFromBitmap, db_postprocess.cpp:231-317. pred f32 [H,W]; bitmap uint8 [H,W].
    Returns list of 4x2 int lists (and, optionally, per-candidate details for parity tests)."""
    min_size, max_candidates = 3, 1000
    pred = np.ascontiguousarray(pred, np.float32)
    bitmap = np.ascontiguousarray(bitmap, np.uint8)
    height, width = bitmap.shape
    contours, _ = cv2.findContours(bitmap, cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)
    num_contours = min(len(contours), max_candidates)
    boxes, details = [], []
    f = np.float32
    for ci in range(num_contours):
        contour = contours[ci].reshape(-1, 2)
        d = {"contour_first": (int(contour[0][0]), int(contour[0][1])), "npts": len(contour), "status": "ok"}
        if return_details:   # what the parity tests need to enumerate exact equal-area ties (tests/db_compare.py)
            d["contour"] = contour.copy()
            d["unclip_ratio"], d["scale"] = unclip_ratio, (width, height, src_w, src_h)
        details.append(d)
        if len(contour) <= 2:
            d["status"] = "le2pts"
            continue
        rect = cv2.minAreaRect(contour)
        array, ssid = get_mini_boxes(rect)
        d["rect"], d["mini"], d["ssid"] = rect, array, float(ssid)
        if ssid < min_size:
            d["status"] = "small"
            continue
        score, cnt = box_score(contour, pred)
        d["score"], d["fill_count"] = score, cnt
        if score < box_thresh:
            d["status"] = "lowscore"
            continue
        rect2, distance, quad, soln = unclip(array, unclip_ratio)
        d["distance"], d["quad"], d["rect2"] = float(distance), quad, rect2
        if rect2[1][1] < 1.001 and rect2[1][0] < 1.001:
            d["status"] = "unclip_empty"
            continue
        cliparray, ssid2 = get_mini_boxes(rect2)
        d["clip"], d["ssid2"] = cliparray, float(ssid2)
        if ssid2 < min_size + 2:
            d["status"] = "small2"
            continue
        out, out_f = [], []
        if use_padding_resize:
            # db_postprocess.cpp:293-302: inverse padding-resize affine map (get_affine_transform :111-135 with
            # float32 points, cv::getAffineTransform in double) and transform_preds (:137-145, double product)
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
            warp = cv2.getAffineTransform(d_tri, s_tri).T          # 3x2, float64
        for j in range(4):
            if use_padding_resize:
                new_pt = np.array([[float(cliparray[j][0]), float(cliparray[j][1]), 1.0]], np.float64) @ warp
                fx, fy = f(new_pt[0, 0]), f(new_pt[0, 1])
            else:
                fx = f(f(cliparray[j][0] / f(width)) * f(src_w))
                fy = f(f(cliparray[j][1] / f(height)) * f(src_h))
            out_f.append((float(fx), float(fy)))
            out.append([int(min(max(roundf(fx), 0.0), float(src_w))),
                        int(min(max(roundf(fy), 0.0), float(src_h)))])
        d["out_f"], d["out"] = out_f, out
        boxes.append(out)
    if return_details:
        return boxes, details
    return boxes


class DBPostProcessOracle(object):
    """db_postprocess.py:10-74 with cpp_speedup=True semantics (the configured path)."""

    def __init__(self, thresh=0.3, box_thresh=0.5, max_candidates=1000, unclip_ratio=1.5,
                 use_dilation=False, score_mode="poly", cpp_speedup=True, out_polygon=False, **kwargs):
        assert score_mode in ["box", "poly"]
        self.thresh, self.box_thresh, self.max_candidates = thresh, box_thresh, max_candidates
        self.unclip_ratio = unclip_ratio
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
                                    self.unclip_ratio, src_w, src_h, use_padding_resize, return_details)
            tmp = out[0] if return_details else out
            d = {"points": np.array(tmp, dtype=np.int16), "scores": [1.0] * len(tmp)}
            if return_details:
                d["details"] = out[1]
            res.append(d)
        return res
