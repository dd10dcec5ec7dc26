"""TEST INFRASTRUCTURE ONLY (oracle). The C++ branch of oracle/db_oracle.py (`semantics="cpp"`,
R/pytocr/postprocess/db_postprocess_fast/src/db_postprocess.cpp:231-317) restated for SPEED: the timing leg of
bench.py (`cpu_baseline`, `--impl reference`). Same OpenCV calls per contour (findContours, minAreaRect, boxPoints,
fillPoly, mean - cv2-python 4.13) and the reference's own compiled Clipper, but everything the reference does in
C++ between those calls (corner ordering, GetContourArea, the int() truncation, UnClip's point loop, the rescale)
runs vectorised over all contours of an image in numpy, and Clipper is entered once per image
(clipper_ref_offset_batch), so that the per-contour Python interpreter overhead of the line-by-line oracle (about
70 of its 80 ms per 736x1280 map) is not billed to the reference. tests/test_oracle_db.py checks that it returns
exactly the boxes of oracle/db_oracle.py."""
import ctypes as C

import cv2
import numpy as np

from . import db_oracle as O

_BATCH = None


def _batch_fn():
    global _BATCH
    if _BATCH is None:
        L = O._clipper()
        if L is None or not hasattr(L, "clipper_ref_offset_batch"):
            _BATCH = False
        else:
            f = L.clipper_ref_offset_batch
            f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
            f.restype = C.c_int
            _BATCH = f
    return _BATCH


def _mini_boxes(rects):
    """GetMiniBoxes for K rectangles at once -> (corners float32 [K,4,2] TL,TR,BR,BL, ssid float32 [K])."""
    K = len(rects)
    pts = np.empty((K, 4, 2), np.float32)
    wh = np.empty((K, 2), np.float32)
    for i, r in enumerate(rects):
        pts[i] = cv2.boxPoints(r)
        wh[i] = r[1]
    order = np.argsort(pts[:, :, 0], axis=1, kind="stable")
    a = np.take_along_axis(pts, order[:, :, None], axis=1)                       # sorted by x (stable)
    sw23 = a[:, 3, 1] <= a[:, 2, 1]
    sw01 = a[:, 1, 1] <= a[:, 0, 1]
    out = np.empty_like(a)
    out[:, 0] = np.where(sw01[:, None], a[:, 1], a[:, 0])
    out[:, 3] = np.where(sw01[:, None], a[:, 0], a[:, 1])
    out[:, 1] = np.where(sw23[:, None], a[:, 3], a[:, 2])
    out[:, 2] = np.where(sw23[:, None], a[:, 2], a[:, 3])
    return out, np.maximum(wh[:, 0], wh[:, 1])


def boxes_from_bitmap(pred, bitmap, box_thresh, unclip_ratio, src_w, src_h, use_padding_resize=False):
    f = np.float32
    fn = _batch_fn()
    if not fn or use_padding_resize:
        return O.boxes_from_bitmap(pred, bitmap, box_thresh, unclip_ratio, src_w, src_h, use_padding_resize)
    pred = np.ascontiguousarray(pred, np.float32)
    bitmap = np.ascontiguousarray(bitmap, np.uint8)
    height, width = bitmap.shape
    contours, _ = cv2.findContours(bitmap, cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)
    contours = [c for c in contours[:1000] if len(c) > 2]
    if not contours:
        return []
    mini, ssid = _mini_boxes([cv2.minAreaRect(c) for c in contours])
    keep = np.nonzero(ssid >= 3)[0]
    scores = np.empty(len(keep))
    for n, i in enumerate(keep):
        scores[n] = O.box_score(contours[i].reshape(-1, 2), pred, line_type=1)[0]
    keep = keep[scores >= box_thresh]
    if not len(keep):
        return []
    box = mini[keep]                                                              # [K,4,2] float32
    # GetContourArea in float32, source order (db_postprocess.cpp:16-32)
    nxt = np.roll(box, -1, axis=1)
    area, dist = np.zeros(len(box), f), np.zeros(len(box), f)
    for i in range(4):
        area = area + (box[:, i, 0] * nxt[:, i, 1] - box[:, i, 1] * nxt[:, i, 0])
        dx, dy = box[:, i, 0] - nxt[:, i, 0], box[:, i, 1] - nxt[:, i, 1]
        dist = dist + np.sqrt(dx * dx + dy * dy)
    area = np.abs((area.astype(np.float64) / 2.0).astype(f))
    distance = (area * f(unclip_ratio)) / dist
    quads = np.ascontiguousarray(np.trunc(box).astype(np.int64))                  # C int() truncation
    deltas = np.ascontiguousarray(distance.astype(np.float64))
    K = len(box)
    cap = 512 * K
    out_xy = np.empty(2 * cap, np.int64)
    start = np.empty(K + 1, np.int32)
    npaths = np.empty(K, np.int32)
    last = np.empty(K, np.int32)
    rc = fn(quads.ctypes.data, deltas.ctypes.data, K, out_xy.ctypes.data, cap, start.ctypes.data,
            npaths.ctypes.data, last.ctypes.data)
    assert rc == 0, "clipper batch capacity exceeded"
    pts_all = out_xy.reshape(-1, 2).astype(np.float32)
    rects2, keep2 = [], []
    for k in range(K):
        pts = pts_all[start[k]:start[k + 1]]
        r2 = cv2.minAreaRect(pts) if len(pts) else ((0.0, 0.0), (1.0, 1.0), 0.0)
        if r2[1][1] < 1.001 and r2[1][0] < 1.001:
            continue
        rects2.append(r2)
        keep2.append(k)
    if not rects2:
        return []
    clip, ssid2 = _mini_boxes(rects2)
    clip = clip[ssid2 >= 5]
    fx = (clip[:, :, 0] / f(width)) * f(src_w)
    fy = (clip[:, :, 1] / f(height)) * f(src_h)

    def rnd(v, hi):   # roundf (half away from zero), clamp, int
        r = np.copysign(np.floor(np.abs(v.astype(np.float64)) + 0.5), v)
        return np.minimum(np.maximum(r, 0.0), float(hi)).astype(np.int64)
    return np.stack([rnd(fx, src_w), rnd(fy, src_h)], axis=-1).tolist()
