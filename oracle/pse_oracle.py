"""TEST INFRASTRUCTURE ONLY (oracle). PSENet / PAN post-processing, CPU restatement of
R/pytocr/postprocess/pse_postprocess.py and pan_postprocess.py (operator level) on top of the
C restatements of pse.pyx / pa.pyx in oracle/c/ocr_oracle.c.

`generate_box` is restated WITHOUT the reference's O(labels*H*W) scans: pixels are grouped by
label with one stable sort, so each label's pixel list is in raster order exactly like
`np.where(label == i)` (pse_postprocess.py:71-73). All arithmetic that is visible in the results
(np.mean in float32 over the raster-ordered pixels :79, cv2.minAreaRect/boxPoints :85-86,
order_points_clockwise utility.py:21-29, /ratio + np.round + clip + int16 :100-102) is done with
the same numpy / cv2 calls on the same data as the reference.
"""
import cv2
import numpy as np

from . import clib


def order_points_clockwise(pts):
    """R/pytocr/utils/utility.py:21-29."""
    rect = np.zeros((4, 2), dtype=np.float32)
    s = pts.sum(axis=1)
    rect[0] = pts[np.argmin(s)]
    rect[2] = pts[np.argmax(s)]
    diff = np.diff(pts, axis=1)
    rect[1] = pts[np.argmin(diff)]
    rect[3] = pts[np.argmax(diff)]
    return rect


def sigmoid_f32(x):
    """F.sigmoid on float32 (pse_postprocess.py:38). Uses torch when importable so that the CPU
    value is torch's own; numpy otherwise (differs by <= 1 ulp)."""
    try:
        import torch
        return torch.sigmoid(torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))).numpy()
    except Exception:  # pragma: no cover
        x = x.astype(np.float32)
        return (1.0 / (1.0 + np.exp(-x))).astype(np.float32)


def upsample_nearest(a, f):
    """F.interpolate(mode='nearest', scale_factor=f) and cv2.resize(INTER_NEAREST) by an integer
    factor are both dst[y,x] = src[y//f, x//f]."""
    if f == 1:
        return a
    return np.repeat(np.repeat(a, f, axis=-2), f, axis=-1)


def generate_box(score, label, shape, min_area, box_thresh, return_details=False):
    """pse_postprocess.py:65-105 / pan_postprocess.py:73-113 (out_polygon=False branch)."""
    src_h, src_w, ratio_h, ratio_w = shape
    H, W = label.shape
    flat = label.ravel()
    order = np.argsort(flat, kind="stable")
    sorted_lab = flat[order]
    label_num = int(flat.max()) + 1 if flat.size else 1
    starts = np.searchsorted(sorted_lab, np.arange(1, label_num + 1), side="left")
    score_flat = score.ravel()
    boxes, scores, details = [], [], []
    for i in range(1, label_num):
        pix = order[starts[i - 1]:starts[i]]  # raster order
        if pix.size == 0:
            continue  # np.where empty -> points.shape[0]==0 < min_area -> continue
        if pix.size < min_area:
            continue
        score_i = np.mean(score_flat[pix])
        if score_i < box_thresh:
            continue
        ys, xs = np.divmod(pix, W)
        points = np.stack([xs, ys], axis=1).astype(np.int64)
        rect = cv2.minAreaRect(points)
        bbox = cv2.boxPoints(rect)
        bbox = order_points_clockwise(bbox)
        pre = bbox.copy()
        bbox[:, 0] = np.clip(np.round(bbox[:, 0] / ratio_w), 0, src_w)
        bbox[:, 1] = np.clip(np.round(bbox[:, 1] / ratio_h), 0, src_h)
        boxes.append(bbox.astype(np.int16))
        scores.append(score_i)
        if return_details:
            details.append({"label": i, "area": int(pix.size), "score": float(score_i),
                            "rect": rect, "box_f": pre, "hull": cv2.convexHull(points.astype(np.int32)).reshape(-1, 2),
                            "box_scaled": np.stack([pre[:, 0] / ratio_w, pre[:, 1] / ratio_h], 1)})
    boxes = np.array(boxes, dtype=np.int16)
    if return_details:
        return boxes, scores, details
    return boxes, scores


class PSEPostProcessOracle(object):
    """pse_postprocess.py:10-105."""

    def __init__(self, thresh=0.5, box_thresh=0.85, min_area=16, scale=4, out_polygon=False,
                 maps_at_processing_res=False, **kwargs):
        assert not out_polygon, "out_polygon is SURVEY 8(f) rank 4 (not on the configured path)"
        self.thresh, self.box_thresh, self.min_area, self.scale = thresh, box_thresh, min_area, scale
        # True: the maps are the tensor AFTER the reference's F.interpolate (:34-36), i.e. the
        # up-sampling step is skipped (bench configs 3/4 enter the path in that state)
        self.maps_at_processing_res = maps_at_processing_res

    def prepare(self, pred):
        """:31-45 -> score f32 [N,H,W], kernels u8 [N,K,H,W] at processing resolution."""
        if hasattr(pred, "detach"):
            pred = pred.detach().cpu().numpy()
        pred = np.asarray(pred, dtype=np.float32)
        if self.maps_at_processing_res:
            self.img_h, self.img_w = pred.shape[2] * self.scale, pred.shape[3] * self.scale
        else:
            self.img_h, self.img_w = pred.shape[2] * 4, pred.shape[3] * 4
            if self.scale != 4:
                pred = upsample_nearest(pred, 4 // self.scale)
        score = sigmoid_f32(pred[:, 0])
        kernels = (pred > self.thresh)
        kernels = (kernels & kernels[:, 0:1]).astype(np.uint8)
        return score, kernels

    def labels(self, kernels_one):
        return clib.pse(kernels_one, self.min_area / (self.scale ** 2))

    def __call__(self, outs_dict, shape_list, return_details=False):
        score, kernels = self.prepare(outs_dict["maps"])
        res = []
        for b in range(score.shape[0]):
            label = label_proc = self.labels(kernels[b])
            sc = score[b]
            if self.scale != 1:
                label = upsample_nearest(label, self.img_h // label.shape[0])
                sc = upsample_nearest(sc, self.img_h // sc.shape[0])
            out = generate_box(sc, label, shape_list[b], self.min_area, self.box_thresh, return_details)
            d = {"points": out[0], "scores": out[1]}
            if return_details:
                d["details"] = out[2]
                d["label"] = label
                d["label_proc"] = label_proc
            res.append(d)
        return res
