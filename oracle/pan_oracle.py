"""TEST INFRASTRUCTURE ONLY (oracle). PAN / PAN++ pixel aggregation, CPU restatement of
R/pytocr/postprocess/pan_postprocess.py:10-113 and pan_postprocess_fast/pa.pyx:14-104."""
import numpy as np

from . import clib
from .pse_oracle import generate_box, sigmoid_f32, upsample_nearest


def pa(kernels, emb, min_area=0):
    """pa.pyx:99-104 + :28-54 (pre-pass in numpy so the float32 np.mean is numpy's own) + BFS in C."""
    kernels = np.ascontiguousarray(kernels, dtype=np.uint8)
    emb = np.ascontiguousarray(emb, dtype=np.float32)
    _, cc = clib.ccl4(kernels[0])
    label_num, label = clib.ccl4(kernels[1])
    min_area = np.float32(min_area)  # C float parameter (pa.pyx:20)
    flat = label.ravel()
    area = np.full((label_num,), -1, dtype=np.float32)
    cnt = np.bincount(flat, minlength=label_num)
    area[1:] = cnt[1:label_num]
    # first raster pixel of every label
    first = np.full((label_num,), -1, dtype=np.int64)
    idx = np.nonzero(flat)[0]
    if idx.size:
        labs = flat[idx]
        # first occurrence per label: reverse assignment keeps the smallest index
        first[labs[::-1]] = idx[::-1]
    flag = np.zeros((label_num,), np.int32)
    mean_emb = np.zeros((label_num, 4), np.float32)
    max_rate = np.float32(1024)
    cc_flat = cc.ravel()
    alive = [False] * label_num
    by_cc = {}
    for i in range(1, label_num):
        if area[i] < min_area:
            continue
        alive[i] = True
        c = int(cc_flat[first[i]])
        for j in by_cc.get(c, ()):  # earlier surviving labels in the same text component, j < i
            rate = area[i] / area[j]  # float32 / float32
            if rate < 1 / max_rate or rate > max_rate:
                flag[i] = 1
                mean_emb[i] = np.mean(emb[:, label == i], axis=1)
                if flag[j] == 0:
                    flag[j] = 1
                    mean_emb[j] = np.mean(emb[:, label == j], axis=1)
        by_cc.setdefault(c, []).append(i)
    dead = np.array([not a for a in alive])
    dead[0] = False
    label = np.where(dead[label], 0, label).astype(np.int32)
    return clib.pa_expand(kernels[0], emb, label, flag, mean_emb), flag, mean_emb


class PANPostProcessOracle(object):
    def __init__(self, thresh=0.5, box_thresh=0.85, min_area=16, min_kernel_area=2.6, scale=4,
                 out_polygon=False, maps_at_processing_res=False, **kwargs):
        assert not out_polygon
        self.maps_at_processing_res = maps_at_processing_res
        self.thresh, self.box_thresh, self.min_area = thresh, box_thresh, min_area
        self.min_kernel_area = min_kernel_area / float(scale ** 2)
        self.scale = scale

    def prepare(self, pred):
        """pan_postprocess.py:32-51."""
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
        kernels = pred[:, :2] > self.thresh
        text = kernels[:, 0:1]
        kernels = np.concatenate([text, kernels[:, 1:2] & text], axis=1).astype(np.uint8)
        emb = (pred[:, 2:] * text.astype(np.float32)).astype(np.float32)
        return score, kernels, emb

    def __call__(self, outs_dict, shape_list, return_details=False):
        score, kernels, emb = self.prepare(outs_dict["maps"])
        res = []
        for b in range(score.shape[0]):
            label, flag, _ = pa(kernels[b], emb[b], self.min_kernel_area)
            label_proc = label
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
                d["flag"] = flag
            res.append(d)
        return res
