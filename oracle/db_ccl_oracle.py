"""TEST INFRASTRUCTURE ONLY (oracle). `BoxesFromBitmap` restated in connected-component terms -
the formulation the CUDA path implements - so that "cv2 contour semantics == CCL semantics" is
itself a CPU-checked statement (tests/test_oracle_db.py), independent of any GPU.

Reference being restated: R/pytocr/postprocess/db_postprocess_fast/src/db_postprocess.cpp:231-317
(control flow), :194-229 (BoxScore incl. fillPoly lineType=1), with the third-party behaviour of
cv::findContours(RETR_LIST, CHAIN_APPROX_SIMPLE) expressed as:

  candidates = outer(C) for every 8-connected foreground component C
             + hole(C,h) for every 4-connected background region h that does not reach the frame
  points(outer C) = pixels of C                 points(hole C,h) = ring = pixels of C 4-adjacent to h
  fill(outer C)   = C + everything C encloses + X_outer(C)
  fill(hole C,h)  = h + islands in h (recursively) + ring + X_hole(C,h)
  X_outer(C) = background pixels e of the region just outside C with left(e) in C and
               (up(e) in C or down(e) in C)                       [4-connected boundary "stairs"]
  X_hole(C,h) = pixels e not already filled with, for dy in {-1,+1}:
               (e.x-1, e.y+dy) in h, (e.x-1, e.y) foreground, (e.x, e.y+dy) foreground
  "contour has <= 2 points"  <=>  C is one pixel or a 1-px-thick straight run (-, |, /, \\)

Rectangles are the fp64 min-area rectangle (geometry_oracle.min_area_rect); unclip uses the
DoOffset restatement (geometry_oracle.do_offset). Everything downstream follows db_oracle.py.
"""
import numpy as np
from scipy import ndimage as ndi

from . import geometry_oracle as G
from .db_oracle import get_contour_area, roundf

_S8 = np.ones((3, 3), np.int32)
_S4 = np.array([[0, 1, 0], [1, 1, 1], [0, 1, 0]], np.int32)


def _shift(m, dy, dx):
    """result[y,x] = m[y+dy, x+dx] (False outside)."""
    o = np.zeros_like(m)
    H, W = m.shape
    ys0, ys1 = max(0, -dy), min(H, H - dy)
    xs0, xs1 = max(0, -dx), min(W, W - dx)
    if ys0 < ys1 and xs0 < xs1:
        o[ys0:ys1, xs0:xs1] = m[ys0 + dy:ys1 + dy, xs0 + dx:xs1 + dx]
    return o


def _first_pixels(lab, n):
    idx = np.nonzero(lab.ravel())[0]
    first = np.full(n + 1, -1, np.int64)
    labs = lab.ravel()[idx]
    first[labs[::-1]] = idx[::-1]
    return first


def mini_box_from_corners(corners):
    """GetMiniBoxes ordering (db_postprocess.cpp:165-190) applied to 4 corner points."""
    order = sorted(range(4), key=lambda i: corners[i][0])
    a = [corners[i] for i in order]
    if a[3][1] <= a[2][1]:
        idx2, idx3 = a[3], a[2]
    else:
        idx2, idx3 = a[2], a[3]
    if a[1][1] <= a[0][1]:
        idx1, idx4 = a[1], a[0]
    else:
        idx1, idx4 = a[0], a[1]
    return np.array([idx1, idx2, idx3, idx4], np.float64)


def candidates(pred, bitmap, line_type=1):
    """Yields dicts: kind, first (x,y) raster-first point of the point set, pts (bool mask of the
    contour point set), fill (bool mask), le2 (the <=2-points rule). line_type=1: the C++ branch's
    cv::fillPoly(..., lineType=1) (4-connected boundary: the X "stair" pixels are part of the fill);
    line_type=8: the Python branch's default LINE_8 fill (db_postprocess.py:193) = the same sets without X."""
    bitmap = bitmap.astype(bool)
    H, W = bitmap.shape
    fg, nf = ndi.label(bitmap, structure=_S8)
    pad = np.pad(~bitmap, 1, constant_values=True)
    bgp, nb = ndi.label(pad, structure=_S4)
    out_label = bgp[0, 0]
    bg = np.where(bgp[1:-1, 1:-1] == out_label, 0, bgp[1:-1, 1:-1])
    ffirst = _first_pixels(fg, nf)
    bfirst = _first_pixels(bg, nb)
    fpar = {}
    for c in range(1, nf + 1):
        y, x = divmod(int(ffirst[c]), W)
        fpar[c] = int(bg[y, x - 1]) if x > 0 else 0          # 0 == OUT
    bpar = {}
    for h in range(1, nb + 1):
        if bfirst[h] < 0:
            continue
        y, x = divmod(int(bfirst[h]), W)
        bpar[h] = int(fg[y - 1, x])
        assert bpar[h] > 0
    fchildren, bchildren = {}, {}
    for h, c in bpar.items():
        fchildren.setdefault(c, []).append(h)
    for c, h in fpar.items():
        if h != 0:
            bchildren.setdefault(h, []).append(c)

    def fill_c(c):
        m = fg == c
        for h in fchildren.get(c, []):
            m |= fill_h(h)
        return m

    def fill_h(h):
        m = bg == h
        for c in bchildren.get(h, []):
            m |= fill_c(c)
        return m

    fgm = fg > 0
    outside = ~fgm & (bg == 0)
    res = []
    for c in range(1, nf + 1):
        C = fg == c
        base = fill_c(c)
        R = (bg == fpar[c]) if fpar[c] != 0 else outside
        X = R & _shift(C, 0, -1) & (_shift(C, -1, 0) | _shift(C, 1, 0))
        if line_type == 8:
            X = np.zeros_like(C)
        ys, xs = np.nonzero(C)
        area = len(ys)
        bw, bh = xs.max() - xs.min() + 1, ys.max() - ys.min() + 1
        le2 = (area == 1 or (bh == 1 and area == bw) or (bw == 1 and area == bh)
               or (bw == bh == area and (len(set(xs - ys)) == 1 or len(set(xs + ys)) == 1)))
        res.append({"kind": "outer", "first": (int(xs[0]), int(ys[0])), "pts": C, "fill": base | X,
                    "le2": bool(le2), "comp": c})
    for h in bpar:
        Hm = bg == h
        C = fg == bpar[h]
        ring = C & (_shift(Hm, 0, 1) | _shift(Hm, 0, -1) | _shift(Hm, 1, 0) | _shift(Hm, -1, 0))
        base = fill_h(h) | ring
        X = np.zeros_like(Hm)
        for dy in (-1, 1):
            X |= _shift(Hm, dy, -1) & _shift(fgm, 0, -1) & _shift(fgm, dy, 0)
        X &= ~base
        if line_type == 8:
            X = np.zeros_like(Hm)
        ys, xs = np.nonzero(ring)
        res.append({"kind": "hole", "first": (int(xs[0]), int(ys[0])), "pts": ring, "fill": base | X,
                    "le2": False, "comp": bpar[h], "hole": h})
    return res


def boxes_from_bitmap_ccl(pred, bitmap, box_thresh, unclip_ratio, src_w, src_h):
    """Same outputs as db_oracle.boxes_from_bitmap(return_details=True) but via CCL semantics and
    fp64 rectangles. Candidate order: reverse raster order of `first` (cv2's outer-contour order)."""
    pred = np.asarray(pred, np.float32)
    H, W = bitmap.shape
    cands = candidates(pred, bitmap)
    cands.sort(key=lambda d: -(d["first"][1] * W + d["first"][0]))
    boxes, details = [], []
    f = np.float32
    for cd in cands:
        d = {"kind": cd["kind"], "contour_first": cd["first"], "status": "ok"}
        details.append(d)
        if cd["le2"]:
            d["status"] = "le2pts"
            continue
        ys, xs = np.nonzero(cd["pts"])
        corners, (w, h) = G.min_area_rect(np.stack([xs, ys], 1))
        ssid = max(w, h)
        mini = mini_box_from_corners(corners).astype(np.float32)
        d["mini"], d["ssid"] = mini, float(ssid)
        if ssid < 3:
            d["status"] = "small"
            continue
        fill = cd["fill"]
        cnt = int(fill.sum())
        score = float(pred[fill].astype(np.float64).sum() / cnt)
        d["score"], d["fill_count"] = score, cnt
        if np.float32(score) < np.float32(box_thresh):
            d["status"] = "lowscore"
            continue
        distance = get_contour_area(mini, unclip_ratio)
        quad = [(int(mini[i][0]), int(mini[i][1])) for i in range(4)]
        off = G.do_offset(quad, float(distance))
        d["distance"], d["quad"] = float(distance), quad
        if len(off) == 0:
            d["status"] = "unclip_empty"
            continue
        corners2, (w2, h2) = G.min_area_rect(off)
        if w2 < 1.001 and h2 < 1.001:
            d["status"] = "unclip_empty"
            continue
        clip = mini_box_from_corners(corners2).astype(np.float32)
        d["clip"], d["ssid2"] = clip, float(max(w2, h2))
        if max(w2, h2) < 5:
            d["status"] = "small2"
            continue
        out, out_f = [], []
        for j in range(4):
            fx = f(f(clip[j][0] / f(W)) * f(src_w))
            fy = f(f(clip[j][1] / f(H)) * f(src_h))
            out_f.append((float(fx), float(fy)))
            out.append([int(min(max(roundf(fx), 0.0), float(src_w))),
                        int(min(max(roundf(fy), 0.0), float(src_h)))])
        d["out_f"], d["out"] = out_f, out
        boxes.append(out)
    return boxes, details
