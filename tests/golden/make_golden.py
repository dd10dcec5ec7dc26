"""Generates tests/golden/*.npz by running the REFERENCE ITSELF in the authoring container.

    python tests/golden/make_golden.py          (needs /root/reference and oracle/_ref built)
    python tests/golden/make_golden.py crops    (reference_crops.npz only: sort_boxes / get_part_img)
    python tests/golden/make_golden.py dbpy     (reference_db_python.npz only: the reference's pure-Python DB branch)

What runs unmodified from /root/reference (imported in place, nothing is copied):
  * pytocr.postprocess.pse_postprocess.PSEPostProcess    (whole operator incl. generate_box)
  * pytocr.postprocess.pan_postprocess.PANPostProcess
  * pytocr.postprocess.rec_postprocess.CTCLabelDecode
  * pytocr.postprocess.db_postprocess.DBPostProcess.__call__ (cpp_speedup=True wrapper semantics)
  * the compiled pse.pyx / pa.pyx (oracle/_ref, built unmodified by oracle/build_ref.py) behind the
    package names the operators import them from
What is stubbed and why: `pyclipper` / `shapely` (absent from the image) -> oracle/ref_shims.py
(the reference's own compiled Clipper behind pyclipper's three calls; GEOS' ring area / length), and
`pytocr.postprocess.db_postprocess_fast.cpp_boxes_from_bitmap` (the C++ module needs OpenCV C++ headers,
absent) -> oracle/db_oracle.boxes_from_bitmap, the cv2-python restatement that calls the reference's own
compiled Clipper. So `reference_outputs.npz`'s DB entries pin the operator wrapper only, while
`reference_db_python.npz` (main_db_python) holds outputs of the reference's UNMODIFIED pure-Python DB branch
(`DBPostProcess(cpp_speedup=False)`, db_postprocess.py:76-194: findContours -> get_mini_boxes -> box_score ->
unclip -> get_mini_boxes -> rescale) - per stage and final - which pin the stages both branches share.

The fixtures are small (inputs + outputs, a few hundred KB) and are checked by
tests/test_golden.py against the oracle (CPU) and against the CUDA path (-m gpu).
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("OCR_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))


def import_reference():
    import torch  # noqa: F401
    from oracle import ref_shims
    ref_shims.install()
    import pa as pa_mod
    import pse as pse_mod
    from oracle import db_oracle

    def cpp_boxes_from_bitmap(pred, bitmap, box_thresh, det_db_unclip_ratio, src_w, src_h, use_padding_resize=False):
        # same signature and uint8 cast as db_postprocess_fast/__init__.py:10-22
        return db_oracle.boxes_from_bitmap(pred, bitmap.astype(np.uint8), box_thresh, det_db_unclip_ratio,
                                           src_w, src_h, use_padding_resize)
    for pkg, attr, fn in (("pytocr.postprocess.pse_postprocess_fast", "pse", pse_mod.pse),
                          ("pytocr.postprocess.pan_postprocess_fast", "pa", pa_mod.pa),
                          ("pytocr.postprocess.db_postprocess_fast", "cpp_boxes_from_bitmap",
                           cpp_boxes_from_bitmap)):
        m = types.ModuleType(pkg)
        setattr(m, attr, fn)
        sys.modules[pkg] = m
    sys.path.insert(0, REF)
    import warnings
    warnings.filterwarnings("ignore")
    from pytocr.postprocess import build_post_process
    return build_post_process


def main():
    import torch
    from pytorchocr_b200 import synth
    build = import_reference()
    out = {}

    # ---- PSE: API-faithful 1/4-resolution head output, the three scales ----
    # (maps are stored as float16 and the reference runs on the up-cast values: exact round trip)
    h, w = 72, 120
    maps = np.stack([synth.pse_maps(7 + i, h, w, n_abs=40, hh_rng=(3, 5), hw_rng=(6, 10)) for i in range(2)])
    maps = maps.astype(np.float16).astype(np.float32)
    sl = np.array([[4 * h, 4 * w, 1.0, 1.0], [432, 600, 1.0 / 1.5, 1.0 / 1.25]], np.float64)
    out["pse_maps"], out["pse_shape"] = maps.astype(np.float16), sl
    for scale in (1, 2, 4):
        op = build({"name": "PSEPostProcess", "thresh": 0, "box_thresh": 0.85, "min_area": 16, "scale": scale})
        res = op({"maps": torch.from_numpy(maps)}, sl)
        for n, r in enumerate(res):
            out["pse_s%d_points_%d" % (scale, n)] = np.asarray(r["points"], np.int16).reshape(-1, 4, 2)
            out["pse_s%d_scores_%d" % (scale, n)] = np.asarray(r["scores"], np.float32)

    # ---- PAN: shipped config (scale 4) and full-res processing (scale 1) ----
    maps = np.stack([synth.pan_maps(11 + i, h, w, n_abs=40, hh_rng=(3, 5), hw_rng=(6, 10)) for i in range(2)])
    maps = maps.astype(np.float16).astype(np.float32)
    out["pan_maps"], out["pan_shape"] = maps.astype(np.float16), sl
    for scale in (1, 2, 4):
        op = build({"name": "PANPostProcess", "thresh": 0, "box_thresh": 0.85, "min_area": 16,
                    "min_kernel_area": 2.6, "scale": scale})
        res = op({"maps": torch.from_numpy(maps)}, sl)
        for n, r in enumerate(res):
            out["pan_s%d_points_%d" % (scale, n)] = np.asarray(r["points"], np.int16).reshape(-1, 4, 2)
            out["pan_s%d_scores_%d" % (scale, n)] = np.asarray(r["scores"], np.float32)

    # ---- CTC: [T,B,C] softmax rows; the reference's real dictionary is data under /root/reference and is
    #      not copied, so the fixture stores the decoded CLASS IDS (via a private-use dictionary) ----
    T, B, Cn = 25, 16, 97
    probs, _ = synth.ctc_probs_numpy(3, T, B, Cn)
    probs[:, 3] = 0.0
    probs[:, 3, 0] = 1.0            # an all-blank line -> ("", nan)
    dict_path = os.path.join(HERE, "_dict_tmp.txt")
    synth.write_char_dict(dict_path, Cn - 1)
    op = build({"name": "CTCLabelDecode", "character_dict_path": dict_path, "use_space_char": False})
    res = op(torch.from_numpy(probs))
    os.remove(dict_path)
    out["ctc_probs"] = probs
    out["ctc_text"] = np.array([r[0] for r in res])
    out["ctc_conf"] = np.array([r[1] for r in res], np.float32)

    # ---- DB: operator wrapper semantics (cpp_speedup=True; box extraction = the restatement) ----
    Hd, Wd = 160, 256
    maps = synth.db_batch(2, seed=5, H=Hd, W=Wd)
    sl_db = np.array([[Hd, Wd, 1.0, 1.0], [240, 320, 1.5, 1.25]], np.float64)
    op = build({"name": "DBPostProcess", "thresh": 0.3, "box_thresh": 0.5, "max_candidates": 1000,
                "unclip_ratio": 1.7, "score_mode": "poly", "cpp_speedup": True})
    res = op({"maps": torch.from_numpy(maps)}, sl_db)
    out["db_maps"], out["db_shape"] = maps.astype(np.float16), sl_db      # fp16 storage: exact round trip below
    res = op({"maps": torch.from_numpy(out["db_maps"].astype(np.float32))}, sl_db)
    for n, r in enumerate(res):
        out["db_points_%d" % n] = np.asarray(r["points"], np.int16).reshape(-1, 4, 2)
        out["db_scores_%d" % n] = np.asarray(r["scores"], np.float32)

    # ---- the compiled Cython modules themselves on adversarial fields: label maps are the bit-exact gate ----
    import cv2
    import pa as pa_mod
    import pse as pse_mod
    rng = np.random.default_rng(4242)
    for i in range(4):
        K = [7, 4, 7, 3][i]
        base = cv2.GaussianBlur(rng.random((72, 96)).astype(np.float32), (0, 0), [1.0, 2.0, 0.6, 1.5][i])
        qs = np.quantile(base, np.linspace(0.35, 0.8, K))
        kernels = np.stack([(base > q) for q in qs])
        if i % 2:   # non-nested: independent random masks inside the text mask
            kernels = (rng.random((K, 72, 96)) > 0.45) & kernels[0]
            kernels = kernels & kernels[0:1]   # the operator multiplies every kernel by the text mask (:41-42)
        kernels = kernels.astype(np.uint8)
        out["exp_pse_kernels_%d" % i] = np.packbits(kernels, axis=None)
        out["exp_pse_shape_%d" % i] = np.array(kernels.shape)
        for ma in (0, 5):
            out["exp_pse_label_%d_ma%d" % (i, ma)] = pse_mod.pse(kernels.copy(), float(ma)).astype(np.int16)
    for i in range(4):
        Hh, Ww = 96, 128
        if i < 2:
            base = cv2.GaussianBlur(rng.random((Hh, Ww)).astype(np.float32), (0, 0), 2.0)
            text = base > np.quantile(base, 0.35)
            kern = (rng.random((Hh, Ww)) > 0.6) & text
        else:       # one text component with a large, a 1-px and a medium kernel: ratio flags + gate
            text = np.zeros((Hh, Ww), bool)
            text[4:60, 4:120] = True
            text[70:90, 10:100] = True
            kern = np.zeros((Hh, Ww), bool)
            kern[6:40, 6:60] = True
            kern[50, 100] = True
            kern[52:55, 70:74] = True
            kern[75:85, 20:60] = True
        inst = rng.integers(0, 4, (Hh, Ww)) if i < 2 else (np.arange(Ww)[None, :] >= 64) + 2 * (np.arange(Hh)[:, None] >= 44)
        centres = np.array([[0, 0, 0, 0], [6, 0, 0, 0], [0, 6, 0, 0], [0, 0, 6, 0]], np.float32)
        emb = (centres[inst].transpose(2, 0, 1) + rng.normal(0, 0.25, (4, Hh, Ww))).astype(np.float16).astype(np.float32)
        kernels = np.stack([text, kern & text]).astype(np.uint8)
        emb_m = emb * text[None].astype(np.float32)
        out["exp_pa_kernels_%d" % i] = np.packbits(kernels, axis=None)
        out["exp_pa_emb_%d" % i] = emb.astype(np.float16)
        for ma in (0.0, 2.6):
            lab = pa_mod.pa(kernels.copy(), emb_m.copy(), ma)
            out["exp_pa_label_%d_ma%d" % (i, int(ma))] = lab.astype(np.int16)

    np.savez_compressed(os.path.join(HERE, "reference_outputs.npz"), **out)
    print("wrote", os.path.join(HERE, "reference_outputs.npz"),
          {k: v.shape for k, v in out.items() if "points" in k})


def main_crops():
    """tests/golden/reference_crops.npz: the reference's own sort_boxes / get_part_img (imported in place from
    /root/reference/pytocr/utils/utility.py) on one small synthetic page."""
    import importlib.util
    from pytorchocr_b200 import synth
    spec = importlib.util.spec_from_file_location("ref_utility", os.path.join(REF, "pytocr", "utils", "utility.py"))
    U = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(U)
    H, W = 200, 320
    img = synth.page_image(31, H, W)
    boxes = synth.page_boxes(32, n=24, H=H, W=W, tall_frac=0.2)
    sorted_boxes = U.sort_boxes(boxes)
    crops = []
    for b in sorted_boxes:
        part = U.get_part_img(img, b)
        if part.shape[0] >= 1.5 * part.shape[1]:          # run_ocr.py:190-191
            part = np.rot90(part, 1)
        crops.append(np.ascontiguousarray(part))
    out = dict(img=img, boxes=boxes, sorted_boxes=np.asarray(sorted_boxes, np.int16),
               dims=np.array([c.shape[:2] for c in crops], np.int32),
               pixels=np.concatenate([c.reshape(-1) for c in crops]))
    np.savez_compressed(os.path.join(HERE, "reference_crops.npz"), **out)
    print("wrote reference_crops.npz", out["dims"].tolist())


def db_python_maps():
    """10 maps 160x256 (stored as float16, exact round trip): 8 synthetic DB pages (rotated regions, holes, low-score
    regions, specks, 1-px runs, border regions) and 2 blurred noise fields (many nested holes, ragged contours)."""
    import cv2
    from pytorchocr_b200 import synth
    Hd, Wd = 160, 256
    maps = [synth.db_map(900 + i, H=Hd, W=Wd, n_regions=200) for i in range(8)]
    rng = np.random.default_rng(77)
    for sigma in (1.5, 2.5):
        f = cv2.GaussianBlur(rng.random((Hd, Wd)).astype(np.float32), (0, 0), sigma)
        f = (f - f.min()) / (f.max() - f.min())
        maps.append(np.clip(0.3 + (f - np.quantile(f, 0.55)) * 6.0, 0, 1).astype(np.float32))
    return np.stack(maps)[:, None].astype(np.float16)


DBPY_CONFIGS = {   # name -> (constructor kwargs, use_padding_resize)
    "poly": (dict(), False),
    "box": (dict(score_mode="box"), False),
    "dilate": (dict(use_dilation=True), False),
    "pad": (dict(), True),
    "cand5": (dict(max_candidates=5), False),
    "r20": (dict(unclip_ratio=2.0, box_thresh=0.6, thresh=0.2), False),
}


def main_db_python():
    """tests/golden/reference_db_python.npz: the reference's pure-Python DB branch, unmodified, per stage and final."""
    import torch
    build = import_reference()
    maps16 = db_python_maps()
    maps = maps16.astype(np.float32)
    N, _, Hd, Wd = maps.shape
    sl = np.array([[Hd, Wd, 1.0, 1.0], [240, 320, 1.5, 1.25], [200, 200, 1.25, 0.78], [97, 301, 0.6, 1.18]] * 3,
                  np.float64)[:N]
    out = {"maps": maps16, "shape": sl}
    for name, (kw, pad) in DBPY_CONFIGS.items():
        cfg = dict({"name": "DBPostProcess", "thresh": 0.3, "box_thresh": 0.5, "max_candidates": 1000,
                    "unclip_ratio": 1.7, "score_mode": "poly", "cpp_speedup": False}, **kw)
        op = build(cfg)
        # record every stage call of the unmodified methods (inputs are reproducible from the maps)
        trace = {"mini_in_n": [], "mini_box": [], "mini_side": [], "score": [], "unclip_in": [], "unclip_n": [],
                 "unclip_pts": []}
        g, b, u = op.get_mini_boxes, op.box_score, op.unclip

        def get_mini_boxes(contour, g=g, t=trace):
            box, side = g(contour)
            t["mini_in_n"].append(len(contour))
            t["mini_box"].append(np.asarray(box, np.float32))
            t["mini_side"].append(np.float32(side))
            return box, side

        def box_score(bitmap, pts, b=b, t=trace):
            s = b(bitmap, pts)
            t["score"].append(s)
            return s

        def unclip(box, u=u, t=trace):
            e = u(box)
            t["unclip_in"].append(np.asarray(box, np.float32))
            t["unclip_n"].append(0 if len(e) != 1 else e.shape[1])
            if len(e) == 1:
                t["unclip_pts"].append(np.asarray(e[0], np.int32))
            return e
        op.get_mini_boxes, op.box_score, op.unclip = get_mini_boxes, box_score, unclip
        res = op({"maps": torch.from_numpy(maps)}, sl, use_padding_resize=pad)
        for n, r in enumerate(res):
            out["%s_points_%d" % (name, n)] = np.asarray(r["points"], np.int16).reshape(-1, 4, 2)
            out["%s_scores_%d" % (name, n)] = np.asarray(r["scores"], np.float64)
        out[name + "_mini_in_n"] = np.asarray(trace["mini_in_n"], np.int32)
        out[name + "_mini_box"] = np.asarray(trace["mini_box"], np.float32).reshape(-1, 4, 2)
        out[name + "_mini_side"] = np.asarray(trace["mini_side"], np.float32)
        out[name + "_score"] = np.asarray(trace["score"], np.float64)
        out[name + "_unclip_in"] = np.asarray(trace["unclip_in"], np.float32).reshape(-1, 4, 2)
        out[name + "_unclip_n"] = np.asarray(trace["unclip_n"], np.int32)
        out[name + "_unclip_pts"] = (np.concatenate(trace["unclip_pts"]) if trace["unclip_pts"]
                                     else np.zeros((0, 2), np.int32))
    # out_polygon=True (db_postprocess.py:98-103,119-122,142): polygons of different lengths end in
    # np.array(boxes, dtype=np.int16) -> ValueError under every numpy version (an explicit numeric dtype never
    # builds a ragged object array). Recorded as evidence that the branch has no behaviour to reproduce.
    op = build({"name": "DBPostProcess", "thresh": 0.3, "box_thresh": 0.5, "unclip_ratio": 1.7,
                "cpp_speedup": False, "out_polygon": True})
    errs = []
    for n in range(N):
        try:
            r = op({"maps": torch.from_numpy(maps[n:n + 1])}, sl[n:n + 1])
            errs.append("ok shape=%s" % (np.asarray(r[0]["points"]).shape,))
        except Exception as e:      # noqa: BLE001
            errs.append("%s: %s" % (type(e).__name__, str(e)[:60]))
    out["out_polygon_outcome"] = np.array(errs)
    np.savez_compressed(os.path.join(HERE, "reference_db_python.npz"), **out)
    print("wrote reference_db_python.npz")
    for name in DBPY_CONFIGS:
        print(" ", name, [len(out["%s_points_%d" % (name, n)]) for n in range(N)], "stage calls:",
              len(out[name + "_mini_side"]), len(out[name + "_score"]), len(out[name + "_unclip_n"]))
    print("  out_polygon:", errs)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "crops":
        main_crops()
    elif len(sys.argv) > 1 and sys.argv[1] == "dbpy":
        main_db_python()
    else:
        main()
