"""Randomised parity stress on a GPU box (not collected by pytest; run by hand):

    python tests/stress_gpu.py [--seconds 120] [--seed 0]

Loops over random sizes / seeds / thresholds of the adversarial generators of the GPU parity tests (blob fields with
merged masks, non-nested kernels, gate scenes, random pages) and checks every result against the oracle with the
same comparators the tests use. Prints one line per failure and a summary; exit code 1 if anything failed."""
import argparse
import os
import sys
import time
import traceback

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import pytest  # noqa: E402
import test_crop_gpu as tc  # noqa: E402
import test_db_gpu as tdb  # noqa: E402
import test_pan_gpu as tpan  # noqa: E402
import test_pse_gpu as tpse  # noqa: E402
from pytorchocr_b200 import synth  # noqa: E402


LAST_TAG = ""


def _db_variant(rng, kw):
    """random stage-2 code path (include/ocrpp.h OCRPP_TUNE_DB_PATH / _SCAN) and reference branch (cpp_speedup)"""
    from pytorchocr_b200 import _lib
    L = _lib.lib()
    path = int(rng.choice([0, 1, 2, 3]))
    scan = int(rng.choice([0, 0, 1, 2]))
    _lib.check(L.ocrpp_set_tuning(_lib.TUNE_DB_PATH, path))
    _lib.check(L.ocrpp_set_tuning(_lib.TUNE_DB_SCAN, scan))
    tag = "path%d scan%d" % (path, scan)
    global LAST_TAG
    LAST_TAG = tag
    if rng.random() < 0.4:
        kw["cpp_speedup"] = False
        kw["score_mode"] = str(rng.choice(["poly", "box"]))
        tag += " python/" + kw["score_mode"]
    LAST_TAG = tag
    return tag


def case_db(rng):
    H, W = int(rng.integers(24, 260)), int(rng.integers(24, 400))
    sig = float(rng.choice([0.45, 0.6, 1.0, 1.5, 2.5, 4.0]))
    p = np.stack([cv2.GaussianBlur(rng.random((H, W)).astype(np.float32), (0, 0), sig) for _ in range(2)])
    lo, hi = np.quantile(p, 0.02), np.quantile(p, 0.98)
    p = np.clip((p - lo) / (hi - lo), 0, 1).astype(np.float32)
    q = float(np.quantile(p, rng.uniform(0.3, 0.7)))
    kw = dict(thresh=q, box_thresh=q + 0.02, max_unmatched=0)
    if rng.random() < 0.3:
        kw["use_dilation"] = True
    tag = _db_variant(rng, kw)
    sl = np.array([[H, W, 1.0, 1.0], [int(H * 1.7), int(W * 0.8), 1.7, 0.8]])
    tdb._check(p[:, None], sl, **kw)
    return "db %dx%d sig=%.2f %s %s" % (H, W, sig, "dil" if "use_dilation" in kw else "", tag)


def case_db_synth(rng):
    H, W = int(rng.integers(64, 400)), int(rng.integers(64, 700))
    maps = synth.db_batch(2, seed=int(rng.integers(1 << 30)), H=H, W=W)
    kw = dict(max_unmatched=0)
    tag = _db_variant(rng, kw)
    tdb._check(maps, np.array([[H, W, 1.0, 1.0]] * 2), **kw)
    return "db_synth %dx%d %s" % (H, W, tag)


def case_pse(rng):
    K = int(rng.integers(2, 8))
    H, W = int(rng.integers(24, 200)), int(rng.integers(24, 260))
    blur = float(rng.choice([0.4, 0.6, 1.0, 1.5, 2.0, 3.0, 5.0]))
    nested = bool(rng.integers(2))
    maps = np.stack([tpse._blob_fields(rng, K, H, W, blur, nested) for _ in range(2)])
    ma = int(rng.choice([0, 5, 16]))
    tpse._check(maps, tpse._shape(2, H, W), maps_at_processing_res=True, min_area=ma, box_thresh=0.5, loose=0.4)
    return "pse K=%d %dx%d blur=%.1f nested=%d min_area=%d" % (K, H, W, blur, nested, ma)


def case_pse_synth(rng):
    H, W = int(rng.integers(96, 420)), int(rng.integers(128, 640))
    maps = np.stack([synth.pse_maps(int(rng.integers(1 << 30)), H, W) for _ in range(2)])
    tpse._check(maps, tpse._shape(2, H, W), maps_at_processing_res=True, loose=0.06)
    return "pse_synth %dx%d" % (H, W)


def case_pan(rng):
    H, W = int(rng.integers(24, 200)), int(rng.integers(24, 260))
    base = cv2.GaussianBlur(rng.random((H, W)).astype(np.float32), (0, 0), float(rng.choice([1.0, 2.0, 4.0])))
    text = base > np.quantile(base, rng.uniform(0.2, 0.6))
    kern = (rng.random((H, W)) > rng.uniform(0.4, 0.9)) & text
    inst = rng.integers(0, 4, (H, W))
    centres = np.array([[0, 0, 0, 0], [6, 0, 0, 0], [0, 6, 0, 0], [0, 0, 6, 0]], np.float32)
    maps = np.empty((1, 6, H, W), np.float32)
    maps[0, 0] = np.where(text, 3.0, -3.0)
    maps[0, 1] = np.where(kern, 3.0, -3.0)
    maps[0, 2:] = centres[inst].transpose(2, 0, 1) + rng.normal(0, 0.25, (4, H, W))
    mka = float(rng.choice([0.0, 2.6]))
    tpan._check(maps, tpan._shape(1, H, W), scale=1, maps_at_processing_res=True, min_kernel_area=mka, min_area=2,
                box_thresh=0.5, loose=0.6)
    return "pan %dx%d mka=%.1f" % (H, W, mka)


def case_pan_synth(rng):
    H, W = int(rng.integers(96, 420)), int(rng.integers(128, 640))
    maps = np.stack([synth.pan_maps(int(rng.integers(1 << 30)), H, W) for _ in range(2)])
    tpan._check(maps, tpan._shape(2, H, W), scale=1, maps_at_processing_res=True, loose=0.06)
    return "pan_synth %dx%d" % (H, W)


def case_crop(rng):
    H, W = int(rng.integers(120, 500)), int(rng.integers(160, 700))
    img = synth.page_image(int(rng.integers(1 << 30)), H, W)
    boxes = synth.page_boxes(int(rng.integers(1 << 30)), n=int(rng.integers(1, 120)), H=H, W=W,
                             tall_frac=float(rng.uniform(0, 0.5)), skew=float(rng.uniform(0, 4)), scale=0.7)
    from oracle import crop_oracle as co

    def usable(b):
        # an empty bounding rectangle makes cv2 raise in the reference as well; collinear / coinciding corners (a box
        # squeezed against the page border) make cv2 fall back to a meaningless SVD solution - both are reported as
        # degenerate by the CUDA path and are not parity cases
        l, t, r, bt = co.crop_rect(b)
        if r - l < 1 or bt - t < 1:
            return False
        dst = np.array([[0, 0], [r - l - 1, 0], [r - l - 1, bt - t - 1], [0, bt - t - 1]], np.float32)
        M = co.perspective_transform((b - [l, t]).astype(np.float32), dst)
        return M is not None and co.invert3(M) is not None
    boxes = boxes[[usable(b) for b in boxes]]
    gb, gc = tc._cropper()(img, boxes)
    tc._check_page(img, boxes, gb, gc)
    return "crop %dx%d n=%d" % (H, W, len(boxes))


_CTC = {}


def case_ctc(rng):
    import tempfile
    import torch
    import test_ctc_gpu as tctc
    T, B, C = int(rng.integers(1, 140)), int(rng.integers(1, 300)), int(rng.choice([2, 37, 97, 513, 6623, 6624]))
    if "ops" not in _CTC:
        _CTC["ops"] = tctc._ops(synth.write_char_dict(os.path.join(tempfile.mkdtemp(), "d.txt"), 6623))
    op, oracle, _ = _CTC["ops"]
    probs, _ = synth.ctc_probs_numpy(int(rng.integers(1 << 30)), T, B, C)
    mode = int(rng.integers(3))
    x = torch.from_numpy(probs).cuda()
    if mode == 1:                                   # fp16 input: the oracle sees the same rounded values
        x = x.half()
        ref = x.float().cpu()
    elif mode == 2:                                 # strided view (class and line padding)
        big = torch.zeros((T, B + 3, C + 5), device="cuda")
        big[:, 1:B + 1, 2:C + 2] = x
        x = big[:, 1:B + 1, 2:C + 2]
        ref = torch.from_numpy(probs)
    else:
        ref = torch.from_numpy(probs)
    tctc._same(op(x), oracle(ref))
    return "ctc T=%d B=%d C=%d mode=%d" % (T, B, C, mode)


def case_db_fp16(rng):
    import torch
    H, W = int(rng.integers(64, 300)), int(rng.integers(64, 500))
    maps = synth.db_batch(2, seed=int(rng.integers(1 << 30)), H=H, W=W)
    big = torch.zeros((2, 1, H + 2, W + 8), dtype=torch.float16, device="cuda")     # strided half view
    big[:, :, 1:H + 1, 8:W + 8] = torch.from_numpy(maps).cuda().half()
    dev = big[:, :, 1:H + 1, 8:W + 8]
    # the oracle gets the float16 array itself, as the reference does (`pred.detach().cpu().numpy()`): numpy then
    # evaluates `pred > thresh` in float16
    tdb._check(dev, np.array([[H, W, 1.0, 1.0]] * 2), oracle_maps=dev.cpu().numpy(), max_unmatched=0)
    return "db_fp16 %dx%d" % (H, W)


CASES = [case_ctc, case_db_fp16, case_db, case_db_synth, case_pse, case_pse_synth, case_pan, case_pan_synth, case_crop]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=120)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--only", default="")
    ap.add_argument("--verbose", action="store_true", help="print every case before it runs (to find a crashing one)")
    ap.add_argument("--first", type=int, default=0, help="first iteration number")
    args = ap.parse_args()
    cases = [c for c in CASES if not args.only or args.only in c.__name__]
    n, fails = {}, []
    t0, it = time.time(), args.first
    while time.time() - t0 < args.seconds:
        case = cases[it % len(cases)]
        seed = args.seed * 1000003 + it
        it += 1
        rng = np.random.default_rng(seed)
        if args.verbose:
            print("run", case.__name__, "it", it - 1, "seed", seed, flush=True)
        try:
            case(rng)
            n[case.__name__] = n.get(case.__name__, 0) + 1
        except (AssertionError, pytest.fail.Exception, Exception) as e:  # noqa: B014
            fails.append((case.__name__, seed, repr(e)[:300]))
            print("FAIL", case.__name__, "seed", seed, LAST_TAG if "db" in case.__name__ else "", repr(e)[:300], flush=True)
            if not isinstance(e, (AssertionError, pytest.fail.Exception)):
                traceback.print_exc()
    print("passed:", n, "failed:", len(fails))
    return 1 if fails else 0


if __name__ == "__main__":
    sys.exit(main())
