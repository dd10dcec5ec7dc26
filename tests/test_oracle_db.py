"""CPU statement that cv2's contour semantics == the CCL formulation the CUDA path implements
(oracle/db_ccl_oracle.py vs the cv2-based line-by-line restatement oracle/db_oracle.py)."""
import cv2
import numpy as np
import pytest

from oracle import db_ccl_oracle, db_oracle
from pytorchocr_b200 import synth


def _match(details_cv, details_ccl):
    """Pairs candidates by the raster-first point of their point set + fill count."""
    key = lambda d: (d["contour_first"], d.get("fill_count", -1))
    return key


@pytest.mark.parametrize("seed", range(6))
def test_fill_masks_random_fields(seed):
    """Every cv2 contour's fillPoly(lineType=1) pixel count and mean == CCL formulation."""
    rng = np.random.default_rng(seed)
    H, W = 48, 64
    p = cv2.GaussianBlur(rng.random((H, W)).astype(np.float32), (0, 0), [0.3, 0.6, 1.0, 1.5, 2.5, 0.45][seed])
    bm = (p > np.quantile(p, [0.4, 0.5, 0.6, 0.45, 0.55, 0.5][seed])).astype(np.uint8)
    contours, _ = cv2.findContours(bm, cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)
    cv = []
    for c in contours:
        c = c.reshape(-1, 2)
        score, cnt = db_oracle.box_score(c, p)
        cv.append((cnt, round(score, 12), len(c) <= 2))
    ccl = []
    for cd in db_ccl_oracle.candidates(p, bm):
        cnt = int(cd["fill"].sum())
        ccl.append((cnt, round(float(p[cd["fill"]].astype(np.float64).sum() / cnt), 12), cd["le2"]))
    assert len(cv) == len(ccl)
    assert sorted(cv) == sorted(ccl)


@pytest.mark.parametrize("seed", range(4))
def test_boxes_synth(seed):
    H, W = 192, 320
    pred = synth.db_map(synth.BASE_SEED + seed, H, W, n_regions=200)
    bm = (pred > 0.3).astype(np.uint8)
    b_cv, d_cv = db_oracle.boxes_from_bitmap(pred, bm, 0.5, 1.7, W, H, return_details=True)
    b_ccl, d_ccl = db_ccl_oracle.boxes_from_bitmap_ccl(pred, bm, 0.5, 1.7, W, H)
    assert len(d_cv) == len(d_ccl)
    from collections import Counter
    assert Counter(d["status"] for d in d_cv) == Counter(d["status"] for d in d_ccl)
    ok_cv = sorted((d for d in d_cv if d["status"] == "ok"), key=lambda d: (d["fill_count"], d["score"]))
    ok_ccl = sorted((d for d in d_ccl if d["status"] == "ok"), key=lambda d: (d["fill_count"], d["score"]))
    n_fragile = 0
    for a, b in zip(ok_cv, ok_ccl):
        assert a["fill_count"] == b["fill_count"]
        assert abs(a["score"] - b["score"]) <= 1e-9 * abs(a["score"])
        assert np.abs(np.asarray(a["mini"]) - np.asarray(b["mini"])).max() < 1e-3
        if a["quad"] != b["quad"]:
            n_fragile += 1      # int() truncation of a corner within float noise of an integer
            continue
        assert np.abs(np.asarray(a["clip"]) - np.asarray(b["clip"])).max() < 1e-3
        fa, fb = np.asarray(a["out_f"]), np.asarray(b["out_f"])
        assert np.abs(fa - fb).max() < 1e-3
        stable = np.abs(fa - np.floor(fa) - 0.5) > 2e-3
        assert np.array_equal(np.asarray(a["out"])[stable], np.asarray(b["out"])[stable])
    assert n_fragile <= 1


def test_fast_port_equals_oracle():
    """bench.py's CPU timing leg (oracle/db_oracle_fast.py: vectorised between the cv2 / Clipper calls) returns exactly
    the boxes of the line-by-line oracle."""
    import numpy as np
    from oracle import db_oracle as O, db_oracle_fast as F
    from pytorchocr_b200 import synth
    if O._clipper() is None:
        pytest.skip("oracle/_ref/libclipper_ref.so not built")
    total = 0
    for seed, (Hh, Ww) in enumerate([(192, 320), (160, 256), (97, 131), (736, 1280)]):
        m = synth.db_map(500 + seed, H=Hh, W=Ww)
        bm = (m > 0.3).astype(np.uint8)
        for sw, sh in ((Ww, Hh), (Ww * 2 + 3, Hh + 11)):
            a = O.boxes_from_bitmap(m, bm, 0.5, 1.7, sw, sh)
            b = F.boxes_from_bitmap(m, bm, 0.5, 1.7, sw, sh)
            assert a == b
            total += len(a)
    assert total > 300
    z = np.zeros((64, 64), np.float32)
    assert F.boxes_from_bitmap(z, z.astype(np.uint8), 0.5, 1.7, 64, 64) == []
