"""Pins the DB oracle to the reference's own code: tests/golden/reference_db_python.npz holds outputs of the
UNMODIFIED pure-Python DB branch (R/pytocr/postprocess/db_postprocess.py:76-194, `cpp_speedup=False`), run in the
authoring container by tests/golden/make_golden.py (`dbpy`) over oracle/ref_shims.py (pyclipper -> the reference's
compiled clipper.cpp, shapely -> GEOS' ring area/length). oracle/db_oracle.py restates both branches with ONE set of
stage functions (findContours order, get_mini_boxes, box_score, Clipper offset, rescale) and explicit switches for the
documented C++ / Python differences; here its Python semantics must reproduce the reference's per-stage values and
final boxes/scores exactly, on 10 maps (8 pages with holes / low-score regions / specks / border regions, 2 noise
fields with nested holes) x 6 configurations (score_mode poly and box, use_dilation, use_padding_resize,
max_candidates, other thresholds)."""
import os

import numpy as np
import pytest

from conftest import ROOT

G = np.load(os.path.join(ROOT, "tests", "golden", "reference_db_python.npz"))
CONFIGS = {
    "poly": (dict(), False),
    "box": (dict(score_mode="box"), False),
    "dilate": (dict(use_dilation=True), False),
    "pad": (dict(), True),
    "cand5": (dict(max_candidates=5), False),
    "r20": (dict(unclip_ratio=2.0, box_thresh=0.6, thresh=0.2), False),
}


def _oracle(name):
    from oracle.db_oracle import DBPostProcessOracle
    kw, pad = CONFIGS[name]
    cfg = dict(dict(thresh=0.3, box_thresh=0.5, max_candidates=1000, unclip_ratio=1.7, score_mode="poly",
                    cpp_speedup=False), **kw)
    return DBPostProcessOracle(**cfg)({"maps": G["maps"].astype(np.float32)}, G["shape"],
                                      use_padding_resize=pad, return_details=True)


@pytest.mark.parametrize("name", list(CONFIGS))
def test_final_boxes_and_scores(name):
    res = _oracle(name)
    assert len(res) == 10
    total = 0
    for n, r in enumerate(res):
        want = G["%s_points_%d" % (name, n)]
        got = np.asarray(r["points"], np.int16).reshape(-1, 4, 2)
        assert got.shape == want.shape, (name, n, got.shape, want.shape)
        assert np.array_equal(got, want), (name, n)
        assert np.array_equal(np.asarray(r["scores"], np.float64), G["%s_scores_%d" % (name, n)]), (name, n)
        total += len(want)
    assert total >= (30 if name == "cand5" else 100)


@pytest.mark.parametrize("name", list(CONFIGS))
def test_stage_values(name):
    """get_mini_boxes (corners + short side), box_score and unclip of every contour, in the reference's call order."""
    res = _oracle(name)
    mini, side, score, un_in, un_n, un_pts, npts = [], [], [], [], [], [], []
    for r in res:
        for d in r["details"]:
            npts.append(d["npts"])
            mini.append(d["mini"]), side.append(d["ssid"])
            if "score" in d:
                score.append(d["score"])
            if "distance" in d:
                un_in.append(d["mini"])
                one = len(d["offset"]) == 1
                un_n.append(len(d["offset"][0]) if one else 0)
                if one:
                    un_pts.append(np.asarray(d["offset"][0], np.int32))
                    npts.append(len(d["offset"][0]))
                    mini.append(d["clip"]), side.append(d["ssid2"])
    assert np.array_equal(np.asarray(npts, np.int32), G[name + "_mini_in_n"])
    assert np.array_equal(np.asarray(mini, np.float32).reshape(-1, 4, 2), G[name + "_mini_box"])
    assert np.array_equal(np.asarray(side, np.float32), G[name + "_mini_side"])
    assert np.array_equal(np.asarray(score, np.float64), G[name + "_score"])
    assert np.array_equal(np.asarray(un_in, np.float32).reshape(-1, 4, 2), G[name + "_unclip_in"])
    assert np.array_equal(np.asarray(un_n, np.int32), G[name + "_unclip_n"])
    assert np.array_equal(np.concatenate(un_pts), G[name + "_unclip_pts"])
    assert len(score) >= 30 and len(un_n) >= 30


def test_fixture_covers_holes_and_skips():
    """the fixture exercises what it claims: hole contours, every skip reason, both score modes differing"""
    from oracle import db_ccl_oracle  # noqa: F401  (hole semantics are checked in test_oracle_db.py)
    res = _oracle("poly")
    statuses = [d["status"] for r in res for d in r["details"]]
    for s in ("ok", "small", "lowscore"):
        assert s in statuses, s
    import cv2
    holes = 0
    for n in range(10):
        bm = (G["maps"][n, 0].astype(np.float32) > 0.3).astype(np.uint8)
        _, hier = cv2.findContours(bm, cv2.RETR_CCOMP, cv2.CHAIN_APPROX_SIMPLE)
        holes += int((hier[0][:, 3] >= 0).sum()) if hier is not None else 0
    assert holes >= 20
    sp = np.concatenate([G["poly_scores_%d" % n] for n in range(8)])
    sb = np.concatenate([G["box_scores_%d" % n] for n in range(8)])
    assert len(sp) == len(sb) and np.abs(sp - sb).max() > 1e-3


def test_out_polygon_has_no_reference_behaviour():
    """DBPostProcess(out_polygon=True) ends in np.array(ragged polygons, dtype=int16) (db_postprocess.py:142) and
    raises on every map of the fixture; recorded by make_golden.py from the unmodified reference."""
    assert all(str(s).startswith("ValueError") for s in G["out_polygon_outcome"])
