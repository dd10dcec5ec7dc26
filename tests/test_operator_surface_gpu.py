"""GPU tests of the operator surface around the kernels: the Distillation wrappers
(R/pytocr/postprocess/db_postprocess.py:197-226, rec_postprocess.py:96-125), `CTCLabelDecode.decode_batch` against
the ORACLE (not against the CUDA path itself), and the hand-off of detection results to the reference's DetMetric
(R/pytocr/metrics/det_metric.py:16-38, which iterates `pred["points"]` next to the ground-truth polygons)."""
import numpy as np
import pytest

from oracle.ctc_oracle import CTCLabelDecodeOracle
from oracle.db_oracle import DBPostProcessOracle
from pytorchocr_b200 import synth

pytestmark = pytest.mark.gpu

DB = dict(thresh=0.3, box_thresh=0.5, max_candidates=1000, unclip_ratio=1.7, score_mode="poly", cpp_speedup=True)


def _build(cfg):
    from pytorchocr_b200.postprocess import build_post_process
    return build_post_process(dict(cfg, cuda_speedup=True), {"use_gpu": True})


def test_distillation_db_postprocess():
    import torch
    H, W = 160, 256
    maps = {"student": synth.db_batch(2, seed=41, H=H, W=W), "teacher": synth.db_batch(2, seed=42, H=H, W=W)}
    sl = np.array([[H, W, 1.0, 1.0], [240, 320, 1.5, 1.25]], np.float64)
    op = _build(dict(DB, name="DistillationDBPostProcess", model_name=["student", "teacher"], key=None))
    res = op({k: {"maps": torch.from_numpy(v).cuda()} for k, v in maps.items()}, sl)
    assert sorted(res) == ["student", "teacher"]
    total = 0
    for k, v in maps.items():
        want = DBPostProcessOracle(**DB)({"maps": v}, sl)
        assert len(res[k]) == len(want) == 2
        for g, w in zip(res[k], want):
            gs = sorted(map(tuple, np.asarray(g["points"]).reshape(-1, 8).tolist()))
            ws = sorted(map(tuple, np.asarray(w["points"]).reshape(-1, 8).tolist()))
            assert len(gs) == len(ws) and g["scores"] == w["scores"]
            assert np.abs(np.array(gs) - np.array(ws)).max() <= 1
            total += len(ws)
    assert total > 20
    # default model_name is ["student"]; a missing model is a KeyError as in the reference
    op1 = _build(dict(DB, name="DistillationDBPostProcess"))
    assert list(op1({"student": {"maps": torch.from_numpy(maps["student"]).cuda()}}, sl)) == ["student"]
    with pytest.raises(KeyError):
        op({"student": {"maps": torch.from_numpy(maps["student"]).cuda()}}, sl)


def test_distillation_ctc_label_decode(tmp_path):
    import torch
    d = synth.write_char_dict(str(tmp_path / "dict.txt"), 96)
    T, B, C = 25, 12, 97
    ps, _ = synth.ctc_probs_numpy(5, T, B, C)
    pt, _ = synth.ctc_probs_numpy(6, T, B, C)
    op = _build({"name": "DistillationCTCLabelDecode", "character_dict_path": d, "use_space_char": False,
                 "model_name": ["student", "teacher"], "key": "head_out"})
    res = op({"student": {"head_out": torch.from_numpy(ps).cuda()}, "teacher": {"head_out": torch.from_numpy(pt).cuda()}})
    oracle = CTCLabelDecodeOracle(d)
    for k, pr in (("student", ps), ("teacher", pt)):
        want = oracle(torch.from_numpy(pr))
        assert [r[0] for r in res[k]] == [w[0] for w in want]
        assert np.allclose([r[1] for r in res[k]], [w[1] for w in want], rtol=1e-6, equal_nan=True)
    # key=None: the model output itself is the tensor; model_name given as a plain string (:110-111)
    op2 = _build({"name": "DistillationCTCLabelDecode", "character_dict_path": d, "model_name": "student"})
    r2 = op2({"student": torch.from_numpy(ps).cuda()})
    assert [r[0] for r in r2["student"]] == [r[0] for r in res["student"]]
    # with labels: (text, decoded label) pairs per model, as the reference returns them
    lab = np.array([[3, 4, 4, 0, 5]] * B)
    r3 = op2({"student": torch.from_numpy(ps).cuda()}, label=lab)
    assert len(r3["student"]) == 2 and r3["student"][1] == oracle.decode(lab)


def test_decode_batch_against_oracle(tmp_path):
    """run_ocr.py:221-224 decodes every crop with B = 1; decode_batch packs crops of different widths into one call.
    The checker is the oracle run crop by crop (the reference's loop), not the CUDA path."""
    import torch
    d = synth.write_char_dict(str(tmp_path / "dict.txt"), 96)
    oracle = CTCLabelDecodeOracle(d)
    op = _build({"name": "CTCLabelDecode", "character_dict_path": d, "use_space_char": False})
    rng = np.random.default_rng(3)
    preds, want = [], []
    for i in range(37):
        T = int(rng.integers(1, 60))
        p, _ = synth.ctc_probs_numpy(100 + i, T, 1, 97)
        if i % 9 == 0:            # a crop that decodes to nothing
            p[:] = 0.0
            p[:, :, 0] = 1.0
        preds.append(torch.from_numpy(p).cuda() if i % 2 else p)
        want.append(oracle(torch.from_numpy(p))[0])
    got = op.decode_batch(preds)
    assert [g[0] for g in got] == [w[0] for w in want]
    assert np.allclose([g[1] for g in got], [w[1] for w in want], rtol=1e-6, equal_nan=True)
    assert any(w[0] == "" and np.isnan(w[1]) for w in want) and sum(len(w[0]) for w in want) > 100


def _quad_iou(a, b):
    """IoU of two convex quads (Sutherland-Hodgman clipping), what DetectionIoUEvaluator computes through shapely."""
    def area(p):
        return 0.5 * abs(sum(p[i][0] * p[(i + 1) % len(p)][1] - p[(i + 1) % len(p)][0] * p[i][1] for i in range(len(p))))

    def clip(subject, cp):
        out = [tuple(map(float, q)) for q in subject]
        e0, e1 = np.subtract(cp[1], cp[0]), np.subtract(cp[2], cp[1])
        if e0[0] * e1[1] - e0[1] * e1[0] < 0:
            cp = cp[::-1]
        for i in range(len(cp)):
            a0, a1 = cp[i], cp[(i + 1) % len(cp)]
            inp, out = out, []
            if not inp:
                break

            def inside(q):
                return (a1[0] - a0[0]) * (q[1] - a0[1]) - (a1[1] - a0[1]) * (q[0] - a0[0]) >= 0
            s = inp[-1]
            for e in inp:
                if inside(e) != inside(s):
                    dx, dy = e[0] - s[0], e[1] - s[1]
                    den = (a1[0] - a0[0]) * dy - (a1[1] - a0[1]) * dx
                    tt = ((a1[0] - a0[0]) * (a0[1] - s[1]) - (a1[1] - a0[1]) * (a0[0] - s[0])) / den
                    out.append((s[0] + tt * dx, s[1] + tt * dy))
                if inside(e):
                    out.append(e)
                s = e
        return out
    inter = clip(list(a), [tuple(map(float, q)) for q in b])
    ia = area(inter) if len(inter) >= 3 else 0.0
    return ia / (area(a) + area(b) - ia + 1e-12)


def test_det_metric_handoff():
    """The structure DetMetric.__call__ consumes (det_metric.py:16-38): `preds` = the operator's list of dicts, zipped
    with batch[2] (gt polygons [N,K,4,2]) and batch[3] (ignore tags); every `det_polyon in pred["points"]` is a [4,2]
    polygon. Ground truth = the oracle's boxes: precision = recall = 1 at the evaluator's IoU 0.5 constraint."""
    import torch
    H, W = 192, 320
    maps = synth.db_batch(3, seed=77, H=H, W=W)
    maps[2] = 0.0                                        # an empty page: points.shape == (0,), iterates to nothing
    sl = np.array([[H, W, 1.0, 1.0]] * 3, np.float64)
    preds = _build(dict(DB, name="DBPostProcess"))({"maps": torch.from_numpy(maps).cuda()}, sl)
    want = DBPostProcessOracle(**DB)({"maps": maps}, sl)
    K = max(len(w["points"]) for w in want)
    gt = np.zeros((3, K, 4, 2), np.int16)
    ignore = np.ones((3, K), bool)
    for n, w in enumerate(want):
        if len(w["points"]):
            gt[n, :len(w["points"])] = w["points"]
            ignore[n, :len(w["points"])] = False
    batch = [None, None, gt, ignore]
    matched = n_gt = n_det = 0
    for pred, gt_polys, tags in zip(preds, batch[2], batch[3]):            # det_metric.py:24-36
        gt_info = [{"points": g, "text": "", "ignore": t} for g, t in zip(gt_polys, tags)]
        det_info = [{"points": dp, "text": ""} for dp in pred["points"]]
        for dinfo in det_info:
            assert np.asarray(dinfo["points"]).shape == (4, 2) and np.asarray(dinfo["points"]).dtype == np.int16
        care = [g for g in gt_info if not g["ignore"]]
        n_gt += len(care)
        n_det += len(det_info)
        used = set()
        for dinfo in det_info:
            for gi, g in enumerate(care):
                if gi not in used and _quad_iou(dinfo["points"].astype(float), g["points"].astype(float)) > 0.5:
                    used.add(gi)
                    matched += 1
                    break
    assert preds[2]["points"].shape == (0,) and preds[2]["scores"] == []
    assert n_gt > 20 and matched == n_gt == n_det
