"""GPU parity: CTC greedy decode through the C-ABI vs the CPU oracle (bit-exact strings, exact
float32 confidences)."""
import math
import os

import numpy as np
import pytest

from oracle.ctc_oracle import CTCLabelDecodeNumpy, CTCLabelDecodeOracle
from pytorchocr_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dict_path(tmp_path_factory):
    return synth.write_char_dict(str(tmp_path_factory.mktemp("d") / "dict6623.txt"), 6623)


def _ops(dict_path):
    from pytorchocr_b200.postprocess import build_post_process
    op = build_post_process({"name": "CTCLabelDecode", "character_dict_path": dict_path,
                             "use_space_char": False, "cuda_speedup": True}, {"use_gpu": True})
    return op, CTCLabelDecodeOracle(dict_path, False), CTCLabelDecodeNumpy(dict_path, False)


def _same(a, b):
    assert len(a) == len(b)
    for (ta, ca), (tb, cb) in zip(a, b):
        assert ta == tb
        if math.isnan(cb):
            assert math.isnan(ca)
        else:
            assert np.float32(ca) == np.float32(cb), (ca, cb)


@pytest.mark.parametrize("T,B,C", [(80, 64, 6623), (80, 33, 6624), (81, 5, 6624), (1, 1, 6623),
                                   (7, 3, 37), (40, 257, 97), (130, 4, 513)])
def test_ctc_matches_oracle(dict_path, T, B, C):
    import torch
    op, oracle, _ = _ops(dict_path)
    probs, _ = synth.ctc_probs_numpy(1234 + T + B + C, T, B, C)
    got = op(torch.from_numpy(probs).cuda())
    want = oracle(torch.from_numpy(probs))
    _same(got, want)
    assert any(len(t) for t, _ in want)


def test_ctc_matches_reference_python_loop(dict_path):
    """small case against the line-by-line numpy/pure-Python restatement"""
    import torch
    op, _, npy = _ops(dict_path)
    probs, _ = synth.ctc_probs_numpy(99, 24, 9, 6623)
    _same(op(torch.from_numpy(probs).cuda()), npy(torch.from_numpy(probs)))


def test_ctc_ties_blank_lines_and_numpy_input(dict_path):
    import torch
    op, oracle, _ = _ops(dict_path)
    T, B, C = 16, 6, 6623
    rng = np.random.default_rng(5)
    p = rng.random((T, B, C)).astype(np.float32) * 0.5
    p[:, 0, 0] = 1.0                      # all blank -> empty string, NaN confidence
    p[:, 1, 77] = 1.0                     # one repeated class -> single char
    p[:, 2, 5] = 2.0; p[:, 2, 6000] = 2.0  # exact ties -> FIRST maximum (5)
    p[::2, 3, 10] = 1.0; p[1::2, 3, 0] = 1.0   # char, blank, char, blank ... -> T/2 chars
    p[:, 4, 6622] = 3.0                   # last class (tail elements)
    p[:, 5, 1] = 3.0; p[3, 5, 2] = 4.0    # head elements
    got = op(torch.from_numpy(p).cuda())
    want = oracle(torch.from_numpy(p))
    _same(got, want)
    assert got[0][0] == "" and math.isnan(got[0][1])
    assert len(got[1][0]) == 1 and len(got[3][0]) == T // 2
    # numpy input is [B,T,C] (reference :80-82 leaves numpy untransposed)
    got_np = op(np.ascontiguousarray(p.transpose(1, 0, 2)))
    _same(got_np, want)
    # tuple input, strided view, label passthrough
    big = torch.from_numpy(p).cuda()
    view = big[:, 1:5]
    _same(op((None, view)), oracle(torch.from_numpy(p[:, 1:5])))
    text, lab = op(big, label=[[1, 2, 0, 2], [0, 0]])
    assert lab[0][0] == op.character[1] + op.character[2] + op.character[2] and lab[1][0] == ""


def test_ctc_fp16(dict_path):
    import torch
    op, oracle, _ = _ops(dict_path)
    probs, _ = synth.ctc_probs_numpy(7, 80, 17, 6623)
    h = torch.from_numpy(probs).half()
    got = op(h.cuda())
    want = oracle(h.float())           # oracle sees the half values upcast
    _same(got, want)


def test_decode_batch_matches_per_crop_decode(tmp_path):
    """run_ocr.py decodes every crop with B = 1 (:221-224); decode_batch packs crops of different widths
    into one call and must give the very same (text, confidence) pairs."""
    import torch
    from pytorchocr_b200.postprocess import build_post_process
    d = synth.write_char_dict(str(tmp_path / "dict.txt"), 96)
    op = build_post_process({"name": "CTCLabelDecode", "character_dict_path": d, "cuda_speedup": True})
    crops = []
    for i, T in enumerate([12, 33, 7, 80, 1, 25]):
        probs, _ = synth.ctc_probs_numpy(40 + i, T, 1, 97)
        crops.append(torch.from_numpy(probs).cuda())
    one_by_one = [op(c)[0] for c in crops]
    packed = op.decode_batch(crops)
    assert [t for t, _ in packed] == [t for t, _ in one_by_one]
    assert np.allclose([c for _, c in packed], [c for _, c in one_by_one], rtol=1e-6, equal_nan=True)
    assert op.decode_batch([]) == []


@pytest.mark.parametrize("B,C,dtype", [(1, 2, "f32"), (37, 2, "f32"), (256, 4, "f16"), (5, 2, "numpy")])
def test_cls_postprocess_matches_reference_semantics(B, C, dtype):
    """R/pytocr/postprocess/cls_postprocess.py:11-20: (label_list[argmax], preds[i, argmax]), first maximum wins."""
    import torch
    from pytorchocr_b200.postprocess import build_post_process
    rng = np.random.default_rng(B)
    preds = rng.random((B, C)).astype(np.float32)
    preds[::3] = preds[::3, :1]                           # ties: every class equal -> index 0
    labels = ["0", "180", "90", "270"][:C]
    op = build_post_process({"name": "ClsPostProcess", "label_list": labels, "cuda_speedup": True})
    if dtype == "numpy":
        x, ref = preds, preds
    else:
        x = torch.from_numpy(preds).cuda()
        x = x.half() if dtype == "f16" else x
        ref = x.float().cpu().numpy()
    got = op(x)
    idx = ref.argmax(axis=1)
    want = [(labels[k], ref[i, k]) for i, k in enumerate(idx)]
    assert [g[0] for g in got] == [w[0] for w in want]
    assert np.array_equal(np.array([g[1] for g in got], np.float32), np.array([w[1] for w in want], np.float32))
    out, lab = op(x, label=[1, 0])
    assert lab == [(labels[1], 1.0), (labels[0], 1.0)] and len(out) == B


def test_ctc_nan_rows_follow_numpy(dict_path):
    """numpy's argmax / max (rec_postprocess.py:83-84) treat NaN as the maximum: the FIRST NaN of a row wins and the
    probability (and with it the line's confidence) is NaN; +inf and -inf in one row are not NaN."""
    import torch
    op, oracle, _ = _ops(dict_path)
    T, B, C = 12, 5, 6623
    rng = np.random.default_rng(8)
    p = rng.random((T, B, C)).astype(np.float32)
    p[3, 0, 4000] = np.nan; p[3, 0, 17] = np.nan        # two NaNs: index 17
    p[5, 1, 0] = np.nan                                 # NaN on the blank class
    p[:, 2, 123] = 5.0; p[7, 2, 6622] = np.nan          # NaN in the scalar tail of the row
    p[2, 3, 9] = np.inf; p[2, 3, 10] = -np.inf          # inf - inf trips the canary, but there is no NaN
    p[:, 4, 1] = 2.0; p[0, 4, 2] = np.nan               # NaN among the head elements
    for dt in (np.float32, np.float16):
        x = p.astype(dt)
        got = op(torch.from_numpy(x).cuda())
        want = oracle(torch.from_numpy(x))
        assert [g[0] for g in got] == [w[0] for w in want], dt
        assert np.allclose([g[1] for g in got], [w[1] for w in want], rtol=2e-3 if dt == np.float16 else 1e-6, equal_nan=True)
    assert math.isnan(got[0][1]) and not math.isnan(got[3][1])
