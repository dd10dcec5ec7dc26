"""Golden vectors produced by running the REFERENCE ITSELF (tests/golden/make_golden.py: the unmodified
operator classes and compiled pse.pyx / pa.pyx from /root/reference) checked against
  * the oracle (CPU, always)          - pins the oracle to the reference without /root/reference
  * the CUDA path (-m gpu, B200 box)  - parity with the reference's own outputs."""
import os

import numpy as np
import pytest

from conftest import ROOT
from pytorchocr_b200 import synth

G = np.load(os.path.join(ROOT, "tests", "golden", "reference_outputs.npz"))
PSE = dict(thresh=0, box_thresh=0.85, min_area=16)
PAN = dict(thresh=0, box_thresh=0.85, min_area=16, min_kernel_area=2.6)
DB = dict(thresh=0.3, box_thresh=0.5, max_candidates=1000, unclip_ratio=1.7, score_mode="poly", cpp_speedup=True)


def _kernels(prefix, i, shape):
    return np.unpackbits(G["%s_kernels_%d" % (prefix, i)])[:int(np.prod(shape))].reshape(shape)


def _same_boxes(res, prefix, exact=True):
    for n, r in enumerate(res):
        want = G["%s_points_%d" % (prefix, n)]
        got = np.asarray(r["points"], np.int16).reshape(-1, 4, 2)
        assert got.shape == want.shape, (prefix, n, got.shape, want.shape)
        if exact:
            assert np.array_equal(got, want), (prefix, n)
        ws = G["%s_scores_%d" % (prefix, n)]
        assert np.allclose(np.asarray(r["scores"], np.float32), ws, rtol=1e-5, atol=1e-7)


# ------------------------------------------------------------------ oracle vs the reference's outputs (CPU)
@pytest.mark.parametrize("scale", [1, 2, 4])
def test_oracle_pse_operator(scale):
    from oracle.pse_oracle import PSEPostProcessOracle
    res = PSEPostProcessOracle(scale=scale, **PSE)({"maps": G["pse_maps"].astype(np.float32)}, G["pse_shape"])
    _same_boxes(res, "pse_s%d" % scale)


@pytest.mark.parametrize("scale", [1, 2, 4])
def test_oracle_pan_operator(scale):
    from oracle.pan_oracle import PANPostProcessOracle
    res = PANPostProcessOracle(scale=scale, **PAN)({"maps": G["pan_maps"].astype(np.float32)}, G["pan_shape"])
    _same_boxes(res, "pan_s%d" % scale)


def test_oracle_expansion_labels():
    from oracle import clib, pan_oracle
    for i in range(4):
        k = _kernels("exp_pse", i, G["exp_pse_shape_%d" % i])
        for ma in (0, 5):
            assert np.array_equal(clib.pse(k, float(ma)), G["exp_pse_label_%d_ma%d" % (i, ma)])
        k = _kernels("exp_pa", i, (2, 96, 128))
        emb = G["exp_pa_emb_%d" % i].astype(np.float32) * k[0][None].astype(np.float32)
        for ma in (0.0, 2.6):
            lab, _, _ = pan_oracle.pa(k, emb, ma)
            assert np.array_equal(lab, G["exp_pa_label_%d_ma%d" % (i, int(ma))])


def test_oracle_ctc(tmp_path):
    from oracle.ctc_oracle import CTCLabelDecodeNumpy, CTCLabelDecodeOracle
    import torch
    probs = G["ctc_probs"]
    d = synth.write_char_dict(str(tmp_path / "dict.txt"), probs.shape[2] - 1)
    for cls in (CTCLabelDecodeNumpy, CTCLabelDecodeOracle):
        res = cls(d)(torch.from_numpy(probs))
        assert [r[0] for r in res] == G["ctc_text"].tolist()
        assert np.allclose(np.array([r[1] for r in res], np.float32), G["ctc_conf"], rtol=1e-6, equal_nan=True)


def test_oracle_db_operator():
    from oracle.db_oracle import DBPostProcessOracle
    res = DBPostProcessOracle(**DB)({"maps": G["db_maps"].astype(np.float32)}, G["db_shape"])
    _same_boxes(res, "db")


# ------------------------------------------------------------------ CUDA path vs the reference's outputs
def _build(cfg):
    from pytorchocr_b200.postprocess import build_post_process
    return build_post_process(dict(cfg, cuda_speedup=True), {"use_gpu": True})


@pytest.mark.gpu
@pytest.mark.parametrize("scale", [1, 2, 4])
def test_cuda_pse_operator(scale):
    import torch
    op = _build(dict(PSE, name="PSEPostProcess", scale=scale))
    res = op({"maps": torch.from_numpy(G["pse_maps"].astype(np.float32)).cuda()}, G["pse_shape"])
    _same_boxes(res, "pse_s%d" % scale)


@pytest.mark.gpu
@pytest.mark.parametrize("scale", [1, 2, 4])
def test_cuda_pan_operator(scale):
    import torch
    op = _build(dict(PAN, name="PANPostProcess", scale=scale))
    res = op({"maps": torch.from_numpy(G["pan_maps"].astype(np.float32)).cuda()}, G["pan_shape"])
    _same_boxes(res, "pan_s%d" % scale)


@pytest.mark.gpu
def test_cuda_expansion_labels():
    """Label maps of the reference's compiled pse() / pa() reproduced bit for bit."""
    import torch
    for i in range(4):
        k = _kernels("exp_pse", i, G["exp_pse_shape_%d" % i])
        K, H, W = k.shape
        maps = torch.from_numpy(np.where(k > 0, 1.0, -1.0).astype(np.float32)[None]).cuda()
        for ma in (0, 5):
            op = _build(dict(PSE, name="PSEPostProcess", scale=1, min_area=ma, maps_at_processing_res=True))
            ex = op.run_device(maps, [[H, W, 1.0, 1.0]], labels=True)[4]
            assert np.array_equal(ex["labels"][0], G["exp_pse_label_%d_ma%d" % (i, ma)]), (i, ma)
        k = _kernels("exp_pa", i, (2, 96, 128))
        maps = np.empty((1, 6, 96, 128), np.float32)
        maps[0, :2] = np.where(k > 0, 1.0, -1.0)
        maps[0, 2:] = G["exp_pa_emb_%d" % i].astype(np.float32)
        for ma in (0.0, 2.6):
            op = _build(dict(PAN, name="PANPostProcess", scale=1, min_kernel_area=ma, maps_at_processing_res=True))
            ex = op.run_device(torch.from_numpy(maps).cuda(), [[96, 128, 1.0, 1.0]], labels=True)[4]
            assert np.array_equal(ex["labels"][0], G["exp_pa_label_%d_ma%d" % (i, int(ma))]), (i, ma)


@pytest.mark.gpu
def test_cuda_ctc(tmp_path):
    import torch
    probs = G["ctc_probs"]
    d = synth.write_char_dict(str(tmp_path / "dict.txt"), probs.shape[2] - 1)
    res = _build({"name": "CTCLabelDecode", "character_dict_path": d, "use_space_char": False})(torch.from_numpy(probs).cuda())
    assert [r[0] for r in res] == G["ctc_text"].tolist()
    assert np.allclose(np.array([r[1] for r in res], np.float32), G["ctc_conf"], rtol=1e-6, equal_nan=True)


@pytest.mark.gpu
def test_cuda_db_operator():
    import torch
    res = _build(dict(DB, name="DBPostProcess"))({"maps": torch.from_numpy(G["db_maps"].astype(np.float32)).cuda()},
                                                 G["db_shape"])
    for n, r in enumerate(res):
        want = G["db_points_%d" % n].reshape(-1, 8)
        got = np.asarray(r["points"], np.int16).reshape(-1, 8)
        assert got.shape == want.shape
        a = np.array(sorted(map(tuple, got.tolist())))
        b = np.array(sorted(map(tuple, want.tolist())))
        diff = np.abs(a - b).max(1)
        assert (diff > 0).sum() <= 1 and diff.max() <= 1   # <= 1 box on a rounding discontinuity
        assert r["scores"] == [1.0] * len(got)
