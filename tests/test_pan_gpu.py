"""GPU parity: PAN / PAN++ post-processing through the C-ABI vs the oracle (oracle/pan_oracle.py,
pinned against the reference's own compiled pa.pyx in test_oracle_vs_reference.py)."""
import cv2
import numpy as np
import pytest

from expand_compare import compare_image, merge
from oracle.pan_oracle import PANPostProcessOracle
from pytorchocr_b200 import synth

pytestmark = pytest.mark.gpu

CFG = dict(thresh=0, box_thresh=0.85, min_area=16, min_kernel_area=2.6, scale=4)   # det_r18_pan.yml:62-69


def _op(**kw):
    from pytorchocr_b200.postprocess import build_post_process
    cfg = dict(CFG, name="PANPostProcess", cuda_speedup=True)
    cfg.update(kw)
    return build_post_process(cfg, {"use_gpu": True})


def _check(maps, shape_list, loose=0.02, expect_flags=False, **kw):
    import torch
    op = _op(**kw)
    dev = torch.from_numpy(maps).cuda()
    boxes, scores, counts, status, ex = op.run_device(dev, shape_list, boxes_f=True, labels=True)
    ocfg = dict(CFG)
    ocfg.update(kw)
    want = PANPostProcessOracle(**ocfg)({"maps": maps}, shape_list, return_details=True)
    tot = {}
    flags = 0
    for n in range(len(want)):
        assert np.array_equal(want[n]["label_proc"], ex["labels"][n]), "label map differs (image %d)" % n
        k = int(counts[n])
        merge(tot, compare_image(boxes[n, :k], ex["boxes_f"][n, :k], scores[n, :k], want[n], np.asarray(shape_list, np.float64)[n]))
        flags += int(want[n]["flag"].sum())
    assert tot.get("tie", 0) + tot.get("ordering", 0) <= max(1, loose * tot.get("n", 0)), sorted(tot.items())
    if expect_flags:
        assert flags >= 2, "the embedding gate was not exercised"
    return want, counts


def _shape(N, H, W):
    return np.array([[H, W, 1.0, 1.0]] * N, np.float64)


@pytest.mark.parametrize("H,W", [(192, 320), (97, 131), (256, 1000)])
def test_pan_synth_processing_res(H, W):
    maps = np.stack([synth.pan_maps(synth.BASE_SEED + i + H, H, W, n_regions=200) for i in range(2)])
    want, counts = _check(maps, _shape(2, H, W), scale=1, maps_at_processing_res=True)
    assert counts.sum() > 0


@pytest.mark.parametrize("scale", [1, 2, 4])
def test_pan_reference_scales(scale):
    """API-faithful: 1/4-resolution head output [N,6,h,w]; shipped config is scale=4."""
    h, w = 48, 80
    maps = np.stack([synth.pan_maps(11 + i, h, w, n_abs=24, hh_rng=(3, 6), hw_rng=(6, 12)) for i in range(2)])
    H, W = 4 * h, 4 * w
    sl = np.array([[H, W, 1.0, 1.0], [int(H * 1.5), int(W * 1.25), 1.0 / 1.5, 1.0 / 1.25]], np.float64)
    _check(maps, sl, scale=scale, loose=0.1)


def test_pan_full_size():
    """BASELINE.json config 4 shape: [.,6,736,1280] at processing resolution."""
    maps = np.stack([synth.pan_maps(synth.BASE_SEED + i) for i in range(2)])
    want, counts = _check(maps, _shape(2, 736, 1280), scale=1, maps_at_processing_res=True)
    assert counts.min() > 100


def _gate_scene(rng, H=96, W=128):
    """One text component holding a large kernel (> 1024 x the tiny one), a 1-px kernel and a medium
    one, plus a second text component; embeddings cluster per instance so the distance-3 gate both
    passes and blocks claims."""
    text = np.zeros((H, W), bool)
    text[4:60, 4:120] = True
    text[70:90, 10:100] = True
    kern = np.zeros((H, W), bool)
    kern[6:40, 6:60] = True        # 1836 px
    kern[50, 100] = True           # 1 px -> ratio > 1024: both flagged
    kern[52:55, 70:74] = True      # 12 px
    kern[75:85, 20:60] = True
    inst = np.zeros((H, W), np.int64)
    inst[:, 64:] = 1
    inst[44:, :] += 2
    centres = np.array([[0, 0, 0, 0], [6, 0, 0, 0], [0, 6, 0, 0], [0, 0, 6, 0]], np.float32)
    emb = centres[inst].transpose(2, 0, 1) + rng.normal(0, 0.25, (4, H, W))
    maps = np.empty((6, H, W), np.float32)
    maps[0] = np.where(text, 4.0, -4.0) + rng.normal(0, 0.3, (H, W))
    maps[1] = np.where(kern, 4.0, -4.0)
    maps[2:] = emb
    return maps


@pytest.mark.parametrize("seed", range(4))
def test_pan_flagged_kernels_and_gate(seed):
    rng = np.random.default_rng(300 + seed)
    maps = np.stack([_gate_scene(rng), _gate_scene(rng)])
    for mka in (0.0, 2.6, 10.0):
        _check(maps, _shape(2, 96, 128), scale=1, maps_at_processing_res=True, min_kernel_area=mka,
               min_area=4, box_thresh=0.5, loose=0.3, expect_flags=(mka <= 1.0))


@pytest.mark.parametrize("seed", range(4))
def test_pan_adversarial_fields(seed):
    rng = np.random.default_rng(200 + seed)
    H, W = 64, 80
    base = cv2.GaussianBlur(rng.random((H, W)).astype(np.float32), (0, 0), 2.0)
    text = base > np.quantile(base, 0.35)
    kern = (rng.random((H, W)) > 0.6) & text
    inst = rng.integers(0, 4, (H, W))
    centres = np.array([[0, 0, 0, 0], [6, 0, 0, 0], [0, 6, 0, 0], [0, 0, 6, 0]], np.float32)
    maps = np.empty((1, 6, H, W), np.float32)
    maps[0, 0] = np.where(text, 3.0, -3.0)
    maps[0, 1] = np.where(kern, 3.0, -3.0)
    maps[0, 2:] = centres[inst].transpose(2, 0, 1) + rng.normal(0, 0.25, (4, H, W))
    for mka in (0.0, 2.6):
        _check(maps, _shape(1, H, W), scale=1, maps_at_processing_res=True, min_kernel_area=mka,
               min_area=2, box_thresh=0.5, loose=0.5)


def test_pan_empty_and_api():
    import torch
    H, W = 64, 96
    z = np.full((2, 6, H, W), -4.0, np.float32)
    z[1, :2, 10:40, 10:80] = 4.0
    op = _op(scale=1, maps_at_processing_res=True)
    res = op({"maps": torch.from_numpy(z).cuda()}, _shape(2, H, W))
    assert res[0]["points"].shape == (0,) and res[0]["scores"] == []
    assert np.array_equal(res[1]["points"][0], np.array([[10, 10], [79, 10], [79, 39], [10, 39]]))
    assert isinstance(res[1]["scores"][0], np.float32)
