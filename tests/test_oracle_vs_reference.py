"""Pins the C/numpy oracle against the reference ITSELF: its unmodified Cython modules
(pse.pyx / pa.pyx compiled into oracle/_ref by oracle/build_ref.py) and cv2's CCL."""
import cv2
import numpy as np
import pytest

from oracle import clib, pan_oracle


def _fields(rng, H, W, K, blur):
    base = rng.random((H, W)).astype(np.float32)
    base = cv2.GaussianBlur(base, (0, 0), blur)
    qs = np.quantile(base, np.linspace(0.35, 0.8, K))
    # channel 0 = largest (text), last = smallest kernel; nested like real PSE maps
    return np.stack([(base > q) for q in qs]).astype(np.uint8)


@pytest.mark.parametrize("seed", range(6))
def test_ccl4_matches_cv2(seed):
    rng = np.random.default_rng(seed)
    img = (rng.random((61, 83)) > [0.3, 0.5, 0.6, 0.45, 0.55, 0.7][seed]).astype(np.uint8)
    n_cv, lab_cv = cv2.connectedComponents(img, connectivity=4)
    n, lab = clib.ccl4(img)
    assert n == n_cv
    assert np.array_equal(lab, lab_cv)


@pytest.mark.parametrize("seed", range(8))
def test_pse_matches_reference_cython(ref_modules, seed):
    pse_ref, _ = ref_modules
    rng = np.random.default_rng(100 + seed)
    K = [7, 7, 3, 2, 7, 5, 7, 4][seed]
    kernels = _fields(rng, 72, 96, K, [1.0, 2.0, 0.6, 1.5, 3.0, 0.4, 2.5, 1.2][seed])
    if seed % 2:  # non-nested adversarial: independent random masks times text
        kernels = ((rng.random((K, 72, 96)) > 0.45) & (kernels[0] > 0)).astype(np.uint8)
    for min_area in (0.0, 1.0, 5.0, 16.0):
        ref = pse_ref.pse(kernels.copy(), min_area)
        got = clib.pse(kernels, min_area)
        assert np.array_equal(ref, got), (seed, min_area)


@pytest.mark.parametrize("seed", range(8))
def test_pa_matches_reference_cython(ref_modules, seed):
    _, pa_ref = ref_modules
    rng = np.random.default_rng(200 + seed)
    H, W = 64, 80
    text = _fields(rng, H, W, 1, 2.0)[0]
    kern = ((rng.random((H, W)) > 0.6) & (text > 0)).astype(np.uint8)
    if seed >= 4:
        # large + tiny kernels in one text component so the 1024 ratio flag fires
        H, W = 96, 128
        text = np.zeros((H, W), np.uint8)
        text[4:60, 4:120] = 1
        kern = np.zeros((H, W), np.uint8)
        kern[6:40, 6:60] = 1          # 34*54 = 1836 px
        kern[50, 100] = 1             # 1 px  -> ratio 1836 > 1024
        kern[52:55, 70:74] = 1
    kernels = np.stack([text, kern])
    inst = rng.integers(0, 4, (H, W))
    centres = np.array([[0, 0, 0, 0], [6, 0, 0, 0], [0, 6, 0, 0], [0, 0, 6, 0]], np.float32)
    emb = (centres[inst].transpose(2, 0, 1) + rng.normal(0, 0.25, (4, H, W))).astype(np.float32)
    emb *= text[None].astype(np.float32)
    for min_area in (0.0, 2.6, 10.0):
        ref = pa_ref.pa(kernels.copy(), emb.copy(), min_area)
        got, flag, _ = pan_oracle.pa(kernels, emb, min_area)
        assert np.array_equal(ref, got), (seed, min_area)
        if seed >= 4 and min_area == 0.0:
            assert flag.sum() >= 2  # the gate path was exercised
