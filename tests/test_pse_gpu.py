"""GPU parity: PSENet post-processing through the C-ABI vs the oracle (oracle/pse_oracle.py, whose
expansion is pinned against the reference's own compiled pse.pyx in test_oracle_vs_reference.py)."""
import cv2
import numpy as np
import pytest

from expand_compare import compare_image, merge
from oracle.pse_oracle import PSEPostProcessOracle
from pytorchocr_b200 import synth

pytestmark = pytest.mark.gpu

CFG = dict(thresh=0, box_thresh=0.85, min_area=16, scale=1)   # det_r50_pse.yml:56-62


def _op(**kw):
    from pytorchocr_b200.postprocess import build_post_process
    cfg = dict(CFG, name="PSEPostProcess", cuda_speedup=True)
    cfg.update(kw)
    return build_post_process(cfg, {"use_gpu": True})


def _check(maps, shape_list, loose=0.02, **kw):
    import torch
    op = _op(**kw)
    dev = torch.from_numpy(maps).cuda() if isinstance(maps, np.ndarray) else maps
    boxes, scores, counts, status, ex = op.run_device(dev, shape_list, boxes_f=True, labels=True)
    ref_in = maps if isinstance(maps, np.ndarray) else maps.float().cpu().numpy()
    ocfg = dict(CFG)
    ocfg.update(kw)
    want = PSEPostProcessOracle(**ocfg)({"maps": ref_in}, shape_list, return_details=True)
    tot = {}
    for n in range(len(want)):
        assert np.array_equal(want[n]["label_proc"], ex["labels"][n]), "label map differs (image %d)" % n
        k = int(counts[n])
        merge(tot, compare_image(boxes[n, :k], ex["boxes_f"][n, :k], scores[n, :k], want[n], np.asarray(shape_list, np.float64)[n]))
    assert tot.get("tie", 0) + tot.get("ordering", 0) <= max(1, loose * tot.get("n", 0)), sorted(tot.items())
    return want, counts


def _shape(N, H, W):
    return np.array([[H, W, 1.0, 1.0]] * N, np.float64)


@pytest.mark.parametrize("H,W", [(192, 320), (97, 131), (64, 64), (256, 1000)])
def test_pse_synth_processing_res(H, W):
    maps = np.stack([synth.pse_maps(synth.BASE_SEED + i + H, H, W, n_regions=200) for i in range(3)])
    sl = np.array([[H, W, 1.0, 1.0], [2 * H, 2 * W, 2.0, 2.0], [H // 2 + 3, W // 2 + 5, 0.5, 0.5]], np.float64)
    want, counts = _check(maps, sl, maps_at_processing_res=True)
    assert counts.sum() > 0 or H * W < 10000


@pytest.mark.parametrize("scale", [1, 2, 4])
def test_pse_reference_scales(scale):
    """API-faithful: 1/4-resolution head output, up-sampled by 4 // scale, labels up-sampled by scale."""
    h, w = 48, 80
    maps = np.stack([synth.pse_maps(7 + i, h, w, n_abs=24, hh_rng=(3, 6), hw_rng=(6, 12)) for i in range(2)])
    H, W = 4 * h, 4 * w
    sl = np.array([[H, W, 1.0, 1.0], [int(H * 1.5), int(W * 1.25), 1.0 / 1.5, 1.0 / 1.25]], np.float64)
    _check(maps, sl, scale=scale, loose=0.1)


def test_pse_full_size():
    """BASELINE.json config 3 shape: 7 kernel maps at 736x1280 (2 images)."""
    maps = np.stack([synth.pse_maps(synth.BASE_SEED + i) for i in range(2)])
    want, counts = _check(maps, _shape(2, 736, 1280), maps_at_processing_res=True)
    assert counts.min() > 100


def _blob_fields(rng, K, H, W, blur, nested):
    base = cv2.GaussianBlur(rng.random((H, W)).astype(np.float32), (0, 0), blur)
    qs = np.quantile(base, np.linspace(0.35, 0.8, K))
    m = np.stack([(base > q) for q in qs])
    if not nested:
        m = (rng.random((K, H, W)) > 0.45) & m[0]
    return np.where(m, 3.0, -3.0).astype(np.float32) + rng.normal(0, 0.3, (K, H, W)).astype(np.float32)


@pytest.mark.parametrize("seed", range(8))
def test_pse_adversarial_fields(seed):
    """Contested expansion: blob fields with merged text masks and non-nested kernels (the same
    family that pins the oracle against the reference's Cython module)."""
    rng = np.random.default_rng(100 + seed)
    K = [7, 7, 3, 2, 7, 5, 7, 4][seed]
    H, W = 72, 96
    maps = np.stack([_blob_fields(rng, K, H, W, [1.0, 2.0, 0.6, 1.5, 3.0, 0.4, 2.5, 1.2][seed], seed % 2 == 0)
                     for _ in range(2)])
    for min_area in (0, 5, 16):
        _check(maps, _shape(2, H, W), maps_at_processing_res=True, min_area=min_area, box_thresh=0.5, loose=0.3)


def test_pse_one_component_whole_image():
    """One text component covering the whole image with many seeds: the general (large) path."""
    rng = np.random.default_rng(5)
    H, W, K = 160, 224, 4
    m = np.zeros((K, H, W), bool)
    m[0] = True
    for k in range(1, K):
        m[k] = cv2.GaussianBlur(rng.random((H, W)).astype(np.float32), (0, 0), 3.0) > [0, 0.49, 0.5, 0.51][k]
    maps = np.where(m, 3.0, -3.0).astype(np.float32)[None]
    _check(maps, _shape(1, H, W), maps_at_processing_res=True, min_area=4, box_thresh=0.5, loose=0.3)


def test_pse_empty_and_api():
    import torch
    H, W = 64, 96
    z = np.full((2, 7, H, W), -4.0, np.float32)
    z[1, :, 10:40, 10:80] = 4.0
    op = _op(maps_at_processing_res=True)
    res = op({"maps": torch.from_numpy(z).cuda()}, _shape(2, H, W))
    assert res[0]["points"].shape == (0,) and res[0]["scores"] == []
    assert res[1]["points"].shape == (1, 4, 2) and res[1]["points"].dtype == np.int16
    assert np.array_equal(res[1]["points"][0], np.array([[10, 10], [79, 10], [79, 39], [10, 39]]))
    with pytest.raises(AssertionError):
        op({"maps": z}, _shape(2, H, W))           # the reference asserts a torch.Tensor (:30)
    # fp16 maps and a strided view
    half = torch.from_numpy(z).half().cuda()
    res16 = op({"maps": half}, _shape(2, H, W))
    assert np.array_equal(res16[1]["points"], res[1]["points"])
    wide = torch.full((2, 9, H, W + 4), -4.0, device="cuda")
    wide[:, 1:8, :, :W] = torch.from_numpy(z).cuda()
    resv = op({"maps": wide[:, 1:8, :, :W]}, _shape(2, H, W))
    assert np.array_equal(resv[1]["points"], res[1]["points"])


def test_pse_capacity_retry():
    rng = np.random.default_rng(0)
    H, W = 64, 64
    maps = np.where(rng.random((1, 3, H, W)) > 0.4, 3.0, -3.0).astype(np.float32)   # ~H*W/4 runs
    _check(maps, _shape(1, H, W), maps_at_processing_res=True, max_runs=32, max_boxes=2, min_area=2,
           box_thresh=0.5, loose=0.5)


def test_pse_text_run_table_overflows_alone():
    """Only the TEXT mask overflows the run table (striped text, a few compact kernels) and the workspace holds
    stale data: the first pass must bail out cleanly (found by tests/stress_gpu.py: the seed lookup walked a text
    table that had not been written) and the retry must match the oracle."""
    import torch
    H, W = 64, 128
    maps = np.full((2, 3, H, W), -3.0, np.float32)
    maps[:, 0, :, ::2] = 3.0                       # text: 64 one-pixel runs per row = 4096 runs
    maps[:, 0, 20:30, 40:80] = 3.0
    maps[:, 1:, 22:28, 44:60] = 3.0                # one kernel blob inside a solid text block
    maps[:, 1:, 10:14, 8] = 3.0                    # and a 1-px-wide kernel on a stripe
    junk = torch.full((64 << 20,), 0x7f7f7f7f, dtype=torch.int32, device="cuda")   # poison the allocator's pool
    del junk
    _check(maps, _shape(2, H, W), maps_at_processing_res=True, max_runs=1024, min_area=2, box_thresh=0.5, loose=0.5)


def test_pse_row_extent_blocks_reserved_and_deferred():
    """Row extents are filled in the single statistics pass from blocks reserved per seed (rows of its text
    component); seeds whose block does not fit the reserved half of the storage take the exact two-pass path. One
    tall text column with 8 stacked kernels and a run capacity that lets only some reservations through exercises
    both paths in one image."""
    H, W = 72, 64
    maps = np.full((2, 3, H, W), -3.0, np.float32)
    maps[:, 0, 4:68, 20:40] = 3.0                   # one text component of 64 rows
    for i in range(8):
        maps[:, 1:, 6 + 8 * i:10 + 8 * i, 24:36] = 3.0   # 8 kernels, each grows into ~8 rows of the column
    maps[1, 0, 30:34, 44:60] = 3.0                  # second image: an extra short component with its own kernel
    maps[1, 1:, 31:33, 46:58] = 3.0
    _check(maps, _shape(2, H, W), maps_at_processing_res=True, max_runs=100, min_area=2, box_thresh=0.5, loose=0.5)


def test_pse_host_batch_is_uploaded_in_chunks():
    """CPU-tensor maps of a large batch take the chunked upload path: same results as the device-tensor path."""
    import torch
    H, W = 96, 128
    maps = np.stack([synth.pse_maps(60 + i, H, W) for i in range(9)])
    sl = _shape(9, H, W)
    op = _op(maps_at_processing_res=True)
    op.upload_chunk_bytes = 2 * maps[0].nbytes            # 9 images -> chunks of 2 (5 chunks, unequal sizes)
    want = op({"maps": torch.from_numpy(maps).cuda()}, sl)
    got = op({"maps": torch.from_numpy(maps).pin_memory()}, sl)
    assert len(got) == len(want) == 9
    for g, w in zip(got, want):
        assert np.array_equal(g["points"], w["points"]) and np.allclose(g["scores"], w["scores"])
