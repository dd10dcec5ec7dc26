"""GPU parity: DB / DB++ box extraction through the C-ABI vs the cv2-based oracle
(oracle/db_oracle.py). Labels bit-exact after canonical renumbering, scores 1e-5 relative,
vertices 1e-3 px pre-rounding, integer boxes exact outside the rounding band."""
import cv2
import numpy as np
import pytest
from scipy import ndimage as ndi

from db_compare import canonical_labels, compare_image, merge
from oracle.db_oracle import DBPostProcessOracle
from pytorchocr_b200 import synth

pytestmark = pytest.mark.gpu

CFG = dict(thresh=0.3, box_thresh=0.5, max_candidates=1000, unclip_ratio=1.7, score_mode="poly", cpp_speedup=True)


@pytest.fixture(autouse=True, params=["image_smem", "image_global", "chain", "image_scan1", "image_scan2"])
def db_path(request):
    """Every test of this module runs on each of the library's three equivalent stage-2 code paths (include/ocrpp.h
    OCRPP_TUNE_DB_PATH): one CTA per image with the tables in shared memory (the default for maps up to 4 Mpx), the
    same kernel with the tables in the global workspace (what an image takes whose tables do not fit), and the
    run-parallel multi-kernel chain (what larger maps take). The default map scan in front of the one-kernel stage is
    db_scan4_kernel (bulk-copy fed); "image_scan1" runs it behind the warp-per-row db_scan_kernel (what dilation,
    unaligned maps and the chain take), "image_scan2" behind the opt-in two-phase scan (db_scan2/3_kernel)."""
    from pytorchocr_b200 import _lib
    L = _lib.lib()
    _lib.check(L.ocrpp_set_tuning(_lib.TUNE_DB_PATH, {"image_smem": 1, "image_global": 2, "chain": 3, "image_scan1": 1, "image_scan2": 1}[request.param]))
    _lib.check(L.ocrpp_set_tuning(_lib.TUNE_DB_SCAN, {"image_scan1": 1, "image_scan2": 2}.get(request.param, 0)))
    yield request.param
    _lib.check(L.ocrpp_set_tuning(_lib.TUNE_DB_PATH, 0))
    _lib.check(L.ocrpp_set_tuning(_lib.TUNE_DB_SCAN, 0))


def _op(**kw):
    from pytorchocr_b200.postprocess import build_post_process
    cfg = dict(CFG, name="DBPostProcess", cuda_speedup=True)
    cfg.update(kw)
    return build_post_process(cfg, {"use_gpu": True})


def _check(maps, shape_list, op=None, oracle_maps=None, loose=0.02, max_unmatched=None, **kw):
    """Runs the CUDA path and the cv2-based oracle on the same maps. `loose` = fraction of boxes
    allowed to sit on one of the reference's own discontinuities (see db_compare.py)."""
    import torch
    op = op or _op(**kw)
    dev_in = torch.from_numpy(maps).cuda() if isinstance(maps, np.ndarray) else maps
    boxes, scores, counts, status, ex = op.run_device(dev_in, shape_list, boxes_f=True, labels=True)
    ref_in = oracle_maps if oracle_maps is not None else (maps if isinstance(maps, np.ndarray) else maps.float().cpu().numpy())
    ocfg = dict(CFG)
    ocfg.update(kw)
    want = DBPostProcessOracle(**ocfg)({"maps": ref_in}, shape_list, return_details=True)
    tot = {}
    for n in range(len(want)):
        k = int(counts[n])
        merge(tot, compare_image(boxes[n, :k], ex["boxes_f"][n, :k], scores[n, :k], want[n]["details"]))
        seg = ref_in[n, 0] > ocfg["thresh"]
        if ocfg.get("use_dilation"):
            seg = cv2.dilate(seg.astype(np.uint8), np.array([[1, 1], [1, 1]], np.uint8)) > 0
        fg, _ = ndi.label(seg, structure=np.ones((3, 3)))
        assert np.array_equal(canonical_labels(fg), ex["labels"][n]), "label map differs"
    n_boxes = max(1, tot.get("n_oracle", 0))
    off = tot.get("n_oracle", 0) - tot.get("exact", 0) + tot.get("unmatched_gpu", 0)
    unmatched = tot.get("unmatched_gpu", 0) + tot.get("unmatched_oracle", 0)
    if max_unmatched is not None:
        # stress mode: any number of boxes on a VERIFIED discontinuity of the reference, none unexplained
        if unmatched > max_unmatched:
            pytest.fail("unexplained boxes: %s" % sorted(tot.items()))
    elif off > max(2, loose * n_boxes) or unmatched > max(1, 0.25 * loose * n_boxes):
        pytest.fail("too many boxes off the oracle: %s" % sorted(tot.items()))
    return want, counts


@pytest.mark.parametrize("H,W", [(192, 320), (97, 131), (64, 64), (256, 1280), (33, 1000)])
def test_db_synth_small(H, W):
    maps = synth.db_batch(3, seed=synth.BASE_SEED + H, H=H, W=W)
    shape_list = np.array([[H, W, 1.0, 1.0], [H * 2, W * 2, 2.0, 2.0], [H // 2 + 7, W // 2 + 3, 0.5, 0.5]], np.float64)
    want, counts = _check(maps, shape_list)
    assert counts.sum() > 0 or H * W < 40000


def test_db_full_size_cfg1():
    """BASELINE.json config 1: one 1x1x736x1280 map, seed 20221001."""
    maps = synth.db_batch(1)
    want, counts = _check(maps, np.array([[736, 1280, 1.0, 1.0]]))
    assert 150 <= counts[0] <= 260


def test_db_batch_full_size():
    maps = synth.db_batch(4, seed=synth.BASE_SEED + 17)
    _check(maps, np.array([[736, 1280, 1.0, 1.0]] * 4), loose=0.02)


@pytest.mark.parametrize("seed", range(6))
def test_db_random_blob_fields(seed):
    """Adversarial topology: many holes, islands in holes, diagonal pinches, specks."""
    rng = np.random.default_rng(seed)
    H, W = 120, 152
    sig = [0.6, 1.0, 1.5, 2.5, 0.45, 2.0][seed]
    p = cv2.GaussianBlur(rng.random((2, H, W)).astype(np.float32), (0, 0), sig)
    p = np.stack([cv2.GaussianBlur(rng.random((H, W)).astype(np.float32), (0, 0), sig) for _ in range(2)])
    lo, hi = np.quantile(p, 0.02), np.quantile(p, 0.98)
    p = np.clip((p - lo) / (hi - lo), 0, 1).astype(np.float32)
    q = float(np.quantile(p, [0.45, 0.5, 0.55, 0.6, 0.5, 0.4][seed]))
    maps = p[:, None]
    _check(maps, np.array([[H, W, 1.0, 1.0]] * 2), thresh=q, box_thresh=q + 0.02, loose=0.35)


def test_db_nested_rings_and_edges():
    H, W = 96, 128
    m = np.full((H, W), 0.02, np.float32)
    m[8:88, 8:120] = 0.9       # big block
    m[16:80, 16:112] = 0.05    # hole
    m[24:72, 24:104] = 0.8     # island in the hole
    m[32:64, 32:96] = 0.1      # hole in the island
    m[40:56, 40:88] = 0.95     # island in that hole
    m[0:5, 0:9] = 0.9          # touches the corner
    m[90:96, 50:70] = 0.9      # touches the bottom edge
    m[2, 100] = 0.9            # single pixel
    m[60:70, 124] = 0.9        # vertical 1-px run
    for i in range(6):
        m[85 + i if 85 + i < H else H - 1, 2 + i] = 0.9   # diagonal 1-px run
    maps = m[None, None]
    _check(maps, np.array([[H, W, 1.0, 1.0]]), box_thresh=0.3)


def test_db_empty_and_full():
    H, W = 64, 96
    z = np.zeros((2, 1, H, W), np.float32)
    z[1] = 1.0
    want, counts = _check(z, np.array([[H, W, 1.0, 1.0]] * 2))
    assert counts[0] == 0
    op = _op()
    import torch
    res = op({"maps": torch.from_numpy(z).cuda()}, np.array([[H, W, 1.0, 1.0]] * 2))
    assert res[0]["points"].shape == (0,) and res[0]["scores"] == []
    assert res[1]["points"].dtype == np.int16


def test_db_input_kinds_and_fp16():
    import torch
    H, W = 160, 256
    maps = synth.db_batch(2, seed=5, H=H, W=W)
    sl = np.array([[H, W, 1.0, 1.0]] * 2)
    op = _op()
    ref = DBPostProcessOracle(**CFG)({"maps": maps}, sl)
    def same(res):
        for r, w in zip(res, ref):
            a = np.array(sorted(map(tuple, r["points"].reshape(-1, 8).tolist())))
            b = np.array(sorted(map(tuple, w["points"].reshape(-1, 8).tolist())))
            assert a.shape == b.shape
            diff = np.abs(a - b).max(1)
            assert (diff > 0).sum() <= 1 and diff.max() <= 1   # <= 1 box on a rounding discontinuity
            assert r["scores"] == w["scores"]
    same(op({"maps": maps}, sl))                                   # numpy (TRT path, infer_det_trt.py:148-151)
    same(op({"maps": torch.from_numpy(maps)}, sl))                 # CPU tensor
    multi = torch.zeros((2, 3, H, W + 8), dtype=torch.float32, device="cuda")
    multi[:, 0, :, :W] = torch.from_numpy(maps[:, 0]).cuda()
    same(op({"maps": multi[:, :, :, :W]}, sl))                     # strided view, channel 0 of 3
    half = torch.from_numpy(maps).half()
    _check(half.cuda(), sl, oracle_maps=half.numpy())              # fp16 map: numpy thresholds it in float16, as we do


def test_db_run_capacity_retry_and_bad_values():
    import torch
    from pytorchocr_b200 import _lib
    H, W = 64, 64
    rng = np.random.default_rng(0)
    noise = (rng.random((1, 1, H, W)) > 0.5).astype(np.float32)   # ~H*W/2 runs >> default capacity floor
    op = _op(max_runs=64)
    want, counts = _check(noise, np.array([[H, W, 1.0, 1.0]]), op=op, loose=0.4)
    bad = noise.copy()
    bad[0, 0, 3, 3] = np.nan
    with pytest.raises(_lib.OcrppError):
        _op()({"maps": torch.from_numpy(bad).cuda()}, np.array([[H, W, 1.0, 1.0]]))


@pytest.mark.parametrize("src_hw", [(480, 640), (900, 600), (512, 512)])
def test_db_use_padding_resize(src_hw):
    """db_postprocess.cpp:293-302: boxes mapped back through the inverse padding-resize affine transform."""
    import torch
    H = W = 256                                   # padded square map
    maps = synth.db_batch(2, seed=11, H=H, W=W)
    src_h, src_w = src_hw
    sl = np.array([[src_h, src_w, 1.0, 1.0]] * 2, np.float64)
    op = _op()
    got = op({"maps": torch.from_numpy(maps).cuda()}, sl, use_padding_resize=True)
    want = DBPostProcessOracle(**CFG)({"maps": maps}, sl, use_padding_resize=True)
    plain = op({"maps": torch.from_numpy(maps).cuda()}, sl)
    n_diff = 0
    for g, w, q in zip(got, want, plain):
        a = np.array(sorted(map(tuple, g["points"].reshape(-1, 8).tolist())))
        b = np.array(sorted(map(tuple, w["points"].reshape(-1, 8).tolist())))
        assert a.shape == b.shape and len(a) > 0
        diff = np.abs(a - b).max(1)
        assert (diff > 0).sum() <= 1 and diff.max() <= 1       # <= 1 box on a rounding discontinuity
        n_diff += int(not np.array_equal(g["points"], q["points"]))
    assert n_diff > 0 or src_h == src_w                       # the flag changes the mapping unless the source is square


@pytest.mark.parametrize("H,W,dtype", [(192, 320, "f32"), (97, 131, "f32"), (256, 1280, "f32"), (33, 1000, "f16"),
                                       (160, 256, "f16")])
def test_db_use_dilation(H, W, dtype):
    """db_postprocess.py:52-55: contours come from cv2.dilate(segmentation, [[1,1],[1,1]])."""
    import torch
    maps = synth.db_batch(3, seed=synth.BASE_SEED + 5 * H, H=H, W=W)
    sl = np.array([[H, W, 1.0, 1.0], [H * 2, W * 2, 2.0, 2.0], [H // 2 + 7, W // 2 + 3, 0.5, 0.5]], np.float64)
    if dtype == "f16":
        dev = torch.from_numpy(maps).cuda().half()
        _check(dev, sl, oracle_maps=dev.cpu().numpy(), use_dilation=True, loose=0.05)
    else:
        # dilated blobs have longer axis-parallel hull edges, so more mini-boxes land within cv2's float32
        # noise of an integer corner (the "truncation" class of db_compare.py) than in the undilated tests
        _check(maps, sl, use_dilation=True, loose=0.05)


def test_db_use_dilation_blob_field():
    """Dilation merges diagonal / one-pixel-gap neighbours: adversarial topology through the dilated scan."""
    rng = np.random.default_rng(4)
    H, W = 120, 152
    f = ndi.gaussian_filter(rng.standard_normal((2, 1, H, W)), (0, 0, 1.2, 1.2))
    maps = (1.0 / (1.0 + np.exp(-6.0 * f / f.std()))).astype(np.float32)
    _check(maps, np.array([[H, W, 1.0, 1.0]] * 2), use_dilation=True, loose=0.35)


def test_db_thin_diagonal_boxes():
    """Two-pixel-wide 45-degree staircases: the mini-box is an exactly right-angled thin rectangle with an unclip
    distance < 1, where Clipper's `cosA > 0` test sees products that cancel exactly. A fused multiply-add leaves a
    rounding residual there and drops the round joins (found by tests/stress_gpu.py)."""
    H, W = 96, 128
    m = np.full((H, W), 0.05, np.float32)
    shape = [(0, 0), (1, 0), (1, 1), (1, 2), (2, 2), (3, 2), (3, 3)]      # the blob of the failing stress case
    k = 0
    for flip_x in (0, 1):
        for flip_y in (0, 1):
            for transpose in (0, 1):
                x0, y0 = 10 + 28 * (k % 4), 12 + 40 * (k // 4)
                for (dx, dy) in shape:
                    dx, dy = (3 - dx if flip_x else dx), (3 - dy if flip_y else dy)
                    if transpose:
                        dx, dy = dy, dx
                    m[y0 + dy, x0 + dx] = 0.9
                k += 1
    want, counts = _check(m[None, None], np.array([[H, W, 1.0, 1.0]]), loose=0.0)
    assert counts[0] >= 8


def test_db_host_batch_is_uploaded_in_chunks():
    """numpy / CPU-tensor maps of a large batch take the chunked upload path (H2D of chunk i+1 overlaps the kernels of
    chunk i): same results as the device-tensor path, also with chunks of unequal size."""
    import torch
    H, W = 96, 160
    maps = synth.db_batch(37, seed=5, H=H, W=W)
    sl = np.array([[H + n, W + 2 * n, 1.0, 1.0] for n in range(37)], np.float64)
    op = _op()
    op.upload_chunk = 8                                    # 37 images -> 5 chunks of 7/8 images
    want = op({"maps": torch.from_numpy(maps).cuda()}, sl)
    for host in (maps, torch.from_numpy(maps).pin_memory()):
        got = op({"maps": host}, sl)
        assert len(got) == len(want)
        for g, w in zip(got, want):
            assert np.array_equal(g["points"], w["points"]) and np.array_equal(g["box_scores"], w["box_scores"])


@pytest.mark.parametrize("dtype", ["float32", "float16"])
def test_db_scan_variants_agree_over_widths(dtype, db_path):
    """The bulk-copy fed map scan (db_scan4_kernel: lane-contiguous chunks of 2..16 sixteen-byte cells, rows that do
    not fill the last lanes, the widest row a 64-pixel chunk allows and the first one past it) against the
    warp-per-row db_scan_kernel on the same maps: identical boxes, scores and counts, bit for bit. The maps hold the
    values the fixed-point trick treats specially (0, -0, 1, the threshold itself, its two neighbours)."""
    import torch
    from pytorchocr_b200 import _lib
    if db_path != "image_smem":
        pytest.skip("compares the two scans itself")
    L = _lib.lib()
    H = 40
    rng = np.random.default_rng(11)
    th = np.float32(0.3)
    special = np.array([0.0, -0.0, 1.0, th, np.nextafter(th, np.float32(1)), np.nextafter(th, np.float32(0))], np.float32)
    op = _op()
    for W in (8, 40, 120, 128, 136, 640, 1288, 2040, 2048, 2056):
        base = cv2.GaussianBlur(rng.random((H, W)).astype(np.float32), (0, 0), 2.0)
        base = (base - base.min()) / (base.max() - base.min())
        pick = rng.random((H, W)) < 0.05
        base[pick] = rng.choice(special, size=int(pick.sum()))
        maps = torch.from_numpy(base[None, None]).cuda()
        if dtype == "float16":
            maps = maps.half()
        sl = np.array([[H, W, 1.0, 1.0]])
        out = []
        for scan in (0, 1):
            _lib.check(L.ocrpp_set_tuning(_lib.TUNE_DB_SCAN, scan))
            boxes, scores, counts, status, _ = op.run_device(maps, sl)
            out.append((boxes.copy(), scores.copy(), counts.copy(), status.copy()))
        _lib.check(L.ocrpp_set_tuning(_lib.TUNE_DB_SCAN, 0))
        for a, b in zip(*out):
            assert np.array_equal(a, b), "scan variants differ at W=%d" % W
        assert not (int(out[0][3][0]) & _lib.IMG_VALUE_OUT_OF_RANGE)
    # a value above 1 and a negative value are reported by both scans; a negative value that rounds away in `f + 1.0f` is not
    for bad, ok in ((np.float32(1.001), False), (np.float32(-1e-3), False), (np.float32(-1e-12), True)):
        m = np.full((1, 1, H, 128), 0.6, np.float32)
        m[0, 0, 7, 77] = bad
        for scan in (0, 1):
            _lib.check(L.ocrpp_set_tuning(_lib.TUNE_DB_SCAN, scan))
            if ok:
                op({"maps": torch.from_numpy(m).cuda()}, np.array([[H, 128, 1.0, 1.0]]))
            else:
                with pytest.raises(_lib.OcrppError):
                    op({"maps": torch.from_numpy(m).cuda()}, np.array([[H, 128, 1.0, 1.0]]))
        _lib.check(L.ocrpp_set_tuning(_lib.TUNE_DB_SCAN, 0))


@pytest.mark.parametrize("ratio", [3.0, 5.0, 8.0])
def test_db_large_unclip_ratio_defers_to_the_generic_path(ratio):
    """Unclip polygons with more points than the per-candidate buffer of db_geometry_kernel (tall regions offset by a
    large distance: Clipper's round joins grow with the distance) are handed to the generic routine AFTER their hull was
    built - from that hull, not from row extents that only ever lived in the stage-2 kernel's shared memory."""
    H, W = 256, 512
    m = np.full((2, 1, H, W), 0.05, np.float32)
    rng = np.random.default_rng(3)
    for n in range(2):
        for i in range(4):
            x0, y0 = 40 + 110 * i, 40 + 30 * n
            h, w = 44 + 5 * i, 70 + 6 * i
            m[n, 0, y0:y0 + h, x0:x0 + w] = 0.9
            m[n, 0, y0 + 3:y0 + 6, x0 + w:x0 + w + 9] = 0.85          # a bump: not a plain rectangle
        m[n, 0, 180:190, 30:480] = 0.8                                  # a long thin line
    m += rng.random(m.shape).astype(np.float32) * 0.04
    _check(m, np.array([[H, W, 1.0, 1.0]] * 2), unclip_ratio=ratio, loose=0.3)


def test_db_call_is_cuda_graph_capturable(db_path):
    """One ocrpp_db_postprocess call (memset, fork / join onto the library's auxiliary streams, all kernels) can be
    captured into a CUDA graph and replayed - what bench.py does for the batch-256 step, and what a serving loop
    would do. Replays on new map contents give the same outputs as direct calls."""
    import torch
    from pytorchocr_b200 import _lib
    L = _lib.lib()
    N, H, W = 64, 96, 160            # >= 64 images: the call forks into two sub-batch pipelines
    op = _op()
    sl = np.array([[H, W, 1.0, 1.0]] * N)
    a = torch.from_numpy(synth.db_batch(N, seed=71, H=H, W=W)).cuda()
    b = torch.from_numpy(synth.db_batch(N, seed=72, H=H, W=W)).cuda()
    maps = a.clone()
    op.run_device(maps, sl)                      # allocates and caches the buffers for this shape
    buf = next(iter(op._cache.values())); key = next(iter(op._cache))
    o_box, o_sc, o_cnt, o_st = buf["offs"]
    base = buf["out_dev"].data_ptr()

    def call(stream):
        _lib.check(L.ocrpp_db_postprocess(maps.data_ptr(), _lib.F32, N, H, W, maps.stride(0), maps.stride(2),
                                          buf["wh_dev"].data_ptr(), CFG["thresh"], CFG["box_thresh"], CFG["unclip_ratio"],
                                          key[5], key[4], 0, 0, base + o_box, base + o_sc, base + o_cnt, base + o_st,
                                          None, None, buf["ws"].data_ptr(), buf["ws_bytes"], stream.cuda_stream))

    def direct(src):
        maps.copy_(src)
        call(torch.cuda.current_stream())
        torch.cuda.synchronize()
        return buf["out_dev"].cpu().numpy().copy()

    want_a, want_b = direct(a), direct(b)
    assert not np.array_equal(want_a, want_b)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, capture_error_mode="thread_local"):
        call(torch.cuda.current_stream())
    for src, want in ((a, want_a), (b, want_b), (a, want_a)):
        maps.copy_(src)
        buf["out_dev"].zero_()
        g.replay()
        torch.cuda.synchronize()
        got = buf["out_dev"].cpu().numpy()
        nb = o_sc   # boxes region; scores / counts / status follow
        cnt = got[o_cnt:o_cnt + 4 * N].view(np.int32)
        assert np.array_equal(cnt, want[o_cnt:o_cnt + 4 * N].view(np.int32))
        cap = key[5]
        gb, wb = got[:nb].view(np.int16).reshape(N, cap, 8), want[:nb].view(np.int16).reshape(N, cap, 8)
        for n in range(N):
            assert np.array_equal(gb[n, :cnt[n]], wb[n, :cnt[n]])
