"""GPU parity: text-line crops (ocrpp_crop_boxes through PartImageCropper) vs the oracle, i.e. the reference's
own cv2 calls (oracle/crop_oracle.py). uint8 pixels: bit-exact; box order: identical."""
import os

import numpy as np
import pytest

from conftest import ROOT
from oracle import crop_oracle as co
from pytorchocr_b200 import synth

pytestmark = pytest.mark.gpu


def _cropper(**kw):
    from pytorchocr_b200.part_img import PartImageCropper
    return PartImageCropper(**kw)


def _check_page(img, boxes, got_boxes, got_crops, sort=True, rotate=True):
    want_boxes = co.sort_boxes(boxes) if sort else list(boxes)
    assert len(got_boxes) == len(want_boxes)
    for gb, wb, gc in zip(got_boxes, want_boxes, got_crops):
        assert np.array_equal(gb, wb)
        want = co.crop_for_rec(img, wb) if rotate else co.get_part_img(img, wb)
        assert gc.shape == want.shape, (gc.shape, want.shape)
        assert np.array_equal(gc, want), (wb.tolist(), np.abs(gc.astype(int) - want.astype(int)).max())


def test_crops_match_reference_golden():
    G = np.load(os.path.join(ROOT, "tests", "golden", "reference_crops.npz"))
    boxes, crops = _cropper()(G["img"], G["boxes"])
    assert np.array_equal(np.asarray(boxes, np.int16), G["sorted_boxes"])
    o = 0
    for c, (rows, cols) in zip(crops, G["dims"]):
        assert c.shape == (rows, cols, 3)
        assert np.array_equal(c.reshape(-1), G["pixels"][o:o + rows * cols * 3])
        o += rows * cols * 3


@pytest.mark.parametrize("H,W,n", [(736, 1280, 200), (300, 400, 120), (97, 131, 9)])
def test_crop_page(H, W, n):
    img = synth.page_image(H, H, W)
    boxes = synth.page_boxes(W, n=n, H=H, W=W, tall_frac=0.25, skew=3.0 if H > 100 else 1.0, scale=1.0 if H > 100 else 0.4)
    gb, gc = _cropper()(img, boxes)
    _check_page(img, boxes, gb, gc)


@pytest.mark.parametrize("sort,rotate", [(False, False), (True, False), (False, True)])
def test_crop_flags(sort, rotate):
    img = synth.page_image(4, 240, 320)
    boxes = synth.page_boxes(8, n=60, H=240, W=320, tall_frac=0.4)
    gb, gc = _cropper(sort=sort, rotate_tall=rotate)(img, boxes)
    _check_page(img, boxes, gb, gc, sort=sort, rotate=rotate)


def test_crop_batch_of_pages_ragged_counts():
    import torch
    H, W = 200, 288
    imgs = np.stack([synth.page_image(20 + i, H, W) for i in range(5)])
    lists = [synth.page_boxes(40 + i, n=k, H=H, W=W) for i, k in enumerate([30, 0, 1, 77, 12])]
    lists[1] = np.zeros((0,), np.int16)                     # the operators' "no boxes" result
    gb, gc = _cropper()(torch.from_numpy(imgs).cuda(), lists)
    for i in range(5):
        _check_page(imgs[i], lists[i].reshape(-1, 4, 2), gb[i], gc[i])


def test_crop_gray_and_four_channels():
    for C in (1, 4):
        img = synth.page_image(9, 160, 200, C=C)
        boxes = synth.page_boxes(10, n=25, H=160, W=200)
        gb, gc = _cropper()(img, boxes)
        for b, c in zip(gb, gc):
            want = co.crop_for_rec(img, b)
            assert np.array_equal(c, want.reshape(c.shape))


def test_crop_boxes_on_the_page_border_and_large_box():
    H, W = 180, 260
    img = synth.page_image(6, H, W)
    boxes = np.array([[[200, 150], [260, 155], [258, 180], [198, 176]],      # right == W, bottom == H
                      [[0, 0], [50, 3], [48, 20], [1, 18]],                    # top-left corner
                      [[2, 3], [255, 8], [250, 175], [6, 170]],                # nearly the whole page (many work items)
                      [[10, 10], [30, 10], [30, 90], [10, 90]]], np.int16)    # tall, axis aligned -> rotated
    gb, gc = _cropper()(img, boxes)
    _check_page(img, boxes, gb, gc)


def test_crop_degenerate_boxes_are_reported():
    img = synth.page_image(6, 100, 100)
    for bad in ([[5, 5], [40, 5], [40, 5], [5, 5]],            # zero height: cv2 raises in the reference
                [[0, 0], [10, 10], [20, 20], [30, 30]]):       # collinear
        with pytest.raises(ValueError):
            _cropper()(img, np.array([bad], np.int16))


def test_crop_arena_retry_and_device_chain_from_db():
    """run_device consumes DBPostProcess.run_device's device buffers (boxes/counts) directly; a too small arena
    is grown from the size the first pass reports."""
    import torch
    from pytorchocr_b200.postprocess import build_post_process
    H, W = 256, 384
    maps = synth.db_batch(3, seed=77, H=H, W=W)
    sl = np.array([[H, W, 1.0, 1.0]] * 3)
    op = build_post_process(dict(name="DBPostProcess", thresh=0.3, box_thresh=0.5, max_candidates=1000, unclip_ratio=1.7,
                                 cuda_speedup=True), {"use_gpu": True})
    dev_maps = torch.from_numpy(maps).cuda()
    res = op({"maps": dev_maps}, sl)
    imgs = torch.from_numpy(np.stack([synth.page_image(i, H, W) for i in range(3)])).cuda()
    cap = max(len(r["points"]) for r in res)
    hb = np.zeros((3, cap, 4, 2), np.int16)
    for n, r in enumerate(res):
        hb[n, :len(r["points"])] = r["points"]
    cnt = torch.tensor([len(r["points"]) for r in res], dtype=torch.int32).cuda()
    cr = _cropper()
    arena, offsets, dims, order, status = cr.run_device(imgs, torch.from_numpy(hb).cuda(), cnt, capacity=1000)
    assert offsets[-1] > 1000 and not (status & 16).any()      # grown after the truncated first pass
    host = arena[:offsets[-1]].cpu().numpy()
    imgs_h = imgs.cpu().numpy()
    for n, r in enumerate(res):
        want_order = co.sort_order(np.asarray(r["points"]))
        assert np.array_equal(order[n, :len(want_order)], want_order)
        for k, b in enumerate(want_order):
            e = n * cap + k
            want = co.crop_for_rec(imgs_h[n], r["points"][b])
            got = host[offsets[e]:offsets[e + 1]].reshape(dims[e][0], dims[e][1], 3)
            assert np.array_equal(got, want)
