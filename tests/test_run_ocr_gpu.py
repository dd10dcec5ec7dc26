"""deploy/run_ocr.py: detector output -> boxes -> crops -> recogniser input -> ONE decode, all on
the device, against the reference's per-box host loop (R/deploy/pytorch/run_ocr.py:181-229) restated with the
reference's own sequence of cv2 calls (oracle/crop_oracle.py, oracle/rec_prep_oracle.py) and the CTC oracle. The
detector and the recogniser are deterministic stubs (the models are not part of the post-processing path); the
recogniser stub treats every batch entry independently, so batch 1 (the loop) and batch K (the chain) agree."""
import numpy as np
import pytest

from oracle import crop_oracle
from oracle.ctc_oracle import CTCLabelDecodeOracle
from oracle.rec_prep_oracle import rec_preprocess
from pytorchocr_b200 import synth

pytestmark = pytest.mark.gpu

NCLS = 97


def _recer(x):
    """[B,C,32,W] float32 -> softmax probabilities [T = W/4, B, NCLS]: a peaked distribution around a class derived
    from the mean of every 32x4 column block; element-wise in B."""
    import torch
    v = torch.nn.functional.avg_pool2d(x.mean(1, keepdim=True), (32, 4))[:, 0, 0]        # [B, T]
    cls = (v + 1.0) * 30.0
    c = torch.arange(NCLS, device=x.device, dtype=torch.float32)
    logits = -((cls[..., None] - c) ** 2) * 3.0
    return torch.softmax(logits, -1).permute(1, 0, 2).contiguous()


@pytest.mark.parametrize("mode,shape", [("GRAY", (1, 32, 320)), ("RGB", (3, 32, 320)), ("BGR", (3, 32, 160))])
def test_chain_equals_reference_loop(tmp_path, mode, shape):
    import torch
    from pytorchocr_b200.deploy.run_ocr import OCRer
    from pytorchocr_b200.postprocess import build_post_process
    H, W, N = 192, 320, 3
    pages = np.stack([synth.page_image(60 + i, H, W) for i in range(N)])
    maps = synth.db_batch(N, seed=70, H=H, W=W)
    maps[2] = 0.0                                            # a page without text
    maps_dev = torch.from_numpy(maps).cuda()
    d = synth.write_char_dict(str(tmp_path / "dict.txt"), NCLS - 1)
    det_post = build_post_process({"name": "DBPostProcess", "thresh": 0.3, "box_thresh": 0.5, "unclip_ratio": 1.7,
                                   "cpp_speedup": True, "cuda_speedup": True})
    rec_post = build_post_process({"name": "CTCLabelDecode", "character_dict_path": d, "use_space_char": False,
                                   "cuda_speedup": True})
    sl = np.array([[H, W, 1.0, 1.0]] * N, np.float64)
    ocr = OCRer(lambda x: {"maps": x}, det_post, _recer, rec_post, rec_image_shape=shape, rec_img_mode=mode, rec_batch=7)
    got = ocr.run_batch(pages, maps_dev, sl)

    # the reference loop, page by page and box by box, from the same detection result
    det = det_post({"maps": maps_dev}, sl)
    oracle = CTCLabelDecodeOracle(d)
    assert len(got) == N and got[2] == []
    n_boxes = 0
    for n in range(N):
        pts = det[n]["points"]
        boxes = crop_oracle.sort_boxes(pts) if len(pts) else []
        assert len(got[n]) == len(boxes)
        for (gbox, gtext, gprob), box in zip(got[n], boxes):
            part = crop_oracle.get_part_img(pages[n], box)
            if part.shape[0] >= 1.5 * part.shape[1]:
                part = np.rot90(part, 1)
            x = torch.from_numpy(rec_preprocess(np.ascontiguousarray(part), mode, shape)).unsqueeze(0).cuda()
            text, prob = oracle(_recer(x))[0]
            assert np.array_equal(gbox, box)
            assert gtext == text, (n, gtext, text)
            assert gprob == round(prob, 2) or abs(gprob - prob) < 6e-3
            n_boxes += 1
    assert n_boxes > 15


def test_single_page_contract(tmp_path):
    """`run` = the reference's OCRer.run result structure for one page."""
    import torch
    from pytorchocr_b200.deploy.run_ocr import OCRer
    from pytorchocr_b200.postprocess import build_post_process
    H, W = 160, 256
    page = synth.page_image(5, H, W)
    maps_dev = torch.from_numpy(synth.db_batch(1, seed=9, H=H, W=W)).cuda()
    d = synth.write_char_dict(str(tmp_path / "dict.txt"), NCLS - 1)
    ocr = OCRer(lambda x: {"maps": x},
                build_post_process({"name": "DBPostProcess", "cpp_speedup": True, "unclip_ratio": 1.7, "cuda_speedup": True}),
                _recer, build_post_process({"name": "CTCLabelDecode", "character_dict_path": d, "cuda_speedup": True}))
    res = ocr.run(page, maps_dev, [H, W, 1.0, 1.0])
    assert len(res) > 3
    for box, text, prob in res:
        assert box.shape == (4, 2) and box.dtype == np.int16 and isinstance(text, str) and isinstance(prob, float)
