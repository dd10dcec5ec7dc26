"""Detection -> crops -> recognition with everything after the detector's forward on the device, batched.

Replaces the per-box host loop of R/deploy/pytorch/run_ocr.py:167-231 (`OCRer.run`):

    reference, per page                                   here, per batch of pages
    ----------------------------------------------------  ------------------------------------------------------------
    det_preds -> .cpu().numpy() -> DBPostProcess           maps stay in HBM -> ocrpp_db_postprocess (boxes stay there too)
    sort_boxes, then per box:                              ocrpp_crop_boxes: sort_boxes + get_part_img + rot90 rule for all
      get_part_img (cv2 warp), rot90 rule                    boxes of all pages into ONE device arena
      cvtColor, RecResizeImg, .to(device)                  ocrpp_rec_preprocess: ONE float32 batch [K, C, 32, W]
      recogniser forward with batch 1                      recogniser forward in chunks of `rec_batch` crops
      CTCLabelDecode on one line                           ONE ocrpp_ctc_greedy over all lines

Same results: the crops and the recogniser input are reproduced bit for bit (cv2's 8-bit arithmetic, see
csrc/crop.cu, csrc/prep.cuh), so a recogniser that does not mix batch entries returns the same probabilities, and
the decoded strings / scores are the reference's. The models themselves (detector, recogniser, their input
transforms) are the caller's torch modules - this package only covers the post-processing path. The optional direction
classifier of the reference loop (run_ocr.py:192-211) is not batched here: pass `clser=None` configurations only.
"""
import numpy as np

from .. import _lib
from ..part_img import PartImageCropper


class OCRer(object):
    def __init__(self, deter, det_post_process, recer, rec_post_process, rec_image_shape=(1, 32, 320),
                 rec_img_mode="GRAY", rec_batch=256, use_padding_resize=False):
        """deter: callable(det_input CUDA tensor) -> dict with "maps" (what the reference's detector returns);
        det_post_process: this package's DBPostProcess / PSEPostProcess / PANPostProcess; recer: callable(float32
        CUDA [B, C, H, W]) -> predictions CTCLabelDecode accepts; rec_post_process: this package's CTCLabelDecode."""
        if rec_img_mode not in _lib.IMG_MODE:
            raise ValueError("rec_img_mode must be GRAY, RGB or BGR")
        self.deter, self.det_post = deter, det_post_process
        self.recer, self.rec_post = recer, rec_post_process
        self.rec_shape = tuple(int(v) for v in rec_image_shape)
        self.rec_img_mode = rec_img_mode
        self.rec_batch = int(rec_batch)
        self.use_padding_resize = bool(use_padding_resize)
        self.cropper = PartImageCropper(sort=True, rotate_tall=True)

    def run_batch(self, imgs, det_input, shape_list):
        """imgs: uint8 BGR pages [N,H,W,3] (numpy or CUDA tensor; what cv2.imdecode returns, stacked); det_input: the
        detector's input batch (CUDA tensor) for the same pages; shape_list: [N,4] (src_h, src_w, ratio_h, ratio_w).
        Returns, per page, the reference's `ocr_res`: a list of [box int16 [4,2], text, round(prob, 2)] in
        sort_boxes order."""
        torch = _lib.require_cuda()
        L = _lib.lib()
        t = torch.from_numpy(np.ascontiguousarray(imgs)) if isinstance(imgs, np.ndarray) else imgs
        if t.dim() == 3:
            t = t.unsqueeze(0)
        if t.dtype != torch.uint8 or t.shape[-1] != 3:
            raise TypeError("pages must be uint8 BGR images [N,H,W,3]")
        t = t.cuda(non_blocking=True) if not t.is_cuda else t
        t = t.contiguous()
        N = t.shape[0]
        with torch.no_grad():
            det_preds = self.deter(det_input)
        # ---- detection post-processing: boxes and counts stay on the device ----
        kw = {"use_padding_resize": True} if self.use_padding_resize else {}
        boxes, _, counts, _, ex = self.det_post.run_device(det_preds["maps"], shape_list, **kw)
        cap = boxes.shape[1]
        total = int(counts.sum())
        if total == 0:
            return [[] for _ in range(N)]
        # ---- all crops of all pages into one arena (sort_boxes order, rot90 rule) ----
        need = 0
        for n in range(N):
            bb = boxes[n, :counts[n]].astype(np.int64)
            if len(bb):
                need += int(((bb[:, :, 0].max(1) - bb[:, :, 0].min(1)) * (bb[:, :, 1].max(1) - bb[:, :, 1].min(1))).sum())
        arena, offsets, dims, order, status = self.cropper.run_device(t, ex["boxes_dev"], ex["counts_dev"],
                                                                      capacity=max(1, need * 3))
        if (status & 8).any():   # OCRPP_IMG_BOX_DEGENERATE: the reference dies inside cv2 on such a box (utility.py:62,72)
            raise ValueError("a detected box has an empty bounding rectangle or lies outside the page")
        meta = self.cropper.last_device_meta
        entries = np.concatenate([n * cap + np.arange(counts[n]) for n in range(N)]).astype(np.int64)
        idx = torch.from_numpy(entries).to(t.device)
        offs_dev = meta["offsets"][:-1][idx].contiguous()
        dims_dev = meta["dims"].view(-1, 2)[idx].contiguous()
        # ---- one recogniser input batch ----
        C, Hh, Ww = self.rec_shape
        K = len(entries)
        batch = torch.empty((K, C, Hh, Ww), dtype=torch.float32, device=t.device)
        stream = torch.cuda.current_stream(t.device)
        _lib.check(L.ocrpp_rec_preprocess(arena.data_ptr(), offs_dev.data_ptr(), dims_dev.data_ptr(), K, 3,
                                          _lib.IMG_MODE[self.rec_img_mode], Hh, Ww, batch.data_ptr(), stream.cuda_stream))
        # ---- recogniser forward (chunked only to bound activation memory) and ONE decode ----
        outs = []
        with torch.no_grad():
            for lo in range(0, K, self.rec_batch):
                outs.append(self.recer(batch[lo:lo + self.rec_batch]))
        if isinstance(outs[0], tuple):
            outs = [o[-1] for o in outs]
        # CTCLabelDecode takes [T, B, C] tensors (rec_postprocess.py:77-83 permutes them to [B, T, C])
        preds = outs[0] if len(outs) == 1 else torch.cat(outs, dim=1)
        decoded = self.rec_post(preds)
        # ---- the reference's result structure ----
        res, k = [], 0
        for n in range(N):
            page = []
            for r in range(int(counts[n])):
                text, prob = decoded[k]
                page.append([boxes[n, order[n, r]].copy(), text, round(prob, 2)])
                k += 1
            res.append(page)
        return res

    def run(self, img, det_input, shape):
        """One page (the reference's `OCRer.run` contract)."""
        return self.run_batch(img[None] if img.ndim == 3 else img, det_input, np.asarray(shape).reshape(1, -1))[0]
