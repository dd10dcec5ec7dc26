"""Text-line crops of detected boxes, computed by libocrpp (sm_100a).

Mirrors what R/deploy/pytorch/run_ocr.py:181-191 does between the detector and the recogniser:
`sort_boxes` (R/pytocr/utils/utility.py:32-50), `get_part_img` per box (:53-78) and the rot90 rule for tall
crops (run_ocr.py:190-191) - for all boxes of a batch of pages in one call. The pixels are cv2's, bit for
bit (see pytorchocr_b200/csrc/crop.cu).

    cropper = PartImageCropper()
    boxes, crops = cropper(img, det_post_result[0]["points"])     # img: uint8 [H,W,3] numpy / CUDA tensor
    # boxes: the sort_boxes order; crops[i] == rot90-rule(get_part_img(img, boxes[i]))

`run_device` is the batch-level entry: it takes the device buffers the detection operators produce
(`DBPostProcess.run_device` boxes/counts) and leaves the crops in a device arena for the recogniser's
preprocessing, without a host round trip.
"""
import numpy as np

from . import _lib

IMG_BOX_DEGENERATE = 8
IMG_CROPS_TRUNCATED = 16


def get_part_imgs(img, boxes, sort=False, rotate_tall=False):
    """[get_part_img(img, b) for b in boxes] (utility.py:53-78) on the GPU; convenience wrapper."""
    return PartImageCropper(sort=sort, rotate_tall=rotate_tall)(img, boxes)[1]


class PartImageCropper(object):
    def __init__(self, sort=True, rotate_tall=True, **kwargs):
        self.sort = bool(sort)
        self.rotate_tall = bool(rotate_tall)
        self._bufs = {}

    # -- batch-level device entry -------------------------------------------------------------------
    def run_device(self, imgs, boxes, counts=None, capacity=None):
        """imgs: CUDA uint8 [N,H,W,C] (rows may be strided); boxes: CUDA int16 [N,cap,4,2]; counts: CUDA int32
        [N] or None. Returns (arena uint8 CUDA, offsets int64 [N*cap+1], dims int32 [N*cap,2], order int32
        [N,cap], status int32 [N]) - the last four as numpy arrays copied from pinned memory."""
        import torch
        L = _lib.lib()
        assert imgs.is_cuda and imgs.dtype == torch.uint8 and imgs.dim() == 4
        assert boxes.is_cuda and boxes.dtype == torch.int16 and boxes.dim() == 4 and boxes.is_contiguous()
        N, H, W, C = imgs.shape
        cap = boxes.shape[1]
        assert boxes.shape[0] == N and imgs.stride(3) == 1 and imgs.stride(2) == C
        if counts is not None:
            assert counts.is_cuda and counts.dtype == torch.int32 and counts.numel() == N and counts.is_contiguous()
        dev = imgs.device
        K = N * cap
        # ONE grow-only set of buffers per device (workspace, device meta block, pinned meta block, crop arena): a page
        # loop calls this with a different (N, cap) for nearly every page, so buffers keyed by the shape would pile up
        ws_bytes = L.ocrpp_crop_workspace_bytes(N, cap)
        meta_elems = 2 * (K + 1) + 2 * K + K + N          # int64 offsets (as 2 x int32), dims, order, status
        buf = self._bufs.get(dev)
        if buf is None:
            buf = self._bufs[dev] = dict(ws=None, meta=None, host=None, arena=None)
        if buf["ws"] is None or buf["ws"].numel() < ws_bytes:
            buf["ws"] = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        if buf["meta"] is None or buf["meta"].numel() < meta_elems:
            buf["meta"] = torch.empty(meta_elems, dtype=torch.int32, device=dev)
            buf["host"] = torch.empty(meta_elems, dtype=torch.int32, pin_memory=True)
        if capacity is None:
            capacity = N * H * W * C if buf["arena"] is None else buf["arena"].numel()
        stream = torch.cuda.current_stream(dev)
        while True:
            if buf["arena"] is None or buf["arena"].numel() < capacity:
                buf["arena"] = torch.empty(int(capacity), dtype=torch.uint8, device=dev)
            arena, meta = buf["arena"], buf["meta"]
            base = meta.data_ptr()
            o_dims = base + 8 * (K + 1)
            o_order = o_dims + 8 * K
            o_status = o_order + 4 * K
            _lib.check(L.ocrpp_crop_boxes(imgs.data_ptr(), N, H, W, C, imgs.stride(0), imgs.stride(1),
                                          boxes.data_ptr(), counts.data_ptr() if counts is not None else None,
                                          cap, 1 if self.sort else 0, 1 if self.rotate_tall else 0,
                                          arena.data_ptr(), arena.numel(), base, o_dims, o_order, o_status,
                                          buf["ws"].data_ptr(), buf["ws"].numel(), stream.cuda_stream))
            buf["host"][:meta_elems].copy_(meta[:meta_elems], non_blocking=True)
            stream.synchronize()
            h = buf["host"].numpy()[:meta_elems]
            # device views of the same block for callers that stay on the device (deploy/run_ocr.py); valid until the next call
            self.last_device_meta = dict(offsets=meta[:2 * (K + 1)].view(torch.int64), dims=meta[2 * (K + 1):2 * (K + 1) + 2 * K])
            offsets = h[:2 * (K + 1)].view(np.int64).copy()
            dims = h[2 * (K + 1):2 * (K + 1) + 2 * K].reshape(K, 2).copy()
            order = h[2 * (K + 1) + 2 * K:2 * (K + 1) + 3 * K].reshape(N, cap).copy()
            status = h[2 * (K + 1) + 3 * K:].copy()
            if (status & IMG_CROPS_TRUNCATED).any():
                capacity = int(offsets[-1])          # exact size, known after the first pass
                continue
            return arena, offsets, dims, order, status

    # -- reference-shaped call -----------------------------------------------------------------------
    def __call__(self, img, boxes):
        """img: uint8 [H,W,C] (numpy or CUDA tensor) or a batch [N,H,W,C]; boxes: int [K,4,2] (or a list of N
        such arrays). Returns (sorted boxes, crops) - lists per page when a batch was given."""
        import torch
        batched = isinstance(boxes, (list, tuple))
        boxes_list = list(boxes) if batched else [boxes]
        t = torch.from_numpy(np.ascontiguousarray(img)) if isinstance(img, np.ndarray) else img
        if t.dim() == 3:
            t = t.unsqueeze(0)
        if t.dtype != torch.uint8:
            raise TypeError("PartImageCropper works on uint8 images (got %s)" % (t.dtype,))
        t = t.cuda(non_blocking=True) if not t.is_cuda else t
        if not t.is_contiguous():
            t = t.contiguous()
        N = t.shape[0]
        if len(boxes_list) != N:
            raise ValueError("%d box lists for %d pages" % (len(boxes_list), N))
        cap = max(1, max(len(b) for b in boxes_list))
        hb = np.zeros((N, cap, 4, 2), np.int16)
        hc = np.zeros(N, np.int32)
        for n, b in enumerate(boxes_list):
            b = np.asarray(b)
            if b.size:
                hb[n, :len(b)] = b.reshape(-1, 4, 2)
                hc[n] = len(b)
        need = 0
        for n in range(N):
            bb = hb[n, :hc[n]].astype(np.int64)
            if len(bb):
                need += int(((bb[:, :, 0].max(1) - bb[:, :, 0].min(1)) * (bb[:, :, 1].max(1) - bb[:, :, 1].min(1))).sum())
        arena, offsets, dims, order, status = self.run_device(
            t, torch.from_numpy(hb).to(t.device), torch.from_numpy(hc).to(t.device), capacity=max(1, need * t.shape[3]))
        if (status & IMG_BOX_DEGENERATE).any():
            # the reference fails inside cv2 for an empty crop (utility.py:62,72) - same contract, clearer message
            raise ValueError("a box has an empty bounding rectangle or lies outside the page")
        host = arena[:int(offsets[-1])].cpu().numpy()
        C = t.shape[3]
        out_boxes, out_crops = [], []
        for n in range(N):
            k = int(hc[n])
            sel = order[n, :k]
            out_boxes.append([np.asarray(boxes_list[n]).reshape(-1, 4, 2)[i] for i in sel])
            crops = []
            for r in range(k):
                e = n * cap + r
                rows, cols = dims[e]
                crops.append(host[offsets[e]:offsets[e] + rows * cols * C].reshape(rows, cols, C))
            out_crops.append(crops)
        if batched:
            return out_boxes, out_crops
        return out_boxes[0], out_crops[0]
