"""DB / DB++ post-processing behind the reference's operator API, computed by libocrpp (sm_100a).

Mirrors R/pytocr/postprocess/db_postprocess.py (DBPostProcess :10-74, DistillationDBPostProcess
:197-226): same ctor kwargs, same `__call__(outs_dict, shape_list, use_padding_resize=False)`, same
return structure (list of {"points": int16 [K,4,2], "scores": [...]}; K == 0 gives points.shape == (0,)).
Both of the reference's branches are reproduced, selected by the reference's own `cpp_speedup` flag:
  * `cpp_speedup: True` (the shipped configs): the semantics of db_postprocess_fast/src/db_postprocess.cpp:231-317 -
    short side = max(w,h), BoxScore over the 4-connected polygon fill, float32 unclip distance, roundf,
    max_candidates fixed at 1000, score_mode ignored, "scores" = [1.0]*K (db_postprocess.py:64-67);
  * `cpp_speedup: False` (the class default): the semantics of DBPostProcess.boxes_from_bitmap (:76-194) - short
    side = min(w,h), BoxScore over the LINE_8 fill of the contour (`score_mode: poly`) or of its mini box
    (`score_mode: box`), float64 unclip distance, np.round, the first `max_candidates` contours in cv2's order, real
    scores.
The BoxScore of every box is additionally returned as "box_scores" (float32 [K]).
`out_polygon: True` has no behaviour to reproduce: the reference ends in np.array(ragged polygons, dtype=int16)
(:142) and raises on every multi-region page (recorded from the unmodified reference in
tests/golden/reference_db_python.npz, tests/test_oracle_db_python.py).

The probability map never leaves the device: it is consumed where the head wrote it
(zero-copy through data_ptr()); only boxes / counts come back, through pinned buffers.
"""
import numpy as np

from .. import _lib


class DBPostProcess(object):
    def __init__(self, thresh=0.3, box_thresh=0.5, max_candidates=1000, unclip_ratio=1.5,
                 use_dilation=False, score_mode="poly", cpp_speedup=False, out_polygon=False,
                 cuda_speedup=True, max_runs=None, out_capacity=384, **kwargs):
        if not cuda_speedup:
            raise _lib.OcrppError("pytorchocr_b200 implements only the CUDA path: set PostProcess.cuda_speedup: True "
                                  "(with the flag off the reference's own DBPostProcess runs)")
        assert score_mode in ["box", "poly"], "Score mode must be in [box, poly] but got: {}".format(score_mode)
        self.use_dilation = bool(use_dilation)   # db_postprocess.py:22: dilation_kernel = [[1,1],[1,1]]
        if out_polygon:
            raise NotImplementedError("out_polygon: the reference's own branch raises ValueError on ragged polygons "
                                      "(db_postprocess.py:142); there is no behaviour to reproduce")
        self.cpp_speedup = bool(cpp_speedup)
        if not self.cpp_speedup and int(max_candidates) > 1000:
            raise ValueError("max_candidates > 1000 is not supported (workspace capacity of the library)")
        self.thresh = thresh
        self.box_thresh = box_thresh
        # db_postprocess.cpp:239 ignores the kwarg (const int max_candidates = 1000); the Python branch honours it
        self.max_candidates = 1000 if self.cpp_speedup else max(0, int(max_candidates))
        self.unclip_ratio = unclip_ratio
        self.min_size = 3
        self.score_mode = score_mode
        self.max_runs = max_runs
        # Output capacity per image of the device->host result block. The reference keeps up to
        # max_candidates (1000) contours; real pages have far fewer, so the block is sized for
        # `out_capacity` candidates and the call is repeated with the full capacity when the library
        # reports OCRPP_IMG_CANDIDATES_TRUNCATED below it (same results, 3x less D2H traffic).
        self.out_capacity = max(1, min(max(1, self.max_candidates), int(out_capacity or self.max_candidates or 1)))
        self._cache = {}

    # -- device plumbing ------------------------------------------------------------------------
    @staticmethod
    def _to_device(pred):
        torch = _lib.require_cuda()
        if isinstance(pred, np.ndarray):
            t = torch.from_numpy(np.ascontiguousarray(pred))
            if t.dtype not in (torch.float32, torch.float16):
                t = t.float()
            return t.pin_memory().cuda(non_blocking=True)
        if not isinstance(pred, torch.Tensor):
            raise TypeError("outs_dict['maps'] must be a torch.Tensor or a numpy array")
        t = pred.detach()
        if not t.is_cuda:
            t = t.pin_memory().cuda(non_blocking=True)
        if t.dtype not in (torch.float32, torch.float16):
            t = t.float()
        return t

    def _buffers(self, device, N, H, W, R, cap):
        torch = _lib.require_cuda()
        key = (str(device), N, H, W, R, cap)
        buf = self._cache.get(key)
        if buf is None:
            while len(self._cache) >= 2:     # two shapes stay resident (the chunked host path uses two chunk sizes)
                self._cache.pop(next(iter(self._cache)))
            L = _lib.lib()
            ws_bytes = L.ocrpp_db_workspace_bytes(N, H, W, R)
            nb, ns = N * cap * 8 * 2, N * cap * 4
            # one device block and one pinned block: [boxes i16 | scores f32 | counts i32 | status i32]
            total = nb + ns + 8 * N
            buf = {
                "ws": torch.empty(ws_bytes, dtype=torch.uint8, device=device),
                "ws_bytes": ws_bytes,
                "out_dev": torch.empty(total, dtype=torch.uint8, device=device),
                "out_host": torch.empty(total, dtype=torch.uint8, pin_memory=True),
                "wh_host": torch.empty((N, 2), dtype=torch.int32, pin_memory=True),
                "wh_dev": torch.empty((N, 2), dtype=torch.int32, device=device),
                "offs": (0, nb, nb + ns, nb + ns + 4 * N),
            }
            self._cache[key] = buf
        return buf

    def _default_runs(self, H, W):
        if self.max_runs is not None:
            return int(self.max_runs)
        return max(4096, (H * W) // 32)

    def run_device(self, pred, shape_list, boxes_f=False, labels=False, use_padding_resize=False):
        """Enqueues the kernels and returns host views (boxes[N,cap,4,2] i16, scores[N,cap] f32,
        counts[N], status[N], extras dict). Blocks until the results are on the host."""
        torch = _lib.require_cuda()
        t = self._to_device(pred)
        if t.dim() != 4:
            raise ValueError("maps must be [N,C,H,W]")
        if t.stride(3) != 1:
            t = t.contiguous()
        N, _, H, W = t.shape
        cap = self.out_capacity
        if N == 0:
            return (np.zeros((0, cap, 4, 2), np.int16), np.zeros((0, cap), np.float32),
                    np.zeros((0,), np.int32), np.zeros((0,), np.int32), {})
        shape = np.asarray(shape_list, dtype=np.float64).reshape(N, -1)
        L = _lib.lib()
        worst = H * (W + 1)
        R = min(self._default_runs(H, W), worst)
        with torch.cuda.device(t.device):
            stream = torch.cuda.current_stream()
            while True:
                buf = self._buffers(t.device, N, H, W, R, cap)
                o_box, o_sc, o_cnt, o_st = buf["offs"]
                buf["wh_host"][:, 0] = torch.from_numpy(shape[:, 1].astype(np.int32))  # src_w
                buf["wh_host"][:, 1] = torch.from_numpy(shape[:, 0].astype(np.int32))  # src_h
                buf["wh_dev"].copy_(buf["wh_host"], non_blocking=True)
                out = buf["out_dev"]
                base = out.data_ptr()
                extras_dev = {}
                bf_ptr = lab_ptr = None
                if boxes_f:
                    extras_dev["boxes_f"] = torch.empty((N, cap, 4, 2), dtype=torch.float32, device=t.device)
                    bf_ptr = extras_dev["boxes_f"].data_ptr()
                if labels:
                    extras_dev["labels"] = torch.empty((N, H, W), dtype=torch.int32, device=t.device)
                    lab_ptr = extras_dev["labels"].data_ptr()
                _lib.check(L.ocrpp_db_postprocess_ex(
                    t.data_ptr(), _lib.F32 if t.dtype == torch.float32 else _lib.F16, N, H, W,
                    t.stride(0), t.stride(2), buf["wh_dev"].data_ptr(),
                    # numpy evaluates `pred > self.thresh` in the dtype of the map (db_postprocess.py:46)
                    float(np.float16(self.thresh)) if t.dtype == torch.float16 else float(self.thresh),
                    float(self.box_thresh), float(self.unclip_ratio), cap, R,
                    1 if self.use_dilation else 0, 1 if use_padding_resize else 0,
                    _lib.DB_SEMANTICS_CPP if self.cpp_speedup else _lib.DB_SEMANTICS_PYTHON,
                    _lib.DB_SCORE_BOX if (self.score_mode == "box" and not self.cpp_speedup) else _lib.DB_SCORE_POLY,
                    base + o_box, base + o_sc, base + o_cnt, base + o_st, bf_ptr, lab_ptr,
                    buf["ws"].data_ptr(), buf["ws_bytes"], stream.cuda_stream))
                buf["out_host"].copy_(out, non_blocking=True)
                stream.synchronize()
                host = buf["out_host"].numpy()
                status = host[o_st:o_st + 4 * N].view(np.int32)
                if (status & _lib.IMG_VALUE_OUT_OF_RANGE).any():
                    raise _lib.OcrppError("DB probability map holds a value outside [0, 1] (or NaN/Inf): not a probability map")
                if (status & _lib.IMG_RUN_OVERFLOW).any():
                    if R >= worst:
                        raise _lib.OcrppError("DB post-process: internal capacity exceeded (image too large for the unclip buffer)")
                    R = min(worst, R * 8)   # capacity retry (still the CUDA path), not a fallback
                    continue
                if cap < self.max_candidates and (status & _lib.IMG_CANDIDATES_TRUNCATED).any():
                    cap = self.out_capacity = self.max_candidates
                    continue
                break
        boxes = host[o_box:o_box + N * cap * 16].view(np.int16).reshape(N, cap, 4, 2)
        scores = host[o_sc:o_sc + N * cap * 4].view(np.float32).reshape(N, cap)
        counts = host[o_cnt:o_cnt + 4 * N].view(np.int32)
        extras = {k: v.cpu().numpy() for k, v in extras_dev.items()}
        # device views of the same results for callers that stay on the device (deploy/run_ocr.py: crops are cut from
        # these without a host round trip); valid until the next call of this operator
        extras["boxes_dev"] = out[o_box:o_box + N * cap * 16].view(torch.int16).view(N, cap, 4, 2)
        extras["counts_dev"] = out[o_cnt:o_cnt + 4 * N].view(torch.int32)
        return boxes, scores, counts, status, extras

    upload_chunk = 64   # images per H2D chunk of the host-input path

    def _call_host_batch(self, maps, shape_list, use_padding_resize):
        """Host maps (numpy / CPU tensor), large batch: the upload is PCIe-bound and ~20x longer than the kernels,
        so the batch is uploaded in chunks on a copy stream and every chunk is post-processed (kernels, D2H of its
        boxes, host assembly) while the next ones are still in flight."""
        torch = _lib.require_cuda()
        t = torch.from_numpy(np.ascontiguousarray(maps)) if isinstance(maps, np.ndarray) else maps.detach()
        if t.dtype not in (torch.float32, torch.float16):
            t = t.float()
        N = t.shape[0]
        dev = torch.device("cuda", torch.cuda.current_device())
        shape = np.asarray(shape_list, dtype=np.float64).reshape(N, -1)
        nchunks = (N + self.upload_chunk - 1) // self.upload_chunk
        bounds = [N * i // nchunks for i in range(nchunks + 1)]
        main = torch.cuda.current_stream(dev)
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(dev)
        full = torch.empty(t.shape, dtype=t.dtype, device=dev)
        self._copy_stream.wait_stream(main)
        events = []
        with torch.cuda.stream(self._copy_stream):
            for lo, hi in zip(bounds[:-1], bounds[1:]):
                full[lo:hi].copy_(t[lo:hi], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._copy_stream)
                events.append(ev)
        full.record_stream(self._copy_stream)
        res = []
        for ev, lo, hi in zip(events, bounds[:-1], bounds[1:]):
            main.wait_event(ev)
            res += self({"maps": full[lo:hi]}, shape[lo:hi], use_padding_resize)
        return res

    def __call__(self, outs_dict, shape_list, use_padding_resize=False):
        maps = outs_dict["maps"]
        on_host = isinstance(maps, np.ndarray) or not getattr(maps, "is_cuda", True)
        if on_host and len(maps) >= 2 * self.upload_chunk:
            return self._call_host_batch(maps, shape_list, use_padding_resize)
        if self.max_candidates == 0:      # Python branch: num_contours = min(len(contours), 0)
            return [{"points": np.array([], dtype=np.int16), "scores": [], "box_scores": np.zeros((0,), np.float32)}
                    for _ in range(len(maps))]
        boxes, scores, counts, _, _ = self.run_device(maps, shape_list, use_padding_resize=use_padding_resize)
        res_batch = []
        for n in range(boxes.shape[0]):
            k = int(counts[n])
            if k == 0:
                pts = np.array([], dtype=np.int16)
            else:
                pts = boxes[n, :k].copy()
            sc = scores[n, :k].copy()
            # cpp_speedup: the wrapper discards the C++ module's scores (db_postprocess.py:64-67)
            res_batch.append({"points": pts, "scores": [1.0] * k if self.cpp_speedup else [float(v) for v in sc],
                              "box_scores": sc})
        return res_batch


class DistillationDBPostProcess(object):
    """R/pytocr/postprocess/db_postprocess.py:197-226."""

    def __init__(self, model_name=["student"], key=None, thresh=0.3, box_thresh=0.5, max_candidates=1000,
                 unclip_ratio=1.5, use_dilation=False, score_mode="poly", cpp_speedup=False,
                 out_polygon=False, cuda_speedup=True, **kwargs):
        self.model_name = model_name
        self.key = key
        self.post_process = DBPostProcess(thresh=thresh, box_thresh=box_thresh, max_candidates=max_candidates,
                                          unclip_ratio=unclip_ratio, use_dilation=use_dilation,
                                          score_mode=score_mode, cpp_speedup=cpp_speedup,
                                          out_polygon=out_polygon, cuda_speedup=cuda_speedup)

    def __call__(self, predicts, shape_list):
        results = {}
        for k in self.model_name:
            results[k] = self.post_process(predicts[k], shape_list=shape_list)
        return results
