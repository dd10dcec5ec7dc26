"""Shared device plumbing of the PSE / PAN operators (libocrpp entry points ocrpp_pse_postprocess /
ocrpp_pan_postprocess, include/ocrpp.h). The head output never leaves the device; boxes, scores
and counts come back through one pinned buffer."""
import numpy as np

from .. import _lib


class ExpandOperator(object):
    """Base of PSEPostProcess / PANPostProcess: same `__call__(outs_dict, shape_list)` contract as
    R/pytocr/postprocess/pse_postprocess.py:28-53 and pan_postprocess.py:30-60."""

    _entry = None            # "pse" | "pan"
    _default_boxes = 512     # output capacity per image; x4 on OCRPP_IMG_CANDIDATES_TRUNCATED

    def _init_common(self, thresh, box_thresh, min_area, scale, out_polygon, cuda_speedup, max_runs,
                     max_boxes, maps_at_processing_res):
        if not cuda_speedup:
            raise _lib.OcrppError("pytorchocr_b200 implements only the CUDA path: set PostProcess.cuda_speedup: True "
                                  "(with the flag off the reference's own operator runs)")
        if out_polygon:
            raise NotImplementedError("out_polygon is not on the configured path (SURVEY.md 8(f) rank 4)")
        if scale not in (1, 2, 4):
            raise ValueError("scale must be 1, 2 or 4")
        self.thresh = thresh
        self.box_thresh = box_thresh
        self.min_area = min_area
        self.out_polygon = out_polygon
        self.scale = scale
        self.max_runs = max_runs
        self.max_boxes = int(max_boxes or self._default_boxes)
        # False (reference semantics): maps are the 1/4-resolution head output and are up-sampled by
        # 4 // scale first. True: maps are ALREADY at processing resolution (the tensor the reference
        # materialises after F.interpolate); used by bench.py for BASELINE.json configs[2] and [3].
        self.maps_at_processing_res = bool(maps_at_processing_res)
        self._cache = {}

    # -- hooks ----------------------------------------------------------------------------------
    def _seed_min_area(self):
        raise NotImplementedError

    def _check_channels(self, C):
        raise NotImplementedError

    # -- device plumbing ------------------------------------------------------------------------
    @staticmethod
    def _to_device(pred):
        torch = _lib.require_cuda()
        if isinstance(pred, np.ndarray):
            pred = torch.from_numpy(np.ascontiguousarray(pred))
        if not isinstance(pred, torch.Tensor):
            raise TypeError("outs_dict['maps'] must be a torch.Tensor")
        t = pred.detach()
        if not t.is_cuda:
            t = t.pin_memory().cuda(non_blocking=True)
        if t.dtype not in (torch.float32, torch.float16):
            t = t.float()
        if t.dim() != 4:
            raise ValueError("maps must be [N,C,H,W]")
        if t.stride(3) != 1:
            t = t.contiguous()
        return t

    def _factors(self):
        fin = 1 if self.maps_at_processing_res else 4 // self.scale
        fout = self.scale
        return fin, fout

    def _workspace_bytes(self, L, N, C, h, w, fin, cap, R, arena):
        if self._entry == "pse":
            return L.ocrpp_pse_workspace_bytes(N, C, h, w, fin, cap, R, arena)
        return L.ocrpp_pan_workspace_bytes(N, h, w, fin, cap, R, arena)

    def _buffers(self, device, N, C, h, w, fin, cap, R, arena):
        torch = _lib.require_cuda()
        key = (str(device), N, C, h, w, fin, cap, R, arena)
        buf = self._cache.get(key)
        if buf is None:
            while len(self._cache) >= 2:     # two shapes stay resident (the chunked host path uses two chunk sizes)
                self._cache.pop(next(iter(self._cache)))
            L = _lib.lib()
            ws_bytes = self._workspace_bytes(L, N, C, h, w, fin, cap, R, arena)
            nb, ns = N * cap * 16, N * cap * 4
            total = nb + ns + 8 * N
            buf = {
                "ws": torch.empty(ws_bytes, dtype=torch.uint8, device=device),
                "ws_bytes": ws_bytes,
                "out_dev": torch.empty(total, dtype=torch.uint8, device=device),
                "out_host": torch.empty(total, dtype=torch.uint8, pin_memory=True),
                "shape_host": torch.empty((N, 4), dtype=torch.float64, pin_memory=True),
                "shape_dev": torch.empty((N, 4), dtype=torch.float64, device=device),
                "offs": (0, nb, nb + ns, nb + ns + 4 * N),
            }
            self._cache[key] = buf
        return buf

    def _call_lib(self, L, t, N, C, h, w, fin, fout, buf, cap, R, arena, bf_ptr, lab_ptr, stream):
        torch = _lib.require_cuda()
        o_box, o_sc, o_cnt, o_st = buf["offs"]
        base = buf["out_dev"].data_ptr()
        dtype = _lib.F32 if t.dtype == torch.float32 else _lib.F16
        tail = (base + o_box, base + o_sc, base + o_cnt, base + o_st, bf_ptr, lab_ptr,
                buf["ws"].data_ptr(), buf["ws_bytes"], stream.cuda_stream)
        if self._entry == "pse":
            return L.ocrpp_pse_postprocess(
                t.data_ptr(), dtype, N, C, h, w, t.stride(0), t.stride(1), t.stride(2), fin, fout,
                buf["shape_dev"].data_ptr(), float(self.thresh), float(self.box_thresh),
                float(self._seed_min_area()), float(self.min_area), cap, R, arena, *tail)
        return L.ocrpp_pan_postprocess(
            t.data_ptr(), dtype, N, h, w, t.stride(0), t.stride(1), t.stride(2), fin, fout,
            buf["shape_dev"].data_ptr(), float(self.thresh), float(self.box_thresh),
            float(self._seed_min_area()), float(self.min_area), cap, R, arena, *tail)

    def run_device(self, pred, shape_list, boxes_f=False, labels=False):
        """Enqueues the kernels and returns host views (boxes[N,cap,4,2] i16, scores[N,cap] f32,
        counts[N], status[N], extras). Blocks until the results are on the host."""
        torch = _lib.require_cuda()
        t = self._to_device(pred)
        N, C, h, w = t.shape
        self._check_channels(C)
        fin, fout = self._factors()
        H, W = h * fin, w * fin
        cap = self.max_boxes
        if N == 0:
            return (np.zeros((0, cap, 4, 2), np.int16), np.zeros((0, cap), np.float32),
                    np.zeros((0,), np.int32), np.zeros((0,), np.int32), {})
        shape = np.ascontiguousarray(np.asarray(shape_list, dtype=np.float64).reshape(N, -1)[:, :4])
        L = _lib.lib()
        worst_runs = H * ((W + 1) // 2)
        R = min(int(self.max_runs) if self.max_runs is not None else max(4096, (H * W) // 32), worst_runs)
        worst_arena = 4 * N * H * W
        arena = min(N * H * W, worst_arena)
        with torch.cuda.device(t.device):
            stream = torch.cuda.current_stream()
            while True:
                buf = self._buffers(t.device, N, C, h, w, fin, cap, R, arena)
                o_box, o_sc, o_cnt, o_st = buf["offs"]
                buf["shape_host"].copy_(torch.from_numpy(shape))
                buf["shape_dev"].copy_(buf["shape_host"], non_blocking=True)
                extras_dev = {}
                bf_ptr = lab_ptr = None
                if boxes_f:
                    extras_dev["boxes_f"] = torch.empty((N, cap, 4, 2), dtype=torch.float32, device=t.device)
                    bf_ptr = extras_dev["boxes_f"].data_ptr()
                if labels:
                    extras_dev["labels"] = torch.empty((N, H, W), dtype=torch.int32, device=t.device)
                    lab_ptr = extras_dev["labels"].data_ptr()
                _lib.check(self._call_lib(L, t, N, C, h, w, fin, fout, buf, cap, R, arena, bf_ptr, lab_ptr, stream))
                buf["out_host"].copy_(buf["out_dev"], non_blocking=True)
                stream.synchronize()
                host = buf["out_host"].numpy()
                status = host[o_st:o_st + 4 * N].view(np.int32)
                if (status & _lib.IMG_RUN_OVERFLOW).any():   # capacity retry (still the CUDA path)
                    if R >= worst_runs and arena >= worst_arena:
                        raise _lib.OcrppError("%s post-process: internal capacity exceeded" % self._entry)
                    R = min(worst_runs, R * 8)
                    arena = worst_arena
                    continue
                if (status & _lib.IMG_CANDIDATES_TRUNCATED).any():
                    cap *= 4
                    self.max_boxes = cap
                    continue
                break
        boxes = host[o_box:o_box + N * cap * 16].view(np.int16).reshape(N, cap, 4, 2)
        scores = host[o_sc:o_sc + N * cap * 4].view(np.float32).reshape(N, cap)
        counts = host[o_cnt:o_cnt + 4 * N].view(np.int32)
        extras = {k: v.cpu().numpy() for k, v in extras_dev.items()}
        return boxes, scores, counts, status, extras

    upload_chunk_bytes = 128 << 20   # H2D chunk of the host-input path

    def _call_host_batch(self, pred, shape_list, per):
        """CPU-tensor maps: the upload is PCIe-bound and an order of magnitude longer than the kernels, so the batch
        goes up in chunks of `per` images on a copy stream and every chunk is post-processed while the next ones are
        still in flight (as DBPostProcess does)."""
        torch = _lib.require_cuda()
        t = pred.detach()
        if t.dtype not in (torch.float32, torch.float16):
            t = t.float()
        N = t.shape[0]
        dev = torch.device("cuda", torch.cuda.current_device())
        shape = np.asarray(shape_list, dtype=np.float64).reshape(N, -1)
        nchunks = (N + per - 1) // per
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
            res += self({"maps": full[lo:hi]}, shape[lo:hi])
        return res

    def __call__(self, outs_dict, shape_list):
        torch = _lib.require_cuda()
        pred = outs_dict["maps"]
        assert isinstance(pred, torch.Tensor)      # pse_postprocess.py:30 / pan_postprocess.py:32
        if not pred.is_cuda and pred.dim() == 4 and pred.shape[0] >= 2:
            per = max(1, self.upload_chunk_bytes // max(1, pred[0].numel() * pred.element_size()))
            if pred.shape[0] >= 2 * per:
                return self._call_host_batch(pred, shape_list, per)
        boxes, scores, counts, _, _ = self.run_device(pred, shape_list)
        res_batch = []
        for n in range(boxes.shape[0]):
            k = int(counts[n])
            pts = boxes[n, :k].copy() if k else np.array([], dtype=np.int16)
            res_batch.append({"points": pts, "scores": [s for s in scores[n, :k].copy()]})
        return res_batch
