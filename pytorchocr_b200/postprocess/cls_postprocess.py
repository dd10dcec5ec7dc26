"""Direction-classifier post-process behind the reference's operator API, computed by libocrpp (sm_100a).

Mirrors R/pytocr/postprocess/cls_postprocess.py:4-20 (`ClsPostProcess(label_list)`, `__call__(preds, label=None)` ->
`[(label_list[argmax], preds[i, argmax])]`). The per-row argmax + max is the CTC path's first kernel
(`ocrpp_ctc_greedy` on the `[B,C]` scores viewed as `[T=1,B,C]`; first maximal index wins, as numpy's argmax does);
used by R/deploy/pytorch/run_ocr.py:206-211 between the crop step and the recogniser."""
import numpy as np

from .. import _lib


class ClsPostProcess(object):
    """ Convert between text-label and text-index """

    def __init__(self, label_list, cuda_speedup=True, **kwargs):
        if not cuda_speedup:
            raise ValueError("pytorchocr_b200 only holds the CUDA path: set PostProcess.cuda_speedup: True "
                             "(with it off the reference's own ClsPostProcess runs)")
        self.label_list = label_list
        self._bufs = {}

    def __call__(self, preds, label=None, *args, **kwargs):
        torch = _lib.require_cuda()
        t = torch.from_numpy(np.ascontiguousarray(preds)) if isinstance(preds, np.ndarray) else preds.detach()
        if t.dim() != 2:
            raise ValueError("ClsPostProcess expects [B, num_classes] scores (got %s)" % (tuple(t.shape),))
        if t.dtype not in (torch.float32, torch.float16):
            t = t.float()
        t = t.cuda() if not t.is_cuda else t
        if t.stride(1) != 1:
            t = t.contiguous()
        B, C = t.shape
        if C > len(self.label_list):
            raise IndexError("preds has %d classes but label_list has %d entries" % (C, len(self.label_list)))
        decode_out = []
        if B > 0:
            with torch.cuda.device(t.device):
                key = (t.device, B)
                buf = self._bufs.get(key)
                if buf is None:
                    buf = dict(i=torch.empty(3 * B, dtype=torch.int32, device=t.device),
                               f=torch.empty(2 * B, dtype=torch.float32, device=t.device),
                               ih=torch.empty(3 * B, dtype=torch.int32, pin_memory=True),
                               fh=torch.empty(2 * B, dtype=torch.float32, pin_memory=True))
                    self._bufs[key] = buf
                i, f = buf["i"], buf["f"]
                stream = torch.cuda.current_stream()
                # T = 1: idx [B,1] | len [B] | raw argmax [B,1] in `i`;  max score [B,1] | conf [B] in `f`
                _lib.check(_lib.lib().ocrpp_ctc_greedy(
                    t.data_ptr(), _lib.F32 if t.dtype == torch.float32 else _lib.F16, 1, B, C, 0, t.stride(0),
                    i.data_ptr(), f.data_ptr(), i.data_ptr() + 4 * B, f.data_ptr() + 4 * B, i.data_ptr() + 8 * B,
                    stream.cuda_stream))
                buf["ih"].copy_(i, non_blocking=True)
                buf["fh"].copy_(f, non_blocking=True)
                stream.synchronize()
            idx = buf["ih"].numpy()[2 * B:3 * B]
            score = buf["fh"].numpy()[:B]
            decode_out = [(self.label_list[int(k)], score[n]) for n, k in enumerate(idx)]
        if label is None:
            return decode_out
        label = [(self.label_list[idx], 1.0) for idx in label]
        return decode_out, label
