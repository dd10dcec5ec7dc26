"""CTC label decode behind the reference's operator API, computed by libocrpp (sm_100a).

Mirrors R/pytocr/postprocess/rec_postprocess.py: BaseRecLabelDecode :5-62, CTCLabelDecode :65-93,
DistillationCTCLabelDecode :96-125 (same ctor kwargs, __call__(preds, label=None), return values).
The per-step argmax/max over classes, the blank/repeat collapse and the confidence mean run on the
GPU; the host only maps the kept class ids to characters."""
import numpy as np

from .. import _lib


class BaseRecLabelDecode(object):
    """Convert between text-label and text-index (dictionary handling as the reference, :8-30)."""

    def __init__(self, character_dict_path=None, use_space_char=False):
        self.beg_str = "sos"
        self.end_str = "eos"
        self.character_str = []
        if character_dict_path is None:
            self.character_str = "0123456789abcdefghijklmnopqrstuvwxyz"
            dict_character = list(self.character_str)
        else:
            with open(character_dict_path, "rb") as fin:
                for line in fin.readlines():
                    self.character_str.append(line.decode("UTF-8").strip("\n").strip("\r\n"))
            if use_space_char:
                self.character_str.append(" ")
            dict_character = list(self.character_str)
        dict_character = self.add_special_char(dict_character)
        self.dict = {char: i for i, char in enumerate(dict_character)}
        self.character = dict_character
        # code-point table for the vectorised id -> str path (index 0 = blank -> NUL terminator)
        single = all(len(c) == 1 for c in dict_character[1:])
        self._codepoints = None
        if single:
            cp = np.zeros(len(dict_character), np.uint32)
            cp[1:] = [ord(c) for c in dict_character[1:]]
            if (cp[1:] != 0).all():
                self._codepoints = cp

    def add_special_char(self, dict_character):
        return dict_character

    def get_ignored_tokens(self):
        return [0]  # for ctc blank

    def decode(self, text_index, text_prob=None, is_remove_duplicate=False):
        """Host-side label -> text conversion (reference :35-59); used for ground-truth labels.
        Predictions do not come through here: they are collapsed on the GPU."""
        result_list = []
        ignored = self.get_ignored_tokens()
        for b in range(len(text_index)):
            chars, confs = [], []
            row = text_index[b]
            for i in range(len(row)):
                if row[i] in ignored:
                    continue
                if is_remove_duplicate and i > 0 and row[i - 1] == row[i]:
                    continue
                chars.append(self.character[int(row[i])])
                confs.append(text_prob[b][i] if text_prob is not None else 1)
            result_list.append(("".join(chars), np.mean(confs) if confs else float("nan")))
        return result_list

    def _strings(self, idx, lens):
        """idx int32 [B,T] (zero beyond lens[b]) -> list of B python strings."""
        B, T = idx.shape
        if self._codepoints is not None and T > 0:
            if idx.size and int(idx.max()) >= len(self._codepoints):
                raise IndexError("class id %d outside the %d-entry dictionary" % (int(idx.max()), len(self._codepoints)))
            cp = np.ascontiguousarray(self._codepoints[idx])
            return cp.view("<U%d" % T).reshape(B).tolist()
        return ["".join(self.character[int(i)] for i in idx[b, :lens[b]]) for b in range(B)]


class CTCLabelDecode(BaseRecLabelDecode):
    """Drop-in for the reference's CTCLabelDecode, selected by `cuda_speedup: True`."""

    def __init__(self, character_dict_path=None, use_space_char=False, cuda_speedup=True, **kwargs):
        super(CTCLabelDecode, self).__init__(character_dict_path, use_space_char)
        if not cuda_speedup:
            raise _lib.OcrppError("pytorchocr_b200 implements only the CUDA path: set PostProcess.cuda_speedup: True "
                                  "(with the flag off the reference's own CTCLabelDecode runs)")
        self._bufs = {}

    def add_special_char(self, dict_character):
        return ["blank"] + dict_character

    # -- device plumbing ------------------------------------------------------------------------
    def _device_view(self, preds):
        """Returns (tensor_on_cuda, T, B, C, stride_t, stride_b). torch [T,B,C] as the CTC head emits
        (rec_ctc_head.py:17-36) or numpy [B,T,C] as the reference also accepts (:80-82)."""
        torch = _lib.require_cuda()
        if isinstance(preds, tuple):
            preds = preds[-1]
        if isinstance(preds, np.ndarray):
            t = torch.from_numpy(np.ascontiguousarray(preds))
            if t.dtype not in (torch.float32, torch.float16):
                t = t.float()
            t = t.pin_memory().cuda(non_blocking=True)     # [B,T,C]
            B, T, C = t.shape
            return t, T, B, C, t.stride(1), t.stride(0)
        if not isinstance(preds, torch.Tensor):
            raise TypeError("preds must be a torch.Tensor, a tuple ending in one, or a numpy array")
        t = preds.detach()
        if not t.is_cuda:
            t = t.pin_memory().cuda(non_blocking=True) if t.device.type == "cpu" else t.cuda()
        if t.dtype not in (torch.float32, torch.float16):
            t = t.float()
        if t.dim() != 3:
            raise ValueError("preds must be [T,B,C]")
        if t.stride(2) != 1:
            t = t.contiguous()
        T, B, C = t.shape
        return t, T, B, C, t.stride(0), t.stride(1)

    def _buffers(self, device, B, T):
        torch = _lib.require_cuda()
        key = (str(device), B, T)
        buf = self._bufs.get(key)
        if buf is None:
            self._bufs.clear()
            # one int32 block [idx B*T | len B] and one float block [prob B*T | conf B] per side
            buf = {
                "i_dev": torch.empty(B * T + B, dtype=torch.int32, device=device),
                "f_dev": torch.empty(B * T + B, dtype=torch.float32, device=device),
                "i_host": torch.empty(B * T + B, dtype=torch.int32, pin_memory=True),
                "c_host": torch.empty(B, dtype=torch.float32, pin_memory=True),
            }
            self._bufs[key] = buf
        return buf

    def decode_device(self, preds):
        """Runs the kernels; returns (idx[B,T] int32 numpy, len[B], conf[B] float32) on the host."""
        torch = _lib.require_cuda()
        t, T, B, C, st, sb = self._device_view(preds)
        if C > len(self.character):
            raise ValueError("preds has %d classes but the dictionary has %d entries" % (C, len(self.character)))
        if B == 0 or T == 0:
            return np.zeros((B, T), np.int32), np.zeros((B,), np.int32), np.full((B,), np.nan, np.float32)
        with torch.cuda.device(t.device):
            buf = self._buffers(t.device, B, T)
            i_dev, f_dev = buf["i_dev"], buf["f_dev"]
            stream = torch.cuda.current_stream().cuda_stream
            L = _lib.lib()
            _lib.check(L.ocrpp_ctc_greedy(
                t.data_ptr(), _lib.F32 if t.dtype == torch.float32 else _lib.F16, T, B, C, st, sb,
                i_dev.data_ptr(), f_dev.data_ptr(), i_dev.data_ptr() + 4 * B * T,
                f_dev.data_ptr() + 4 * B * T, None, stream))
            buf["i_host"].copy_(i_dev, non_blocking=True)
            buf["c_host"].copy_(f_dev[B * T:], non_blocking=True)
            torch.cuda.current_stream().synchronize()
        ih = buf["i_host"].numpy()
        return ih[:B * T].reshape(B, T), ih[B * T:], buf["c_host"].numpy()

    def __call__(self, preds, label=None, *args, **kwargs):
        idx, lens, conf = self.decode_device(preds)
        strings = self._strings(idx, lens)
        text = list(zip(strings, conf.tolist()))
        if label is None:
            return text
        return text, self.decode(label)

    def decode_batch(self, preds_list):
        """Decodes many recogniser outputs with ONE library call: the reference's
        pipeline runs the CTC decode once per text crop with B = 1 (R/deploy/pytorch/run_ocr.py:221-224),
        which is launch-bound. `preds_list` holds per-crop tensors [T_i, 1, C] (or [T_i, C]) with
        different T_i; they are packed into one [T_max, B, C] device tensor whose padding rows are pure
        blank (class 0 = 1.0). Blank steps are dropped by the collapse and do not enter the confidence
        mean, so every crop decodes exactly as `self(pred)[0]` would."""
        torch = _lib.require_cuda()
        if len(preds_list) == 0:
            return []
        items = []
        for t in preds_list:
            if isinstance(t, tuple):
                t = t[-1]
            if isinstance(t, np.ndarray):
                t = torch.from_numpy(t)
            t = t.detach()
            if t.dim() == 3:
                if t.shape[1] != 1:
                    raise ValueError("decode_batch expects per-crop predictions [T, 1, C] or [T, C]")
                t = t[:, 0]
            items.append(t)
        dev = next((t.device for t in items if t.is_cuda), torch.device("cuda", torch.cuda.current_device()))
        C = items[0].shape[1]
        Tmax = max(int(t.shape[0]) for t in items)
        packed = torch.zeros((Tmax, len(items), C), dtype=torch.float32, device=dev)
        packed[:, :, 0] = 1.0
        for b, t in enumerate(items):
            packed[:t.shape[0], b] = t.to(dev, dtype=torch.float32, non_blocking=True)
        return self(packed)


class DistillationCTCLabelDecode(CTCLabelDecode):
    """Reference :96-125 - dict of model outputs -> dict of decoded results."""

    def __init__(self, character_dict_path=None, use_space_char=False, model_name=["student"],
                 key=None, **kwargs):
        super(DistillationCTCLabelDecode, self).__init__(character_dict_path, use_space_char, **kwargs)
        if not isinstance(model_name, list):
            model_name = [model_name]
        self.model_name = model_name
        self.key = key

    def __call__(self, preds, label=None, *args, **kwargs):
        output = dict()
        for name in self.model_name:
            pred = preds[name]
            if self.key is not None:
                pred = pred[self.key]
            output[name] = super().__call__(pred, label=label, *args, **kwargs)
        return output
