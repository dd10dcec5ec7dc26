"""TEST INFRASTRUCTURE ONLY (oracle). CTC greedy decode, CPU restatement of
R/pytocr/postprocess/rec_postprocess.py (BaseRecLabelDecode :5-62, CTCLabelDecode :65-93).

Two forms:
  * CTCLabelDecodeNumpy - the reference's own numpy + pure-Python algorithm restated line by line
    (argmax/max :83-84, decode :35-59). Slow; used on small cases and as the "as shipped" timing.
  * CTCLabelDecodeOracle - same results through the C loop in oracle/c/ocr_oracle.c
    (oracle_ctc_greedy) for sizes where the Python double loop is too slow.
"""
import numpy as np

from . import clib


def load_character(character_dict_path=None, use_space_char=False):
    """rec_postprocess.py:8-30 + CTCLabelDecode.add_special_char :91-93."""
    if character_dict_path is None:
        chars = list("0123456789abcdefghijklmnopqrstuvwxyz")
    else:
        chars = []
        with open(character_dict_path, "rb") as fin:
            for line in fin.readlines():
                chars.append(line.decode("UTF-8").strip("\n").strip("\r\n"))
        if use_space_char:
            chars.append(" ")
    return ["blank"] + chars


def _to_btc(preds):
    """rec_postprocess.py:78-82: tuple -> last; torch [T,B,C] -> numpy [B,T,C] view; numpy stays."""
    if isinstance(preds, tuple):
        preds = preds[-1]
    if hasattr(preds, "detach"):
        preds = preds.detach().cpu().numpy().transpose((1, 0, 2))
    return preds


class CTCLabelDecodeNumpy(object):
    def __init__(self, character_dict_path=None, use_space_char=False, **kwargs):
        self.character = load_character(character_dict_path, use_space_char)

    def decode(self, text_index, text_prob=None, is_remove_duplicate=False):
        result_list = []
        for b in range(len(text_index)):
            char_list, conf_list = [], []
            for i in range(len(text_index[b])):
                if text_index[b][i] in [0]:
                    continue
                if is_remove_duplicate and i > 0 and text_index[b][i - 1] == text_index[b][i]:
                    continue
                char_list.append(self.character[int(text_index[b][i])])
                conf_list.append(text_prob[b][i] if text_prob is not None else 1)
            with np.errstate(all="ignore"):
                import warnings
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    result_list.append(("".join(char_list), np.mean(conf_list)))
        return result_list

    def __call__(self, preds, label=None, *args, **kwargs):
        preds = _to_btc(preds)
        preds_idx = preds.argmax(axis=2)
        preds_prob = preds.max(axis=2)
        text = self.decode(preds_idx, preds_prob, is_remove_duplicate=True)
        if label is None:
            return text
        return text, self.decode(label)


class CTCLabelDecodeOracle(CTCLabelDecodeNumpy):
    def __call__(self, preds, label=None, *args, **kwargs):
        preds = _to_btc(preds)  # [B,T,C] (possibly a transposed view)
        tbc = np.asarray(preds, dtype=np.float32).transpose((1, 0, 2))
        if tbc.strides[2] != 4:
            tbc = np.ascontiguousarray(tbc)
        idx, prob, ln, _ = clib.ctc_greedy(tbc)
        text = []
        import warnings
        for b in range(idx.shape[0]):
            n = int(ln[b])
            chars = "".join(self.character[int(i)] for i in idx[b, :n])
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                conf = np.mean(list(prob[b, :n]))  # :58 np.mean of a list of float32 scalars
            text.append((chars, conf))
        if label is None:
            return text
        return text, self.decode(label)
