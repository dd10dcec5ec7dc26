"""Operator registry mirroring R/pytocr/postprocess/__init__.py:13-30.

`build_post_process(config, global_config)` has the reference's semantics (deep copy, pop `name`,
"None" -> no operator, merge ALL of `Global` into the kwargs, construct by class name). The new
switch is `PostProcess.cuda_speedup` (next to the existing `cpp_speedup`); this package only holds
the CUDA implementations, so the flag must be True here - with it off (the default in the
reference) the reference's own classes run (see INTEGRATION.md for the two-line dispatch patch).
"""
import copy

from .cls_postprocess import ClsPostProcess
from .rec_postprocess import CTCLabelDecode, DistillationCTCLabelDecode

__all__ = ["build_post_process"]

_REGISTRY = {
    "CTCLabelDecode": CTCLabelDecode,
    "DistillationCTCLabelDecode": DistillationCTCLabelDecode,
    "ClsPostProcess": ClsPostProcess,
}

try:  # detection operators (registered as they are built)
    from .db_postprocess import DBPostProcess, DistillationDBPostProcess
    _REGISTRY.update(DBPostProcess=DBPostProcess, DistillationDBPostProcess=DistillationDBPostProcess)
except ImportError:  # pragma: no cover
    pass
try:
    from .pse_postprocess import PSEPostProcess
    _REGISTRY.update(PSEPostProcess=PSEPostProcess)
except ImportError:  # pragma: no cover
    pass
try:
    from .pan_postprocess import PANPostProcess
    _REGISTRY.update(PANPostProcess=PANPostProcess)
except ImportError:  # pragma: no cover
    pass


def build_post_process(config, global_config=None):
    config = copy.deepcopy(config)
    module_name = config.pop("name")
    if module_name == "None":
        return
    if global_config is not None:
        config.update(global_config)
    assert module_name in _REGISTRY, Exception(
        "post process only support {}".format(sorted(_REGISTRY)))
    config.setdefault("cuda_speedup", False)   # reference default: flag absent -> off
    return _REGISTRY[module_name](**config)
