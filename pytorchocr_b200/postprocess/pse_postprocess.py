"""PSENet post-processing behind the reference's operator API, computed by libocrpp (sm_100a).

Mirrors R/pytocr/postprocess/pse_postprocess.py:10-105: same ctor kwargs (`thresh, box_thresh,
min_area, scale, out_polygon, **kwargs`), same `__call__(outs_dict, shape_list)`, same return
structure (list of {"points": int16 [K,4,2] (shape (0,) when empty), "scores": [float32]*K}).
The nearest up-sampling, sigmoid, thresholding, text masking, the connected components and FIFO
expansion of pse_postprocess_fast/pse.pyx and generate_box all run on the device."""
from ._expand_op import ExpandOperator


class PSEPostProcess(ExpandOperator):
    _entry = "pse"

    def __init__(self, thresh=0.5, box_thresh=0.85, min_area=16, scale=4, out_polygon=False,
                 cuda_speedup=True, max_runs=None, max_boxes=None, maps_at_processing_res=False, **kwargs):
        self._init_common(thresh, box_thresh, min_area, scale, out_polygon, cuda_speedup, max_runs,
                          max_boxes, maps_at_processing_res)

    def _seed_min_area(self):
        return self.min_area / (self.scale ** 2)     # pse_postprocess.py:56

    def _check_channels(self, C):
        if not 1 <= C <= 8:
            raise ValueError("PSE maps must have between 1 and 8 kernel channels, got %d" % C)
