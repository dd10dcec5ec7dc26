"""PAN / PAN++ post-processing behind the reference's operator API, computed by libocrpp (sm_100a).

Mirrors R/pytocr/postprocess/pan_postprocess.py:10-113: same ctor kwargs (`thresh, box_thresh,
min_area, min_kernel_area, scale, out_polygon, **kwargs`), same `__call__(outs_dict, shape_list)`
and return structure. Text/kernel thresholding, the two 4-connected labelings, the area-ratio
flags and mean embeddings of pan_postprocess_fast/pa.pyx, the gated expansion and generate_box all
run on the device; the 4 embedding channels are only read where a flagged kernel makes a claim."""
from ._expand_op import ExpandOperator


class PANPostProcess(ExpandOperator):
    _entry = "pan"

    def __init__(self, thresh=0.5, box_thresh=0.85, min_area=16, min_kernel_area=2.6, scale=4,
                 out_polygon=False, cuda_speedup=True, max_runs=None, max_boxes=None,
                 maps_at_processing_res=False, **kwargs):
        self._init_common(thresh, box_thresh, min_area, scale, out_polygon, cuda_speedup, max_runs,
                          max_boxes, maps_at_processing_res)
        self.min_kernel_area = min_kernel_area / float(scale ** 2)   # pan_postprocess.py:25

    def _seed_min_area(self):
        return self.min_kernel_area

    def _check_channels(self, C):
        if C != 6:
            raise ValueError("PAN maps must have 6 channels (text, kernel, 4-d embedding), got %d" % C)
