import sys; sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
from pytorchocr_b200 import synth, _lib
from pytorchocr_b200.postprocess import build_post_process
from oracle.db_oracle import DBPostProcessOracle
from db_compare import compare_image, classify, on_discontinuity
import test_full_size_gpu as T
H, W = 736, 1280
uniq = synth.db_batch(T.UNIQ if hasattr(T, "UNIQ") else 8) if False else None
