"""Development aid: run the DB CUDA path on the parity-test inputs and dump raw outputs to
gpurun_out/db_dump.npz for offline comparison with the oracles."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2
import numpy as np
import torch
from pytorchocr_b200.postprocess import build_post_process
from pytorchocr_b200 import synth

out = {}
def run(name, maps, sl, **kw):
    cfg = dict(name="DBPostProcess", thresh=0.3, box_thresh=0.5, unclip_ratio=1.7, cuda_speedup=True)
    cfg.update(kw)
    op = build_post_process(cfg)
    b, s, c, st, ex = op.run_device(torch.from_numpy(maps).cuda(), sl, boxes_f=True, labels=True)
    out[name + "_maps"] = maps; out[name + "_boxes"] = b.copy(); out[name + "_scores"] = s.copy()
    out[name + "_counts"] = c.copy(); out[name + "_boxes_f"] = ex["boxes_f"]; out[name + "_status"] = st.copy()
    out[name + "_cfg"] = np.array([cfg["thresh"], cfg["box_thresh"], cfg["unclip_ratio"]])
    print(name, c, st)

H, W = 96, 128
m = np.full((H, W), 0.02, np.float32)
m[8:88, 8:120] = 0.9; m[16:80, 16:112] = 0.05; m[24:72, 24:104] = 0.8; m[32:64, 32:96] = 0.1; m[40:56, 40:88] = 0.95
m[0:5, 0:9] = 0.9; m[90:96, 50:70] = 0.9; m[2, 100] = 0.9; m[60:70, 124] = 0.9
for i in range(6):
    m[85 + i if 85 + i < H else H - 1, 2 + i] = 0.9
run("nested", m[None, None], np.array([[H, W, 1.0, 1.0]]), box_thresh=0.3)

for seed in range(6):
    rng = np.random.default_rng(seed)
    H, W = 120, 152
    sig = [0.6, 1.0, 1.5, 2.5, 0.45, 2.0][seed]
    p = cv2.GaussianBlur(rng.random((2, H, W)).astype(np.float32), (0, 0), sig)
    p = np.stack([cv2.GaussianBlur(rng.random((H, W)).astype(np.float32), (0, 0), sig) for _ in range(2)])
    lo, hi = np.quantile(p, 0.02), np.quantile(p, 0.98)
    p = np.clip((p - lo) / (hi - lo), 0, 1).astype(np.float32)
    q = float(np.quantile(p, [0.45, 0.5, 0.55, 0.6, 0.5, 0.4][seed]))
    run("blob%d" % seed, p[:, None], np.array([[H, W, 1.0, 1.0]] * 2), thresh=q, box_thresh=q + 0.02)

maps = synth.db_batch(2, seed=5, H=160, W=256)
run("kinds", maps, np.array([[160, 256, 1.0, 1.0]] * 2))
os.makedirs("gpurun_out", exist_ok=True)
np.savez_compressed("gpurun_out/db_dump.npz", **out)
