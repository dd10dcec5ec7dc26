import sys, os; sys.path.insert(0, "/root/repo")
import numpy as np, torch, cv2
from pytorchocr_b200 import _lib
from pytorchocr_b200.postprocess import build_post_process
L = _lib.lib()
rng = np.random.default_rng(2)
H, W = 120, 152
maps = []
for sig in (0.6, 1.0, 1.5, 2.5, 0.45, 2.0):
    p = cv2.GaussianBlur(rng.random((H, W)).astype(np.float32), (0, 0), sig)
    lo, hi = np.quantile(p, 0.02), np.quantile(p, 0.98)
    maps.append(np.clip((p - lo) / (hi - lo), 0, 1).astype(np.float32))
maps = torch.from_numpy(np.stack(maps)[:, None]).cuda()
N = maps.shape[0]
sl = np.array([[H, W, 1.0, 1.0]] * N)
q = float(np.quantile(maps.cpu().numpy(), 0.5))
res = {}
for path in (3, 2):
    _lib.check(L.ocrpp_set_tuning(0, path))
    for fill in (0, 255, 0x55):
        op = build_post_process({"name": "DBPostProcess", "thresh": q, "box_thresh": q + 0.02, "unclip_ratio": 1.7, "cuda_speedup": True, "max_runs": 32768})
        op.run_device(maps, sl)
        for b in op._cache.values():
            b["ws"].fill_(fill)
        boxes, scores, counts, status, ex = op.run_device(maps, sl)
        res[(path, fill)] = (boxes.copy(), scores.copy(), counts.copy())
        print("path", path, "fill", fill, "counts", counts, "status", status)
ref = res[(3, 0)]
for key, cur in res.items():
    for n in range(N):
        k = ref[2][n]
        if cur[2][n] != k:
            print(key, "image", n, "count", cur[2][n], "vs", k); continue
        d = np.nonzero((cur[0][n, :k] != ref[0][n, :k]).any((1, 2)) | (cur[1][n, :k] != ref[1][n, :k]))[0]
        if len(d):
            print(key, "image", n, "differs at", d[:6], [(float(cur[1][n][i]), float(ref[1][n][i])) for i in d[:3]], cur[0][n][d[0]].tolist())
