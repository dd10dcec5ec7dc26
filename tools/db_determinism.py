"""Development aid: runs the same DB inputs many times on each stage-2 path and reports any run whose outputs differ
from the first one (a race shows up as a difference)."""
import sys; sys.path.insert(0, "/root/repo")
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
import os
for path in ([int(os.environ["ONLY_PATH"])] if "ONLY_PATH" in os.environ else (1, 2, 3)):
    _lib.check(L.ocrpp_set_tuning(0, path))
    op = build_post_process({"name": "DBPostProcess", "thresh": q, "box_thresh": q + 0.02, "unclip_ratio": 1.7, "cpp_speedup": True, "cuda_speedup": True})
    ref = None
    bad = 0
    for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 300):
        boxes, scores, counts, status, ex = op.run_device(maps, sl, boxes_f=True, labels=(it % 2 == 0))
        cur = (boxes.copy(), scores.copy(), counts.copy())
        for n in range(N):
            cur[0][n, counts[n]:] = 0; cur[1][n, counts[n]:] = 0
        if ref is None:
            ref = cur
            print("path", path, "counts", counts)
        elif not all(np.array_equal(a, b) for a, b in zip(ref, cur)):
            bad += 1
            if bad <= 3:
                for n in range(N):
                    if counts[n] != ref[2][n]:
                        print("  iter", it, "image", n, "count", counts[n], "vs", ref[2][n])
                    else:
                        d = np.nonzero((cur[0][n] != ref[0][n]).any((1, 2)) | (cur[1][n] != ref[1][n]))[0]
                        if len(d):
                            print("  iter", it, "image", n, "boxes differ at", d[:5], cur[0][n][d[0]].tolist(), ref[0][n][d[0]].tolist(), cur[1][n][d[0]], ref[1][n][d[0]])
    print("path", path, "differing runs:", bad)
