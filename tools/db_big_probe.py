"""Development aid: which candidates of a bench rank's pages take the generic geometry path (rows, hull size)."""
import sys, os; sys.path.insert(0, "/root/repo")
import numpy as np, torch, cv2
import bench
rank = int(sys.argv[1]) if len(sys.argv) > 1 else 3
maps = np.stack([bench._gen_db(bench.SEED + rank * 256 + i) for i in range(256)])
tall = []
for n, m in enumerate(maps):
    seg = (m > 0.3).astype(np.uint8)
    cs, _ = cv2.findContours(seg, cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)
    for c in cs:
        x, y, w, h = cv2.boundingRect(c)
        if h > 64:
            tall.append((n, h, w, len(c)))
print("pages with a contour taller than 64 rows:", len(set(t[0] for t in tall)), "candidates:", len(tall))
print(sorted(tall, key=lambda t: -t[1])[:20])
