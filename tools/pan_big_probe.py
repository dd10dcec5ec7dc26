"""Development aid (CPU, cv2): sizes of the multi-kernel text components of a bench rank's PAN pages - the work items of
ex_expand_kernel. A component whose padded bounding box exceeds the largest shared-memory tile (53 248 px) takes the
global-memory path; a single one of those is the tail of that rank's step."""
import sys; sys.path.insert(0, "/root/repo")
import numpy as np, cv2
import bench
rank = int(sys.argv[1]) if len(sys.argv) > 1 else 1
npages = int(sys.argv[2]) if len(sys.argv) > 2 else 128
big, tot = [], 0
for i in range(npages):
    text, kern, inst, C = bench._gen_pan_scene(bench.SEED + rank * 128 + i)
    nt, lab, st, _ = cv2.connectedComponentsWithStats((text > 0).astype(np.uint8), connectivity=4)
    kl = lab * (kern > 0)
    for t in range(1, nt):
        x, y, w, h, a = st[t]
        nk = cv2.connectedComponents(((kl[y:y + h, x:x + w] == t)).astype(np.uint8), connectivity=4)[0] - 1
        if nk >= 2:
            tot += 1
            tile = (w + 2) * (h + 2)
            if tile > 18432:
                big.append((i, tile, int(w), int(h), int(a), nk))
print("rank", rank, "multi-kernel components:", tot, "| tiles > 18 432 px:", len(big), "| > 53 248 px (global path):", sum(b[1] > 53248 for b in big))
print(sorted(big, key=lambda b: -b[1])[:12])
