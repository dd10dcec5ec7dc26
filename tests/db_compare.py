"""Shared comparison logic for DB parity tests (GPU result vs the cv2-based oracle)."""
import numpy as np


def canonical_labels(lab):
    """Renumber labels by the raster order of each label's first pixel (0 stays 0)."""
    flat = lab.ravel()
    idx = np.nonzero(flat)[0]
    if idx.size == 0:
        return lab.copy()
    labs = flat[idx]
    first = np.full(int(flat.max()) + 1, -1, np.int64)
    first[labs[::-1]] = idx[::-1]
    present = np.nonzero(first >= 0)[0]
    order = present[np.argsort(first[present])]
    remap = np.zeros(int(flat.max()) + 1, np.int64)
    remap[order] = np.arange(1, len(order) + 1)
    return remap[lab]


def compare_image(boxes, boxes_f, scores, details, tol_px=1e-3, tol_score=1e-5):
    """boxes int16 [K,4,2], boxes_f float32 [K,4,2], scores [K] from the GPU; details from
    oracle.db_oracle.boxes_from_bitmap(return_details=True). Returns the number of 'fragile'
    boxes (oracle corner within float noise of an int()-truncation boundary before unclip)."""
    ok = [d for d in details if d["status"] == "ok"]
    assert len(ok) == len(boxes), (len(ok), len(boxes))
    if not ok:
        return 0
    of = np.array([d["out_f"] for d in ok], np.float64)           # [K,4,2]
    gf = np.asarray(boxes_f, np.float64)
    dist = np.abs(of[:, None] - gf[None]).reshape(len(ok), len(ok), -1).max(-1)
    fragile = 0
    used = set()
    for i in np.argsort(dist.min(1)):
        j = int(np.argmin([dist[i, j] if j not in used else np.inf for j in range(len(ok))]))
        used.add(j)
        d = ok[i]
        assert abs(scores[j] - d["score"]) <= tol_score * abs(d["score"]) + 1e-7, (scores[j], d["score"])
        if dist[i, j] < tol_px:
            stable = np.abs(of[i] - np.floor(of[i]) - 0.5) > 2e-3
            assert np.array_equal(np.asarray(d["out"])[stable], np.asarray(boxes[j], np.int64)[stable]), (d["out"], boxes[j])
        else:
            mini = np.asarray(d["mini"], np.float64)
            near_int = np.abs(mini - np.round(mini)).min() < 2e-3
            assert near_int and dist[i, j] <= 2.5, ("box mismatch", dist[i, j], d["out_f"], gf[j], mini)
            fragile += 1
    return fragile
