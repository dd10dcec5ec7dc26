"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU oracle for the text-line crop that follows the
detection path.

Follows R/pytocr/utils/utility.py:32-50 (`sort_boxes`), :53-78 (`get_part_img`) and the caller's rot90 rule
(R/deploy/pytorch/run_ocr.py:188-191). `get_part_img` is restated twice:

  * `get_part_img`            - the reference's own sequence of cv2 calls (getPerspectiveTransform +
                                warpPerspective, BORDER_REPLICATE, INTER_LINEAR); this is the checker.
  * `get_part_img_restated`   - the same result from first principles (numpy / Python floats), spelling out
                                what those two cv2 calls compute for 8-bit images. It documents the arithmetic
                                the CUDA kernel implements and is pinned bit-exactly against cv2 4.13 in
                                tests/test_oracle_crop.py (transform matrices compared bitwise too).

Parity: pinned against cv2 (above) and against the reference's own `get_part_img` / `sort_boxes` imported from
R/pytocr/utils/utility.py when tests/golden/make_golden.py generated tests/golden/reference_outputs.npz.

cv2 arithmetic restated (OpenCV 4.13 `imgproc`, 8-bit, INTER_LINEAR, fixed-point path):
  getPerspectiveTransform: 8x8 system [x y 1 0 0 0 -xu -yu; 0 0 0 x y 1 -xv -yv] c = [u; v] in float64, solved by
      LU with partial pivoting (first largest pivot, eps = 100*DBL_EPSILON), back substitution, c8 = 1.
  warpPerspective: M <- inverse of the 3x3 by cofactors * (1/det); destination processed in blocks of bw x bh
      (bh = min(16,h), bw = min(1024/bh, w), bh = min(1024/bw, h)); for the pixel x = xb + x1 of a block starting
      at xb:  X0 = M0*xb + M1*y + M2 (likewise Y0, W0);  W = W0 + M6*x1;  W = W ? 32/W : 0;
      X = rint(clamp((X0 + M0*x1)*W)) (ties to even), Y likewise; source cell (X>>5, Y>>5), 5-bit fractions
      (X&31, Y&31); bilinear taps weighted by round((1-fy)(1-fx)*32768) etc. (saturated to int16: a weight of
      1.0 is 32767), each tap's coordinates clamped to the image (BORDER_REPLICATE), result
      (sum + 16384) >> 15.
"""
import cv2
import numpy as np


def sort_boxes(dt_boxes):
    """utility.py:32-50: sort by (y, x) of the first corner, then ONE adjacent-swap pass for boxes whose first
    corners are within 10 px vertically. Returns a list of [4,2] arrays."""
    num_boxes = dt_boxes.shape[0]
    boxes = sorted(dt_boxes, key=lambda b: (b[0][1], b[0][0]))
    for i in range(num_boxes - 1):
        if abs(boxes[i + 1][0][1] - boxes[i][0][1]) < 10 and boxes[i + 1][0][0] < boxes[i][0][0]:
            boxes[i], boxes[i + 1] = boxes[i + 1], boxes[i]
    return boxes


def sort_order(dt_boxes):
    """Permutation form of `sort_boxes`: order[i] = index in dt_boxes of the i-th sorted box."""
    n = dt_boxes.shape[0]
    order = sorted(range(n), key=lambda k: (dt_boxes[k][0][1], dt_boxes[k][0][0]))   # stable, like sorted()
    for i in range(n - 1):
        a, b = dt_boxes[order[i]][0], dt_boxes[order[i + 1]][0]
        if abs(int(b[1]) - int(a[1])) < 10 and b[0] < a[0]:
            order[i], order[i + 1] = order[i + 1], order[i]
    return np.asarray(order, np.int32)


def crop_rect(pts):
    """utility.py:57-61: integer bounding box (left, top, right, bottom) of the box corners."""
    pts = np.asarray(pts).astype(np.float32)
    return int(np.min(pts[:, 0])), int(np.min(pts[:, 1])), int(np.max(pts[:, 0])), int(np.max(pts[:, 1]))


def get_part_img(img, pts):
    """utility.py:53-78, the reference's own cv2 calls."""
    pts = np.asarray(pts).astype(np.float32)
    left, top, right, bottom = crop_rect(pts)
    img_crop = img[top:bottom, left:right, :].copy()
    pts = pts - np.array([left, top], dtype=np.float32)
    w, h = int(right - left), int(bottom - top)
    dst = np.array([[0, 0], [w - 1, 0], [w - 1, h - 1], [0, h - 1]], dtype=np.float32)
    M = cv2.getPerspectiveTransform(pts, dst)
    return cv2.warpPerspective(img_crop, M, (w, h), borderMode=cv2.BORDER_REPLICATE, flags=cv2.INTER_LINEAR)


def crop_for_rec(img, box):
    """run_ocr.py:188-191: the crop, rotated counter-clockwise when it is at least 1.5x taller than wide."""
    part = get_part_img(img, box)
    if part.shape[0] >= 1.5 * part.shape[1]:
        part = np.rot90(part, 1)
    return part


# ------------------------------------------------------------------------------------------------
# first-principles restatement
# ------------------------------------------------------------------------------------------------
def perspective_transform(src, dst):
    """cv2.getPerspectiveTransform(src, dst) bit for bit (see module docstring). src, dst: [4,2] float32."""
    m = 8
    A = [[0.0] * m for _ in range(m)]
    b = [0.0] * m
    for i in range(4):
        sx, sy, dx, dy = float(src[i][0]), float(src[i][1]), float(dst[i][0]), float(dst[i][1])
        A[i][0] = A[i + 4][3] = sx
        A[i][1] = A[i + 4][4] = sy
        A[i][2] = A[i + 4][5] = 1.0
        A[i][6], A[i][7] = -sx * dx, -sy * dx
        A[i + 4][6], A[i + 4][7] = -sx * dy, -sy * dy
        b[i], b[i + 4] = dx, dy
    eps = 2.220446049250313e-16 * 100
    for i in range(m):
        k = i
        for j in range(i + 1, m):
            if abs(A[j][i]) > abs(A[k][i]):
                k = j
        if abs(A[k][i]) < eps:
            return None        # collinear corners: cv2 4.13 falls back to an SVD null-space solution (meaningless
                               # pixels); the CUDA path reports such a box as degenerate instead
        if k != i:
            A[i], A[k] = A[k], A[i]
            b[i], b[k] = b[k], b[i]
        d = -1 / A[i][i]
        for j in range(i + 1, m):
            alpha = A[j][i] * d
            for c in range(i + 1, m):
                A[j][c] += alpha * A[i][c]
            b[j] += alpha * b[i]
    for i in range(m - 1, -1, -1):
        s = b[i]
        for c in range(i + 1, m):
            s -= A[i][c] * b[c]
        b[i] = s / A[i][i]
    return np.array(b + [1.0]).reshape(3, 3)


def invert3(S):
    """cv2.invert of a 3x3 float64 matrix (cofactors times 1/det), bit for bit."""
    S = [[float(v) for v in r] for r in S]
    det = (S[0][0] * (S[1][1] * S[2][2] - S[1][2] * S[2][1]) - S[0][1] * (S[1][0] * S[2][2] - S[1][2] * S[2][0])
           + S[0][2] * (S[1][0] * S[2][1] - S[1][1] * S[2][0]))
    if det == 0:
        return None
    d = 1.0 / det
    t = [(S[1][1] * S[2][2] - S[1][2] * S[2][1]) * d, (S[0][2] * S[2][1] - S[0][1] * S[2][2]) * d,
         (S[0][1] * S[1][2] - S[0][2] * S[1][1]) * d, (S[1][2] * S[2][0] - S[1][0] * S[2][2]) * d,
         (S[0][0] * S[2][2] - S[0][2] * S[2][0]) * d, (S[0][2] * S[1][0] - S[0][0] * S[1][2]) * d,
         (S[1][0] * S[2][1] - S[1][1] * S[2][0]) * d, (S[0][1] * S[2][0] - S[0][0] * S[2][1]) * d,
         (S[0][0] * S[1][1] - S[0][1] * S[1][0]) * d]
    return np.array(t).reshape(3, 3)


def warp_block_width(w, h):
    bh = min(16, h)
    bw = min(1024 // bh, w)
    return bw


def warp_perspective_restated(src, Minv, w, h):
    """cv2.warpPerspective(src, M, (w,h), INTER_LINEAR, BORDER_REPLICATE) for uint8 [hs,ws,C], given inv(M)."""
    hs, ws = src.shape[:2]
    M = [float(v) for v in np.asarray(Minv).reshape(-1)]
    bw = warp_block_width(w, h)
    xs = np.arange(w)
    xb = ((xs // bw) * bw).astype(np.float64)[None, :]
    x1 = (xs % bw).astype(np.float64)[None, :]
    ys = np.arange(h, dtype=np.float64)[:, None]
    X0 = (M[0] * xb + M[1] * ys) + M[2]
    Y0 = (M[3] * xb + M[4] * ys) + M[5]
    W0 = (M[6] * xb + M[7] * ys) + M[8]
    W = W0 + M[6] * x1
    with np.errstate(divide="ignore"):
        W = np.where(W != 0, 32.0 / W, 0.0)
    lim = 2147483647.0
    X = np.rint(np.clip((X0 + M[0] * x1) * W, -lim - 1, lim)).astype(np.int64)
    Y = np.rint(np.clip((Y0 + M[3] * x1) * W, -lim - 1, lim)).astype(np.int64)
    sx, sy = np.clip(X >> 5, -32768, 32767), np.clip(Y >> 5, -32768, 32767)
    ax, ay = (X & 31).astype(np.float32) / np.float32(32), (Y & 31).astype(np.float32) / np.float32(32)
    one = np.float32(1)
    wts = [(one - ay) * (one - ax), (one - ay) * ax, ay * (one - ax), ay * ax]
    wts = [np.clip(np.rint(t * np.float32(32768)), -32768, 32767).astype(np.int64)[..., None] for t in wts]

    def tap(yy, xx):
        return src[np.clip(yy, 0, hs - 1), np.clip(xx, 0, ws - 1)].astype(np.int64)

    v = tap(sy, sx) * wts[0] + tap(sy, sx + 1) * wts[1] + tap(sy + 1, sx) * wts[2] + tap(sy + 1, sx + 1) * wts[3]
    return np.clip((v + (1 << 14)) >> 15, 0, 255).astype(np.uint8)


def get_part_img_restated(img, pts):
    """`get_part_img` without cv2 (same bits)."""
    pts = np.asarray(pts).astype(np.float32)
    left, top, right, bottom = crop_rect(pts)
    img_crop = img[top:bottom, left:right, :]
    pts = pts - np.array([left, top], dtype=np.float32)
    w, h = int(right - left), int(bottom - top)
    dst = np.array([[0, 0], [w - 1, 0], [w - 1, h - 1], [0, h - 1]], dtype=np.float32)
    M = perspective_transform(pts, dst)
    Mi = invert3(M) if M is not None else None
    if Mi is None:
        raise ValueError("degenerate box (collinear corners)")
    return warp_perspective_restated(img_crop, Mi, w, h)
