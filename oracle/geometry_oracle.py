"""TEST INFRASTRUCTURE ONLY (oracle). Pure-Python / numpy restatements of the small geometric pieces:

  * do_offset      - ClipperOffset::DoOffset for ONE closed polygon, jtRound
                     (R/pytocr/postprocess/db_postprocess_fast/src/clipper.cpp:3837-3879 AddPath,
                      :3889-3913 FixOrientations, :3987-4081 DoOffset, :4160-4201 OffsetPoint,
                      :4225-4244 DoRound, :136-140 Round, :393-411 Area/Orientation, :55 arc tolerance).
                     ClipperOffset::Execute then unions the result (:3915-3943); for the convex
                     quads the DB path feeds it the union only drops redundant vertices, so the
                     convex hull - and hence cv::minAreaRect - is unchanged. Pinned against the
                     reference's real Clipper in tests/test_oracle_geometry.py.
  * convex_hull / min_area_rect - fp64 "smallest bounding rectangle over all hull edges", the
                     spec of what the CUDA path computes in place of the third-party
                     cv::minAreaRect + cv::boxPoints (float32 rotating calipers). Pinned against
                     cv2 4.13 in tests/test_oracle_geometry.py (<= 1e-3 px, equal-area ties counted).
"""
import math

import numpy as np

PI = 3.141592653589793238
TWO_PI = PI * 2
DEF_ARC_TOLERANCE = 0.25


def clipper_round(v):
    """clipper.cpp:136-140: cast toward zero of v -/+ 0.5."""
    return int(v - 0.5) if v < 0 else int(v + 0.5)


def _area(poly):
    """clipper.cpp:393-411."""
    n = len(poly)
    if n < 3:
        return 0.0
    a = 0.0
    j = n - 1
    for i in range(n):
        a += (float(poly[j][0]) + poly[i][0]) * (float(poly[j][1]) - poly[i][1])
        j = i
    return -a * 0.5


def _unit_normal(p1, p2):
    """clipper.cpp:3797-3808."""
    if p1[0] == p2[0] and p1[1] == p2[1]:
        return (0.0, 0.0)
    dx = float(p2[0] - p1[0])
    dy = float(p2[1] - p1[1])
    f = 1 * 1.0 / math.sqrt(dx * dx + dy * dy)
    dx *= f
    dy *= f
    return (dy, -dx)


def tied_min_area_rects(points, rel=1e-9, corners=False):
    """All edge-aligned minimum-area rectangles of an integer point set, as cv2 RotatedRect tuples
    ((cx, cy), (w, h), angle_deg), one per hull-edge direction (mod 90 degrees) whose rectangle area equals the
    minimum (exact integer projections, so `equal` is decided exactly). cv2.minAreaRect returns ONE of them,
    chosen by float32 noise in its rotating calipers; the parity tests accept any. `rel` widens
    "equal" to near-ties below cv2's float32 resolution; corners=True returns the exact float64 corner arrays
    [4,2] instead (no float32 boxPoints noise)."""
    import cv2
    pts = np.asarray(points, np.int64).reshape(-1, 2)
    h = cv2.convexHull(pts.astype(np.int32)).reshape(-1, 2).astype(np.int64)
    m = len(h)
    if m < 3:
        return []
    cands = []
    for i in range(m):
        d = h[(i + 1) % m] - h[i]
        l2 = int(d @ d)
        if l2 == 0:
            continue
        s = (h - h[i]) @ d
        t = (h - h[i]) @ np.array([-d[1], d[0]])
        num = int(s.max() - s.min()) * int(t.max() - t.min())          # area * l2, exact
        cands.append((num, l2, i, d, s, t))
    best = min(c[0] / c[1] for c in cands)
    out, seen = [], []
    for num, l2, i, d, s, t in cands:
        if num / l2 > best * (1 + rel) + 1e-12:
            continue
        ang = np.degrees(np.arctan2(float(d[1]), float(d[0]))) % 90.0
        if any(min(abs(ang - a), 90 - abs(ang - a)) < 1e-7 for a in seen):
            continue
        seen.append(ang)
        ln = np.sqrt(l2)
        u = d / ln
        v = np.array([-d[1], d[0]]) / ln
        sc, tc = (s.max() + s.min()) / 2.0 / ln, (t.max() + t.min()) / 2.0 / ln
        if corners:
            s0, s1, t0, t1 = s.min() / ln, s.max() / ln, t.min() / ln, t.max() / ln
            out.append(np.array([h[i] + u * a + v * b for a, b in ((s0, t0), (s1, t0), (s1, t1), (s0, t1))]))
            continue
        c = h[i] + u * sc + v * tc
        out.append(((float(c[0]), float(c[1])), (float((s.max() - s.min()) / ln), float((t.max() - t.min()) / ln)),
                    float(np.degrees(np.arctan2(float(d[1]), float(d[0]))))))
    return out


def do_offset(path, delta, arc_tolerance=0.25):
    """Raw m_destPoly of DoOffset for one etClosedPolygon/jtRound path (list of int (x,y)).
    Returns a list of int (x,y); empty when AddPath rejects the path."""
    path = [(int(p[0]), int(p[1])) for p in path]
    high = len(path) - 1
    if high < 0:
        return []
    while high > 0 and path[0] == path[high]:      # :3845-3846
        high -= 1
    contour = [path[0]]
    for i in range(1, high + 1):                   # :3850-3858
        if contour[-1] != path[i]:
            contour.append(path[i])
    if len(contour) - 1 < 2:                       # :3859-3863 (j < 2)
        return []
    if not (_area(contour) >= 0):                  # FixOrientations :3889-3903 (single path => it is the lowest)
        contour = contour[::-1]
    if -1e-20 < delta < 1e-20:                     # NEAR_ZERO :3993-4003
        return list(contour)
    ad = abs(delta)
    if arc_tolerance <= 0.0:                       # :4009-4013
        y = DEF_ARC_TOLERANCE
    elif arc_tolerance > ad * DEF_ARC_TOLERANCE:
        y = ad * DEF_ARC_TOLERANCE
    else:
        y = arc_tolerance
    steps = PI / math.acos(1 - y / ad)             # :4015
    if steps > ad * PI:
        steps = ad * PI
    m_sin = math.sin(TWO_PI / steps)
    m_cos = math.cos(TWO_PI / steps)
    steps_per_rad = steps / TWO_PI
    if delta < 0.0:
        m_sin = -m_sin
    n = len(contour)
    if delta <= 0 and n < 3:
        return []
    normals = [_unit_normal(contour[j], contour[(j + 1) % n]) for j in range(n)]  # :4066-4071
    dest = []
    k = n - 1
    for j in range(n):                              # :4075-4079 + OffsetPoint :4160-4201
        sin_a = normals[k][0] * normals[j][1] - normals[j][0] * normals[k][1]
        if abs(sin_a * delta) < 1.0:
            cos_a = normals[k][0] * normals[j][0] + normals[j][1] * normals[k][1]
            if cos_a > 0:
                dest.append((clipper_round(contour[j][0] + normals[k][0] * delta),
                             clipper_round(contour[j][1] + normals[k][1] * delta)))
                continue                            # NOTE: returns before `k = j` (:4172)
        elif sin_a > 1.0:
            sin_a = 1.0
        elif sin_a < -1.0:
            sin_a = -1.0
        if sin_a * delta < 0:
            dest.append((clipper_round(contour[j][0] + normals[k][0] * delta),
                         clipper_round(contour[j][1] + normals[k][1] * delta)))
            dest.append(contour[j])
            dest.append((clipper_round(contour[j][0] + normals[j][0] * delta),
                         clipper_round(contour[j][1] + normals[j][1] * delta)))
        else:                                       # DoRound :4225-4244
            a = math.atan2(sin_a, normals[k][0] * normals[j][0] + normals[k][1] * normals[j][1])
            nsteps = max(int(clipper_round(steps_per_rad * abs(a))), 1)
            X, Y = normals[k]
            for _ in range(nsteps):
                dest.append((clipper_round(contour[j][0] + X * delta),
                             clipper_round(contour[j][1] + Y * delta)))
                X2 = X
                X = X * m_cos - m_sin * Y
                Y = X2 * m_sin + Y * m_cos
            dest.append((clipper_round(contour[j][0] + normals[j][0] * delta),
                         clipper_round(contour[j][1] + normals[j][1] * delta)))
        k = j
    return dest


def convex_hull(points):
    """Andrew monotone chain on (x, y); returns hull vertices counter-clockwise in the
    mathematical sense (x right, y up), without collinear points. points: iterable of (x,y)."""
    pts = sorted(set((float(p[0]), float(p[1])) for p in points))
    if len(pts) <= 2:
        return pts

    def cross(o, a, b):
        return (a[0] - o[0]) * (b[1] - o[1]) - (a[1] - o[1]) * (b[0] - o[0])

    lower = []
    for p in pts:
        while len(lower) >= 2 and cross(lower[-2], lower[-1], p) <= 0:
            lower.pop()
        lower.append(p)
    upper = []
    for p in reversed(pts):
        while len(upper) >= 2 and cross(upper[-2], upper[-1], p) <= 0:
            upper.pop()
        upper.append(p)
    return lower[:-1] + upper[:-1]


def min_area_rect(points):
    """fp64 minimum-area enclosing rectangle of a point set.
    Returns (corners[4,2] float64, (w, h)) with w measured along the chosen hull edge.
    Degenerate sets: 1 point -> four equal corners, size (0,0); collinear -> w = length, h = 0."""
    hull = convex_hull(points)
    n = len(hull)
    if n == 0:
        return np.zeros((4, 2)), (0.0, 0.0)
    if n == 1:
        return np.array([hull[0]] * 4, np.float64), (0.0, 0.0)
    H = np.array(hull, np.float64)
    best = None

    def folded(d):
        qx, qy = int(round(d[0])), int(round(d[1]))
        for _ in range(3):
            if qx > 0 and qy >= 0:
                break
            qx, qy = qy, -qx
        return qx, qy

    def better(a, b):
        """a, b = (area, qx, qy, index). Smaller area; exact tie -> larger edge angle modulo 90 degrees
        (cv::minAreaRect keeps the LAST minimum of a quarter-turn caliper sweep); then smaller index."""
        if abs(a[0] - b[0]) > 1e-12 * max(a[0], b[0]):
            return a[0] < b[0]
        l, r = a[2] * b[1], b[2] * a[1]
        if l != r:
            return l > r
        return a[3] < b[3]

    for i in range(n):
        p, q = H[i], H[(i + 1) % n]
        d = q - p
        ln = math.hypot(d[0], d[1])
        u = d / ln
        v = np.array([-u[1], u[0]])
        s = (H - p) @ u
        t = (H - p) @ v
        smin, smax, tmin, tmax = s.min(), s.max(), t.min(), t.max()
        area = (smax - smin) * (tmax - tmin)
        qx, qy = folded(d)
        key = (area, qx, qy, i)
        if best is None or better(key, best[0]):
            best = (key, p, u, v, smin, smax, tmin, tmax)
        if n == 2:
            break
    _, p, u, v, smin, smax, tmin, tmax = best
    corners = np.array([p + u * smin + v * tmin, p + u * smax + v * tmin,
                        p + u * smax + v * tmax, p + u * smin + v * tmax])
    return corners, (smax - smin, tmax - tmin)
