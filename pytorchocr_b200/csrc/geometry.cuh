// Per-candidate geometry shared by the detection operators: convex hull of integer points,
// minimum-area enclosing rectangle, GetMiniBoxes ordering, DB unclip (Clipper DoOffset for one
// closed quad with round joins) and the final rescale/round/clamp.
//
// Everything here is a plain function of its arguments (no global state) and is marked
// OCRPP_HD so that tests/host_shim can compile the SAME code with g++ and compare it with the
// Python oracle on the CPU box; the product only ever calls it from device code.
//
// Reference being replaced:
//   cv::minAreaRect + cv::boxPoints (third-party float32 rotating calipers) as called at
//     R/pytocr/postprocess/db_postprocess_fast/src/db_postprocess.cpp:61,164,259 and
//     R/pytocr/postprocess/pse_postprocess.py:85-86 -> min_area_rect() below, exact integer
//     projections + fp64, "smallest bounding rectangle over all hull edges"
//   DBPostProcessor::GetMiniBoxes  db_postprocess.cpp:159-192
//   DBPostProcessor::GetContourArea db_postprocess.cpp:16-32 (float32, source op order, no FMA)
//   DBPostProcessor::UnClip         db_postprocess.cpp:34-64
//   ClipperOffset AddPath/FixOrientations/DoOffset/OffsetPoint/DoRound
//     R/.../src/clipper.cpp:3837-3879, 3889-3913, 3987-4081, 4160-4201, 4225-4244, Round :136-140
//   order_points_clockwise R/pytocr/utils/utility.py:21-29
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define OCRPP_HD __host__ __device__ __forceinline__
#else
#define OCRPP_HD inline
#endif

namespace ocrpp {
namespace geom {

struct P2i {
  int x, y;
};

struct Rect {        // result of min_area_rect
  double cx[4], cy[4];  // corners p+u*smin+v*tmin, p+u*smax+v*tmin, p+u*smax+v*tmax, p+u*smin+v*tmax
  double w, h;          // w along the chosen hull edge, h across it
};

// float32 arithmetic exactly as written (no FMA contraction) on both host and device
OCRPP_HD float fmul(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);
#else
  volatile float r = a * b;
  return r;
#endif
}
OCRPP_HD float fadd(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  volatile float r = a + b;
  return r;
#endif
}
OCRPP_HD float fsub(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fsub_rn(a, b);
#else
  volatile float r = a - b;
  return r;
#endif
}
OCRPP_HD float fdiv(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fdiv_rn(a, b);
#else
  volatile float r = a / b;
  return r;
#endif
}
// float64 likewise: the reference's Clipper is plain x86-64 code (every product and sum rounded on its own). A
// contracted a*b+c*d leaves a residual of one rounding error, which is enough to flip Clipper's `cosA > 0` test
// for an exactly right-angled corner (found by tests/stress_gpu.py on a thin 45-degree box).
OCRPP_HD double dmul(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b);
#else
  volatile double r = a * b;
  return r;
#endif
}
OCRPP_HD double dadd(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dadd_rn(a, b);
#else
  volatile double r = a + b;
  return r;
#endif
}
OCRPP_HD double dsub(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dsub_rn(a, b);
#else
  volatile double r = a - b;
  return r;
#endif
}
OCRPP_HD float fsqrt(float a) {
#if defined(__CUDA_ARCH__)
  return __fsqrt_rn(a);
#else
  return sqrtf(a);
#endif
}

OCRPP_HD long long cross(const P2i& o, const P2i& a, const P2i& b) {
  return (long long)(a.x - o.x) * (b.y - o.y) - (long long)(a.y - o.y) * (b.x - o.x);
}

// Andrew monotone chain over points that are ALREADY sorted lexicographically by (y, x)
// (row extents come out of the labelling in that order). `pts[0..n)` in, hull out in `out`
// (capacity >= n + 1), strictly convex (collinear points dropped). Returns the hull size.
// `out` may not alias `pts`.
OCRPP_HD int hull_sorted(const P2i* pts, int n, P2i* out) {
  if (n <= 1) {
    if (n == 1) out[0] = pts[0];
    return n;
  }
  int k = 0;
  for (int i = 0; i < n; ++i) {
    if (i > 0 && pts[i].x == pts[i - 1].x && pts[i].y == pts[i - 1].y) continue;
    while (k >= 2 && cross(out[k - 2], out[k - 1], pts[i]) <= 0) --k;
    out[k++] = pts[i];
  }
  if (k == 1) return 1;
  const int lo = k + 1;
  for (int i = n - 2; i >= 0; --i) {
    if (pts[i].x == pts[i + 1].x && pts[i].y == pts[i + 1].y) continue;
    while (k >= lo && cross(out[k - 2], out[k - 1], pts[i]) <= 0) --k;
    out[k++] = pts[i];
  }
  return k - 1;  // last point == first point
}

// Candidate rectangle flush with hull edge i; exact integer projections.
// Returns area*len^2 pieces so that callers can compare areas in double.
struct EdgeFit {
  long long smin, smax, tmin, tmax;
  long long len2;
  long long qx, qy;  // edge direction folded by quarter turns into qx > 0, qy >= 0
  double area;
};

// Which of two edge-aligned rectangles wins. Smaller area; on an exact tie the edge whose
// direction has the larger angle modulo 90 degrees (cv::minAreaRect's rotating calipers sweep a
// quarter turn starting axis-aligned and keep the LAST minimum, so e.g. a diagonal rectangle beats
// an equal-area axis-aligned one); then the smaller edge index.
OCRPP_HD bool fit_better(const EdgeFit& a, int ia, const EdgeFit& b, int ib) {
  const double m = a.area > b.area ? a.area : b.area;
  const double d = a.area - b.area;
  if ((d < 0 ? -d : d) > 1e-12 * m) return a.area < b.area;
  const long long l = a.qy * b.qx, r = b.qy * a.qx;
  if (l != r) return l > r;
  return ia < ib;
}

OCRPP_HD EdgeFit fit_edge(const P2i* h, int n, int i) {
  const P2i p = h[i], q = h[(i + 1 == n) ? 0 : i + 1];
  const long long dx = q.x - p.x, dy = q.y - p.y;
  EdgeFit f;
  f.smin = f.tmin = 0x7fffffffffffffffLL;
  f.smax = f.tmax = -0x7fffffffffffffffLL;
  for (int j = 0; j < n; ++j) {
    const long long vx = h[j].x - p.x, vy = h[j].y - p.y;
    const long long s = vx * dx + vy * dy, t = vy * dx - vx * dy;
    f.smin = s < f.smin ? s : f.smin;
    f.smax = s > f.smax ? s : f.smax;
    f.tmin = t < f.tmin ? t : f.tmin;
    f.tmax = t > f.tmax ? t : f.tmax;
  }
  f.len2 = dx * dx + dy * dy;
  f.area = (double)(f.smax - f.smin) * (double)(f.tmax - f.tmin) / (double)f.len2;
  long long qx = dx, qy = dy;
  for (int t = 0; t < 3 && !(qx > 0 && qy >= 0); ++t) {  // rotate by -90 degrees: (x,y) -> (y,-x)
    const long long tx = qy;
    qy = -qx;
    qx = tx;
  }
  f.qx = qx;
  f.qy = qy;
  return f;
}

OCRPP_HD void rect_from_fit(const P2i* h, int n, int i, const EdgeFit& f, Rect* r) {
  const P2i p = h[i], q = h[(i + 1 == n) ? 0 : i + 1];
  const double dx = (double)(q.x - p.x), dy = (double)(q.y - p.y);
  const double il2 = 1.0 / (double)f.len2;
  const double s0 = (double)f.smin * il2, s1 = (double)f.smax * il2;
  const double t0 = (double)f.tmin * il2, t1 = (double)f.tmax * il2;
  // u = (dx,dy)/len, v = (-dy,dx)/len ; corner = p + u*(s/len) + v*(t/len)
  const double ss[4] = {s0, s1, s1, s0}, tt[4] = {t0, t0, t1, t1};
  for (int k = 0; k < 4; ++k) {
    r->cx[k] = (double)p.x + dx * ss[k] - dy * tt[k];
    r->cy[k] = (double)p.y + dy * ss[k] + dx * tt[k];
  }
  const double len = sqrt((double)f.len2);
  r->w = (double)(f.smax - f.smin) / len;
  r->h = (double)(f.tmax - f.tmin) / len;
}

// Sequential reference formulation (the device path parallelises the edge loop over a warp and
// must pick the same edge, see fit_better).
OCRPP_HD void min_area_rect(const P2i* h, int n, Rect* r) {
  if (n == 1) {
    for (int k = 0; k < 4; ++k) {
      r->cx[k] = h[0].x;
      r->cy[k] = h[0].y;
    }
    r->w = r->h = 0.0;
    return;
  }
  int best = 0;
  EdgeFit bf = fit_edge(h, n, 0);
  const int ne = n == 2 ? 1 : n;
  for (int i = 1; i < ne; ++i) {
    EdgeFit f = fit_edge(h, n, i);
    if (fit_better(f, i, bf, best)) {
      bf = f;
      best = i;
    }
  }
  rect_from_fit(h, n, best, bf, r);
}

// cv::boxPoints corner ORDER for a rectangle found by cv::minAreaRect (angle in [-90, 0)): index 0 is
// the lowest corner (largest y; the right one when two share it, i.e. axis-aligned), then clockwise
// on screen: left-most, top-most, right-most. Rect's own corner sequence is already screen-clockwise,
// so this is a rotation of the indices. The order only matters where a later argmin/argmax or stable
// sort over the corners meets an exact tie (45-degree "diamonds": order_points_clockwise then repeats
// a corner exactly as the reference does).
OCRPP_HD void cv_box_order(const Rect& r, double* ox, double* oy) {
  int i0 = 0;
  for (int i = 1; i < 4; ++i) {
    const double dy = r.cy[i] - r.cy[i0];
    if (dy > 1e-7 || (dy >= -1e-7 && r.cx[i] > r.cx[i0])) i0 = i;
  }
  for (int k = 0; k < 4; ++k) {
    ox[k] = r.cx[(i0 + k) & 3];
    oy[k] = r.cy[(i0 + k) & 3];
  }
}

// GetMiniBoxes ordering (db_postprocess.cpp:165-190): stable sort by x, then
// [TL, TR, BR, BL] from the y order inside the left pair and the right pair.
OCRPP_HD void mini_box(const float* cx, const float* cy, float* ox, float* oy) {
  int idx[4] = {0, 1, 2, 3};
  for (int i = 1; i < 4; ++i) {  // stable insertion sort on x
    const int v = idx[i];
    int j = i - 1;
    while (j >= 0 && cx[idx[j]] > cx[v]) {
      idx[j + 1] = idx[j];
      --j;
    }
    idx[j + 1] = v;
  }
  int i1, i2, i3, i4;
  if (cy[idx[3]] <= cy[idx[2]]) {
    i2 = idx[3];
    i3 = idx[2];
  } else {
    i2 = idx[2];
    i3 = idx[3];
  }
  if (cy[idx[1]] <= cy[idx[0]]) {
    i1 = idx[1];
    i4 = idx[0];
  } else {
    i1 = idx[0];
    i4 = idx[1];
  }
  ox[0] = cx[i1]; oy[0] = cy[i1];
  ox[1] = cx[i2]; oy[1] = cy[i2];
  ox[2] = cx[i3]; oy[2] = cy[i3];
  ox[3] = cx[i4]; oy[3] = cy[i4];
}

// order_points_clockwise (utility.py:21-29): [argmin(x+y), argmin(y-x), argmax(x+y), argmax(y-x)],
// first index wins ties (numpy argmin/argmax), float32 sums/differences.
OCRPP_HD void order_points_clockwise(const float* cx, const float* cy, float* ox, float* oy) {
  int a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  float s[4], d[4];
  for (int i = 0; i < 4; ++i) {
    s[i] = fadd(cx[i], cy[i]);
    d[i] = fsub(cy[i], cx[i]);
  }
  for (int i = 1; i < 4; ++i) {
    if (s[i] < s[a0]) a0 = i;
    if (s[i] > s[a2]) a2 = i;
    if (d[i] < d[a1]) a1 = i;
    if (d[i] > d[a3]) a3 = i;
  }
  ox[0] = cx[a0]; oy[0] = cy[a0];
  ox[1] = cx[a1]; oy[1] = cy[a1];
  ox[2] = cx[a2]; oy[2] = cy[a2];
  ox[3] = cx[a3]; oy[3] = cy[a3];
}

// GetContourArea (db_postprocess.cpp:16-32): float32, source order.
OCRPP_HD float unclip_distance(const float* bx, const float* by, float unclip_ratio) {
  float area = 0.0f, dist = 0.0f;
  for (int i = 0; i < 4; ++i) {
    const int j = (i + 1) & 3;
    area = fadd(area, fsub(fmul(bx[i], by[j]), fmul(by[i], bx[j])));
    const float dx = fsub(bx[i], bx[j]), dy = fsub(by[i], by[j]);
    dist = fadd(dist, fsqrt(fadd(fmul(dx, dx), fmul(dy, dy))));
  }
  area = fabsf((float)((double)area / 2.0));
  return fdiv(fmul(area, unclip_ratio), dist);
}

OCRPP_HD long long clipper_round(double v) { return v < 0 ? (long long)(v - 0.5) : (long long)(v + 0.5); }

// ClipperOffset for ONE closed polygon of 4 integer points, jtRound, ArcTolerance 0.25.
// Writes the raw m_destPoly into out[0..cap) and returns the number of points, 0 when AddPath
// rejects the path (< 3 distinct points), or -1 when cap is too small.
OCRPP_HD int do_offset_quad(const P2i* quad, double delta, P2i* out, int cap) {
  const double kPi = 3.141592653589793238, kTwoPi = kPi * 2;
  P2i c[4];
  int high = 3;
  while (high > 0 && quad[0].x == quad[high].x && quad[0].y == quad[high].y) --high;  // :3845-3846
  int n = 1;
  c[0] = quad[0];
  for (int i = 1; i <= high; ++i)  // :3850-3858
    if (c[n - 1].x != quad[i].x || c[n - 1].y != quad[i].y) c[n++] = quad[i];
  if (n - 1 < 2) return 0;  // :3859-3863
  {  // Orientation / FixOrientations :393-411, :3889-3903
    double a = 0;
    for (int i = 0, j = n - 1; i < n; ++i) {
      a += ((double)c[j].x + c[i].x) * ((double)c[j].y - c[i].y);
      j = i;
    }
    if (!(-a * 0.5 >= 0)) {
      for (int i = 0, j = n - 1; i < j; ++i, --j) {
        const P2i t = c[i];
        c[i] = c[j];
        c[j] = t;
      }
    }
  }
  if (delta > -1e-20 && delta < 1e-20) {  // NEAR_ZERO :3993-4003
    if (n > cap) return -1;
    for (int i = 0; i < n; ++i) out[i] = c[i];
    return n;
  }
  const double ad = fabs(delta);
  double y = 0.25;  // ArcTolerance = 0.25 (ctor default) ; :4009-4013
  if (0.25 > ad * 0.25) y = ad * 0.25;
  double steps = kPi / acos(1 - y / ad);  // :4015
  if (steps > ad * kPi) steps = ad * kPi;
  double m_sin = sin(kTwoPi / steps);
  const double m_cos = cos(kTwoPi / steps);
  const double steps_per_rad = steps / kTwoPi;
  if (delta < 0.0) m_sin = -m_sin;
  if (delta <= 0 && n < 3) return 0;
  double nx[4], ny[4];
  for (int j = 0; j < n; ++j) {  // GetUnitNormal :3797-3808
    const P2i p1 = c[j], p2 = c[(j + 1 == n) ? 0 : j + 1];
    double dx = (double)(p2.x - p1.x), dy = (double)(p2.y - p1.y);
    const double f = 1 * 1.0 / sqrt(dx * dx + dy * dy);
    dx *= f;
    dy *= f;
    nx[j] = dy;
    ny[j] = -dx;
  }
  int m = 0;
  int k = n - 1;
#define OCRPP_OFS(C, N) clipper_round(dadd((double)(C), dmul((N), delta)))
#define OCRPP_EMIT(X, Y)                                  \
  do {                                                    \
    if (m >= cap) return -1;                              \
    out[m].x = (int)(X);                                  \
    out[m].y = (int)(Y);                                  \
    ++m;                                                  \
  } while (0)
  for (int j = 0; j < n; ++j) {  // OffsetPoint :4160-4201
    double sin_a = dsub(dmul(nx[k], ny[j]), dmul(nx[j], ny[k]));
    if (fabs(sin_a * delta) < 1.0) {
      const double cos_a = dadd(dmul(nx[k], nx[j]), dmul(ny[j], ny[k]));
      if (cos_a > 0) {
        OCRPP_EMIT(OCRPP_OFS(c[j].x, nx[k]), OCRPP_OFS(c[j].y, ny[k]));
        continue;  // returns before `k = j` (:4172)
      }
    } else if (sin_a > 1.0) sin_a = 1.0;
    else if (sin_a < -1.0) sin_a = -1.0;
    if (sin_a * delta < 0) {
      OCRPP_EMIT(OCRPP_OFS(c[j].x, nx[k]), OCRPP_OFS(c[j].y, ny[k]));
      OCRPP_EMIT(c[j].x, c[j].y);
      OCRPP_EMIT(OCRPP_OFS(c[j].x, nx[j]), OCRPP_OFS(c[j].y, ny[j]));
    } else {  // DoRound :4225-4244
      const double a = atan2(sin_a, dadd(dmul(nx[k], nx[j]), dmul(ny[k], ny[j])));
      long long ns = clipper_round(dmul(steps_per_rad, fabs(a)));
      if (ns < 1) ns = 1;
      double X = nx[k], Y = ny[k];
      for (long long i = 0; i < ns; ++i) {
        OCRPP_EMIT(OCRPP_OFS(c[j].x, X), OCRPP_OFS(c[j].y, Y));
        const double X2 = X;
        X = dsub(dmul(X, m_cos), dmul(m_sin, Y));
        Y = dadd(dmul(X2, m_sin), dmul(Y, m_cos));
      }
      OCRPP_EMIT(OCRPP_OFS(c[j].x, nx[j]), OCRPP_OFS(c[j].y, ny[j]));
    }
    k = j;
  }
#undef OCRPP_EMIT
#undef OCRPP_OFS
  return m;
}

// In-place lexicographic (y, x) insertion sort; offset polygons have a few dozen points.
OCRPP_HD void sort_points_yx(P2i* p, int n) {
  for (int i = 1; i < n; ++i) {
    const P2i v = p[i];
    int j = i - 1;
    while (j >= 0 && (p[j].y > v.y || (p[j].y == v.y && p[j].x > v.x))) {
      p[j + 1] = p[j];
      --j;
    }
    p[j + 1] = v;
  }
}

OCRPP_HD float roundf_half_away(float v) { return roundf(v); }  // C roundf: half away from zero

// Final coordinates of one unclipped corner (db_postprocess.cpp:291-311). Default: scale by src/map size in
// float32. use_padding_resize: the inverse of the "pad to a square, then resize" affine map
// (get_affine_transform(center, max(src_w, src_h), map_height, inv=1) + transform_preds, :111-145,293-302):
// cv::getAffineTransform on those three point pairs is, in exact arithmetic, the similarity
//   x' = s*x + max(cx - cy, 0) - ... i.e.  x' = s*x + (cx >= cy ? 0 : cx - cy),  y' = s*y + (cx >= cy ? cy - cx : 0)
// with s = max(src_w, src_h) / map_height and (cx, cy) = (src_w/2, src_h/2); it is evaluated in double like the
// reference's matrix product and cast to float32 (cv::Point2f) before roundf.
OCRPP_HD void db_rescale(float mx, float my, int W, int H, float sw, float sh, int use_padding_resize,
                         float* fx, float* fy) {
  if (!use_padding_resize) {
    *fx = fmul(fdiv(mx, (float)W), sw);
    *fy = fmul(fdiv(my, (float)H), sh);
    return;
  }
  const double cx = (double)(float)((double)sw / 2.0), cy = (double)(float)((double)sh / 2.0);
  const double m = sw > sh ? (double)sw : (double)sh;
  const double s = m / (double)H;
  const double tx = cx >= cy ? 0.0 : cx - cy, ty = cx >= cy ? cy - cx : 0.0;
  *fx = (float)(s * (double)mx + tx);
  *fy = (float)(s * (double)my + ty);
}

// np.round (half to even) on a double, as used by PSE/PAN generate_box (pse_postprocess.py:100-101)
OCRPP_HD double round_half_even(double v) {
  return nearbyint(v);
}

}  // namespace geom
}  // namespace ocrpp
