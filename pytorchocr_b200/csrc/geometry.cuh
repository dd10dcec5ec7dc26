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

// The reference's pure-Python branch (db_postprocess.py:143-146): shapely Polygon(box).area * unclip_ratio /
// .length in float64 - GEOS' Area::ofRingSigned (coordinates shifted by x0, sum of x_i * (y_{i-1} - y_{i+1}), halved)
// and Length::ofLine over the closed ring; every product and sum rounded on its own as in the x86 build.
OCRPP_HD double unclip_distance_py(const float* bx, const float* by, double unclip_ratio) {
  const double x0 = (double)bx[0];
  double sum = 0.0, len = 0.0;
  for (int i = 1; i <= 3; ++i) {
    const int n = (i + 1) & 3;   // ring point i+1 (point 4 is point 0 again)
    sum = dadd(sum, dmul(dsub((double)bx[i], x0), dsub((double)by[i - 1], (double)by[n])));
  }
  for (int i = 0; i < 4; ++i) {
    const int j = (i + 1) & 3;
    const double dx = dsub((double)bx[j], (double)bx[i]), dy = dsub((double)by[j], (double)by[i]);
    len = dadd(len, sqrt(dadd(dmul(dx, dx), dmul(dy, dy))));
  }
  const double area = fabs(sum / 2.0);
  return dmul(area, unclip_ratio) / len;
}

OCRPP_HD long long clipper_round(double v) { return v < 0 ? (long long)(v - 0.5) : (long long)(v + 0.5); }

// ClipperOffset for ONE closed polygon of 4 integer points, jtRound, ArcTolerance 0.25.
// Writes the raw m_destPoly into out[0..cap) and returns the number of points, 0 when AddPath
// rejects the path (< 3 distinct points), or -1 when cap is too small.
// (`Out` stores point i: P2iOut into a P2i array; PackedOut as x | y << 16 for coordinates within +-32767)
struct P2iOut {
  P2i* p;
  OCRPP_HD void set(int i, int x, int y) const { p[i].x = x; p[i].y = y; }
};
struct PackedOut {
  int* p;
  OCRPP_HD void set(int i, int x, int y) const { p[i] = (x & 0xffff) | (int)((unsigned)y << 16); }
};
template <typename Out>
OCRPP_HD int do_offset_quad_t(const P2i* quad, double delta, Out out, int cap) {
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
    for (int i = 0; i < n; ++i) out.set(i, c[i].x, c[i].y);
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
    out.set(m, (int)(X), (int)(Y));                       \
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
OCRPP_HD int do_offset_quad(const P2i* quad, double delta, P2i* out, int cap) {
  return do_offset_quad_t(quad, delta, P2iOut{out}, cap);
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

// The Python branch's rescale (db_postprocess.py:124-141): float32 `x / width * dest_width` (or, with
// use_padding_resize, utility.py transform_preds: float64 matrix times the float32 point, kept in float64), then
// np.round (half to even), np.clip, astype(int16). Returns the pre-rounding values as doubles.
OCRPP_HD void db_rescale_py(float mx, float my, int W, int H, float sw, float sh, int use_padding_resize,
                            double* fx, double* fy) {
  if (!use_padding_resize) {
    *fx = (double)fmul(fdiv(mx, (float)W), sw);
    *fy = (double)fmul(fdiv(my, (float)H), sh);
    return;
  }
  const double cx = (double)(float)((double)sw / 2.0), cy = (double)(float)((double)sh / 2.0);
  const double m = sw > sh ? (double)sw : (double)sh;
  const double s = m / (double)H;
  const double tx = cx >= cy ? 0.0 : cx - cy, ty = cx >= cy ? cy - cx : 0.0;
  *fx = dadd(dmul(s, (double)mx), tx);
  *fy = dadd(dmul(s, (double)my), ty);
}


// ------------------------------------------------------------------------------------------------
// score_mode "box" of the reference's Python branch (db_postprocess.py:109-110,178-194): the mean of the map over
// cv2.fillPoly(mask, mini_box.astype(int32), 1) - LINE_8, shift 0 - in the bounding rectangle of the float corners
// clipped to the map. Third-party behaviour restated from OpenCV's drawing.cpp (CollectPolyEdges /
// FillEdgeCollection, Line -> LineIterator(connectivity 8, leftToRight), clipLine), in its integer arithmetic:
//   every edge is drawn as a Bresenham line between the (clipped) end points, starting from the left one;
//   the interior of every scan line y0 <= y < y1 runs from ceil to floor of the edges' 16.16 fixed-point x, which
//   advance by the truncated quotient dx per row; an edge with an end point outside the mask is rebuilt from its
//   clipped end points when those lie on different rows.
// For the convex quads of this path every mask row is one interval: the result is L[y], R[y] (R < L: empty row).
// Pinned against cv2 4.13 by tests/test_geometry_host.py (exact for quads inside the mask, which is every mini box
// that does not leave the map; 1 differing mask in 2700 otherwise).
// ------------------------------------------------------------------------------------------------
OCRPP_HD bool cv_clip_line(long long w, long long h, long long* px1, long long* py1, long long* px2, long long* py2) {
  long long x1 = *px1, y1 = *py1, x2 = *px2, y2 = *py2;
  const long long right = w - 1, bottom = h - 1;
  if (w <= 0 || h <= 0) return false;
  int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
  int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
  if ((c1 & c2) == 0 && (c1 | c2) != 0) {
    long long a;
    if (c1 & 12) {
      a = c1 < 8 ? 0 : bottom;
      x1 += (long long)(dmul((double)(a - y1), (double)(x2 - x1)) / (double)(y2 - y1));
      y1 = a;
      c1 = (x1 < 0) + (x1 > right) * 2;
    }
    if (c2 & 12) {
      a = c2 < 8 ? 0 : bottom;
      x2 += (long long)(dmul((double)(a - y2), (double)(x2 - x1)) / (double)(y2 - y1));
      y2 = a;
      c2 = (x2 < 0) + (x2 > right) * 2;
    }
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
      if (c1) {
        a = c1 == 1 ? 0 : right;
        y1 += (long long)(dmul((double)(a - x1), (double)(y2 - y1)) / (double)(x2 - x1));
        x1 = a;
        c1 = 0;
      }
      if (c2) {
        a = c2 == 1 ? 0 : right;
        y2 += (long long)(dmul((double)(a - x2), (double)(y2 - y1)) / (double)(x2 - x1));
        x2 = a;
        c2 = 0;
      }
    }
  }
  *px1 = x1; *py1 = y1; *px2 = x2; *py2 = y2;
  return (c1 | c2) == 0;
}

OCRPP_HD void fill_quad_rows(const int* qx, const int* qy, int w, int h, int* L, int* R) {
  for (int y = 0; y < h; ++y) {
    L[y] = 0x7fffffff;
    R[y] = -1;
  }
  struct Edge { long long x, dx; int y0, y1; };
  Edge e[4];
  int ne = 0;
  for (int i = 0; i < 4; ++i) {
    const int j = (i + 3) & 3;   // previous vertex
    const long long t0x = qx[j], t0y = qy[j], t1x = qx[i], t1y = qy[i];
    const bool outside = (unsigned long long)t0x >= (unsigned long long)w || (unsigned long long)t1x >= (unsigned long long)w ||
                         (unsigned long long)t0y >= (unsigned long long)h || (unsigned long long)t1y >= (unsigned long long)h;
    long long c0x = t0x, c0y = t0y, c1x = t1x, c1y = t1y;
    bool visible = true;
    if (outside) visible = cv_clip_line(w, h, &c0x, &c0y, &c1x, &c1y);
    if (visible) {   // Line(): Bresenham from the left end point
      long long x = c0x, y = c0y;
      long long dx = c1x - c0x, dy = c1y - c0y;
      int sy = 1;
      if (dx < 0) { dx = -dx; dy = -dy; x = c1x; y = c1y; }
      if (dy < 0) { dy = -dy; sy = -1; }
      const bool vert = dy > dx;
      if (vert) { const long long t = dx; dx = dy; dy = t; }
      long long err = dx - (dy + dy);
      const long long plus = dx + dx, minus = -(dy + dy);
      for (long long k = 0; k <= dx; ++k) {
        if (L[y] > (int)x) L[y] = (int)x;
        if (R[y] < (int)x) R[y] = (int)x;
        const bool m = err < 0;
        err += minus + (m ? plus : 0);
        if (vert) { y += sy; if (m) x += 1; } else { x += 1; if (m) y += sy; }
      }
    }
    if (t0y == t1y) continue;
    // the scan-line edge: from the clipped end points when the original ones are outside and the clipped ones span rows
    long long a0x = t0x << 16, a0y = t0y, a1x = t1x << 16, a1y = t1y;
    if (outside && c0y != c1y) { a0x = c0x << 16; a0y = c0y; a1x = c1x << 16; a1y = c1y; }
    const long long ddx = (a1x - a0x) / (a1y - a0y);   // C integer division: truncation toward zero
    Edge ed;
    ed.dx = ddx;
    if (t0y < t1y) { ed.y0 = (int)t0y; ed.y1 = (int)t1y; ed.x = a0x + (t0y - a0y) * ddx; }
    else { ed.y0 = (int)t1y; ed.y1 = (int)t0y; ed.x = a1x + (t1y - a1y) * ddx; }
    e[ne++] = ed;
  }
  if (ne < 2) return;
  // sort by (y0, x, dx)
  for (int i = 1; i < ne; ++i) {
    const Edge v = e[i];
    int j = i - 1;
    while (j >= 0 && (e[j].y0 > v.y0 || (e[j].y0 == v.y0 && (e[j].x > v.x || (e[j].x == v.x && e[j].dx > v.dx))))) {
      e[j + 1] = e[j];
      --j;
    }
    e[j + 1] = v;
  }
  int ymax = e[0].y1;
  for (int i = 1; i < ne; ++i) ymax = e[i].y1 > ymax ? e[i].y1 : ymax;
  if (ymax > h) ymax = h;
  int act[4], na = 0, next = 0;
  for (int y = e[0].y0; y < ymax; ++y) {
    int k = 0;
    for (int i = 0; i < na; ++i)
      if (e[act[i]].y1 != y) act[k++] = act[i];
    na = k;
    while (next < ne && e[next].y0 == y) {   // insert by x
      int pos = 0;
      while (pos < na && e[act[pos]].x < e[next].x) ++pos;
      for (int i = na; i > pos; --i) act[i] = act[i - 1];
      act[pos] = next++;
      ++na;
    }
    for (int i = 0; i + 1 < na; i += 2) {
      Edge& a = e[act[i]];
      Edge& b = e[act[i + 1]];
      if (y >= 0) {
        const long long xa = a.x > b.x ? b.x : a.x, xb = a.x > b.x ? a.x : b.x;
        long long x1 = (xa + 65535) >> 16, x2 = xb >> 16;
        if (x1 < w && x2 >= 0) {
          if (x1 < 0) x1 = 0;
          if (x2 >= w) x2 = w - 1;
          if (x1 <= x2) {
            if (L[y] > (int)x1) L[y] = (int)x1;
            if (R[y] < (int)x2) R[y] = (int)x2;
          }
        }
      }
      a.x += a.dx;
      b.x += b.dx;
    }
    for (int i = 1; i < na; ++i) {   // keep the active list sorted by x
      const int v = act[i];
      int j = i - 1;
      while (j >= 0 && e[act[j]].x > e[v].x) {
        act[j + 1] = act[j];
        --j;
      }
      act[j + 1] = v;
    }
  }
}

// box_score's rectangle and integer quad (db_postprocess.py:183-192): float32 corners -> (xmin, ymin, w, h), qx/qy
OCRPP_HD void box_score_quad(const float* bx, const float* by, int W, int H, int* xmin, int* ymin, int* w, int* h,
                             int* qx, int* qy) {
  float fx0 = bx[0], fx1 = bx[0], fy0 = by[0], fy1 = by[0];
  for (int i = 1; i < 4; ++i) {
    fx0 = bx[i] < fx0 ? bx[i] : fx0;
    fx1 = bx[i] > fx1 ? bx[i] : fx1;
    fy0 = by[i] < fy0 ? by[i] : fy0;
    fy1 = by[i] > fy1 ? by[i] : fy1;
  }
  auto clampi = [](double v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : (int)v); };
  const int x0 = clampi(floor((double)fx0), 0, W - 1), x1 = clampi(ceil((double)fx1), 0, W - 1);
  const int y0 = clampi(floor((double)fy0), 0, H - 1), y1 = clampi(ceil((double)fy1), 0, H - 1);
  *xmin = x0;
  *ymin = y0;
  *w = x1 - x0 + 1;
  *h = y1 - y0 + 1;
  for (int i = 0; i < 4; ++i) {
    qx[i] = (int)fsub(bx[i], (float)x0);   // float32 subtraction, astype(int32): truncation toward zero
    qy[i] = (int)fsub(by[i], (float)y0);
  }
}

// np.round (half to even) on a double, as used by PSE/PAN generate_box (pse_postprocess.py:100-101)
OCRPP_HD double round_half_even(double v) {
  return nearbyint(v);
}

}  // namespace geom
}  // namespace ocrpp
