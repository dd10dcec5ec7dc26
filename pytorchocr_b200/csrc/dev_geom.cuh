// Warp- and group-parallel minimum-area rectangle / convex hull helpers on top of geometry.cuh,
// shared by the per-candidate geometry kernels of db.cu and expand.cu.
#pragma once
#include "common.cuh"
#include "geometry.cuh"

namespace ocrpp {
namespace {

using geom::P2i;

struct WarpBest {
  double area;
  int idx;
};

// warp-parallel min_area_rect: lanes take hull edges, same choice as geom::min_area_rect
__device__ void warp_min_area_rect(const P2i* h, int n, geom::Rect* r, int lane) {
  if (n == 1) {
    geom::min_area_rect(h, n, r);
    return;
  }
  const int ne = n == 2 ? 1 : n;
  geom::EdgeFit bf;
  bf.area = 1e300;
  bf.qx = 1;
  bf.qy = 0;
  int bi = 0x7fffffff;
  for (int i = lane; i < ne; i += 32) {
    const geom::EdgeFit f = geom::fit_edge(h, n, i);
    if (geom::fit_better(f, i, bf, bi)) {
      bf = f;
      bi = i;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    geom::EdgeFit of;
    of.area = __shfl_xor_sync(0xffffffffu, bf.area, o);
    of.qx = __shfl_xor_sync(0xffffffffu, bf.qx, o);
    of.qy = __shfl_xor_sync(0xffffffffu, bf.qy, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (oi != 0x7fffffff && (bi == 0x7fffffff || geom::fit_better(of, oi, bf, bi))) {
      bf.area = of.area;
      bf.qx = of.qx;
      bf.qy = of.qy;
      bi = oi;
    }
  }
  const geom::EdgeFit f = geom::fit_edge(h, n, bi);  // every lane recomputes the winner
  geom::rect_from_fit(h, n, bi, f, r);
}

#ifndef OCRPP_GEO_GRP
#define OCRPP_GEO_GRP 4
#endif
constexpr int kGrp = OCRPP_GEO_GRP;           // lanes per candidate (a power of two >= 4: one lane per corner in the offset).
                                              // 4 (eight candidates per warp) halves the warp instructions of the serial sections
                                              // (mini box, distance, chains, stores) against 8: db_geometry_kernel 0.125 -> 0.117 ms
__device__ __forceinline__ int pk(int x, int y) { return (x & 0xffff) | (y << 16); }
__device__ __forceinline__ int pkx(int v) { return (int)(short)(v & 0xffff); }
__device__ __forceinline__ int pky(int v) { return v >> 16; }

__device__ __forceinline__ int cross32(int o, int a, int b) {
  return (pkx(a) - pkx(o)) * (pky(b) - pky(o)) - (pky(a) - pky(o)) * (pkx(b) - pkx(o));
}

// monotone chain over packed points sorted by (y, x); same result as geom::hull_sorted
__device__ int hull_sorted32(const int* pts, int n, int* out) {
  if (n <= 1) {
    if (n == 1) out[0] = pts[0];
    return n;
  }
  int k = 0;
  for (int i = 0; i < n; ++i) {
    const int q = pts[i];
    if (i > 0 && q == pts[i - 1]) continue;
    while (k >= 2 && cross32(out[k - 2], out[k - 1], q) <= 0) --k;
    out[k++] = q;
  }
  if (k == 1) return 1;
  const int lo = k + 1;
  for (int i = n - 2; i >= 0; --i) {
    const int q = pts[i];
    if (q == pts[i + 1]) continue;
    while (k >= lo && cross32(out[k - 2], out[k - 1], q) <= 0) --k;
    out[k++] = q;
  }
  return k - 1;
}

// hull_sorted32 for row extents: pts = (left_0, right_0, left_1, right_1, ...) of consecutive rows. Same hull, half
// the visits: with y ascending a kept turn bulges towards +x, so the first pass can only keep RIGHT extents (a
// left extent of a later row is popped again by the right extent of its own row, and so is everything it popped),
// and the second pass only LEFT extents.
__device__ int hull_row_extents32(const int* pts, int nrows, int* out) {
  int k = 0;
  int prev = pts[0];
  out[k++] = prev;
  for (int i = 0; i < nrows; ++i) {
    const int q = pts[2 * i + 1];
    if (q == prev) continue;
    prev = q;
    while (k >= 2 && cross32(out[k - 2], out[k - 1], q) <= 0) --k;
    out[k++] = q;
  }
  if (k == 1) return 1;
  const int lo = k + 1;
  for (int i = nrows - 1; i >= 0; --i) {
    const int q = pts[2 * i];
    if (q == prev) continue;
    prev = q;
    while (k >= lo && cross32(out[k - 2], out[k - 1], q) <= 0) --k;
    out[k++] = q;
  }
  return k - 1;
}

struct Fit32 {
  int smin, smax, tmin, tmax, len2, qx, qy, idx;
  double area;
};

__device__ __forceinline__ bool fit32_better(const Fit32& a, const Fit32& b) {
  if (b.idx == 0x7fffffff) return a.idx != 0x7fffffff;
  if (a.idx == 0x7fffffff) return false;
  const double m = fmax(a.area, b.area);
  if (fabs(a.area - b.area) > 1e-12 * m) return a.area < b.area;
  const long long l = (long long)a.qy * b.qx, r = (long long)b.qy * a.qx;
  if (l != r) return l > r;
  return a.idx < b.idx;
}

// group-parallel min-area rectangle over packed hull points in shared memory; identical choice
// to geom::min_area_rect (exact integer projections, fit_better ordering)
__device__ void group_min_area_rect(const int* h, int n, geom::Rect* r, int gl, unsigned gmask) {
  if (n == 1) {
    for (int q = 0; q < 4; ++q) {
      r->cx[q] = pkx(h[0]);
      r->cy[q] = pky(h[0]);
    }
    r->w = r->h = 0.0;
    return;
  }
  const int ne = n == 2 ? 1 : n;
  Fit32 best;
  best.idx = 0x7fffffff;
  best.area = 1e300;
  best.qx = 1;
  best.qy = 0;
  best.smin = best.smax = best.tmin = best.tmax = 0;
  best.len2 = 1;
  for (int i = gl; i < ne; i += kGrp) {
    const int p0 = h[i], p1 = h[i + 1 == n ? 0 : i + 1];
    const int px = pkx(p0), py = pky(p0);
    const int dx = pkx(p1) - px, dy = pky(p1) - py;
    Fit32 f;
    f.smin = f.tmin = 0x7fffffff;
    f.smax = f.tmax = -0x7fffffff;
    for (int j = 0; j < n; ++j) {
      const int v = h[j];
      const int vx = pkx(v) - px, vy = pky(v) - py;
      const int sv = vx * dx + vy * dy, tv = vy * dx - vx * dy;
      f.smin = min(f.smin, sv);
      f.smax = max(f.smax, sv);
      f.tmin = min(f.tmin, tv);
      f.tmax = max(f.tmax, tv);
    }
    f.len2 = dx * dx + dy * dy;
    f.area = (double)((long long)(f.smax - f.smin) * (long long)(f.tmax - f.tmin)) / (double)f.len2;
    int qx = dx, qy = dy;
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      if (!(qx > 0 && qy >= 0)) {
        const int tx = qy;
        qy = -qx;
        qx = tx;
      }
    }
    f.qx = qx;
    f.qy = qy;
    f.idx = i;
    if (fit32_better(f, best)) best = f;
  }
#pragma unroll
  for (int o = kGrp / 2; o > 0; o >>= 1) {
    Fit32 of;
    of.smin = __shfl_xor_sync(gmask, best.smin, o);
    of.smax = __shfl_xor_sync(gmask, best.smax, o);
    of.tmin = __shfl_xor_sync(gmask, best.tmin, o);
    of.tmax = __shfl_xor_sync(gmask, best.tmax, o);
    of.len2 = __shfl_xor_sync(gmask, best.len2, o);
    of.qx = __shfl_xor_sync(gmask, best.qx, o);
    of.qy = __shfl_xor_sync(gmask, best.qy, o);
    of.idx = __shfl_xor_sync(gmask, best.idx, o);
    of.area = __shfl_xor_sync(gmask, best.area, o);
    if (fit32_better(of, best)) best = of;
  }
  // rect_from_fit (geometry.cuh) on the winning edge
  const int i = best.idx;
  const int p0 = h[i], p1 = h[i + 1 == n ? 0 : i + 1];
  const double dx = (double)(pkx(p1) - pkx(p0)), dy = (double)(pky(p1) - pky(p0));
  const double il2 = 1.0 / (double)best.len2;
  const double s0 = (double)best.smin * il2, s1 = (double)best.smax * il2;
  const double t0 = (double)best.tmin * il2, t1 = (double)best.tmax * il2;
  const double ss[4] = {s0, s1, s1, s0}, tt[4] = {t0, t0, t1, t1};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    r->cx[q] = (double)pkx(p0) + dx * ss[q] - dy * tt[q];
    r->cy[q] = (double)pky(p0) + dy * ss[q] + dx * tt[q];
  }
  const double len = sqrt((double)best.len2);
  r->w = (double)(best.smax - best.smin) / len;
  r->h = (double)(best.tmax - best.tmin) / len;
}

// ClipperOffset of one closed integer quad (geom::do_offset_quad) with ONE LANE PER CORNER: lanes 0..3 of
// the group compute their corner's normals, angle and arc points at the same time and write them at their
// rank; the result (points and order) is identical to the sequential routine. Inputs are group-uniform.
// Anything off the common path (fewer than 4 distinct points, a near-straight corner that Clipper skips
// without advancing k, near-zero delta) takes the sequential routine on lane 0. Returns the point count
// on every lane (-1: cap too small).
// `out` receives the points packed as x | y << 16 (coordinates of this path are within +-16384 plus the distance).
__device__ int group_do_offset_quad(const P2i* quad, double delta, int* out, int cap, int gl, unsigned gmask) {
  const double kPi = 3.141592653589793238, kTwoPi = kPi * 2;
  bool serial = false;
  // AddPath duplicate stripping (clipper.cpp:3845-3864): the parallel path needs 4 distinct points
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int j = (i + 1) & 3;
    serial |= quad[i].x == quad[j].x && quad[i].y == quad[j].y;
  }
  serial |= (delta > -1e-20 && delta < 1e-20) || !(delta > 0.0);
  P2i c[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) c[i] = quad[i];
  {  // FixOrientations (clipper.cpp:3889-3903)
    double a = 0;
#pragma unroll
    for (int i = 0, j = 3; i < 4; j = i, ++i) a += ((double)c[j].x + c[i].x) * ((double)c[j].y - c[i].y);
    if (!(-a * 0.5 >= 0)) {
      const P2i t0 = c[0], t1 = c[1];
      c[0] = c[3];
      c[1] = c[2];
      c[2] = t1;
      c[3] = t0;
    }
  }
  const double ad = fabs(delta);
  double y = 0.25;
  if (0.25 > ad * 0.25) y = ad * 0.25;
  double steps = kPi / acos(1 - y / ad);
  if (steps > ad * kPi) steps = ad * kPi;
  const double m_sin = sin(kTwoPi / steps), m_cos = cos(kTwoPi / steps);
  const double steps_per_rad = steps / kTwoPi;
  // lane j (< 4) owns corner j: normal of edge j -> j+1 and of edge k -> j with k = j - 1
  const int j = gl & 3, k = (j + 3) & 3;
  double njx, njy, nkx, nky;
  {
    const P2i p1 = c[j], p2 = c[(j + 1) & 3];
    double dx = (double)(p2.x - p1.x), dy = (double)(p2.y - p1.y);
    const double f = 1 * 1.0 / sqrt(dx * dx + dy * dy);
    dx *= f;
    dy *= f;
    njx = dy;
    njy = -dx;
  }
  // normal k from the lane that owns corner k (same group)
  const int src = ((threadIdx.x & 31) & ~(kGrp - 1)) + k;
  nkx = __shfl_sync(gmask, njx, src);
  nky = __shfl_sync(gmask, njy, src);
  // products and sums rounded one by one, as in the reference's x86 build (see geom::dmul)
  double sin_a = geom::dsub(geom::dmul(nkx, njy), geom::dmul(njx, nky));
  const double cos_a = geom::dadd(geom::dmul(nkx, njx), geom::dmul(njy, nky));
  int kind = 1, ns = 0;  // 1: round join, 2: concave (3 points)
  if (fabs(sin_a * delta) < 1.0) {
    if (cos_a > 0) kind = 0;  // Clipper emits one point and does NOT advance k: sequential routine
  } else if (sin_a > 1.0) sin_a = 1.0;
  else if (sin_a < -1.0) sin_a = -1.0;
  double ang = 0;
  if (kind == 1) {
    if (sin_a * delta < 0) kind = 2;
    else {
      ang = atan2(sin_a, geom::dadd(geom::dmul(nkx, njx), geom::dmul(nky, njy)));
      long long n = geom::clipper_round(geom::dmul(steps_per_rad, fabs(ang)));
      if (n < 1) n = 1;
      ns = n > 100000 ? 100000 : (int)n;
    }
  }
  serial |= (__ballot_sync(gmask, gl < 4 && kind == 0) & gmask) != 0;
  if (serial) {
    int m = 0;
    if (gl == 0) m = geom::do_offset_quad_t(quad, delta, geom::PackedOut{out}, cap);
    return __shfl_sync(gmask, m, (threadIdx.x & 31) & ~(kGrp - 1));
  }
  const int cnt = gl < 4 ? (kind == 2 ? 3 : ns + 1) : 0;
  // exclusive prefix of the four counts (corner order)
  int off = 0, total = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int cq = __shfl_sync(gmask, cnt, ((threadIdx.x & 31) & ~(kGrp - 1)) + q);
    off += q < j ? cq : 0;
    total += cq;
  }
  if (total > cap) return -1;
  auto ofs = [delta](int cc, double nn) { return (int)geom::clipper_round(geom::dadd((double)cc, geom::dmul(nn, delta))); };
  if (gl < 4) {
    int* o = out + off;
    if (kind == 2) {
      o[0] = pk(ofs(c[j].x, nkx), ofs(c[j].y, nky));
      o[1] = pk(c[j].x, c[j].y);
      o[2] = pk(ofs(c[j].x, njx), ofs(c[j].y, njy));
    } else {
      double X = nkx, Y = nky;
      for (int i = 0; i < ns; ++i) {
        o[i] = pk(ofs(c[j].x, X), ofs(c[j].y, Y));
        const double X2 = X;
        X = geom::dsub(geom::dmul(X, m_cos), geom::dmul(m_sin, Y));
        Y = geom::dadd(geom::dmul(X2, m_sin), geom::dmul(Y, m_cos));
      }
      o[ns] = pk(ofs(c[j].x, njx), ofs(c[j].y, njy));
    }
  }
  return total;
}

}  // namespace
}  // namespace ocrpp
