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

constexpr int kGrp = 8;                       // lanes per candidate
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

}  // namespace
}  // namespace ocrpp
