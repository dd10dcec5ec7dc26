// DB box extraction, stage 2: everything between the map scan and the per-candidate geometry as ONE kernel,
// one CTA per image (included by db.cu inside namespace ocrpp::{anonymous}).
//
// db_scan_kernel leaves, per image row, the run starts with cumulative pixel sums. This kernel turns them into
//   run table -> run union-find (both polarities) -> dense component ids -> component statistics ->
//   component/hole tree -> fill sums -> row extents, hole rings, stair pixels -> candidate order ->
//   triage (<= 2-point rule, BoxScore) -> convex hull of the row extents
// with every table in SHARED memory (8 bytes per run, 68 bytes per component, 8 bytes per candidate row; carved
// as the counts become known), so that the dependent look-ups (binary searches in the run table, parent
// chains, atomics of the reductions) run at shared-memory latency and the ten launches of the run-parallel
// chain (db_runs .. db_hull) become one. An image whose tables do not fit the CTA's shared memory is processed
// by the same code with the tables in the global workspace (`kSmem = false`), by the same CTA.
//
// What keeps one SM busy instead of waiting:
//   * global loads are issued in batches before their first use (run starts of 4 rows, cumulative sums of 4 runs,
//     the pixels under stair / ring positions are first collected in a list and then loaded all at once);
//   * runs alternate between foreground and background inside a row, so a thread owns a PAIR of runs and
//     handles "the foreground one" / "the background one" of its pair: no lane idles on the polarity test;
//   * every background run that touches the frame is linked straight to the first such run (R0): the outside
//     region (half of all runs, chains as long as the image is high) is flat from the start;
//   * 64-bit sums are two native 32-bit shared-memory atomics with an exact carry (no CAS loop);
//   * the two monotone chains of a candidate's hull run in place in its row-extent arrays, one thread each.
//
// Results handed to db_geometry_kernel, per candidate k (cv2 order) of image n, ko = n * maxc + k:
//   res_keep[ko] 0 dropped | 2 deferred to db_geometry_big_run | 3 hull ready,  res_score[ko] BoxScore,
//   cand_off[ko] slice of the hull scratch / global row extents, cand_y0[ko], cand_nrows[ko], hull_n[ko]
#pragma once

#ifndef OCRPP_IMG_THREADS
#define OCRPP_IMG_THREADS 1024
#endif
constexpr int kImgThreads = OCRPP_IMG_THREADS;   // a multiple of 64
constexpr int kImgChainRows = 512;               // tallest candidate whose hull is built here (in place, two threads)

struct ImgTables {
  int* rowptr;            // [H+1]
  uint16_t *xs, *yf;      // [nr]  first pixel | row + polarity << 15   (last pixel = next run's first - 1)
  int* par;               // [nr]  union-find parent, later ~component id
  unsigned* flag;         // [(nr+31)/32]
  // components, dense ids in raster order of the first pixel
  int *croot, *area, *xmin, *xmax, *ymax, *dmin, *dmax, *smin, *smax, *ecnt, *cpar, *cflag, *rowoff;
  unsigned long long *sum, *esum;
  int *ext_l, *ext_r;     // [etot]
  int* clist;             // [maxc] candidate -> component
  int* hcnt;              // [2*maxc] points of the two hull halves of every candidate
  int2* tasks;            // [ntask_cap] (component, y << 16 | x): add that pixel to the component's extra count / sum
  int ntask_cap;
};

// last run of [l, h) whose first pixel is <= x
template <typename XS>
__device__ __forceinline__ int run_at_range(const XS* xs, int l, int h, int x) {
  while (h - l > 1) {
    const int m = (l + h) >> 1;
    if ((int)xs[m] <= x) l = m; else h = m;
  }
  return l;
}

// *acc += v with two 32-bit atomics: the thread whose addition wraps the low word carries into the high word
template <bool kSmem>
__device__ __forceinline__ void add64(unsigned long long* acc, unsigned long long v) {
  if (kSmem) {
    unsigned* w = reinterpret_cast<unsigned*>(acc);
    const unsigned lo = (unsigned)v;
    unsigned hi = (unsigned)(v >> 32);
    const unsigned old = atomicAdd(w, lo);
    hi += (old + lo < old) ? 1u : 0u;
    if (hi) atomicAdd(w + 1, hi);
  } else {
    atomicAdd(acc, v);
  }
}

#ifdef OCRPP_IMG_ASSERT   // development aid: bounds checks that name the failing line
#define IMG_CHK(cond) do { if (!(cond)) { printf("IMG_CHK failed line %d img %d tid %d: %s\n", __LINE__, n, (int)threadIdx.x, #cond); __trap(); } } while (0)
#else
#define IMG_CHK(cond) do { } while (0)
#endif

#ifdef OCRPP_IMG_CLK   // development aid: per-stage cycle counts of image n in the last three slots of boxes_f_out
#define IMG_CLK(i) do { __syncthreads(); if (threadIdx.x == 0 && p.boxes_f_out) p.boxes_f_out[((size_t)n * p.maxc + p.maxc - 3) * 8 + (i)] = (float)(clock64() - clk0); } while (0)
#else
#define IMG_CLK(i) do { } while (0)
#endif

template <typename T, bool kSmem>
__device__ bool db_image_run(const DbParams& p, const int n, char* smem, const size_t smem_bytes) {
  __shared__ int s_etot, s_r0, s_ntask;
  const int tid = threadIdx.x, nt = kImgThreads, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
#ifdef OCRPP_IMG_CLK
  const long long clk0 = clock64();
#endif
  const int H = p.H, W = p.W;
  const int32_t* rc = p.srow_cnt + (size_t)n * H;
  const size_t ro = (size_t)n * p.R;
  ImgTables t;
  size_t used = 0;
  auto take = [&](size_t bytes) -> char* {
    used = (used + 7) & ~(size_t)7;
    char* q = smem + used;
    used += bytes;
    return q;
  };
  t.rowptr = kSmem ? reinterpret_cast<int*>(take(sizeof(int) * (H + 1))) : p.rowptr + (size_t)n * (H + 1);
  if (tid == 0) {
    s_r0 = 0x7fffffff;
    s_ntask = 0;
  }

  // ---- A: row counts -> rowptr, R0 (first background run that touches the frame); run table ----
  int nr;
  {
    const int chunk = (H + nt - 1) / nt;
    const int y0 = min(H, tid * chunk), y1 = min(H, y0 + chunk);
    int local = 0, over = 0;
    for (int y = y0; y < y1; ++y) {
      const int c = rc[y] & 0x7fffffff;
      over |= c > p.cap;
      local += c;
    }
    int total;
    int base = block_exclusive_scan(local, &total);
    int r0 = 0x7fffffff;
    for (int y = y0; y < y1; ++y) {
      t.rowptr[y] = base;
      const int c = rc[y] & 0x7fffffff, first = (unsigned)rc[y] >> 31;
      int cand = 0x7fffffff;
      if (!first) cand = base;                                            // first run is background: x = 0
      else if (y == 0 || y == H - 1) cand = c > 1 ? base + 1 : cand;      // second run of a frame row
      else if ((c - 1) & 1) cand = base + c - 1;                          // last run is background: x = W-1
      r0 = min(r0, cand);
      base += c;
    }
    if (r0 != 0x7fffffff) atomicMin(&s_r0, r0);
    if (tid == 0) t.rowptr[H] = total;
    over = __syncthreads_or(over);
    if (over || total > p.R) {
      if (tid == 0) {
        atomicOr(&p.imgflags[n], OCRPP_IMG_RUN_OVERFLOW);
        p.ncand[n] = 0;
      }
      return true;
    }
    nr = total;
  }
  const int R0 = s_r0;
  IMG_CLK(0);
  const int nwords = (nr + 31) / 32;
  if (kSmem) {
    t.xs = reinterpret_cast<uint16_t*>(take(sizeof(uint16_t) * (nr + 1)));
    t.yf = reinterpret_cast<uint16_t*>(take(sizeof(uint16_t) * (nr + 1)));
    t.par = reinterpret_cast<int*>(take(sizeof(int) * nr));
    t.flag = reinterpret_cast<unsigned*>(take(sizeof(unsigned) * nwords));
    if (used > smem_bytes) return false;
  } else {
    t.xs = p.run_xs + ro;
    t.yf = p.run_yf + ro;
    t.par = p.par + ro;
    t.flag = p.flagw + (size_t)n * (p.R / 32 + 1);
  }
  const unsigned long long* scum = p.scum + (size_t)n * H * (p.cap + 1);
  {
    // a warp takes 8 rows per step and issues the loads of all of them (32 entries each: cap >= 32, entries past a
    // row's count are ignored) before the first store
    constexpr int kRows = 8;
    for (int yb = warp * kRows; yb < H; yb += nw * kRows) {
      unsigned long long e[kRows];
      int cnt[kRows];
#pragma unroll
      for (int k = 0; k < kRows; ++k) {
        const int y = min(yb + k, H - 1);
        e[k] = scum[(size_t)y * (p.cap + 1) + lane];
        cnt[k] = rc[y];
      }
#pragma unroll
      for (int k = 0; k < kRows; ++k) {
        const int y = yb + k;
        if (y < H) {
          const int c = cnt[k] & 0x7fffffff, first = (unsigned)cnt[k] >> 31;
          const int rbase = t.rowptr[y];
          if (lane < c) {
            t.xs[rbase + lane] = (uint16_t)(e[k] >> 48);
            t.yf[rbase + lane] = (uint16_t)(y | ((first ^ (lane & 1)) << 15));
          }
          for (int j = lane + 32; j < c; j += 32) {
            t.xs[rbase + j] = (uint16_t)(scum[(size_t)y * (p.cap + 1) + j] >> 48);
            t.yf[rbase + j] = (uint16_t)(y | ((first ^ (j & 1)) << 15));
          }
        }
      }
    }
  }
  for (int w = tid; w < nwords; w += nt) t.flag[w] = 0u;
  __syncthreads();
  IMG_CLK(1);

  // last pixel of run r of a row that ends at run index `rend`
  auto xe_of = [&](int r, int rend) { return r + 1 < rend ? (int)t.xs[r + 1] - 1 : W - 1; };
  // first run of row y-1 that overlaps run r (same polarity; foreground 8-connected: [xs-1, xe+1], background
  // 4-connected: [xs, xe]), or -1; same-polarity runs alternate, so the further ones are q+2, q+4, ... while their
  // first pixel is <= *hi_out and they are < *b_out
  auto first_overlap = [&](int y, int fg, int xsr, int xer, int* hi_out, int* a_out, int* b_out) {
    const int b = t.rowptr[y], a = t.rowptr[y - 1];
    const int lo = xsr - fg, hi = xer + fg;
    int q = run_at_range(t.xs, a, b, max(lo, 0));
    if ((t.yf[q] >> 15) != fg) ++q;
    *hi_out = hi;
    *a_out = a;
    *b_out = b;
    return (q < b && (int)t.xs[q] <= hi) ? q : -1;
  };

  // ---- B: union-find. pass 1: every run points at its first overlapping run of the row above (no atomics).
  //      Background runs on the frame all point at R0 instead; one that also overlaps a run which is NOT on the
  //      frame (not the first / last run of its row, not in row 0) merges with all its overlaps in pass 3 ----
  for (int r = tid; r < nr; r += nt) {
    const int yfr = t.yf[r], y = yfr & 0x7fff, fg = yfr >> 15;
    const int xsr = t.xs[r], xer = xe_of(r, t.rowptr[y + 1]);
    const bool frame = !fg && (y == 0 || y == H - 1 || xsr == 0 || xer == W - 1);
    int first = frame ? R0 : r;
    if (y > 0) {
      int hi, a, b;
      const int q = first_overlap(y, fg, xsr, xer, &hi, &a, &b);
      if (q >= 0) {
        bool more;
        if (!frame) {
          first = q;
          more = q + 2 < b && (int)t.xs[q + 2] <= hi;
        } else {
          more = false;
          if (y > 1)
            for (int qq = q; qq < b && (int)t.xs[qq] <= hi; qq += 2) more |= qq != a && qq != b - 1;
        }
        if (more) atomicOr(&t.flag[r >> 5], 1u << (r & 31));
      }
    }
    IMG_CHK(first >= 0 && first <= r);
    t.par[r] = first;
  }
  __syncthreads();
  IMG_CLK(2);
  // pass 2: pointer jumping
  while (true) {
    int changed = 0;
    for (int r = tid; r < nr; r += nt) {
      const int q = t.par[r];
      const int g = t.par[q];
      if (g != q) {
        t.par[r] = g;
        changed = 1;
      }
    }
    if (!__syncthreads_or(changed)) break;
  }
  IMG_CLK(3);
  // pass 3: further overlaps merge chains (atomicMin linking, smallest run index wins). The flagged runs are few
  // and scattered: they are first compacted into a list in the still unused part of the table memory
  {
    auto merge_run = [&](int r) {
      const int yfr = t.yf[r], y = yfr & 0x7fff, fg = yfr >> 15;
      const int xsr = t.xs[r], xer = xe_of(r, t.rowptr[y + 1]);
      const bool frame = !fg && (y == 0 || y == H - 1 || xsr == 0 || xer == W - 1);
      int hi, a, b;
      const int q0 = first_overlap(y, fg, xsr, xer, &hi, &a, &b);
      for (int q = frame ? q0 : q0 + 2; q < b && (int)t.xs[q] <= hi; q += 2) uf_union_s(t.par, q, r);
    };
    int* flist = kSmem ? reinterpret_cast<int*>(smem + ((used + 7) & ~(size_t)7))
                       : reinterpret_cast<int*>(p.hull + (size_t)n * p.E * 4);
    const int flist_cap = kSmem ? (int)((smem_bytes - ((used + 7) & ~(size_t)7)) / sizeof(int)) : p.E * 8;
    const int wchunk = (nwords + nt - 1) / nt;
    const int w0 = min(nwords, tid * wchunk), w1 = min(nwords, w0 + wchunk);
    int local = 0;
    for (int w = w0; w < w1; ++w) local += __popc(t.flag[w]);
    int nflag;
    int pos = block_exclusive_scan(local, &nflag);
    if (nflag <= flist_cap) {
      for (int w = w0; w < w1; ++w)
        for (unsigned bits = t.flag[w]; bits; bits &= bits - 1) flist[pos++] = w * 32 + __ffs(bits) - 1;
      __syncthreads();
      for (int i = tid; i < nflag; i += nt) merge_run(flist[i]);
    } else {
      for (int w = warp; w < nwords; w += nw)
        if ((t.flag[w] >> lane) & 1u) merge_run(w * 32 + lane);
    }
  }
  __syncthreads();
  IMG_CLK(4);
  // pass 4: flatten, then dense component ids (raster order of the first pixel)
  int ncomp;
  {
    const int chunk = (nr + nt - 1) / nt;
    const int lo = min(nr, tid * chunk), hi = min(nr, lo + chunk);
    // read-only walks: a path-halving store of another thread could land after the owner's store of the root
    for (int r = tid; r < nr; r += nt) {
      int x = r, q;
      while ((q = reinterpret_cast<volatile int*>(t.par)[x]) != x) x = q;
      t.par[r] = x;
    }
    __syncthreads();
    // the outside region: every frame run hangs off R0, whose own root may be an earlier interior run
    const int root_out = R0 != 0x7fffffff ? t.par[R0] : -1;
    IMG_CLK(5);
    int local = 0;
    for (int r = lo; r < hi; ++r) local += t.par[r] == r ? 1 : 0;
    int id = block_exclusive_scan(local, &ncomp);
    if (kSmem) {
      const size_t nc = ncomp;
      t.sum = reinterpret_cast<unsigned long long*>(take(8 * nc));
      t.esum = reinterpret_cast<unsigned long long*>(take(8 * nc));
      int* const blk = reinterpret_cast<int*>(take(4 * 13 * nc));
      t.croot = blk; t.area = blk + nc; t.xmin = blk + 2 * nc; t.xmax = blk + 3 * nc; t.ymax = blk + 4 * nc;
      t.dmin = blk + 5 * nc; t.dmax = blk + 6 * nc; t.smin = blk + 7 * nc; t.smax = blk + 8 * nc;
      t.ecnt = blk + 9 * nc; t.cpar = blk + 10 * nc; t.cflag = blk + 11 * nc; t.rowoff = blk + 12 * nc;
      t.clist = reinterpret_cast<int*>(take(4 * (size_t)p.maxc));
      t.hcnt = reinterpret_cast<int*>(take(8 * (size_t)p.maxc));
      if (used > smem_bytes) return false;
    } else {
      t.croot = p.cand_root + ro; t.area = p.area + ro; t.xmin = p.xmin + ro; t.xmax = p.xmax + ro;
      t.ymax = p.ymax + ro; t.dmin = p.dmin + ro; t.dmax = p.dmax + ro; t.smin = p.smin + ro; t.smax = p.smax + ro;
      t.ecnt = p.fcnt + ro; t.cpar = p.cpar + ro; t.cflag = p.cflag + ro; t.rowoff = p.rowoff + ro;
      t.sum = reinterpret_cast<unsigned long long*>(p.sum + ro);
      t.esum = reinterpret_cast<unsigned long long*>(p.fsum + ro);
      t.clist = p.cand + (size_t)n * p.maxc;
      t.hcnt = p.hcnt + (size_t)n * p.maxc * 2;
    }
    for (int r = lo; r < hi; ++r) {
      if (t.par[r] != r) continue;
      const int c = id++;
      t.croot[c] = r;
      t.cflag[c] = r == root_out ? kOutFlag : 0;
      t.area[c] = 0;
      t.xmin[c] = 0x7fffffff; t.xmax[c] = -1; t.ymax[c] = -1;
      t.dmin[c] = 0x7fffffff; t.dmax[c] = -0x7fffffff;
      t.smin[c] = 0x7fffffff; t.smax[c] = -0x7fffffff;
      t.sum[c] = 0ull; t.esum[c] = 0ull; t.ecnt[c] = 0;
      t.cpar[c] = -1; t.rowoff[c] = -1;
      t.par[r] = ~c;
    }
    __syncthreads();
    for (int r = tid; r < nr; r += nt) {
      const int v = t.par[r];
      if (v >= 0) t.par[r] = t.par[v];
    }
    __syncthreads();
  }
  const int cout = R0 != 0x7fffffff ? ~t.par[R0] : -1;   // the (merged) outside region
  IMG_CHK(R0 == 0x7fffffff || (R0 >= 0 && R0 < nr && cout >= 0 && cout < ncomp));
  for (int r = tid; r < nr; r += nt) IMG_CHK(~t.par[r] >= 0 && ~t.par[r] < ncomp);
  IMG_CLK(6);

  // A thread owns the run pair (2i, 2i+1): `pick(i, pol)` is the run of polarity pol in it (the first one when
  // both have it: then *twice is set and the caller also handles 2i+1), or -1.
  const int npairs = (nr + 1) >> 1;
  auto pick = [&](int i, int pol, bool* twice) {
    const int r0 = 2 * i, r1 = 2 * i + 1;
    const bool m0 = (t.yf[r0] >> 15) == pol, m1 = r1 < nr && (t.yf[r1] >> 15) == pol;
    *twice = m0 && m1;
    return m0 ? r0 : (m1 ? r1 : -1);
  };

  // ---- C: per-component reductions from the per-run sums (no pixel is read again) ----
  {
    constexpr unsigned long long kCumMask = (1ull << 48) - 1ull;
    auto stat = [&](int r, int c, int y, unsigned long long e0, unsigned long long e1) {
      const int fg = t.yf[r] >> 15;
      const int a = t.xs[r], b = xe_of(r, t.rowptr[y + 1]);
      const unsigned long long rs = ((e1 & kCumMask) - (e0 & kCumMask)) << 9;   // 2^-23 units -> 32.32
      atomicAdd(&t.area[c], b - a + 1);
      add64<kSmem>(&t.sum[c], rs);
      atomicMin(&t.xmin[c], a);
      atomicMax(&t.xmax[c], b);
      atomicMax(&t.ymax[c], y);
      if (fg && a == b) {   // diagonal extents feed the "<= 2 contour points" rule only (one pixel per row)
        atomicMin(&t.dmin[c], a - y);
        atomicMax(&t.dmax[c], a - y);
        atomicMin(&t.smin[c], a + y);
        atomicMax(&t.smax[c], a + y);
      }
    };
    // all runs that are not part of the outside region, 4 per thread and step, loads first
    constexpr int kU = 4;
    for (int r0 = tid; r0 < nr; r0 += nt * kU) {
      unsigned long long e0[kU], e1[kU];
      int cc[kU], yy[kU];
#pragma unroll
      for (int k = 0; k < kU; ++k) {
        const int r = r0 + k * nt;
        cc[k] = -1;
        e0[k] = e1[k] = 0ull;
        yy[k] = 0;
        if (r < nr) {
          const int c = ~t.par[r];
          if (c != cout) {
            const int y = t.yf[r] & 0x7fff;
            const unsigned long long* sc = scum + (size_t)y * (p.cap + 1) + (r - t.rowptr[y]);
            e0[k] = sc[0];
            e1[k] = sc[1];
            cc[k] = c;
            yy[k] = y;
          }
        }
      }
#pragma unroll
      for (int k = 0; k < kU; ++k)
        if (cc[k] >= 0) stat(r0 + k * nt, cc[k], yy[k], e0[k], e1[k]);
    }
  }
  __syncthreads();
  IMG_CLK(7);

  // ---- D: parent links of the component / hole tree + row-extent slots ----
  const int cchunk = (ncomp + nt - 1) / nt;
  {
    const int lo = min(ncomp, tid * cchunk), hi = min(ncomp, lo + cchunk);
    int local = 0;
    for (int c = lo; c < hi; ++c) {
      const int r = t.croot[c];
      const int yfr = t.yf[r], y = yfr & 0x7fff, fg = yfr >> 15;
      int need = 0;
      if (fg) {
        // region LEFT of the component's first pixel: previous run of the same row (background)
        if (t.xs[r] > 0) {
          const int h = ~t.par[r - 1];
          if (h != cout) t.cpar[c] = h;
        }
        need = t.ymax[c] - y + 2;          // rows + 1
      } else if (c != cout) {
        IMG_CHK(y > 0 && y < H - 1 && t.xs[r] > 0);
        // pixel ABOVE the hole's first pixel is foreground and belongs to the enclosing component
        t.cpar[c] = ~t.par[run_at_range(t.xs, t.rowptr[y - 1], t.rowptr[y], (int)t.xs[r])];
        need = t.ymax[c] - y + 4;          // ring rows ymin-1 .. ymax+1, + 1
      }
      IMG_CHK(need >= 0 && need <= H + 4);
      t.rowoff[c] = need;
      local += need;
    }
    int etot;
    int base = block_exclusive_scan(local, &etot);
    for (int c = lo; c < hi; ++c) {
      const int need = t.rowoff[c];
      t.rowoff[c] = need ? base : -1;
      base += need;
    }
    if (tid == 0) s_etot = etot;
    __syncthreads();
  }
  const int etot = s_etot;
  if (kSmem) {
    t.ext_l = reinterpret_cast<int*>(take(sizeof(int) * etot));
    t.ext_r = reinterpret_cast<int*>(take(sizeof(int) * etot));
    if (used > smem_bytes || etot > p.E) return false;
    used = (used + 7) & ~(size_t)7;
    t.tasks = reinterpret_cast<int2*>(smem + used);
    t.ntask_cap = (int)((smem_bytes - used) / sizeof(int2));
  } else {
    if (etot > p.E) {   // cannot happen with E = 4R + 4; fail loudly
      if (tid == 0) {
        atomicOr(&p.imgflags[n], OCRPP_IMG_RUN_OVERFLOW);
        p.ncand[n] = 0;
      }
      return true;
    }
    t.ext_l = p.ext_l + (size_t)n * p.E;
    t.ext_r = p.ext_r + (size_t)n * p.E;
    t.tasks = reinterpret_cast<int2*>(p.hull + (size_t)n * p.E * 4);   // the hull scratch is not in use yet
    t.ntask_cap = p.E * 4;
  }
  for (int i = tid; i < etot; i += nt) {
    t.ext_l[i] = 0x7fffffff;
    t.ext_r[i] = -1;
  }
  IMG_CLK(8);
  // ---- E: every candidate adds its own (count, sum) to all its ancestors: fill = own + descendants ----
  for (int c = tid; c < ncomp; c += nt) {
    if (t.rowoff[c] < 0) continue;
    const int cnt = t.area[c];
    const unsigned long long s = t.sum[c];
    int a = t.cpar[c];
    while (a >= 0) {
      IMG_CHK(a < ncomp);
      atomicAdd(&t.ecnt[a], cnt);
      add64<kSmem>(&t.esum[a], s);
      a = t.cpar[a];
    }
  }
  __syncthreads();
  IMG_CLK(9);

  // ---- F: row extents of every candidate's point set, hole rings, stair pixels ----
  const long long img = n * p.stride_n;
  auto px = [&](int x, int y) { return (unsigned long long)to_fixed(load_px<T>(p.maps, img + y * p.stride_h + x)); };
  {
    // pixel (x,y) counts towards the fill of component c: queued (one shared-memory atomic per warp-wide group of
    // requests), its value is loaded later, all pixels at once
    auto task = [&](int c, int x, int y) {
      IMG_CHK(c >= 0 && c < ncomp && x >= 0 && x < W && y >= 0 && y < H);
      const unsigned m = __activemask();
      const int leader = __ffs(m) - 1;
      int base = 0;
      if (lane == leader) base = atomicAdd(&s_ntask, __popc(m));
      base = __shfl_sync(m, base, leader);
      const int slot = base + __popc(m & ((1u << lane) - 1u));
      if (slot < t.ntask_cap) {
        t.tasks[slot] = make_int2(c, (y << 16) | x);
      } else {
        atomicAdd(&t.ecnt[c], 1);
        add64<kSmem>(&t.esum[c], px(x, y));
      }
    };
    // component id of pixel (x,y), polarity in *fg
    auto comp_at = [&](int x, int y, int* fg) {
      const int q = run_at_range(t.xs, t.rowptr[y], t.rowptr[y + 1], x);
      *fg = t.yf[q] >> 15;
      return ~t.par[q];
    };
    auto is_fg = [&](int x, int y) {
      return (int)(t.yf[run_at_range(t.xs, t.rowptr[y], t.rowptr[y + 1], x)] >> 15);
    };
    auto in_hole = [&](int x, int y, int h) {
      if (x < 0 || y < 0 || x >= W || y >= H) return false;
      int fg;
      const int c = comp_at(x, y, &fg);
      return !fg && c == h;
    };
    auto fg_run = [&](int r) {
      const int y = t.yf[r] & 0x7fff;
      const int root = ~t.par[r];
      const int i = t.rowoff[root] + (y - (t.yf[t.croot[root]] & 0x7fff));
      IMG_CHK(t.rowoff[root] >= 0 && i >= 0 && i < etot);
      atomicMin(&t.ext_l[i], (int)t.xs[r]);
      atomicMax(&t.ext_r[i], xe_of(r, t.rowptr[y + 1]));
    };
    // (a) stair pixel of the OUTER contour of the component to the left of background run r: e = (a, y)
    auto bg_run = [&](int r) {
      const int y = t.yf[r] & 0x7fff;
      const int a = t.xs[r];
      if (a == 0) return;
      const bool updn = (y > 0 && is_fg(a, y - 1)) || (y < H - 1 && is_fg(a, y + 1));
      if (!updn) return;
      const int root = ~t.par[r];
      const int C = ~t.par[r - 1];
      const int pc = t.cpar[C];
      if (pc < 0 ? root == cout : root == pc) task(C, a, y);
    };
    for (int i = tid; i < npairs; i += nt) {
      bool twice;
      const int r = pick(i, 1, &twice);
      if (r >= 0) fg_run(r);
      if (twice) fg_run(2 * i + 1);
    }
    IMG_CLK(12);
    if (p.stairs) {
      for (int i = tid; i < npairs; i += nt) {
        bool twice;
        const int r = pick(i, 0, &twice);
        if (r >= 0) bg_run(r);
        if (twice) bg_run(2 * i + 1);
      }
    }
    IMG_CLK(15);
    // (b) hole runs (background, not the outside region; few): ring pixels - each counted once, owner = the first of
    //     its up / left / right / down neighbours that lies in the hole - with their row extents, and (c) the stair
    //     pixels of the hole contour. One WARP per hole run: lanes take the pixels of the run (one thread walking a
    //     run alone would be the critical path of the whole image), lane 0 the two ends.
    // the hole runs are first collected in a list (they cluster in a few rows), then dealt to the warps in turn
    int* hlist = reinterpret_cast<int*>(t.hcnt);   // [2*maxc] ints, not in use before stage G
    const int hcap = 2 * p.maxc;
    if (tid == 0) s_etot = 0;
    __syncthreads();
    for (int rr = tid; rr < nr; rr += nt) {
      if (!(t.yf[rr] >> 15) && ~t.par[rr] != cout) {
        const int slot = atomicAdd(&s_etot, 1);
        if (slot < hcap) hlist[slot] = rr;
      }
    }
    __syncthreads();
    const int nhole = s_etot;
    const bool listed = nhole <= hcap;
    IMG_CLK(16);
    for (int w = warp; w < (listed ? nhole : nwords); w += nw) {
      const int rr = w * 32 + lane;
      const bool hole = !listed && rr < nr && !(t.yf[rr] >> 15) && ~t.par[rr] != cout;
      for (unsigned todo = listed ? 1u : __ballot_sync(0xffffffffu, hole); todo; todo &= todo - 1) {
        const int r = listed ? hlist[w] : w * 32 + __ffs(todo) - 1;
        const int y = t.yf[r] & 0x7fff;
        const int a = t.xs[r], b = xe_of(r, t.rowptr[y + 1]);
        const int h = ~t.par[r];
        const int C = t.cpar[h];  // enclosing foreground component: the ring consists of ITS pixels only
        const int off = t.rowoff[h];
        const int y0 = (t.yf[t.croot[h]] & 0x7fff) - 1;
        IMG_CHK(C >= 0 && C < ncomp && off >= 0 && y > 0 && y < H - 1 && a > 0 && b < W - 1);
        auto in_C = [&](int x, int yy) {
          int f;
          const int c = comp_at(x, yy, &f);
          return f && c == C;
        };
        auto ring = [&](int x, int yy) {
          IMG_CHK(off + yy - y0 >= 0 && off + yy - y0 < etot);
          task(h, x, yy);
          atomicMin(&t.ext_l[off + yy - y0], x);
          atomicMax(&t.ext_r[off + yy - y0], x);
        };
        // hole runs never touch the frame: a-1, b+1, y-1, y+1 are inside the image
        for (int x = a + lane; x <= b; x += 32) {
          if (in_C(x, y + 1)) ring(x, y + 1);  // its UP neighbour is in h: always the owner
          if (in_C(x, y - 1)) {                // p = (x, y-1): down neighbour in h
            if (!in_hole(x, y - 2, h) && !in_hole(x - 1, y - 1, h) && !in_hole(x + 1, y - 1, h)) ring(x, y - 1);
          }
        }
        if (lane == 0) {
          // p = (b+1, y): left neighbour in h; owner unless its up neighbour is in h
          if (in_C(b + 1, y) && !in_hole(b + 1, y - 1, h)) ring(b + 1, y);
        } else if (lane == 1) {
          // p = (a-1, y): right neighbour in h; owner unless up or left neighbour is in h
          if (in_C(a - 1, y) && !in_hole(a - 1, y - 1, h) && !in_hole(a - 2, y, h)) ring(a - 1, y);
        } else if ((lane == 2 || lane == 3) && p.stairs) {
          // (c) o = (b, y) is the last pixel of a hole run, q = (b+1, y) is foreground; for dy in {-1,+1}:
          //     p = (b, y+dy) in C => the hole contour steps diagonally p <-> q and the 4-connected
          //     boundary also paints e = (b+1, y+dy)
          const int dy = lane == 2 ? -1 : 1;
          bool take_it = in_C(b, y + dy);
          const int ex = b + 1, ey = y + dy;
          // the same e is produced from o' = (b, y-2) with dy=+1 when that qualifies: count it there
          if (take_it && dy == -1 && in_hole(b, y - 2, h) && is_fg(b + 1, y - 2)) take_it = false;
          if (take_it) {
            int f;
            const int ce = comp_at(ex, ey, &f);
            // foreground e already belongs to the ring when one of its other neighbours is in h;
            // background e may itself be hole background
            if (f ? (in_hole(ex + 1, ey, h) || in_hole(ex, ey + dy, h)) : ce == h) take_it = false;
          }
          if (take_it) task(h, ex, ey);
        }
      }
    }
    __syncthreads();
  }
  IMG_CLK(10);

  // ---- G: candidates in cv2 order (reverse raster order of the first point); hulls; triage ----
  int ncand;
  {
    const int hi = ncomp - tid * cchunk, lo = max(0, hi - cchunk);
    int local = 0;
    for (int c = hi - 1; c >= lo; --c) local += t.rowoff[c] >= 0 ? 1 : 0;
    int total;
    int k = block_exclusive_scan(local, &total);
    if (tid == 0) {
      p.ncand[n] = min(total, p.maxc);
      if (total > p.maxc) atomicOr(&p.imgflags[n], OCRPP_IMG_CANDIDATES_TRUNCATED);
    }
    for (int c = hi - 1; c >= lo && k < p.maxc; --c)
      if (t.rowoff[c] >= 0) t.clist[k++] = c;
    ncand = min(total, p.maxc);
  }
  // the queued pixels: up to kU loads per thread are issued now and consumed after the hull chains, which do not need
  // the scores (DRAM serves these scattered 32-byte reads slowly; the chains hide it)
  const int ntask = min(s_ntask, t.ntask_cap);
  constexpr int kU = 4;
  int2 tk[kU];
  float tv[kU];   // raw values: converting here would wait for the loads
#pragma unroll
  for (int k = 0; k < kU; ++k) {
    const int i = tid + k * nt;
    tk[k] = make_int2(-1, 0);
    tv[k] = 0.f;
    if (i < ntask) {
      tk[k] = t.tasks[i];
      tv[k] = load_px<T>(p.maps, img + (long long)((unsigned)tk[k].y >> 16) * p.stride_h + (tk[k].y & 0xffff));
    }
  }
  __syncthreads();   // clist
  IMG_CLK(14);
  // Hull = monotone chain over (ext_l[i], y0+i), (ext_r[i], y0+i) (dev_geom.cuh: hull_row_extents32): the first half
  // walks the RIGHT extents downwards, the second the LEFT extents upwards from the last point of the first. The two
  // halves are independent, so they run as two threads (in different warps: no divergence between them); each
  // half's stack is a prefix of what it has visited and is kept IN PLACE in the extent array it walks (as packed
  // points), top-of-stack points in registers. Chains run for every candidate the non-score rules keep.
  constexpr int kHalf = kImgThreads / 2;
  for (int kb = 0; kb < ncand; kb += kHalf) {
    const int side = tid / kHalf, k = kb + (tid - side * kHalf);
    int c = 0, off = 0, nrows = 0, y0 = 0, verdict = 0;   // verdict 0 dropped | 2 big | 3 hull
    int first_pt = 0, turn_pt = 0;
    if (k < ncand) {
      c = t.clist[k];
      IMG_CHK(c >= 0 && c < ncomp);
      off = t.rowoff[c];
      const int yfr = t.yf[t.croot[c]];
      const int fg = yfr >> 15, y_first = yfr & 0x7fff;
      const int ymax = t.ymax[c], area = t.area[c];
      verdict = 3;
      if (fg && p.skip2) {  // "contour has <= 2 points" (db_postprocess.cpp:255-257)
        const int bw = t.xmax[c] - t.xmin[c] + 1, bh = ymax - y_first + 1;
        const bool diag = (bw == bh && bw == area) && (t.dmin[c] == t.dmax[c] || t.smin[c] == t.smax[c]);
        if (area == 1 || (bh == 1 && area == bw) || (bw == 1 && area == bh) || diag) verdict = 0;
      }
      nrows = fg ? (ymax - y_first + 1) : (ymax - y_first + 3);
      y0 = fg ? y_first : y_first - 1;
      // The chains run in place in the extent arrays, whatever the candidate's height; db_geometry_kernel takes any
      // hull of up to kFastPts vertices. Only very tall candidates go to the generic routine (one warp, mostly one
      // lane: ~100 us for a single candidate, which is then the tail of the whole step - with the limit at 64 rows
      // one 65-row blob on one page of a 256-page batch cost that rank 20 % of its step).
      if (verdict && (nrows > kImgChainRows || W >= 16384 || H >= 16384)) verdict = 2;
      if (verdict == 3) {
        IMG_CHK(off >= 0 && off + nrows < etot + 1 && nrows >= 1);
        first_pt = pk(t.ext_l[off], y0);                            // out[0] of the chain
        turn_pt = pk(t.ext_r[off + nrows - 1], y0 + nrows - 1);     // last point of the first half
      }
    }
    __syncthreads();   // both halves have read the two shared points before either writes in place
    int cntk = 0;      // side 0: points of the first half (with first_pt); side 1: pushes of the second half
    if (verdict == 3) {
      const int last = nrows - 1;
      if (!side) {
        int* ext_r = t.ext_r + off;
        // stack: out[0] = first_pt (register only), out[j] = ext_r[j - 1] for j >= 1
        int kk = 1, a = 0, b = first_pt, prev = first_pt;
        for (int j = 0; j < nrows; ++j) {
          const int q = pk(ext_r[j], y0 + j);
          if (q == prev) continue;   // single-pixel first row
          prev = q;
          while (kk >= 2 && cross32(a, b, q) <= 0) {
            --kk;
            b = a;
            if (kk >= 2) a = kk == 2 ? first_pt : ext_r[kk - 3];
          }
          ext_r[kk - 1] = q;
          ++kk;
          a = b;
          b = q;
        }
        cntk = kk;
      } else {
        int* ext_l = t.ext_l + off;
        // stack above the base: push j lives at ext_l[last - j]
        int kk = 0, a = 0, b = turn_pt, prev = turn_pt;
        for (int j = last; j >= 0; --j) {
          const int q = pk(ext_l[j], y0 + j);
          if (q == prev) continue;   // single-pixel last row
          prev = q;
          while (kk >= 1 && cross32(a, b, q) <= 0) {
            --kk;
            b = a;
            if (kk >= 1) a = kk == 1 ? turn_pt : ext_l[last - (kk - 2)];
          }
          ext_l[last - kk] = q;
          ++kk;
          a = b;
          b = q;
        }
        cntk = kk;
      }
    } else if (verdict == 2 && kSmem) {   // db_geometry_big_run reads the row extents from the global workspace
      int* g = (side ? p.ext_r : p.ext_l) + (size_t)n * p.E + off;
      const int* e = (side ? t.ext_r : t.ext_l) + off;
      for (int j = 0; j < nrows; ++j) g[j] = e[j];
    }
    if (k < ncand) t.hcnt[side * p.maxc + k] = cntk;
    if (kb == 0) IMG_CLK(17);
    if (kb == 0) {
      // the pixels requested above have arrived by now; further ones (more than kU per thread) in the plain way
#pragma unroll
      for (int q = 0; q < kU; ++q) asm volatile("" : "+f"(tv[q]));   // keeps the conversion (= the wait) down here
#pragma unroll
      for (int q = 0; q < kU; ++q)
        if (tk[q].x >= 0) {
          atomicAdd(&t.ecnt[tk[q].x], 1);
          add64<kSmem>(&t.esum[tk[q].x], (unsigned long long)to_fixed(tv[q]));
        }
      for (int i = tid + kU * nt; i < ntask; i += nt) {
        const int2 e = t.tasks[i];
        atomicAdd(&t.ecnt[e.x], 1);
        add64<kSmem>(&t.esum[e.x], px(e.y & 0xffff, (int)((unsigned)e.y >> 16)));
      }
    }
    __syncthreads();   // chain lengths, chains in place, fill sums complete
    if (kb == 0) IMG_CLK(18);
    if (k < ncand) {
      const size_t ko = (size_t)n * p.maxc + k;
      // BoxScore: a low score drops the candidate whatever its rectangle is
      const unsigned long long tot = t.sum[c] + t.esum[c];
      const int cnt = t.area[c] + t.ecnt[c];
      const double mean = ((double)(long long)tot / kFixScale) / (double)cnt;
      const float score = (float)mean;
      if (!p.score_box && db_low_score(p, mean)) verdict = 0;
      const int k1 = t.hcnt[k], k2 = t.hcnt[p.maxc + k];
      if (!side) {
        p.res_keep[ko] = verdict;
        if (verdict) {
          p.res_score[ko] = score;
          p.cand_off[ko] = off;
          p.cand_y0[ko] = y0;
          p.cand_nrows[ko] = nrows;
        }
        if (verdict == 2) {
          const int slot = atomicAdd(&p.nbig[n], 1);
          if (slot < p.maxc) p.big[(size_t)n * p.maxc + slot] = k;
        }
      }
      if (verdict == 3) {
        int* gout = reinterpret_cast<int*>(p.hull + ((size_t)n * p.E + off) * 4);   // >= 8 * (nrows + 1) ints
        // a single point (one row, one pixel): k1 == 1 and the reference routine does not run the second half
        if (!side) {
          gout[0] = first_pt;
          for (int j = 1; j < k1; ++j) gout[j] = t.ext_r[off + j - 1];
          p.hull_n[ko] = k1 > 1 ? k1 + k2 - 1 : 1;
        } else if (k1 > 1) {
          for (int j = 0; j < k2 - 1; ++j) gout[k1 + j] = t.ext_l[off + nrows - 1 - j];
        }
      }
    }
  }
  IMG_CLK(11);

  // ---- debug: canonical 8-connected foreground label map (id = 1 + rank of the first raster pixel) ----
  if (p.labels_dbg) {
    __syncthreads();
    const int lo = min(ncomp, tid * cchunk), hi = min(ncomp, lo + cchunk);
    int local = 0;
    for (int c = lo; c < hi; ++c) local += t.yf[t.croot[c]] >> 15;
    int total;
    int id = block_exclusive_scan(local, &total);
    for (int c = lo; c < hi; ++c)
      if (t.yf[t.croot[c]] >> 15) t.dmin[c] = ++id;
    __syncthreads();
    int32_t* lab = p.labels_dbg + (size_t)n * H * W;
    for (int r = tid; r < nr; r += nt) {
      const int yfr = t.yf[r], y = yfr & 0x7fff;
      if (!(yfr >> 15)) continue;
      const int v = t.dmin[~t.par[r]];
      const int b = xe_of(r, t.rowptr[y + 1]);
      for (int x = t.xs[r]; x <= b; ++x) lab[(size_t)y * W + x] = v;
    }
  }
  return true;
}

template <typename T>
__global__ void __launch_bounds__(kImgThreads) db_image_kernel(DbParams p, int smem_bytes, int mode) {
  extern __shared__ __align__(16) char s_img[];
  const int n = blockIdx.x + p.n0;
  // mode 0: shared-memory tables, global tables when they do not fit; 1: global tables only (test hook)
  if (mode == 0 && db_image_run<T, true>(p, n, s_img, (size_t)smem_bytes)) return;
  __syncthreads();
  db_image_run<T, false>(p, n, nullptr, 0);
}
