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
// Results handed to db_geometry_kernel, per candidate k (cv2 order) of image n, ko = n * maxc + k:
//   res_keep[ko] 0 dropped | 2 deferred to db_geometry_big_kernel | 3 hull ready,  res_score[ko] BoxScore,
//   cand_off[ko] slice of the hull scratch / global row extents, cand_y0[ko], cand_nrows[ko], hull_n[ko]
#pragma once

constexpr int kImgThreads = 1024;

struct ImgTables {
  int* rowptr;            // [H+1]
  uint16_t *xs, *yf;      // [nr]  first pixel | row + polarity << 15   (last pixel = next run's first - 1)
  int* par;               // [nr]  union-find parent, later ~component id
  unsigned* flag;         // [(nr+31)/32]
  // components, dense ids in raster order of the first pixel
  int *croot, *area, *xmin, *xmax, *ymax, *dmin, *dmax, *smin, *smax, *ecnt, *cpar, *cflag, *rowoff;
  unsigned long long *sum, *esum;
  int *ext_l, *ext_r;     // [etot]
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

template <typename T, bool kSmem>
__device__ bool db_image_run(const DbParams& p, const int n, char* smem, const size_t smem_bytes) {
  __shared__ int s_etot;
  const int tid = threadIdx.x, nt = kImgThreads, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
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

  // ---- A: row counts -> rowptr; run table ----
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
    for (int y = y0; y < y1; ++y) {
      t.rowptr[y] = base;
      base += rc[y] & 0x7fffffff;
    }
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
  const int nwords = (nr + 31) / 32;
  if (kSmem) {
    t.xs = reinterpret_cast<uint16_t*>(take(sizeof(uint16_t) * nr));
    t.yf = reinterpret_cast<uint16_t*>(take(sizeof(uint16_t) * nr));
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
  for (int y = warp; y < H; y += nw) {
    const int rbase = t.rowptr[y];
    const int c = rc[y] & 0x7fffffff;
    const int first = (unsigned)rc[y] >> 31;
    const unsigned long long* sc = scum + (size_t)y * (p.cap + 1);
    for (int j = lane; j < c; j += 32) {
      t.xs[rbase + j] = (uint16_t)(sc[j] >> 48);
      t.yf[rbase + j] = (uint16_t)(y | ((first ^ (j & 1)) << 15));
    }
  }
  for (int w = tid; w < nwords; w += nt) t.flag[w] = 0u;
  __syncthreads();

  // last pixel of run r of a row that ends at run index `rend`
  auto xe_of = [&](int r, int rend) { return r + 1 < rend ? (int)t.xs[r + 1] - 1 : W - 1; };
  // first run of row y-1 that overlaps run r (same polarity; foreground 8-connected: [xs-1, xe+1], background
  // 4-connected: [xs, xe]), or -1; *more: a second one exists (same-polarity runs alternate: q, q+2, ...)
  auto first_overlap = [&](int r, int y, int fg, int* hi_out, int* b_out) {
    const int b = t.rowptr[y], a = t.rowptr[y - 1];
    const int lo = (int)t.xs[r] - fg, hi = xe_of(r, t.rowptr[y + 1]) + fg;
    int q = run_at_range(t.xs, a, b, max(lo, 0));
    if ((t.yf[q] >> 15) != fg) ++q;
    *hi_out = hi;
    *b_out = b;
    return (q < b && (int)t.xs[q] <= hi) ? q : -1;
  };

  // ---- B: union-find. pass 1: every run points at its first overlapping run of the row above (no atomics) ----
  for (int r = tid; r < nr; r += nt) {
    const int yfr = t.yf[r], y = yfr & 0x7fff, fg = yfr >> 15;
    int first = r;
    if (y > 0) {
      int hi, b;
      const int q = first_overlap(r, y, fg, &hi, &b);
      if (q >= 0) {
        first = q;
        if (q + 2 < b && (int)t.xs[q + 2] <= hi) atomicOr(&t.flag[r >> 5], 1u << (r & 31));
      }
    }
    t.par[r] = first;
  }
  __syncthreads();
  // pass 2: pointer jumping (the outside region forms chains as long as the image is high)
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
  // pass 3: further overlaps merge chains (atomicMin linking, smallest run index wins)
  for (int r = tid; r < nr; r += nt) {
    if (!((t.flag[r >> 5] >> (r & 31)) & 1u)) continue;
    const int yfr = t.yf[r], y = yfr & 0x7fff, fg = yfr >> 15;
    int hi, b;
    const int q0 = first_overlap(r, y, fg, &hi, &b);
    for (int q = q0 + 2; q < b && (int)t.xs[q] <= hi; q += 2) uf_union_s(t.par, q, r);
  }
  __syncthreads();
  // pass 4: flatten; background regions touching the frame are OUT (collected in the bitmap, per root)
  for (int w = tid; w < nwords; w += nt) t.flag[w] = 0u;
  __syncthreads();
  for (int r = tid; r < nr; r += nt) {
    const int root = uf_find_s(t.par, r);
    t.par[r] = root;
    const int yfr = t.yf[r], y = yfr & 0x7fff, fg = yfr >> 15;
    if (!fg && (y == 0 || y == H - 1 || t.xs[r] == 0 || xe_of(r, t.rowptr[y + 1]) == W - 1)) {
      const unsigned bit = 1u << (root & 31);
      if (!(reinterpret_cast<volatile unsigned*>(t.flag)[root >> 5] & bit)) atomicOr(&t.flag[root >> 5], bit);
    }
  }
  __syncthreads();

  // ---- dense component ids (raster order of the first pixel) ----
  int ncomp;
  {
    const int chunk = (nr + nt - 1) / nt;
    const int lo = min(nr, tid * chunk), hi = min(nr, lo + chunk);
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
      if (used > smem_bytes) return false;
    } else {
      t.croot = p.cand_root + ro; t.area = p.area + ro; t.xmin = p.xmin + ro; t.xmax = p.xmax + ro;
      t.ymax = p.ymax + ro; t.dmin = p.dmin + ro; t.dmax = p.dmax + ro; t.smin = p.smin + ro; t.smax = p.smax + ro;
      t.ecnt = p.fcnt + ro; t.cpar = p.cpar + ro; t.cflag = p.cflag + ro; t.rowoff = p.rowoff + ro;
      t.sum = reinterpret_cast<unsigned long long*>(p.sum + ro);
      t.esum = reinterpret_cast<unsigned long long*>(p.fsum + ro);
    }
    for (int r = lo; r < hi; ++r) {
      if (t.par[r] != r) continue;
      const int c = id++;
      t.croot[c] = r;
      t.cflag[c] = ((t.flag[r >> 5] >> (r & 31)) & 1u) ? kOutFlag : 0;
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

  // ---- C: per-component reductions from the per-run sums (no pixel is read again) ----
  constexpr unsigned long long kCumMask = (1ull << 48) - 1ull;
  for (int r = tid; r < nr; r += nt) {
    const int c = ~t.par[r];
    const int yfr = t.yf[r], y = yfr & 0x7fff, fg = yfr >> 15;
    if (!fg && (t.cflag[c] & kOutFlag)) continue;
    const int a = t.xs[r], b = xe_of(r, t.rowptr[y + 1]);
    const unsigned long long* sc = scum + (size_t)y * (p.cap + 1) + (r - t.rowptr[y]);
    const unsigned long long rs = ((sc[1] & kCumMask) - (sc[0] & kCumMask)) << 9;   // 2^-23 units -> 32.32
    atomicAdd(&t.area[c], b - a + 1);
    atomicAdd(&t.sum[c], rs);
    atomicMin(&t.xmin[c], a);
    atomicMax(&t.xmax[c], b);
    atomicMax(&t.ymax[c], y);
    if (fg && a == b) {   // diagonal extents feed the "<= 2 contour points" rule only (one pixel per row)
      atomicMin(&t.dmin[c], a - y);
      atomicMax(&t.dmax[c], a - y);
      atomicMin(&t.smin[c], a + y);
      atomicMax(&t.smax[c], a + y);
    }
  }
  __syncthreads();

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
          if (!(t.cflag[h] & kOutFlag)) t.cpar[c] = h;
        }
        need = t.ymax[c] - y + 2;          // rows + 1
      } else if (!(t.cflag[c] & kOutFlag)) {
        // pixel ABOVE the hole's first pixel is foreground and belongs to the enclosing component
        t.cpar[c] = ~t.par[run_at_range(t.xs, t.rowptr[y - 1], t.rowptr[y], (int)t.xs[r])];
        need = t.ymax[c] - y + 4;          // ring rows ymin-1 .. ymax+1, + 1
      }
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
  }
  for (int i = tid; i < etot; i += nt) {
    t.ext_l[i] = 0x7fffffff;
    t.ext_r[i] = -1;
  }
  // ---- E: every candidate adds its own (count, sum) to all its ancestors: fill = own + descendants ----
  for (int c = tid; c < ncomp; c += nt) {
    if (t.rowoff[c] < 0) continue;
    const int cnt = t.area[c];
    const unsigned long long s = t.sum[c];
    int a = t.cpar[c];
    while (a >= 0) {
      atomicAdd(&t.ecnt[a], cnt);
      atomicAdd(&t.esum[a], s);
      a = t.cpar[a];
    }
  }
  __syncthreads();

  // ---- F: row extents of every candidate's point set, hole rings, stair pixels ----
  {
    const long long img = n * p.stride_n;
    auto px = [&](int x, int y) { return (unsigned long long)to_fixed(load_px<T>(p.maps, img + y * p.stride_h + x)); };
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
    for (int r = tid; r < nr; r += nt) {
      const int yfr = t.yf[r], y = yfr & 0x7fff, fg = yfr >> 15;
      const int a = t.xs[r], b = xe_of(r, t.rowptr[y + 1]);
      const int root = ~t.par[r];
      if (fg) {
        const int off = t.rowoff[root];
        const int i = off + (y - (t.yf[t.croot[root]] & 0x7fff));
        atomicMin(&t.ext_l[i], a);
        atomicMax(&t.ext_r[i], b);
        continue;
      }
      // ---- background run ----
      // (a) stair pixel of the OUTER contour of the component to the left: e = (a, y)
      if (a > 0 && p.stairs) {
        const int C = ~t.par[r - 1];
        const bool updn = (y > 0 && is_fg(a, y - 1)) || (y < H - 1 && is_fg(a, y + 1));
        if (updn) {
          const int pc = t.cpar[C];
          const bool outer_region = pc < 0 ? (t.cflag[root] & kOutFlag) != 0 : (root == pc);
          if (outer_region) {
            atomicAdd(&t.ecnt[C], 1);
            atomicAdd(&t.esum[C], px(a, y));
          }
        }
      }
      if (t.cflag[root] & kOutFlag) continue;
      // (b) hole run: ring pixels (each counted once: owner = first of up/left/right/down neighbour that lies
      //     in the hole) and their row extents; (c) stair pixels of the hole contour.
      const int h = root;
      const int C = t.cpar[h];  // enclosing foreground component: the ring consists of ITS pixels only
      const int off = t.rowoff[h];
      const int y0 = (t.yf[t.croot[h]] & 0x7fff) - 1;
      int cnt = 0;
      unsigned long long s = 0;
      auto in_C = [&](int x, int yy) {
        int f;
        const int c = comp_at(x, yy, &f);
        return f && c == C;
      };
      auto ring = [&](int x, int yy) {
        ++cnt;
        s += px(x, yy);
        atomicMin(&t.ext_l[off + yy - y0], x);
        atomicMax(&t.ext_r[off + yy - y0], x);
      };
      // hole runs never touch the frame: a-1, b+1, y-1, y+1 are inside the image
      for (int x = a; x <= b; ++x) {
        if (in_C(x, y + 1)) ring(x, y + 1);  // its UP neighbour is in h: always the owner
        if (in_C(x, y - 1)) {                // p = (x, y-1): down neighbour in h
          if (!in_hole(x, y - 2, h) && !in_hole(x - 1, y - 1, h) && !in_hole(x + 1, y - 1, h)) ring(x, y - 1);
        }
      }
      if (in_C(b + 1, y) && !in_hole(b + 1, y - 1, h)) ring(b + 1, y);
      if (in_C(a - 1, y) && !in_hole(a - 1, y - 1, h) && !in_hole(a - 2, y, h)) ring(a - 1, y);
      if (p.stairs) {
        for (int dy = -1; dy <= 1; dy += 2) {
          if (!in_C(b, y + dy)) continue;
          const int ex = b + 1, ey = y + dy;
          if (dy == -1) {
            if (in_hole(b, y - 2, h) && is_fg(b + 1, y - 2)) continue;
          }
          int f;
          const int ce = comp_at(ex, ey, &f);
          if (f) {
            if (in_hole(ex + 1, ey, h) || in_hole(ex, ey + dy, h)) continue;
          } else {
            if (ce == h) continue;  // e itself is hole background
          }
          ++cnt;
          s += px(ex, ey);
        }
      }
      if (cnt) {
        atomicAdd(&t.ecnt[h], cnt);
        atomicAdd(&t.esum[h], s);
      }
    }
  }
  __syncthreads();

  // ---- G: candidates in cv2 order (reverse raster order of the first point), triage, hull ----
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
    for (int c = hi - 1; c >= lo; --c) {
      const int off = t.rowoff[c];
      if (off < 0) continue;
      if (k >= p.maxc) break;
      const size_t ko = (size_t)n * p.maxc + k;
      ++k;
      const int yfr = t.yf[t.croot[c]];
      const int fg = yfr >> 15, y_first = yfr & 0x7fff;
      const int ymax = t.ymax[c], area = t.area[c];
      p.res_keep[ko] = 0;
      if (fg && p.skip2) {  // "contour has <= 2 points" (db_postprocess.cpp:255-257)
        const int bw = t.xmax[c] - t.xmin[c] + 1, bh = ymax - y_first + 1;
        const bool diag = (bw == bh && bw == area) && (t.dmin[c] == t.dmax[c] || t.smin[c] == t.smax[c]);
        if (area == 1 || (bh == 1 && area == bw) || (bw == 1 && area == bh) || diag) continue;
      }
      // BoxScore: a low score drops the candidate whatever its rectangle is
      const unsigned long long tot = t.sum[c] + t.esum[c];
      const int cnt = area + t.ecnt[c];
      const float score = (float)(((double)(long long)tot / kFixScale) / (double)cnt);
      if (score < p.box_thresh) continue;
      const int nrows = fg ? (ymax - y_first + 1) : (ymax - y_first + 3);
      const int y0 = fg ? y_first : y_first - 1;
      p.res_score[ko] = score;
      p.cand_off[ko] = off;
      p.cand_y0[ko] = y0;
      p.cand_nrows[ko] = nrows;
      const int* ext_l = t.ext_l + off;
      const int* ext_r = t.ext_r + off;
      if (nrows > kFastRows || W >= 16384 || H >= 16384) {
        const int slot = atomicAdd(&p.nbig[n], 1);
        if (slot < p.maxc) p.big[(size_t)n * p.maxc + slot] = k - 1;
        p.res_keep[ko] = 2;
        if (kSmem) {   // db_geometry_big_kernel reads the row extents from the global workspace
          int* gl = p.ext_l + (size_t)n * p.E + off;
          int* gr = p.ext_r + (size_t)n * p.E + off;
          for (int i = 0; i < nrows; ++i) {
            gl[i] = ext_l[i];
            gr[i] = ext_r[i];
          }
        }
        continue;
      }
      int* gout = reinterpret_cast<int*>(p.hull + ((size_t)n * p.E + off) * 4);   // >= 8 * (nrows + 1) ints
      int out[2 * kFastRows + 2];   // dynamically indexed => thread-local memory, which L1 caches write-back
      // monotone chain over (ext_l[i], y0+i), (ext_r[i], y0+i), sorted by (y, x): the first pass can only keep
      // RIGHT extents, the second only LEFT extents (dev_geom.cuh: hull_row_extents32); the two top-of-stack
      // points stay in registers
      const int last = nrows - 1;
      int kk = 0, a = 0, b = 0;   // a = out[kk-2], b = out[kk-1]
      int prev = pk(ext_l[0], y0);
      out[kk++] = prev;
      b = prev;
      for (int i = 0; i < nrows; ++i) {
        const int q = pk(ext_r[i], y0 + i);
        if (q == prev) continue;   // single-pixel first row
        prev = q;
        while (kk >= 2 && cross32(a, b, q) <= 0) {
          --kk;
          b = a;
          if (kk >= 2) a = out[kk - 2];
        }
        out[kk++] = q;
        a = b;
        b = q;
      }
      int hn = kk;
      if (kk > 1) {
        const int lo2 = kk + 1;
        for (int i = last; i >= 0; --i) {
          const int q = pk(ext_l[i], y0 + i);
          if (q == prev) continue;   // single-pixel last row
          prev = q;
          while (kk >= lo2 && cross32(a, b, q) <= 0) {
            --kk;
            b = a;
            a = out[kk - 2];
          }
          out[kk++] = q;
          a = b;
          b = q;
        }
        hn = kk - 1;
      }
      for (int i = 0; i < hn; ++i) gout[i] = out[i];
      p.hull_n[ko] = hn;
      p.res_keep[ko] = 3;
    }
  }

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
