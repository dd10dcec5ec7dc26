// PSENet progressive scale expansion and PAN / PAN++ pixel aggregation on sm_100a.
//
// Replaces, for a whole batch and without leaving the device:
//   R/pytocr/postprocess/pse_postprocess.py:28-53     upsample / sigmoid / threshold / text masking / D2H
//   R/pytocr/postprocess/pse_postprocess_fast/pse.pyx:13-69   cv2 CCL + FIFO multi-level expansion
//   R/pytocr/postprocess/pse_postprocess.py:55-105    label/score upsample + generate_box
//   R/pytocr/postprocess/pan_postprocess.py:30-113 and pan_postprocess_fast/pa.pyx:14-104 (same shape,
//     two masks, area-ratio flags, mean embeddings, one expansion level gated by embedding distance)
//
// The reference's expansion is defined by a sequential FIFO queue: a contested pixel belongs to
// whichever neighbour is popped first. That order is reproduced exactly (DESIGN.md "expansion"):
//   * pops are processed in blocks of up to kExThreads queue entries; every entry r proposes
//     key = 4*r + direction to its free neighbours with atomicMin, the smallest key owns the pixel,
//     and the owners append their pixels to the next wave with an ordered (scan) compaction, which
//     is exactly the order in which the sequential queue would have pushed them;
//   * entries that claimed nothing go, in pop order, to the next level's queue (pse.pyx:47,58-60);
//   * queue order only matters between pixels of ONE 4-connected component of the text mask (all
//     kernels are multiplied by it), so every text component is an independent work item handled by
//     one CTA; entries that can never claim again (no free text neighbour) are dropped, which does
//     not change the relative order of the others.
// Labels (ids of cv2.connectedComponents(connectivity=4)) = 1 + rank of the seed component's first
// raster pixel; inside the pipeline a label is carried as (root run index + 1) of the seed mask.
//
// Data layout per image at processing resolution H x W (H = h*upsample_in):
//   kb  u8  [H*W]   bit k = kernel k after text masking (PSE: k < K; PAN: bit0 text, bit1 kernel)
//   st  u32 [H*W]   state of text pixels: kUnset | proposal key (< 2^31) | kLabelBit (+ kGateBit) + label
//   1-bit masks of text and seed kernel -> run tables -> run union-find (4-connectivity)
// Algorithmic bytes per image: C*h*w*sizeof(elem), the head output, read once by ex_binarize_kernel
// (plus channel 0 again at text pixels for the score mean, and embeddings at gated pixels only).
#include "common.cuh"

#include <cstdlib>
#include <mutex>
#include "dev_common.cuh"
#include "dev_geom.cuh"
#include "geometry.cuh"

namespace ocrpp {
namespace {

constexpr uint32_t kUnset = 0xffffffffu;
constexpr uint32_t kLabelBit = 0x80000000u;
constexpr uint32_t kGateBit = 0x20000000u;    // PAN: the label is "flagged" (pa.pyx:42-54), claims are gated
constexpr uint32_t kLabelMask = 0x1fffffffu;  // label = root run index of the seed mask + 1
constexpr double kFix = 4294967296.0;  // 2^32

__device__ __forceinline__ bool is_labelled(uint32_t v) { return (v & kLabelBit) && v != kUnset; }

enum { kModePse = 0, kModePan = 1 };

struct ExParams {
  const void* maps;
  long long stride_n, stride_c, stride_h;
  const double* shape;  // [N,4] src_h, src_w, ratio_h, ratio_w
  int N, C, K, h, w, fin, fout, H, W, Wd, R, E, maxc, mode, seed_bit;
  int n0;  // first image of the sub-batch a launch covers (work lists / counters / arena are per sub-batch)
  float thresh, box_thresh, min_area_seed, min_area_box;
  long long arena_cap;
  // per-image maps
  uint8_t* kb;      // [N][H*W]
  uint32_t* st;     // [N][H*W]
  uint32_t* bits;   // [N][2][H*Wd]
  int32_t* rowptr;  // [N][2][H+1]
  int32_t* rowcnt;  // [N][2][H] run starts per row, written by ex_binarize_kernel
  uint16_t *run_xs, *run_xe, *run_y;  // [N][2][R]
  int32_t* par;     // [N][2][R]
  // text-root slots [N][R]
  int32_t *t_area, *t_xmin, *t_xmax, *t_ymax, *t_seen, *t_amin, *t_amax, *t_nseed, *t_lab;
  // seed-root slots [N][R]
  int32_t *s_area, *s_cc, *s_cid, *s_alive, *s_flag;
  double* s_emb;    // [N][R][4] embedding sums of flagged kernels
  int32_t *l_area, *l_ymin, *l_ymax, *l_rowoff;
  int32_t *l_rowbase, *l_row0;   // [R] row-extent block reserved at seeding time (rows of the seed's text component)
  int32_t* extdefer;              // [N] some label of the image could not reserve its block: second pass needed
  long long* l_sum;
  int32_t *ext_l, *ext_r;  // [N][E]
  P2i* hull;               // [N][8*E]
  int32_t *nruns;          // [N][2]
  int32_t *ext_alloc, *imgflags, *ncand;  // [N]
  int32_t* cand;           // [N][maxc]
  int32_t* res_keep;
  int16_t* res_box;
  float* res_boxf;
  float* res_score;
  // batch-global
  int2* work;              // [4][N*R] (image, text root) by tile class: tiny, small, big, huge
  int32_t* g_nwork;        // [4]
  int32_t* g_next;         // [4]
  unsigned long long* g_arena_used;  // [1]
  uint32_t* arena;         // [arena_cap]
  // outputs
  int16_t* boxes_out;
  float* scores_out;
  int32_t* counts_out;
  int32_t* status_out;
  float* boxes_f_out;
  int32_t* labels_dbg;
};

template <typename T>
__device__ __forceinline__ float ex_load(const ExParams& p, int n, int c, int y, int x) {
  // value of channel c at PROCESSING-resolution pixel (y,x): nearest upsample by fin
  // (F.interpolate(mode="nearest", scale_factor=fin) == src[y / fin][x / fin])
  const long long off = n * p.stride_n + c * p.stride_c + (long long)(y / p.fin) * p.stride_h + (x / p.fin);
  return load_scalar<T>(reinterpret_cast<const T*>(p.maps) + off);
}

// ------------------------------------------------------------------------------------------------
// E1: threshold + text masking -> kb bytes and the two 1-bit masks. One warp per row, 4 pixels per
// lane per iteration; 128-bit loads per channel when the input is at processing resolution.
// ------------------------------------------------------------------------------------------------
constexpr int kBinWarps = 8;
constexpr int kMaxK = 8;
constexpr int kTinyCap = 4096;    // expansion, tiny tiles: pixels of the padded bounding box (9 CTAs of 24 KB per SM)
constexpr int kSmallCap = 7936;   // small tiles (4 CTAs of 47 KB per SM)
constexpr int kBigCap = 18432;    // big tiles (merged text regions): 2 CTAs of 88 KB per SM

// kPre: 128-pixel groups whose loads are issued before the first of them is processed. With K channels a lane has
// K 128-bit loads in flight per group: 7 for PSE, but only 2 for PAN (3.5 KB vs 1 KB per warp) - too little
// memory-level parallelism to stream at the HBM rate, so few-channel inputs prefetch four groups (kKmax = 2).
template <typename T, bool kVec, int kPre, int kKmax>
__global__ void __launch_bounds__(kBinWarps * 32) ex_binarize_kernel(ExParams p) {
  const int n = blockIdx.y + p.n0;
  const int y = blockIdx.x * kBinWarps + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (y >= p.H) return;
  uint8_t* kb = p.kb + ((size_t)n * p.H + y) * p.W;
  uint32_t* tb = p.bits + ((size_t)(n * 2 + 0) * p.H + y) * p.Wd;
  uint32_t* sb = p.bits + ((size_t)(n * 2 + 1) * p.H + y) * p.Wd;
  const int K = p.K;
  unsigned tprev = 0, sprev = 0;   // last word of the previous 128-pixel group (run starts need its top bit)
  int tcnt = 0, scnt = 0;          // run starts of this row, counted by the lanes that own a word
  for (int xb = 0; xb < p.W; xb += 128 * kPre) {
    uint4 v[kPre][kKmax];
    if (kVec) {  // fin == 1, float rows 16-byte aligned, W % 4 == 0
#pragma unroll
      for (int u = 0; u < kPre; ++u) {
        const int x = xb + u * 128 + lane * 4;
        const float* base = reinterpret_cast<const float*>(p.maps) + n * p.stride_n + (long long)y * p.stride_h + x;
#pragma unroll
        for (int k = 0; k < kKmax; ++k)
          if (k < K && x < p.W) v[u][k] = kPre > 1 ? ldg_stream_u4_ordered(base + k * p.stride_c) : ldg_stream_u4(base + k * p.stride_c);
      }
    }
#pragma unroll
    for (int u = 0; u < kPre; ++u) {
    const int x0 = xb + u * 128;
    if (x0 >= p.W) break;
    const int x = x0 + lane * 4;
    unsigned b[4] = {0u, 0u, 0u, 0u};
    if (x < p.W) {
      if (kVec) {
#pragma unroll
        for (int k = 0; k < kKmax; ++k)
          if (k < K) {
            b[0] |= (__uint_as_float(v[u][k].x) > p.thresh ? 1u : 0u) << k;
            b[1] |= (__uint_as_float(v[u][k].y) > p.thresh ? 1u : 0u) << k;
            b[2] |= (__uint_as_float(v[u][k].z) > p.thresh ? 1u : 0u) << k;
            b[3] |= (__uint_as_float(v[u][k].w) > p.thresh ? 1u : 0u) << k;
          }
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (x + i < p.W)
            for (int k = 0; k < K; ++k) b[i] |= (ex_load<T>(p, n, k, y, x + i) > p.thresh ? 1u : 0u) << k;
      }
    }
    unsigned tn = 0, sn = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (!(b[i] & 1u)) b[i] = 0u;  // every kernel is multiplied by the text mask (pse_postprocess.py:41-42)
      tn |= (b[i] & 1u) << i;
      sn |= ((b[i] >> p.seed_bit) & 1u) << i;
    }
    if (x < p.W) {
      if (kVec) {
        *reinterpret_cast<uint32_t*>(kb + x) = b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (x + i < p.W) kb[x + i] = (uint8_t)b[i];
      }
    }
    unsigned tw = tn << (4 * (lane & 7)), sw = sn << (4 * (lane & 7));
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      tw |= __shfl_xor_sync(0xffffffffu, tw, o);
      sw |= __shfl_xor_sync(0xffffffffu, sw, o);
    }
    const int wi = (x0 >> 5) + (lane >> 3);
    // word before mine: the previous owner lane's word, or the last word of the previous group
    unsigned tp = __shfl_up_sync(0xffffffffu, tw, 8), sp = __shfl_up_sync(0xffffffffu, sw, 8);
    if (lane < 8) {
      tp = tprev;
      sp = sprev;
    }
    if ((lane & 7) == 0 && wi < p.Wd) {
      tb[wi] = tw;
      sb[wi] = sw;
      tcnt += __popc(tw & ~((tw << 1) | (tp >> 31)));
      scnt += __popc(sw & ~((sw << 1) | (sp >> 31)));
    }
    tprev = __shfl_sync(0xffffffffu, tw, 24);
    sprev = __shfl_sync(0xffffffffu, sw, 24);
    }
  }
  tcnt = __reduce_add_sync(0xffffffffu, tcnt);
  scnt = __reduce_add_sync(0xffffffffu, scnt);
  if (lane == 0) {
    p.rowcnt[(size_t)(n * 2 + 0) * p.H + y] = tcnt;
    p.rowcnt[(size_t)(n * 2 + 1) * p.H + y] = scnt;
  }
}

// ------------------------------------------------------------------------------------------------
// E2: 1-bit mask -> table of foreground runs in raster order (row counts come from ex_binarize_kernel).
// Several CTAs per (image, mask).
// ------------------------------------------------------------------------------------------------
constexpr int kRunThreads = 512;

__device__ __forceinline__ unsigned run_starts(unsigned w, unsigned prev) { return w & ~((w << 1) | (prev >> 31)); }
__device__ __forceinline__ unsigned run_ends(unsigned w, unsigned next) { return w & ~((w >> 1) | (next << 31)); }

// run count of mask m of image n for every kernel AFTER ex_runs_kernel: when either mask overflowed the run
// table, the other one's table is complete but the tables no longer describe the same image (a seed's text
// run is looked up in a table that was never written), so the image is treated as empty from here on - it is
// reported through OCRPP_IMG_RUN_OVERFLOW and its outputs are dropped anyway
__device__ __forceinline__ int ex_nruns(const ExParams& p, int n, int m) {
  return (p.imgflags[n] & OCRPP_IMG_RUN_OVERFLOW) ? 0 : p.nruns[n * 2 + m];
}

__global__ void __launch_bounds__(kRunThreads) ex_runs_kernel(ExParams p) {
  extern __shared__ int s_rowcnt[];  // [H+1]
  const int n = blockIdx.y + p.n0, m = blockIdx.z;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = kRunThreads / 32;
  const uint32_t* bits = p.bits + (size_t)(n * 2 + m) * p.H * p.Wd;
  const int32_t* rc = p.rowcnt + (size_t)(n * 2 + m) * p.H;
  int32_t* rowptr = p.rowptr + (size_t)(n * 2 + m) * (p.H + 1);
  const size_t ro = (size_t)(n * 2 + m) * p.R;
  const size_t so = (size_t)n * p.R;
  // every CTA of the image scans the row counts (cheap); each converts its share of the rows
  const int chunk = (p.H + kRunThreads - 1) / kRunThreads;
  const int y0 = threadIdx.x * chunk, y1 = min(p.H, y0 + chunk);
  int local = 0;
  for (int y = y0; y < y1; ++y) local += rc[y];
  int total;
  int base = block_exclusive_scan(local, &total);
  for (int y = y0; y < y1; ++y) {
    s_rowcnt[y] = base;
    base += rc[y];
  }
  if (threadIdx.x == 0) s_rowcnt[p.H] = total;
  __syncthreads();
  if (total > p.R) {
    if (blockIdx.x == 0) {
      if (threadIdx.x == 0) {
        atomicOr(&p.imgflags[n], OCRPP_IMG_RUN_OVERFLOW);
        p.nruns[n * 2 + m] = 0;
      }
      for (int y = threadIdx.x; y <= p.H; y += kRunThreads) rowptr[y] = 0;
    }
    return;
  }
  if (blockIdx.x == 0) {
    for (int y = threadIdx.x; y <= p.H; y += kRunThreads) rowptr[y] = s_rowcnt[y];
    if (threadIdx.x == 0) p.nruns[n * 2 + m] = total;
  }

  for (int y = blockIdx.x * nw + warp; y < p.H; y += gridDim.x * nw) {
    const int rbase = s_rowcnt[y];
    int carry_s = 0, carry_e = 0;
    for (int k0 = 0; k0 < p.Wd; k0 += 32) {
      const int k = k0 + lane;
      unsigned S = 0, Eb = 0;
      if (k < p.Wd) {
        const unsigned w = bits[(size_t)y * p.Wd + k];
        const unsigned pw = k > 0 ? bits[(size_t)y * p.Wd + k - 1] : 0u;
        const unsigned nx = k + 1 < p.Wd ? bits[(size_t)y * p.Wd + k + 1] : 0u;
        S = run_starts(w, pw);
        Eb = run_ends(w, nx);
      }
      const int cs = __popc(S), ce = __popc(Eb);
      int is = cs, ie = ce;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int ts = __shfl_up_sync(0xffffffffu, is, o), te = __shfl_up_sync(0xffffffffu, ie, o);
        if (lane >= o) {
          is += ts;
          ie += te;
        }
      }
      int js = carry_s + is - cs, je = carry_e + ie - ce;
      carry_s += __shfl_sync(0xffffffffu, is, 31);
      carry_e += __shfl_sync(0xffffffffu, ie, 31);
      while (S) {
        const int bpos = __ffs(S) - 1;
        S &= S - 1;
        const size_t r = ro + rbase + js;
        p.run_xs[r] = (uint16_t)(k * 32 + bpos);
        p.run_y[r] = (uint16_t)y;
        p.par[r] = rbase + js;
        const size_t q = so + rbase + js;
        if (m == 0) {
          p.t_area[q] = 0; p.t_xmin[q] = 0x7fffffff; p.t_xmax[q] = -1; p.t_ymax[q] = -1;
          p.t_seen[q] = 0; p.t_amin[q] = 0x7fffffff; p.t_amax[q] = 0; p.t_nseed[q] = 0; p.t_lab[q] = 0;
        } else {
          p.s_area[q] = 0; p.s_cc[q] = -1; p.s_cid[q] = 0; p.s_alive[q] = 0; p.s_flag[q] = 0;
          p.l_area[q] = 0; p.l_ymin[q] = 0x7fffffff; p.l_ymax[q] = -1; p.l_rowoff[q] = -1; p.l_rowbase[q] = -1; p.l_sum[q] = 0;
          if (p.mode == kModePan) {
            p.s_emb[q * 4 + 0] = 0.0; p.s_emb[q * 4 + 1] = 0.0; p.s_emb[q * 4 + 2] = 0.0; p.s_emb[q * 4 + 3] = 0.0;
          }
        }
        ++js;
      }
      while (Eb) {
        const int bpos = __ffs(Eb) - 1;
        Eb &= Eb - 1;
        p.run_xe[ro + rbase + je] = (uint16_t)(k * 32 + bpos);
        ++je;
      }
    }
  }
}

constexpr int kImgCtas = 8;   // minimum CTAs per image of the run-parallel kernels (more for small batches)
constexpr int kRunBlk = 256;
#define EX_FOR_EACH_RUN(r, nr) for (int r = blockIdx.x * kRunBlk + threadIdx.x; r < (nr); r += gridDim.x * kRunBlk)

// E3: 4-connectivity: link every run with the runs of the row above that overlap [xs, xe].
__global__ void __launch_bounds__(kRunBlk) ex_link_kernel(ExParams p) {
  const int n = blockIdx.y + p.n0, m = blockIdx.z;
  const int nr = ex_nruns(p, n, m);
  const size_t ro = (size_t)(n * 2 + m) * p.R;
  const int32_t* rowptr = p.rowptr + (size_t)(n * 2 + m) * (p.H + 1);
  const uint16_t *xs = p.run_xs + ro, *xe = p.run_xe + ro, *ry = p.run_y + ro;
  int32_t* par = p.par + ro;
  EX_FOR_EACH_RUN(r, nr) {
    const int y = ry[r];
    if (y == 0) continue;
    const int lo = xs[r], hi = xe[r];
    int l = rowptr[y - 1], h = rowptr[y];
    const int b = h;
    while (l < h) {
      const int mid = (l + h) >> 1;
      if ((int)xe[mid] < lo) l = mid + 1; else h = mid;
    }
    for (int q = l; q < b && (int)xs[q] <= hi; ++q) uf_union(par, q, r);
  }
}

// E4: flatten + per-root reductions (text: area and bounding box; seed: area).
__global__ void __launch_bounds__(kRunBlk) ex_flatten_kernel(ExParams p) {
  const int n = blockIdx.y + p.n0, m = blockIdx.z;
  const int nr = ex_nruns(p, n, m);
  const size_t ro = (size_t)(n * 2 + m) * p.R, so = (size_t)n * p.R;
  const uint16_t *xs = p.run_xs + ro, *xe = p.run_xe + ro, *ry = p.run_y + ro;
  int32_t* par = p.par + ro;
  EX_FOR_EACH_RUN(r, nr) {
    const int root = uf_find(par, r);
    par[r] = root;
    const int len = (int)xe[r] - (int)xs[r] + 1;
    if (m == 0) {
      atomicAdd(&p.t_area[so + root], len);
      atomicMin(&p.t_xmin[so + root], (int)xs[r]);
      atomicMax(&p.t_xmax[so + root], (int)xe[r]);
      atomicMax(&p.t_ymax[so + root], (int)ry[r]);
    } else {
      atomicAdd(&p.s_area[so + root], len);
    }
  }
}

// run of `mask m` in row y that contains pixel x (the pixel is known to be set in that mask)
__device__ __forceinline__ int ex_run_at(const int32_t* rowptr, const uint16_t* xs, int y, int x) {
  int l = rowptr[y], h = rowptr[y + 1];
  while (h - l > 1) {
    const int mid = (l + h) >> 1;
    if ((int)xs[mid] <= x) l = mid; else h = mid;
  }
  return l;
}

// E5: seed components: cv2 label ids, min-area filter (pse.pyx:21-23, pa.pyx:33-37), enclosing text
// component, work items, and (PAN) the per-text-component extreme kernel areas. One CTA per image.
__global__ void __launch_bounds__(kRunThreads) ex_seed_kernel(ExParams p) {
  const int n = blockIdx.x + p.n0;
  const int nr = ex_nruns(p, n, 1);
  const size_t to = (size_t)(n * 2 + 0) * p.R, ro = (size_t)(n * 2 + 1) * p.R, so = (size_t)n * p.R;
  const int32_t* t_rowptr = p.rowptr + (size_t)(n * 2 + 0) * (p.H + 1);
  const int chunk = (nr + kRunThreads - 1) / kRunThreads;
  const int lo = min(nr, (int)threadIdx.x * chunk), hi = min(nr, lo + chunk);
  int local = 0, need = 0;
  for (int r = lo; r < hi; ++r) local += p.par[ro + r] == r ? 1 : 0;
  int total;
  int id = block_exclusive_scan(local, &total);
  for (int r = lo; r < hi; ++r) {
    if (p.par[ro + r] != r) continue;
    p.s_cid[so + r] = ++id;
    const int area = p.s_area[so + r];
    if ((float)area < p.min_area_seed) continue;
    const int y = p.run_y[ro + r], x = p.run_xs[ro + r];
    const int cc = p.par[to + ex_run_at(t_rowptr, p.run_xs + to, y, x)];
    p.s_cc[so + r] = cc;
    p.s_alive[so + r] = 1;
    need += p.t_ymax[so + cc] - (int)p.run_y[to + cc] + 2;   // rows of the text component + 1 (see below)
    p.t_seen[so + cc] = 1;
    atomicAdd(&p.t_nseed[so + cc], 1);
    p.t_lab[so + cc] = r + 1;   // only read when the component owns exactly one seed
    if (p.mode == kModePan) {
      atomicMin(&p.t_amin[so + cc], area);
      atomicMax(&p.t_amax[so + cc], area);
    }
  }
  {
    // A label never leaves its text component, so the component's rows bound the label's rows: every surviving
    // seed reserves a row-extent block of that many rows now, and the ONE pass over the label map after the
    // expansion fills area, score sum, row range and row extents together (the exact per-label allocation needs
    // the row range first and with it a second pass; it remains for the seeds whose block does not fit).
    // Reservations use the first half of the extent storage only - the second half keeps the capacity the exact
    // allocation always had. This CTA is the image's only allocator at this point: block scan, no atomics.
    int reserved;
    int off = block_exclusive_scan(need, &reserved);
    const int limit = p.E / 2;
    for (int r = lo; r < hi; ++r) {
      if (p.par[ro + r] != r || !p.s_alive[so + r]) continue;
      const int cc = p.s_cc[so + r];
      const int y0 = p.run_y[to + cc], nrows = p.t_ymax[so + cc] - y0 + 1;
      if (off + nrows + 1 <= limit) {
        p.l_rowbase[so + r] = off;
        p.l_row0[so + r] = y0;
      } else {
        p.extdefer[n] = 1;
      }
      off += nrows + 1;
    }
    reserved = min(reserved, limit);
    if (threadIdx.x == 0) p.ext_alloc[n] = reserved;
    for (int i = threadIdx.x; i < reserved; i += kRunThreads) {
      p.ext_l[(size_t)n * p.E + i] = 0x7fffffff;
      p.ext_r[(size_t)n * p.E + i] = -1;
    }
  }
  __syncthreads();
  // PAN: a text component that owns exactly ONE surviving kernel needs no expansion at all: the only
  // level runs over the text mask itself (pa.pyx:72), every queued pixel tries all four neighbours, the
  // component is 4-connected, nobody contests it and a label is only gated when a second kernel shares
  // its component (pa.pyx:42-54), so every pixel ends with that kernel's label; ex_paint_kernel writes it
  // directly and only components with two or more kernels become work items of ex_expand_kernel.
  // PSE has no such shortcut: a pixel that claimed a neighbour is not revisited at later levels
  // (pse.pyx:47,58-60), so which text pixels stay unlabelled depends on the pop order even for one seed.
  const int nt = ex_nruns(p, n, 0);
  const int min_seeds = p.mode == kModePan ? 2 : 1;
  for (int r = threadIdx.x; r < nt; r += kRunThreads) {
    if (p.par[to + r] != r || p.t_nseed[so + r] < min_seeds) continue;
    const long long tile_px = (long long)(p.t_xmax[so + r] - p.t_xmin[so + r] + 3) *
                              (p.t_ymax[so + r] - (int)p.run_y[to + r] + 3);
    const int cls = tile_px <= kTinyCap ? 0 : (tile_px <= kSmallCap ? 1 : (tile_px <= kBigCap ? 2 : 3));
    p.work[(size_t)cls * p.N * p.R + atomicAdd(&p.g_nwork[cls], 1)] = make_int2(n, r);
  }
}

// E6 (PAN): area-ratio flags (pa.pyx:42-54; float32 rate vs 1024 == exact integer compare for areas
// < 2^20) and embedding sums of flagged kernels over their ORIGINAL pixels.
template <typename T>
__global__ void __launch_bounds__(kRunBlk) ex_pan_flag_kernel(ExParams p) {
  const int n = blockIdx.y + p.n0;
  const int nr = ex_nruns(p, n, 1);
  const size_t ro = (size_t)(n * 2 + 1) * p.R, so = (size_t)n * p.R;
  EX_FOR_EACH_RUN(r, nr) {
    const int root = p.par[ro + r];
    if (!p.s_alive[so + root]) continue;
    const long long a = p.s_area[so + root];
    const int cc = p.s_cc[so + root];
    const long long amin = p.t_amin[so + cc], amax = p.t_amax[so + cc];
    const bool flag = (a * 1024 < amax) || (a > amin * 1024);
    if (!flag) continue;
    if (r == root) p.s_flag[so + root] = 1;
    const int y = p.run_y[ro + r];
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    for (int x = p.run_xs[ro + r]; x <= (int)p.run_xe[ro + r]; ++x)
#pragma unroll
      for (int c = 0; c < 4; ++c) s[c] += (double)ex_load<T>(p, n, 2 + c, y, x);
#pragma unroll
    for (int c = 0; c < 4; ++c) atomicAdd(&p.s_emb[(so + root) * 4 + c], s[c]);
  }
}

// E7: paint the state map: every text pixel -> kUnset, or (PAN) directly the label when its component
// owns exactly one kernel (m == 0; components without a seed are painted too: the bounding-box scan of a
// neighbouring component reads them), then the surviving seed pixels -> their label (m == 1).
// One warp per run.
__global__ void __launch_bounds__(kRunBlk) ex_paint_kernel(ExParams p, int m) {
  const int n = blockIdx.y + p.n0;
  const int nr = ex_nruns(p, n, m);
  const size_t ro = (size_t)(n * 2 + m) * p.R, so = (size_t)n * p.R;
  uint32_t* st = p.st + (size_t)n * p.H * p.W;
  // 8 lanes per run (one 32-byte sector per store instruction): text runs are a few dozen pixels long and
  // the kernel is bound by the latency of the per-run look-ups, so four runs per warp are in flight
  const int gl = threadIdx.x & 7, gpb = kRunBlk / 8;
  for (int r = blockIdx.x * gpb + (threadIdx.x >> 3); r < nr; r += gridDim.x * gpb) {
    const int root = p.par[ro + r];
    uint32_t v;
    if (m == 0) {
      v = (p.mode == kModePan && p.t_nseed[so + root] == 1) ? (kLabelBit | (uint32_t)p.t_lab[so + root]) : kUnset;
    } else {
      if (!p.s_alive[so + root]) continue;
      v = kLabelBit | (p.s_flag[so + root] ? kGateBit : 0u) | (uint32_t)(root + 1);
    }
    const size_t row = (size_t)p.run_y[ro + r] * p.W;
    for (int x = p.run_xs[ro + r] + gl; x <= (int)p.run_xe[ro + r]; x += 8) st[row + x] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// E8: the expansion. Persistent CTAs pull text components from the work list. A component whose
// padded bounding box fits the shared-memory tile is expanded entirely in shared memory (state,
// kernel bits and the four queues); anything larger, or a queue overflow, takes the global-memory
// path, which has the same structure with the queues in a batch-wide arena.
// ------------------------------------------------------------------------------------------------
constexpr int kStrip = 8;
constexpr int kSmallThreads = 128;
constexpr int kBigThreads = 256;
constexpr int kHugeCap = 53248;   // huge tiles (many merged regions): 1 CTA of 224 KB per SM
constexpr int kHugeThreads = 512;
constexpr int kTinyThreads = 128;
constexpr int kTinyList = 1024;   // entries per queue, tiny tiles
constexpr int kListCap = 2048;    // entries per queue, small and big tiles
constexpr size_t ex_smem_bytes(int cap, int list) { return (size_t)cap * 4 + (size_t)list * 2 * 4; }

__device__ __forceinline__ uint32_t pack_yx(int y, int x) { return ((uint32_t)y << 16) | (uint32_t)x; }

struct ExItem {
  int n, a, x0, x1, y0, y1, area;
};

// pa.pyx:86-87: a flagged label only claims pixels whose embedding lies within distance 3 of its mean
struct ExMean {
  float m[4];
};
__device__ __forceinline__ ExMean ex_mean_emb(const ExParams& p, size_t sr) {   // float32 mean (pa.pyx:50,54)
  ExMean e;
  const double ar = (double)p.s_area[sr];
#pragma unroll
  for (int c = 0; c < 4; ++c) e.m[c] = (float)(p.s_emb[sr * 4 + c] / ar);
  return e;
}
template <typename T>
__device__ __forceinline__ bool ex_gate_blocks(const ExParams& p, int n, const ExMean& e, int ty, int tx) {
  float ss = 0.f;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float d = __fsub_rn(ex_load<T>(p, n, 2 + c, ty, tx), e.m[c]);
    ss = __fadd_rn(ss, __fmul_rn(d, d));
  }
  return __fsqrt_rn(ss) > 3.f;
}

// ---- shared-memory path. Returns false (nothing written to global memory) when a queue overflows.
// One 32-bit cell per pixel of the padded bounding box: kernel bits in the top byte, and in the low 24
// bits the state: kCellUnset | proposal key (< 2^23) | kCellLabel + label (+ kCellGated, + kCellClaimed
// when the label was written by this expansion). The top byte is the same for every proposal to a pixel, so
// atomicMin on the whole cell orders proposals by key.
constexpr uint32_t kCellUnset = 0x00ffffffu;
constexpr uint32_t kCellLabel = 0x00800000u;
constexpr uint32_t kCellClaimed = 0x00400000u;
constexpr uint32_t kCellGated = 0x00200000u;     // copy of kGateBit
constexpr uint32_t kCellLabMask = 0x001fffffu;   // labels (root run index + 1) must stay below 2^21

__device__ __forceinline__ bool cell_labelled(uint32_t c) {
  return (c & kCellLabel) && (c & 0x00ffffffu) != kCellUnset;
}

// exclusive prefix over the CTA (thread order) of a packed pair of small counters; `s_tot` holds one
// slot per warp and is double-buffered by the caller so that a single barrier per call is enough
template <int kThreads>
__device__ __forceinline__ int ex_block_prefix(int v, int* s_tot, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_tot[warp] = inc;
  __syncthreads();
  int before = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) {
    const int t = s_tot[w];
    before += w < warp ? t : 0;
    tot += t;
  }
  *total = tot;
  return before + inc - v;
}

template <typename T, int kTileCap, int kListCap, int kExThreads>
__device__ bool ex_expand_tile(const ExParams& p, const ExItem& it, unsigned char* smem) {
  __shared__ int s_tot[2][kExThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = kExThreads / 32;
  const int W = p.W, H = p.H;
  const int tw = it.x1 - it.x0 + 3, th = it.y1 - it.y0 + 3;  // 1-pixel border of non-text
  uint32_t* cell = reinterpret_cast<uint32_t*>(smem);
  uint16_t* lists = reinterpret_cast<uint16_t*>(cell + kTileCap);
  const size_t so = (size_t)it.n * p.R;
  const uint8_t* kb = p.kb + (size_t)it.n * H * W;
  uint32_t* st = p.st + (size_t)it.n * H * W;
  // load the tile (rows are contiguous in global memory; both loads are issued unconditionally)
  // (four 32-pixel chunks per step, all eight loads in flight before the first cell is written: a big tile is
  // hundreds of steps per warp and each used to wait for its own two loads)
  constexpr int kLoadU = 4;
  for (int ty = warp; ty < th; ty += nwarps) {
    const int gy = it.y0 - 1 + ty;
    const bool row_in = ty > 0 && ty < th - 1;
    for (int tx0 = lane; tx0 < tw; tx0 += 32 * kLoadU) {
      unsigned kv[kLoadU];
      uint32_t sv[kLoadU];
#pragma unroll
      for (int u = 0; u < kLoadU; ++u) {
        const int tx = tx0 + 32 * u;
        kv[u] = 0u;
        sv[u] = 0u;
        if (row_in && tx > 0 && tx < tw - 1) {   // interior = inside the image
          const size_t g = (size_t)gy * W + (it.x0 - 1 + tx);
          kv[u] = kb[g];
          sv[u] = __ldcg(st + g);                // garbage where kv has no text bit
        }
      }
#pragma unroll
      for (int u = 0; u < kLoadU; ++u) {
        const int tx = tx0 + 32 * u;
        if (tx >= tw) break;
        uint32_t c = 0;
        if (kv[u] & 1u)
          c = (kv[u] << 24) | (is_labelled(sv[u]) ? (kCellLabel | ((sv[u] & kGateBit) ? kCellGated : 0u) | (sv[u] & kCellLabMask))
                                                  : kCellUnset);
        cell[ty * tw + tx] = c;
      }
    }
  }
  __syncthreads();
  const int nb4[4] = {-tw, tw, -1, 1};  // pse.pyx:29-30: up, down, left, right
  uint16_t *Qc = lists, *Qn = lists + kListCap, *X = lists + 2 * kListCap, *Y = lists + 3 * kListCap;
  int flip = 0;
  // initial queue: surviving seed pixels of this component in raster order that still have a free
  // text neighbour
  int nq = 0;
  {
    const int bw = tw - 2;
    const int spr = (bw + kStrip - 1) / kStrip;
    const int total = spr * (th - 2);
    for (int s0 = 0; s0 < total; s0 += kExThreads) {
      const int s = s0 + tid;
      unsigned mask = 0;
      int base = 0;
      if (s < total) {
        const int ty = 1 + s / spr, tx0 = 1 + (s % spr) * kStrip;
        base = ty * tw + tx0;
#pragma unroll
        for (int i = 0; i < kStrip; ++i) {
          if (tx0 + i > bw) break;
          const int q = base + i;
          const uint32_t v = cell[q];
          if (!cell_labelled(v)) continue;
          if (p.s_cc[so + (v & kCellLabMask) - 1] != it.a) continue;
          bool alive = false;
#pragma unroll
          for (int j = 0; j < 4; ++j) alive |= (cell[q + nb4[j]] & 0x01ffffffu) == (0x01000000u | kCellUnset);
          if (alive) mask |= 1u << i;
        }
      }
      int tot;
      int pos = nq + ex_block_prefix<kExThreads>(__popc(mask), s_tot[flip], &tot);
      flip ^= 1;
      if (nq + tot > kListCap) return false;   // uniform
      while (mask) {
        const int i = __ffs(mask) - 1;
        mask &= mask - 1;
        Qc[pos++] = (uint16_t)(base + i);
      }
      nq += tot;
    }
    __syncthreads();
  }
  for (int level = (p.mode == kModePse ? p.K - 2 : 0); level >= 0 && nq > 0; --level) {
    const uint16_t* wave = Qc;
    int nw = nq, nqn = 0;
    uint16_t* nxt = X;
    const uint32_t lvl_bit = 0x01000000u << level;
    while (nw > 0) {
      if (nw <= 32) {
        // Small frontier (the long tail of thin / merged components, where a wave is a handful of pixels): warp 0
        // runs the waves alone, with warp-level synchronisation and a shuffle prefix, for as long as they fit a
        // warp - a wave then costs a few hundred cycles instead of three block barriers. Same claims, same order.
        __shared__ int s_small[5];
        if (warp == 0) {
          const uint16_t* w_ = wave;
          uint16_t* n_ = nxt;
          int nw_ = nw, nqn_ = nqn;
          bool over = false;
          while (nw_ > 0 && nw_ <= 32) {
            const int r = lane;
            const bool active = r < nw_;
            int q = 0;
            uint32_t lab = 0;
            if (active) {
              q = w_[r];
              lab = cell[q] & (kCellLabMask | kCellGated);
              const bool gated = (lab & kCellGated) != 0;
              ExMean mean{};
              if (gated) mean = ex_mean_emb(p, so + (lab & kCellLabMask) - 1);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int t = q + nb4[j];
                const uint32_t c = reinterpret_cast<volatile uint32_t*>(cell)[t];
                if (!(c & lvl_bit) || cell_labelled(c)) continue;
                if (gated) {
                  const int ty = t / tw, tx = t - ty * tw;
                  if (ex_gate_blocks<T>(p, it.n, mean, it.y0 - 1 + ty, it.x0 - 1 + tx)) continue;
                }
                atomicMin(cell + t, (c & 0xff000000u) | (uint32_t)(r * 4 + j));
              }
            }
            __syncwarp();
            unsigned win = 0;
            bool alive = false;
            if (active) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint32_t c = reinterpret_cast<volatile uint32_t*>(cell)[q + nb4[j]];
                if (!(c & 0x01000000u)) continue;
                const uint32_t pay = c & 0x00ffffffu;
                if (pay == kCellUnset) alive = true;
                else if (pay == (uint32_t)(r * 4 + j)) win |= 1u << j;
              }
            }
            const bool requeue = active && win == 0 && alive;
            const int mine = __popc(win);
            int inc = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
              const int t = __shfl_up_sync(0xffffffffu, inc, o);
              if (lane >= o) inc += t;
            }
            const int totw = __shfl_sync(0xffffffffu, inc, 31);
            const unsigned rq = __ballot_sync(0xffffffffu, requeue);
            if (totw > kListCap || nqn_ + __popc(rq) > kListCap) {
              over = true;
              break;
            }
            if (active) {
              int pos = inc - mine;
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (win & (1u << j)) {
                  const int t = q + nb4[j];
                  n_[pos++] = (uint16_t)t;
                  cell[t] = (cell[t] & 0xff000000u) | kCellLabel | kCellClaimed | lab;
                }
              if (requeue) Qn[nqn_ + __popc(rq & ((1u << lane) - 1u))] = (uint16_t)q;
            }
            nqn_ += __popc(rq);
            __syncwarp();
            w_ = n_;
            nw_ = totw;
            n_ = (n_ == X) ? Y : X;
          }
          if (lane == 0) {
            s_small[0] = nw_;
            s_small[1] = nqn_;
            s_small[2] = w_ == X ? 1 : (w_ == Y ? 2 : 0);
            s_small[3] = n_ == X ? 1 : 2;
            s_small[4] = over ? 1 : 0;
          }
        }
        __syncthreads();
        const int sw = s_small[2];
        nw = s_small[0];
        nqn = s_small[1];
        if (sw) wave = sw == 1 ? X : Y;
        nxt = s_small[3] == 1 ? X : Y;
        const bool over = s_small[4] != 0;
        __syncthreads();   // s_small may be rewritten by the next small phase
        if (over) return false;
        continue;
      }
      int nn = 0;
      for (int c0 = 0; c0 < nw; c0 += kExThreads) {
        const int r = c0 + tid;
        const bool active = r < nw;
        int q = 0;
        uint32_t lab = 0;
        if (active) {
          q = wave[r];
          lab = cell[q] & (kCellLabMask | kCellGated);
          const bool gated = (lab & kCellGated) != 0;
          ExMean mean{};
          if (gated) mean = ex_mean_emb(p, so + (lab & kCellLabMask) - 1);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int t = q + nb4[j];
            const uint32_t c = reinterpret_cast<volatile uint32_t*>(cell)[t];
            if (!(c & lvl_bit) || cell_labelled(c)) continue;
            if (gated) {
              const int ty = t / tw, tx = t - ty * tw;
              if (ex_gate_blocks<T>(p, it.n, mean, it.y0 - 1 + ty, it.x0 - 1 + tx)) continue;
            }
            atomicMin(cell + t, (c & 0xff000000u) | (uint32_t)(r * 4 + j));
          }
        }
        __syncthreads();
        unsigned win = 0;
        bool alive = false;
        if (active) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t c = cell[q + nb4[j]];
            if (!(c & 0x01000000u)) continue;
            const uint32_t pay = c & 0x00ffffffu;
            if (pay == kCellUnset) alive = true;
            else if (pay == (uint32_t)(r * 4 + j)) win |= 1u << j;
          }
        }
        const bool requeue = active && win == 0 && alive;
        int tot;
        const int ex = ex_block_prefix<kExThreads>(__popc(win) | (requeue ? 0x10000 : 0), s_tot[flip], &tot);
        flip ^= 1;
        if (nn + (tot & 0xffff) > kListCap || nqn + (tot >> 16) > kListCap) return false;  // uniform
        if (active) {
          int pos = nn + (ex & 0xffff);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (win & (1u << j)) {
              const int t = q + nb4[j];
              nxt[pos++] = (uint16_t)t;
              cell[t] = (cell[t] & 0xff000000u) | kCellLabel | kCellClaimed | lab;
            }
          if (requeue) Qn[nqn + (ex >> 16)] = (uint16_t)q;
        }
        nn += tot & 0xffff;
        nqn += tot >> 16;
        __syncthreads();
      }
      wave = nxt;
      nw = nn;
      nxt = (nxt == X) ? Y : X;
    }
    uint16_t* tmp = Qc;
    Qc = Qn;
    Qn = tmp;
    nq = nqn;
  }
  __syncthreads();
  // write the labels claimed by this expansion back to the global state map
  for (int ty = 1 + warp; ty < th - 1; ty += nwarps) {
    const size_t grow = (size_t)(it.y0 - 1 + ty) * W + (it.x0 - 1);
    for (int tx = 1 + lane; tx < tw - 1; tx += 32) {
      const uint32_t c = cell[ty * tw + tx];
      if ((c & kCellClaimed) && cell_labelled(c))
        st[grow + tx] = kLabelBit | ((c & kCellGated) ? kGateBit : 0u) | (c & kCellLabMask);
    }
  }
  return true;
}

// ---- global-memory path (any size)
template <typename T, int kExThreads>
__device__ void ex_expand_global(const ExParams& p, const ExItem& it) {
  __shared__ long long s_base;
  const int tid = threadIdx.x;
  const int H = p.H, W = p.W;
  const int dy4[4] = {-1, 1, 0, 0}, dx4[4] = {0, 0, -1, 1};  // pse.pyx:29-30: up, down, left, right
  const int n = it.n, a = it.a, x0 = it.x0, x1 = it.x1, y0 = it.y0, y1 = it.y1, area = it.area;
  const size_t so = (size_t)n * p.R;
  __syncthreads();
  if (tid == 0) {
    const unsigned long long b = atomicAdd(p.g_arena_used, 4ull * (unsigned long long)area);
    if ((long long)(b + 4ull * area) <= p.arena_cap) s_base = (long long)b;
    else {
      s_base = -1;
      atomicOr(&p.imgflags[n], OCRPP_IMG_RUN_OVERFLOW);
    }
  }
  __syncthreads();
  if (s_base < 0) return;
  uint32_t* Qc = p.arena + s_base;
  uint32_t* Qn = Qc + area;
  uint32_t* X = Qn + area;
  uint32_t* Y = X + area;
  const uint8_t* kb = p.kb + (size_t)n * H * W;
  uint32_t* st = p.st + (size_t)n * H * W;

  // ---- initial queue: surviving seed pixels of this text component in raster order (pse.pyx:33-37),
  //      minus the ones that have no free text neighbour (they can never claim anything)
  int nq = 0;
  {
    const int spr = (x1 - x0 + kStrip) / kStrip;
    const int total = spr * (y1 - y0 + 1);
    for (int s0 = 0; s0 < total; s0 += kExThreads) {
      const int s = s0 + tid;
      unsigned mask = 0;
      int y = 0, xb = 0;
      if (s < total) {
        y = y0 + s / spr;
        xb = x0 + (s % spr) * kStrip;
#pragma unroll
        for (int i = 0; i < kStrip; ++i) {
          const int x = xb + i;
          if (x > x1) break;
          const size_t q = (size_t)y * W + x;
          if (!(kb[q] & 1u)) continue;
          const uint32_t v = __ldcg(st + q);
          if (!is_labelled(v)) continue;
          if (p.s_cc[so + (v & kLabelMask) - 1] != a) continue;
          bool alive = false;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int ty = y + dy4[j], tx = x + dx4[j];
            if (ty < 0 || ty >= H || tx < 0 || tx >= W) continue;
            const size_t t = (size_t)ty * W + tx;
            if ((kb[t] & 1u) && __ldcg(st + t) == kUnset) alive = true;
          }
          if (alive) mask |= 1u << i;
        }
      }
      int tot;
      int pos = nq + block_exclusive_scan(__popc(mask), &tot);
      while (mask) {
        const int i = __ffs(mask) - 1;
        mask &= mask - 1;
        Qc[pos++] = pack_yx(y, xb + i);
      }
      nq += tot;
    }
    __syncthreads();  // queue entries visible to every thread of the CTA
  }

  // ---- levels (PSE: K-2 .. 0, level K-1 is a no-op that only re-queues every seed in order,
  //      SURVEY H2; PAN: the text mask only, pa.pyx:72)
  for (int level = (p.mode == kModePse ? p.K - 2 : 0); level >= 0 && nq > 0; --level) {
    const uint32_t* wave = Qc;
    int nw = nq, nqn = 0;
    uint32_t* nxt = X;
    while (nw > 0) {
      int nn = 0;
      for (int c0 = 0; c0 < nw; c0 += kExThreads) {
        const int r = c0 + tid;
        const bool active = r < nw;
        int qy = 0, qx = 0;
        uint32_t lab = 0, q = 0;
        if (active) {
          q = wave[r];
          qy = q >> 16;
          qx = q & 0xffffu;
          lab = __ldcg(st + (size_t)qy * W + qx);
          // phase 1: propose key 4r+j to every free neighbour that is inside kernel `level`
          const bool gated = (lab & kGateBit) != 0;
          ExMean mean{};
          if (gated) mean = ex_mean_emb(p, so + (lab & kLabelMask) - 1);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int ty = qy + dy4[j], tx = qx + dx4[j];
            if (ty < 0 || ty >= H || tx < 0 || tx >= W) continue;
            const size_t t = (size_t)ty * W + tx;
            if (!((kb[t] >> level) & 1u)) continue;
            if (is_labelled(__ldcg(st + t))) continue;
            if (gated && ex_gate_blocks<T>(p, n, mean, ty, tx)) continue;
            atomicMin(st + t, (uint32_t)(r * 4 + j));
          }
        }
        __syncthreads();
        // phase 2: owners collect their pixels; alive = some text neighbour is still free
        unsigned win = 0;
        bool alive = false;
        if (active) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int ty = qy + dy4[j], tx = qx + dx4[j];
            if (ty < 0 || ty >= H || tx < 0 || tx >= W) continue;
            const size_t t = (size_t)ty * W + tx;
            if (!(kb[t] & 1u)) continue;
            const uint32_t v = __ldcg(st + t);
            if (v == kUnset) alive = true;
            else if (v == (uint32_t)(r * 4 + j)) win |= 1u << j;
          }
        }
        const bool requeue = active && win == 0 && alive;  // is_edge (pse.pyx:47,58-60)
        int tot;
        const int ex = block_exclusive_scan(__popc(win) | (requeue ? 0x10000 : 0), &tot);
        if (active) {
          int pos = nn + (ex & 0xffff);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (win & (1u << j)) {
              const int ty = qy + dy4[j], tx = qx + dx4[j];
              nxt[pos++] = pack_yx(ty, tx);
              st[(size_t)ty * W + tx] = lab;
            }
          if (requeue) Qn[nqn + (ex >> 16)] = q;
        }
        nn += tot & 0xffff;
        nqn += tot >> 16;
        __syncthreads();  // labels visible to the next block of pops
      }
      wave = nxt;
      nw = nn;
      nxt = (nxt == X) ? Y : X;
    }
    uint32_t* tmp = Qc;
    Qc = Qn;
    Qn = tmp;
    nq = nqn;
  }
}

// kClass selects the work list: 0 tiny, 1 small, 2 big, 3 huge and beyond (by the pixel count of the
// padded bounding box)
template <typename T, int kTileCap, int kList, int kExThreads, int kClass>
__global__ void __launch_bounds__(kExThreads) ex_expand_kernel(ExParams p) {
  extern __shared__ __align__(16) unsigned char ex_smem[];
  __shared__ int s_item;
  const int tid = threadIdx.x;
  const int2* work = p.work + (size_t)kClass * p.N * p.R;
  int32_t* next = p.g_next + kClass;
  const int nwork = p.g_nwork[kClass];
  const bool tile_ok = p.R < (1 << 21);   // labels must fit the 21-bit cell field
  while (true) {
    __syncthreads();
    if (tid == 0) s_item = atomicAdd(next, 1);
    __syncthreads();
    const int item = s_item;
    if (item >= nwork) return;
    ExItem it;
    it.n = work[item].x;
    it.a = work[item].y;
    const size_t to = (size_t)(it.n * 2 + 0) * p.R, so = (size_t)it.n * p.R;
    it.x0 = p.t_xmin[so + it.a];
    it.x1 = p.t_xmax[so + it.a];
    it.y0 = p.run_y[to + it.a];
    it.y1 = p.t_ymax[so + it.a];
    it.area = p.t_area[so + it.a];
    const long long tile_px = (long long)(it.x1 - it.x0 + 3) * (it.y1 - it.y0 + 3);
    if (tile_ok && tile_px <= kTileCap && ex_expand_tile<T, kTileCap, kList, kExThreads>(p, it, ex_smem)) continue;
    ex_expand_global<T, kExThreads>(p, it);
  }
}

// ------------------------------------------------------------------------------------------------
// E9: per-label reductions over the final label map, one warp per text run.
//   PASS 1: area, sum of sigmoid(text logit) in 32.32 fixed point, first/last row
//   PASS 2: row extents (the hull of a label only needs its leftmost/rightmost pixel per row)
// ------------------------------------------------------------------------------------------------
template <typename T, int PASS>
__global__ void __launch_bounds__(kRunBlk) ex_stats_kernel(ExParams p) {
  const int n = blockIdx.y + p.n0;
  const int nr = ex_nruns(p, n, 0);
  const size_t ro = (size_t)(n * 2 + 0) * p.R, so = (size_t)n * p.R;
  const uint32_t* st = p.st + (size_t)n * p.H * p.W;
  const int lane = threadIdx.x & 31, wpb = kRunBlk / 32;
  if (PASS == 2 && !p.extdefer[n]) return;   // every label of the image got its extents in pass 1
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < nr; r += gridDim.x * wpb) {
    if (!p.t_seen[so + p.par[ro + r]]) continue;
    const int y = p.run_y[ro + r], a = p.run_xs[ro + r], b = p.run_xe[ro + r];
    for (int xb = a; xb <= b; xb += 32) {
      const int x = xb + lane;
      uint32_t lab = 0;
      unsigned long long fx = 0;
      if (x <= b) {
        const uint32_t v = st[(size_t)y * p.W + x];
        if (is_labelled(v)) {
          lab = v & kLabelMask;
          if (PASS == 1) {
            const float logit = ex_load<T>(p, n, 0, y, x);
            const float sc = 1.f / (1.f + expf(-logit));  // F.sigmoid (pse_postprocess.py:38)
            fx = (unsigned long long)__float2ll_rn(sc * 4294967296.0f);
          }
        }
      }
      unsigned todo = __ballot_sync(0xffffffffu, lab != 0);
      while (todo) {
        const int leader = __ffs(todo) - 1;
        const uint32_t L = __shfl_sync(0xffffffffu, lab, leader);
        const unsigned mk = __ballot_sync(0xffffffffu, lab == L);
        if (lab == L) {
          const size_t q = so + L - 1;
          if (PASS == 1) {
            const unsigned lo = __reduce_add_sync(mk, (unsigned)(fx & 0xffffu));
            const unsigned hi = __reduce_add_sync(mk, (unsigned)(fx >> 16));
            const int xmin = __reduce_min_sync(mk, x), xmax = __reduce_max_sync(mk, x);
            if (lane == leader) {
              atomicAdd(&p.l_area[q], __popc(mk));
              atomicAdd((unsigned long long*)&p.l_sum[q], ((unsigned long long)hi << 16) + lo);
              atomicMin(&p.l_ymin[q], y);
              atomicMax(&p.l_ymax[q], y);
              const int base = p.l_rowbase[q];
              if (base >= 0) {
                const size_t e = (size_t)n * p.E + base + (y - p.l_row0[q]);
                atomicMin(&p.ext_l[e], xmin);
                atomicMax(&p.ext_r[e], xmax);
              }
            }
          } else if (p.l_rowbase[q] < 0) {
            const int xmin = __reduce_min_sync(mk, x), xmax = __reduce_max_sync(mk, x);
            const int off = p.l_rowoff[q];
            if (lane == leader && off >= 0) {
              const size_t e = (size_t)n * p.E + off + (y - p.l_ymin[q]);
              atomicMin(&p.ext_l[e], xmin);
              atomicMax(&p.ext_r[e], xmax);
            }
          }
        }
        todo &= ~mk;
      }
    }
  }
}

// E10: candidates = surviving labels in label order (generate_box iterates i = 1..max(label),
// pse_postprocess.py:70); row-extent slots. One CTA per image.
__global__ void __launch_bounds__(kRunThreads) ex_cand_kernel(ExParams p) {
  const int n = blockIdx.x + p.n0;
  const int nr = ex_nruns(p, n, 1);
  const size_t ro = (size_t)(n * 2 + 1) * p.R, so = (size_t)n * p.R;
  const int chunk = (nr + kRunThreads - 1) / kRunThreads;
  const int lo = min(nr, (int)threadIdx.x * chunk), hi = min(nr, lo + chunk);
  auto is_cand = [&](int r) { return p.par[ro + r] == r && p.s_alive[so + r] && p.l_area[so + r] > 0; };
  int local = 0;
  for (int r = lo; r < hi; ++r) local += is_cand(r) ? 1 : 0;
  int total;
  int rank = block_exclusive_scan(local, &total);
  for (int r = lo; r < hi; ++r) {
    if (!is_cand(r)) continue;
    if (rank < p.maxc) {
      p.cand[(size_t)n * p.maxc + rank] = r;
      const int nrows = p.l_ymax[so + r] - p.l_ymin[so + r] + 1;
      if (p.l_rowbase[so + r] >= 0) {   // reserved at seeding time and already filled: rows start at the label's first row
        p.l_rowoff[so + r] = p.l_rowbase[so + r] + (p.l_ymin[so + r] - p.l_row0[so + r]);
        ++rank;
        continue;
      }
      const int off = atomicAdd(&p.ext_alloc[n], nrows + 1);
      if (off + nrows + 1 <= p.E) {
        p.l_rowoff[so + r] = off;
        for (int i = 0; i < nrows; ++i) {
          p.ext_l[(size_t)n * p.E + off + i] = 0x7fffffff;
          p.ext_r[(size_t)n * p.E + off + i] = -1;
        }
      } else {
        atomicOr(&p.imgflags[n], OCRPP_IMG_RUN_OVERFLOW);
      }
    }
    ++rank;
  }
  if (threadIdx.x == 0) {
    p.ncand[n] = min(total, p.maxc);
    if (total > p.maxc) atomicOr(&p.imgflags[n], OCRPP_IMG_CANDIDATES_TRUNCATED);
  }
}

// E11: generate_box (pse_postprocess.py:65-105 / pan_postprocess.py:73-113): the area and score filters,
// then hull -> min-area rectangle -> corner order -> rescale for labels of up to kFastRows rows, FOUR
// labels per warp (groups of 8 lanes, 32-bit integer projections, points packed in shared memory).
constexpr int kFastRows = 128;
constexpr int kGeoThreads = 16 * kGrp;

__device__ __forceinline__ void ex_emit_box(const ExParams& p, int n, size_t ko, const geom::Rect& rect, float score) {
  double bx[4], by[4];
  geom::cv_box_order(rect, bx, by);  // corner order of cv2.boxPoints
  float cx[4], cy[4], ox[4], oy[4];
  for (int q = 0; q < 4; ++q) {
    cx[q] = (float)bx[q];
    cy[q] = (float)by[q];
  }
  geom::order_points_clockwise(cx, cy, ox, oy);  // utility.py:21-29
  const double src_h = p.shape[4 * n + 0], src_w = p.shape[4 * n + 1];
  const double ratio_h = p.shape[4 * n + 2], ratio_w = p.shape[4 * n + 3];
  for (int q = 0; q < 4; ++q) {  // :100-102 (float64 division under numpy 2, np.round = half to even)
    const double fx = (double)ox[q] / ratio_w, fy = (double)oy[q] / ratio_h;
    p.res_boxf[ko * 8 + 2 * q] = (float)fx;
    p.res_boxf[ko * 8 + 2 * q + 1] = (float)fy;
    p.res_box[ko * 8 + 2 * q] = (int16_t)(int)fmin(fmax(geom::round_half_even(fx), 0.0), src_w);
    p.res_box[ko * 8 + 2 * q + 1] = (int16_t)(int)fmin(fmax(geom::round_half_even(fy), 0.0), src_h);
  }
  p.res_score[ko] = score;
  p.res_keep[ko] = 1;
}

__global__ void __launch_bounds__(kGeoThreads) ex_geometry_fast_kernel(ExParams p) {
  constexpr int kGroups = kGeoThreads / kGrp;
  __shared__ int s_a[kGroups][2 * kFastRows];
  __shared__ int s_b[kGroups][2 * kFastRows + 2];
  const int n = blockIdx.y + p.n0;
  const int g = threadIdx.x / kGrp, gl = threadIdx.x % kGrp;
  const unsigned gmask = ((1u << kGrp) - 1u) << ((threadIdx.x & 31) / kGrp * kGrp);
  const int k = blockIdx.x * kGroups + g;
  if (k >= p.ncand[n]) return;
  const size_t so = (size_t)n * p.R;
  const size_t ko = (size_t)n * p.maxc + k;
  const int c = p.cand[ko];
  const int f = p.fout;
  const long long area = (long long)p.l_area[so + c] * f * f;  // label upsampled by fout (:58-62)
  int verdict = -1;                                            // 0 drop, 2 defer, -1 go on
  if ((double)area < (double)p.min_area_box) verdict = 0;      // points.shape[0] < min_area (:75)
  // np.mean(score[ind]) (:79): float32 pairwise sum in the reference, exact fixed point here
  const float score = (float)(((double)p.l_sum[so + c] / kFix) / (double)p.l_area[so + c]);
  if (verdict < 0 && score < p.box_thresh) verdict = 0;        // :80
  const int ymin = p.l_ymin[so + c], nrows = p.l_ymax[so + c] - ymin + 1;
  const int off = p.l_rowoff[so + c];
  if (verdict < 0 && (f != 1 || nrows > kFastRows || off < 0 || p.W >= 16384 || p.H >= 16384)) verdict = 2;
  if (verdict >= 0) {
    if (gl == 0) p.res_keep[ko] = verdict;
    return;
  }
  const int32_t* ext_l = p.ext_l + (size_t)n * p.E + off;
  const int32_t* ext_r = p.ext_r + (size_t)n * p.E + off;
  int* A = s_a[g];
  int* B = s_b[g];
  for (int i = gl; i < nrows; i += kGrp) {
    A[2 * i] = pk(ext_l[i], ymin + i);
    A[2 * i + 1] = pk(ext_r[i], ymin + i);
  }
  __syncwarp(gmask);
  int hn = 0;
  if (gl == 0) hn = hull_row_extents32(A, nrows, B);
  hn = __shfl_sync(gmask, hn, 0, kGrp);
  __syncwarp(gmask);
  geom::Rect rect;
  group_min_area_rect(B, hn, &rect, gl, gmask);   // cv2.minAreaRect + boxPoints (:85-86)
  if (gl == 0) ex_emit_box(p, n, ko, rect, score);
}

// E11b: generate_box for the candidates the fast kernel below defers (more than kFastRows rows, an
// up-sampled label map, or coordinates beyond the packed 16-bit range), one warp per label.
constexpr int kGeoWarps = 4;
constexpr int kSmallRows = 64;

__global__ void __launch_bounds__(kGeoWarps * 32) ex_geometry_kernel(ExParams p) {
  __shared__ P2i s_pts[kGeoWarps][4 * kSmallRows];
  __shared__ P2i s_hull[kGeoWarps][4 * kSmallRows + 2];
  const int n = blockIdx.y + p.n0;
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t so = (size_t)n * p.R;
  const int nc = p.ncand[n];
  const int f = p.fout;
  for (int k = blockIdx.x * kGeoWarps + wib; k < nc; k += gridDim.x * kGeoWarps) {
    const size_t ko = (size_t)n * p.maxc + k;
    const int c = p.cand[ko];
    if (p.res_keep[ko] != 2) continue;   // decided by ex_geometry_fast_kernel
    __syncwarp();
    if (lane == 0) p.res_keep[ko] = 0;
    const float score = (float)(((double)p.l_sum[so + c] / kFix) / (double)p.l_area[so + c]);
    const int off = p.l_rowoff[so + c];
    if (off < 0) continue;  // extent arena exhausted: flagged in ex_cand_kernel
    const int ymin = p.l_ymin[so + c], nrows = p.l_ymax[so + c] - ymin + 1;
    const int32_t* ext_l = p.ext_l + (size_t)n * p.E + off;
    const int32_t* ext_r = p.ext_r + (size_t)n * p.E + off;
    const int ppr = f == 1 ? 2 : 4;  // points per row: pixel-block corners of the upsampled label
    P2i *pts, *hull;
    if (nrows <= kSmallRows) {
      pts = s_pts[wib];
      hull = s_hull[wib];
    } else {
      pts = p.hull + ((size_t)n * p.E + off) * 8;
      hull = pts + 4 * nrows;
    }
    for (int i = lane; i < nrows; i += 32) {
      const int xl = ext_l[i] * f, xr = ext_r[i] * f + f - 1, yy = (ymin + i) * f;
      if (f == 1) {
        pts[2 * i] = P2i{xl, yy};
        pts[2 * i + 1] = P2i{xr, yy};
      } else {
        pts[4 * i] = P2i{xl, yy};
        pts[4 * i + 1] = P2i{xr, yy};
        pts[4 * i + 2] = P2i{xl, yy + f - 1};
        pts[4 * i + 3] = P2i{xr, yy + f - 1};
      }
    }
    __syncwarp();
    int hn = 0;
    if (lane == 0) hn = geom::hull_sorted(pts, ppr * nrows, hull);
    hn = __shfl_sync(0xffffffffu, hn, 0);
    __syncwarp();
    geom::Rect rect;
    warp_min_area_rect(hull, hn, &rect, lane);  // cv2.minAreaRect + boxPoints (:85-86)
    if (lane == 0) ex_emit_box(p, n, ko, rect, score);
  }
}

// E12: ordered compaction of the kept boxes, one CTA per image.
__global__ void __launch_bounds__(kRunThreads) ex_compact_kernel(ExParams p) {
  const int n = blockIdx.x + p.n0;
  const int nc = p.ncand[n];
  const size_t ko = (size_t)n * p.maxc;
  const int chunk = (nc + kRunThreads - 1) / kRunThreads;
  const int lo = min(nc, (int)threadIdx.x * chunk), hi = min(nc, lo + chunk);
  int local = 0;
  for (int k = lo; k < hi; ++k) local += p.res_keep[ko + k];
  int total;
  int pos = block_exclusive_scan(local, &total);
  for (int k = lo; k < hi; ++k) {
    if (!p.res_keep[ko + k]) continue;
    for (int q = 0; q < 8; ++q) {
      p.boxes_out[(ko + pos) * 8 + q] = p.res_box[(ko + k) * 8 + q];
      if (p.boxes_f_out) p.boxes_f_out[(ko + pos) * 8 + q] = p.res_boxf[(ko + k) * 8 + q];
    }
    p.scores_out[ko + pos] = p.res_score[ko + k];
    ++pos;
  }
  if (threadIdx.x == 0) {
    const int fl = p.imgflags[n];
    p.counts_out[n] = (fl & OCRPP_IMG_RUN_OVERFLOW) ? 0 : total;
    p.status_out[n] = fl;
  }
}

// debug / parity: the label map pse()/pa() return (cv2 ids), at processing resolution
__global__ void __launch_bounds__(kRunBlk) ex_labels_kernel(ExParams p) {
  const int n = blockIdx.y + p.n0;
  const int nr = ex_nruns(p, n, 0);
  const size_t ro = (size_t)(n * 2 + 0) * p.R, so = (size_t)n * p.R;
  const uint32_t* st = p.st + (size_t)n * p.H * p.W;
  int32_t* lab = p.labels_dbg + (size_t)n * p.H * p.W;
  const int lane = threadIdx.x & 31, wpb = kRunBlk / 32;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < nr; r += gridDim.x * wpb) {
    if (!p.t_seen[so + p.par[ro + r]]) continue;
    const size_t row = (size_t)p.run_y[ro + r] * p.W;
    for (int x = p.run_xs[ro + r] + lane; x <= (int)p.run_xe[ro + r]; x += 32) {
      const uint32_t v = st[row + x];
      if (is_labelled(v)) lab[row + x] = p.s_cid[so + (v & kLabelMask) - 1];
    }
  }
}

size_t ex_carve(ExParams& p, void* ws) {
  Carver c{(char*)ws, 0};
  const size_t N = p.N, R = p.R, E = p.E, HW = (size_t)p.H * p.W;
  p.g_nwork = c.take<int32_t>(64);  // g_nwork[4] | g_next[4] | g_arena_used (2 words), cleared together
  p.g_next = p.g_nwork ? p.g_nwork + 4 : nullptr;
  p.g_arena_used = p.g_nwork ? reinterpret_cast<unsigned long long*>(p.g_nwork + 8) : nullptr;
  p.nruns = c.take<int32_t>(6 * N);  // nruns[2N] | ext_alloc | imgflags | ncand | extdefer, cleared together
  p.ext_alloc = p.nruns ? p.nruns + 2 * N : nullptr;
  p.imgflags = p.nruns ? p.nruns + 3 * N : nullptr;
  p.ncand = p.nruns ? p.nruns + 4 * N : nullptr;
  p.extdefer = p.nruns ? p.nruns + 5 * N : nullptr;
  p.kb = c.take<uint8_t>(N * HW);
  p.st = c.take<uint32_t>(N * HW);
  p.bits = c.take<uint32_t>(N * 2 * p.H * p.Wd);
  p.rowptr = c.take<int32_t>(N * 2 * (p.H + 1));
  p.rowcnt = c.take<int32_t>(N * 2 * p.H);
  p.run_xs = c.take<uint16_t>(N * 2 * R);
  p.run_xe = c.take<uint16_t>(N * 2 * R);
  p.run_y = c.take<uint16_t>(N * 2 * R);
  p.par = c.take<int32_t>(N * 2 * R);
  p.t_area = c.take<int32_t>(N * R);
  p.t_xmin = c.take<int32_t>(N * R);
  p.t_xmax = c.take<int32_t>(N * R);
  p.t_ymax = c.take<int32_t>(N * R);
  p.t_seen = c.take<int32_t>(N * R);
  p.t_amin = c.take<int32_t>(N * R);
  p.t_amax = c.take<int32_t>(N * R);
  p.t_nseed = c.take<int32_t>(N * R);
  p.t_lab = c.take<int32_t>(N * R);
  p.s_area = c.take<int32_t>(N * R);
  p.s_cc = c.take<int32_t>(N * R);
  p.s_cid = c.take<int32_t>(N * R);
  p.s_alive = c.take<int32_t>(N * R);
  p.s_flag = c.take<int32_t>(N * R);
  p.s_emb = c.take<double>(p.mode == kModePan ? N * R * 4 : 4);
  p.l_area = c.take<int32_t>(N * R);
  p.l_ymin = c.take<int32_t>(N * R);
  p.l_ymax = c.take<int32_t>(N * R);
  p.l_rowoff = c.take<int32_t>(N * R);
  p.l_rowbase = c.take<int32_t>(N * R);
  p.l_row0 = c.take<int32_t>(N * R);
  p.l_sum = c.take<long long>(N * R);
  p.ext_l = c.take<int32_t>(N * E);
  p.ext_r = c.take<int32_t>(N * E);
  p.hull = c.take<P2i>(N * E * 8);
  p.cand = c.take<int32_t>(N * p.maxc);
  p.res_keep = c.take<int32_t>(N * p.maxc);
  p.res_box = c.take<int16_t>(N * p.maxc * 8);
  p.res_boxf = c.take<float>(N * p.maxc * 8);
  p.res_score = c.take<float>(N * p.maxc);
  p.work = c.take<int2>(4 * N * R);
  p.arena = c.take<uint32_t>((size_t)p.arena_cap);
  return align_up(c.off, 256);
}

int ex_resolve_runs(int H, int W, int max_runs) {
  const long long worst = (long long)H * ((W + 1) / 2);  // foreground runs only
  if (max_runs <= 0 || max_runs > worst) return (int)worst;
  return max_runs;
}

long long ex_resolve_arena(int N, int H, int W, long long arena_elems) {
  const long long worst = 4ll * N * H * W;
  if (arena_elems <= 0 || arena_elems > worst) return worst;
  return arena_elems;
}

void ex_fill(ExParams& p, int mode, int N, int C, int K, int h, int w, int fin, int fout, int max_boxes,
             int max_runs, long long arena_elems) {
  p.mode = mode;
  p.N = N; p.C = C; p.K = K; p.h = h; p.w = w; p.fin = fin; p.fout = fout;
  p.H = h * fin; p.W = w * fin; p.Wd = (p.W + 31) / 32;
  p.R = ex_resolve_runs(p.H, p.W, max_runs);
  p.E = 2 * (2 * p.R + 4);   // first half: blocks reserved at seeding time; second half: exact per-label blocks
  p.maxc = max_boxes;
  p.seed_bit = mode == kModePse ? K - 1 : 1;
  p.arena_cap = ex_resolve_arena(N, p.H, p.W, arena_elems);
}

// auxiliary streams / events of the current device (created once, never destroyed)
// The auxiliary streams / events below are shared per device by every instantiation (float, __half) of the pipeline:
// one fork/join section at a time, guarded by ONE mutex per set.
std::mutex g_ex_enqueue_mu, g_ex_split_enqueue_mu;

struct ExAux {
  cudaStream_t st[3];
  cudaEvent_t fork, join[3];
};

ExAux* ex_aux() {
  static std::mutex mu;   // first use may come from several host threads
  std::lock_guard<std::mutex> lock(mu);
  static ExAux aux[64];
  static bool ready[64] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (!ready[dev]) {
    for (int i = 0; i < 3; ++i) {
      if (cudaStreamCreateWithFlags(&aux[dev].st[i], cudaStreamNonBlocking) != cudaSuccess) return nullptr;
      if (cudaEventCreateWithFlags(&aux[dev].join[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
    }
    if (cudaEventCreateWithFlags(&aux[dev].fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    ready[dev] = true;
  }
  return &aux[dev];
}

// the whole chain for images [p.n0, p.n0 + N) on stream s; `fork` = run the four tile-class kernels on
// auxiliary streams (single-chain mode), otherwise back to back on s (another sub-batch fills the tails)
template <typename T>
int ex_pipeline(ExParams p, int N, cudaStream_t s, bool vec, ProfileScope* prof, bool fork) {
  {
    dim3 grid((p.H + kBinWarps - 1) / kBinWarps, N);
    if (vec && p.K <= 2) ex_binarize_kernel<T, true, 4, 2><<<grid, kBinWarps * 32, 0, s>>>(p);
    else if (vec) ex_binarize_kernel<T, true, 1, kMaxK><<<grid, kBinWarps * 32, 0, s>>>(p);
    else ex_binarize_kernel<T, false, 1, kMaxK><<<grid, kBinWarps * 32, 0, s>>>(p);
    OCRPP_LAUNCHED();
    if (prof) prof->mark("ex_binarize");
  }
  const int ictas = max(kImgCtas, min(64, (kNumSMs * 4 + N - 1) / N));   // fill the GPU at small batch sizes too
  ex_runs_kernel<<<dim3(ictas, N, 2), kRunThreads, sizeof(int) * (p.H + 1), s>>>(p);
  OCRPP_LAUNCHED();
  if (prof) prof->mark("ex_runs");
  const dim3 rgrid2(ictas, N, 2), rgrid(ictas, N);
  ex_link_kernel<<<rgrid2, kRunBlk, 0, s>>>(p);
  OCRPP_LAUNCHED();
  if (prof) prof->mark("ex_link");
  ex_flatten_kernel<<<rgrid2, kRunBlk, 0, s>>>(p);
  OCRPP_LAUNCHED();
  if (prof) prof->mark("ex_flatten");
  ex_seed_kernel<<<N, kRunThreads, 0, s>>>(p);
  OCRPP_LAUNCHED();
  if (prof) prof->mark("ex_seed");
  if (p.mode == kModePan) {
    ex_pan_flag_kernel<T><<<rgrid, kRunBlk, 0, s>>>(p);
    OCRPP_LAUNCHED();
  }
  ex_paint_kernel<<<rgrid, kRunBlk, 0, s>>>(p, 0);
  OCRPP_LAUNCHED();
  ex_paint_kernel<<<rgrid, kRunBlk, 0, s>>>(p, 1);
  OCRPP_LAUNCHED();
  if (prof) prof->mark("ex_paint");
  {
    auto tiny_k = ex_expand_kernel<T, kTinyCap, kTinyList, kTinyThreads, 0>;
    auto small_k = ex_expand_kernel<T, kSmallCap, kListCap, kSmallThreads, 1>;
    auto big_k = ex_expand_kernel<T, kBigCap, kListCap, kBigThreads, 2>;
    auto huge_k = ex_expand_kernel<T, kHugeCap, kListCap, kHugeThreads, 3>;
    // opt in to > 48 KB of dynamic shared memory (a per-device function attribute: set on every call)
    OCRPP_CUDA(cudaFuncSetAttribute(huge_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ex_smem_bytes(kHugeCap, kListCap)));
    OCRPP_CUDA(cudaFuncSetAttribute(small_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ex_smem_bytes(kSmallCap, kListCap)));
    OCRPP_CUDA(cudaFuncSetAttribute(big_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ex_smem_bytes(kBigCap, kListCap)));
    // The four tile classes are independent persistent kernels: fork them onto auxiliary streams so that
    // the tail of one (a few long items) overlaps the bulk of the next, and join before the statistics.
    ExAux* aux = fork ? ex_aux() : nullptr;
    OCRPP_CHECK_ARG(!fork || aux != nullptr, "expand: cannot create auxiliary streams");
    cudaStream_t s1 = s, s2 = s, s3 = s;
    std::unique_lock<std::mutex> lock(g_ex_enqueue_mu, std::defer_lock);
    if (aux) {
      lock.lock();
      OCRPP_CUDA(cudaEventRecord(aux->fork, s));
      for (int i = 0; i < 3; ++i) OCRPP_CUDA(cudaStreamWaitEvent(aux->st[i], aux->fork, 0));
      s1 = aux->st[0];
      s2 = aux->st[1];
      s3 = aux->st[2];
    }
    int frc = OCRPP_OK;   // inside the fork/join section errors are collected: the join must always be enqueued
    auto launched = [&frc](const char* what) {
      g_launch_count.fetch_add(1, std::memory_order_relaxed);
      const cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess && frc == OCRPP_OK) frc = set_error(OCRPP_ERR_CUDA, "%s launch failed: %s", what, cudaGetErrorString(e));
    };
    huge_k<<<kNumSMs, kHugeThreads, ex_smem_bytes(kHugeCap, kListCap), s>>>(p);   // long items first
    launched("ex_expand (huge tiles)");
    big_k<<<kNumSMs * 2, kBigThreads, ex_smem_bytes(kBigCap, kListCap), s1>>>(p);
    launched("ex_expand (big tiles)");
    small_k<<<kNumSMs * 4, kSmallThreads, ex_smem_bytes(kSmallCap, kListCap), s2>>>(p);
    launched("ex_expand (small tiles)");
    tiny_k<<<kNumSMs * 9, kTinyThreads, ex_smem_bytes(kTinyCap, kTinyList), s3>>>(p);
    launched("ex_expand (tiny tiles)");
    if (aux)
      for (int i = 0; i < 3; ++i) {
        cudaEventRecord(aux->join[i], aux->st[i]);
        cudaStreamWaitEvent(s, aux->join[i], 0);
      }
    if (frc != OCRPP_OK) return frc;
    if (debug_sync()) OCRPP_CUDA(cudaDeviceSynchronize());
  }
  if (prof) prof->mark("ex_expand");
  ex_stats_kernel<T, 1><<<rgrid, kRunBlk, 0, s>>>(p);
  OCRPP_LAUNCHED();
  if (prof) prof->mark("ex_stats");
  ex_cand_kernel<<<N, kRunThreads, 0, s>>>(p);
  OCRPP_LAUNCHED();
  if (prof) prof->mark("ex_cand");
  ex_stats_kernel<T, 2><<<rgrid, kRunBlk, 0, s>>>(p);
  OCRPP_LAUNCHED();
  if (prof) prof->mark("ex_extents");
  ex_geometry_fast_kernel<<<dim3((p.maxc + kGeoThreads / kGrp - 1) / (kGeoThreads / kGrp), N), kGeoThreads, 0, s>>>(p);
  OCRPP_LAUNCHED();
  ex_geometry_kernel<<<dim3(4, N), kGeoWarps * 32, 0, s>>>(p);
  OCRPP_LAUNCHED();
  if (prof) prof->mark("ex_geometry");
  ex_compact_kernel<<<N, kRunThreads, 0, s>>>(p);
  OCRPP_LAUNCHED();
  if (prof) prof->mark("ex_compact");
  if (p.labels_dbg) {
    OCRPP_CUDA(cudaMemsetAsync(p.labels_dbg + (size_t)p.n0 * p.H * p.W, 0, sizeof(int32_t) * (size_t)N * p.H * p.W, s));
    ex_labels_kernel<<<rgrid, kRunBlk, 0, s>>>(p);
    OCRPP_LAUNCHED();
  }
  return OCRPP_OK;
}

// per sub-batch i of nsplit: own work lists (a slice of each class region), counters and arena slice
ExParams ex_sub_params(const ExParams& p, int i, int nsplit, int n0) {
  ExParams q = p;
  q.n0 = n0;
  q.g_nwork = p.g_nwork + 16 * i;
  q.g_next = q.g_nwork + 4;
  q.g_arena_used = reinterpret_cast<unsigned long long*>(q.g_nwork + 8);
  q.work = p.work + (size_t)n0 * p.R;
  q.arena_cap = p.arena_cap / nsplit;
  q.arena = p.arena + (size_t)q.arena_cap * i;
  return q;
}

struct ExSplitAux {
  cudaStream_t st[3];
  cudaEvent_t fork, join[3];
};

ExSplitAux* ex_split_aux() {
  static std::mutex mu;   // first use may come from several host threads
  std::lock_guard<std::mutex> lock(mu);
  static ExSplitAux aux[64];
  static bool ready[64] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (!ready[dev]) {
    for (int i = 0; i < 3; ++i) {
      if (cudaStreamCreateWithFlags(&aux[dev].st[i], cudaStreamNonBlocking) != cudaSuccess) return nullptr;
      if (cudaEventCreateWithFlags(&aux[dev].join[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
    }
    if (cudaEventCreateWithFlags(&aux[dev].fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    ready[dev] = true;
  }
  return &aux[dev];
}

template <typename T>
int ex_launch(ExParams& p, cudaStream_t s, bool vec) {
  OCRPP_CUDA(cudaMemsetAsync(p.g_nwork, 0, sizeof(int32_t) * 64, s));
  OCRPP_CUDA(cudaMemsetAsync(p.nruns, 0, sizeof(int32_t) * 6 * p.N, s));
  const int N = p.N;
  // Large batches run as 2 or 4 independent sub-batch pipelines on separate streams, so that the
  // bandwidth-bound binarisation of one overlaps the latency-bound expansion of another. (Not while
  // per-phase profiling is on: the event marks describe one whole-batch chain.)
  int nsplit = N >= 64 ? 4 : 1;   // measured: 2 sub-batches lose to the single chain with forked tile classes
  static const int forced = [] { const char* e = getenv("OCRPP_EX_SPLIT"); return e ? atoi(e) : 0; }();   // tuning aid
  if (forced > 0) nsplit = forced > 4 ? 4 : forced;
  ExSplitAux* aux = (nsplit > 1 && nsplit <= N && !profile_on()) ? ex_split_aux() : nullptr;
  if (!aux) {
    ProfileScope prof(s);
    return ex_pipeline<T>(ex_sub_params(p, 0, 1, 0), N, s, vec, &prof, true);
  }
  std::lock_guard<std::mutex> lock(g_ex_split_enqueue_mu);
  OCRPP_CUDA(cudaEventRecord(aux->fork, s));
  int rc = OCRPP_OK;
  for (int i = 0; i < nsplit; ++i) {
    cudaStream_t si = i == 0 ? s : aux->st[i - 1];
    if (i > 0 && cudaStreamWaitEvent(si, aux->fork, 0) != cudaSuccess) rc = set_error(OCRPP_ERR_CUDA, "expand: cudaStreamWaitEvent failed");
    const int lo = (int)((long long)N * i / nsplit), hi = (int)((long long)N * (i + 1) / nsplit);
    if (rc == OCRPP_OK) rc = ex_pipeline<T>(ex_sub_params(p, i, nsplit, lo), hi - lo, si, vec, nullptr, false);
    if (i > 0) {   // always join, also after a failed launch: the auxiliary stream may still touch the caller's buffers
      cudaEventRecord(aux->join[i - 1], si);
      cudaStreamWaitEvent(s, aux->join[i - 1], 0);
    }
  }
  if (rc != OCRPP_OK) return rc;
  return OCRPP_OK;
}

int ex_postprocess(int mode, const void* maps_dev, int dtype, int N, int C, int K, int h, int w,
                   int64_t stride_n, int64_t stride_c, int64_t stride_h, int fin, int fout,
                   const double* shape_dev, float thresh, float box_thresh, float min_area_seed,
                   float min_area_box, int max_boxes, int max_runs, int64_t arena_elems,
                   int16_t* boxes_out_dev, float* scores_out_dev, int32_t* counts_out_dev,
                   int32_t* status_out_dev, float* boxes_f_out_dev, int32_t* labels_dbg_dev,
                   void* workspace_dev, size_t workspace_bytes, void* stream) {
  const char* nm = mode == kModePse ? "pse" : "pan";
  OCRPP_CHECK_ARG(dtype == OCRPP_F32 || dtype == OCRPP_F16, "%s: dtype must be OCRPP_F32 or OCRPP_F16", nm);
  OCRPP_CHECK_ARG(N >= 0 && h > 0 && w > 0, "%s: bad shape N=%d h=%d w=%d", nm, N, h, w);
  OCRPP_CHECK_ARG(fin >= 1 && fin <= 4 && fout >= 1 && fout <= 4, "%s: upsample factors must be in [1,4]", nm);
  OCRPP_CHECK_ARG(K >= 1 && K <= kMaxK, "%s: kernel_num must be in [1,%d]", nm, kMaxK);
  OCRPP_CHECK_ARG((long long)h * fin * fout < 32768 && (long long)w * fin * fout < 32768,
                  "%s: output resolution must be below 32768 x 32768", nm);
  OCRPP_CHECK_ARG(max_boxes > 0, "%s: max_boxes must be positive", nm);
  if (N == 0) return OCRPP_OK;
  OCRPP_CHECK_ARG(maps_dev && shape_dev && boxes_out_dev && scores_out_dev && counts_out_dev && status_out_dev && workspace_dev,
                  "%s: null pointer argument", nm);
  ExParams p{};
  ex_fill(p, mode, N, C, K, h, w, fin, fout, max_boxes, max_runs, arena_elems);
  p.maps = maps_dev; p.stride_n = stride_n; p.stride_c = stride_c; p.stride_h = stride_h; p.shape = shape_dev;
  p.thresh = thresh; p.box_thresh = box_thresh; p.min_area_seed = min_area_seed; p.min_area_box = min_area_box;
  const size_t need = ex_carve(p, workspace_dev);
  if (need > workspace_bytes)
    return set_error(OCRPP_ERR_WORKSPACE_TOO_SMALL, "%s: workspace needs %zu bytes, got %zu", nm, need, workspace_bytes);
  p.boxes_out = boxes_out_dev; p.scores_out = scores_out_dev; p.counts_out = counts_out_dev;
  p.status_out = status_out_dev; p.boxes_f_out = boxes_f_out_dev; p.labels_dbg = labels_dbg_dev;
  cudaStream_t s = (cudaStream_t)stream;
  const bool vec = dtype == OCRPP_F32 && fin == 1 && ((uintptr_t)maps_dev % 16 == 0) && stride_n % 4 == 0 &&
                   stride_c % 4 == 0 && stride_h % 4 == 0 && p.W % 4 == 0;
  if (dtype == OCRPP_F32) return ex_launch<float>(p, s, vec);
  return ex_launch<__half>(p, s, false);
}

}  // namespace
}  // namespace ocrpp

extern "C" size_t ocrpp_pse_workspace_bytes(int N, int K, int h, int w, int upsample_in, int max_boxes,
                                            int max_runs, int64_t arena_elems) {
  using namespace ocrpp;
  if (N <= 0 || h <= 0 || w <= 0 || upsample_in < 1 || max_boxes <= 0) return 0;
  ExParams p{};
  ex_fill(p, kModePse, N, K, K, h, w, upsample_in, 1, max_boxes, max_runs, arena_elems);
  return ex_carve(p, nullptr);
}

extern "C" int ocrpp_pse_postprocess(const void* maps_dev, int dtype, int N, int K, int h, int w,
                                     int64_t stride_n, int64_t stride_c, int64_t stride_h,
                                     int upsample_in, int upsample_out, const double* shape_dev,
                                     float thresh, float box_thresh, float min_area_seed,
                                     float min_area_box, int max_boxes, int max_runs, int64_t arena_elems,
                                     int16_t* boxes_out_dev, float* scores_out_dev,
                                     int32_t* counts_out_dev, int32_t* status_out_dev,
                                     float* boxes_f_out_dev, int32_t* labels_dbg_dev,
                                     void* workspace_dev, size_t workspace_bytes, void* stream) {
  return ocrpp::ex_postprocess(ocrpp::kModePse, maps_dev, dtype, N, K, K, h, w, stride_n, stride_c, stride_h,
                               upsample_in, upsample_out, shape_dev, thresh, box_thresh, min_area_seed,
                               min_area_box, max_boxes, max_runs, arena_elems, boxes_out_dev, scores_out_dev,
                               counts_out_dev, status_out_dev, boxes_f_out_dev, labels_dbg_dev,
                               workspace_dev, workspace_bytes, stream);
}

extern "C" size_t ocrpp_pan_workspace_bytes(int N, int h, int w, int upsample_in, int max_boxes,
                                            int max_runs, int64_t arena_elems) {
  using namespace ocrpp;
  if (N <= 0 || h <= 0 || w <= 0 || upsample_in < 1 || max_boxes <= 0) return 0;
  ExParams p{};
  ex_fill(p, kModePan, N, 6, 2, h, w, upsample_in, 1, max_boxes, max_runs, arena_elems);
  return ex_carve(p, nullptr);
}

extern "C" int ocrpp_pan_postprocess(const void* maps_dev, int dtype, int N, int h, int w,
                                     int64_t stride_n, int64_t stride_c, int64_t stride_h,
                                     int upsample_in, int upsample_out, const double* shape_dev,
                                     float thresh, float box_thresh, float min_kernel_area,
                                     float min_area_box, int max_boxes, int max_runs, int64_t arena_elems,
                                     int16_t* boxes_out_dev, float* scores_out_dev,
                                     int32_t* counts_out_dev, int32_t* status_out_dev,
                                     float* boxes_f_out_dev, int32_t* labels_dbg_dev,
                                     void* workspace_dev, size_t workspace_bytes, void* stream) {
  return ocrpp::ex_postprocess(ocrpp::kModePan, maps_dev, dtype, N, 6, 2, h, w, stride_n, stride_c, stride_h,
                               upsample_in, upsample_out, shape_dev, thresh, box_thresh, min_kernel_area,
                               min_area_box, max_boxes, max_runs, arena_elems, boxes_out_dev, scores_out_dev,
                               counts_out_dev, status_out_dev, boxes_f_out_dev, labels_dbg_dev,
                               workspace_dev, workspace_bytes, stream);
}
