// DB / DB++ box extraction on sm_100a.
//
// Replaces, for a whole batch and without leaving the device:
//   R/pytocr/postprocess/db_postprocess.py:43-46     D2H + `pred > thresh`
//   R/pytocr/postprocess/db_postprocess_fast/src/db_postprocess.cpp:231-317  BoxesFromBitmap
//     (cv::findContours RETR_LIST / CHAIN_APPROX_SIMPLE, minAreaRect, BoxScore :194-229,
//      UnClip :34-64, GetMiniBoxes :159-192, rescale/roundf/clamp :303-311)
//
// cv::findContours is restated in connected-component terms (oracle/db_ccl_oracle.py):
//   candidates = outer(C) for every 8-connected foreground component C
//              + hole(C,h) for every 4-connected background region h that does not reach the frame
//   points(outer C) = C,  points(hole C,h) = ring = pixels of C 4-adjacent to h
//   fill(outer C)   = C + everything C encloses + "stair" pixels of the 4-connected fillPoly boundary
//   fill(hole C,h)  = h + islands in h + ring + stair pixels
//
// Data layout: the map is read ONCE at full HBM rate (db_scan_kernel) into a 1-bit/pixel mask, the
// table of horizontal RUNS of equal bits and the pixel sum of every run; everything after works on
// those runs (a few thousand per image, L2 resident):
// run union-find for both polarities, per-component reductions keyed by the root run, a parent tree
// between foreground components and holes, per-component row extents -> convex hull -> min-area
// rectangle -> unclip -> rectangle -> rescale. Scores are accumulated in 32.32 fixed point (exact for
// probabilities >= 2^-9, order independent => deterministic). Algorithmic bytes per image:
// H*W*sizeof(elem), the probability map, read once by db_scan_kernel.
#include "common.cuh"

#include <cstdlib>
#include <mutex>
#include "dev_common.cuh"
#include "dev_geom.cuh"
#include "geometry.cuh"
#include "db_scan4.cuh"

namespace ocrpp {
namespace {

using geom::P2i;

constexpr int kOutFlag = 1;
constexpr double kFixScale = 4294967296.0;  // 2^32

struct DbParams {
  // input
  const void* maps;
  long long stride_n, stride_h;
  const int32_t* src_wh;
  int N, H, W, Wd, R, maxc, cap, epl, pad_resize, n0;   // n0: first image of the sub-batch this launch covers
  float thresh, box_thresh, unclip_ratio;
  // per-image workspace (index with n * count)
  uint32_t* bits;        // [H*Wd] lane-major bit mask (see db_scan_kernel)
  uint32_t* rawbits;     // [H*Wd] raw threshold bits (use_dilation only)
  int dilate;
  unsigned long long* scum;  // [H*(cap+1)] per-row run starts: x << 48 | cumulative row sum before x (2^-23 units)
  int32_t* srow_cnt;     // [H] runs in the row | first pixel bit << 31
  long long* run_sum;    // [R] pixel sum of every run (32.32 fixed point)
  int32_t* rowptr;       // [H+1]
  uint16_t *run_xs, *run_xe, *run_yf;  // [R]
  int32_t* par;          // [R]
  int32_t *area, *xmin, *xmax, *ymax, *dmin, *dmax, *smin, *smax;  // [R]
  long long* sum;        // [R]
  int32_t* fcnt;         // [R]
  long long* fsum;       // [R]
  int32_t* xcnt;         // [R]
  long long* xsum;       // [R]
  int32_t *cpar, *cflag, *rowoff;  // [R]
  int32_t *ext_l, *ext_r;          // [E]
  P2i* hull;                       // [4*E]
  int E;
  int32_t *nruns, *ext_alloc, *imgflags, *ncand, *nbig;  // [1] per image
  int32_t* big;          // [maxc] candidates deferred to the generic geometry kernel
  int32_t* cand;         // [maxc]
  int32_t* res_keep;     // [maxc]
  int32_t* hull_n;       // [maxc] vertices of the precomputed hull (db_image_kernel / db_hull_kernel)
  int32_t *cand_off, *cand_y0, *cand_nrows;   // [maxc] candidate -> slice of the extent/hull scratch, first row, rows
  int32_t* hcnt;         // [2*maxc] hull half lengths (db_image_kernel, global-table variant)
  int32_t* cand_root;    // [R] root run of every dense component id (db_image_kernel, global-table variant)
  uint32_t* flagw;       // [R/32+1] run flag words (db_image_kernel, global-table variant)
  int dtype;             // OCRPP_F32 | OCRPP_F16 (element type of maps)
  int sem;               // 0: the C++ branch's semantics (cpp_speedup: True); 1: the pure-Python branch's
  int score_box;         // Python branch, score_mode "box": BoxScore over the filled mini box instead of the contour
  double box_thresh_d, unclip_ratio_d;   // the Python branch compares / multiplies in float64
  int stairs, skip2;     // reference-branch switches: 4-connected fillPoly boundary (stair pixels), <= 2-point skip
  int16_t* res_box;      // [maxc*8]
  float* res_boxf;       // [maxc*8]
  float* res_score;      // [maxc]
  // outputs
  int16_t* boxes_out;
  float* scores_out;
  int32_t* counts_out;
  int32_t* status_out;
  float* boxes_f_out;
  int32_t* labels_dbg;
};

__device__ __forceinline__ long long to_fixed(float v) { return __float2ll_rn(v * 4294967296.0f); }

template <typename T>
__device__ __forceinline__ float load_px(const void* maps, long long off) {
  return load_scalar<T>(reinterpret_cast<const T*>(maps) + off);
}

// ------------------------------------------------------------------------------------------------
// K1: the ONE pass over the probability map. One warp per image row, kEpl pixels per lane per group
// (128-bit loads). Per group of 32*kEpl pixels it produces, from kEpl ballots:
//   * the 1-bit mask, stored "lane-major": word g*kEpl + k, bit l  <->  pixel x = g*32*kEpl + l*kEpl + k
//   * the run boundaries of the row (both polarities): transitions are XORs of the ballot words
//   * the running sum of the row in 2^-23 fixed point (probabilities are in [0,1], checked): one
//     32-bit warp prefix scan per group; every run start stores the cumulative sum BEFORE it, so the
//     sum over any run is the difference of two adjacent entries and no later kernel re-reads pixels
// into a per-row segment (one 64-bit entry per run: x << 48 | cumulative sum) of fixed capacity (row_cap runs; overflow is flagged and the caller retries).
// ------------------------------------------------------------------------------------------------
constexpr int kBinWarps = 8;
#ifndef OCRPP_SCAN_MIN_CTAS
#define OCRPP_SCAN_MIN_CTAS 4
#endif
constexpr int kScanMinCtas = OCRPP_SCAN_MIN_CTAS;   // resident CTAs per SM the scan kernel is compiled for (register cap)

template <typename T, int kEpl>
struct RowLoader;
template <>
struct RowLoader<float, 4> {
  __device__ static __forceinline__ void load(const float* row, int x, int W, float* v) {
    const uint4 u = x < W ? ldg_stream_u4(row + x) : make_uint4(0, 0, 0, 0);
    v[0] = __uint_as_float(u.x); v[1] = __uint_as_float(u.y); v[2] = __uint_as_float(u.z); v[3] = __uint_as_float(u.w);
  }
};
template <>
struct RowLoader<__half, 8> {
  __device__ static __forceinline__ void load(const __half* row, int x, int W, float* v) {
    const uint4 u = x < W ? ldg_stream_u4(row + x) : make_uint4(0, 0, 0, 0);
    v[0] = h2f_lo(u.x); v[1] = h2f_hi(u.x); v[2] = h2f_lo(u.y); v[3] = h2f_hi(u.y);
    v[4] = h2f_lo(u.z); v[5] = h2f_hi(u.z); v[6] = h2f_lo(u.w); v[7] = h2f_hi(u.w);
  }
};
template <typename T>
struct RowLoader<T, 1> {  // unaligned / odd-width fallback: one pixel per lane
  __device__ static __forceinline__ void load(const T* row, int x, int W, float* v) {
    v[0] = x < W ? load_scalar<T>(row + x) : 0.f;
  }
};

// one group of 32*kEpl pixels; kFull = every pixel of the group is inside the row.
// kDilate: the mask is not `f > thresh` but its 2x2 dilation (use_dilation, db_postprocess.py:52-55:
// cv2.dilate with a [[1,1],[1,1]] kernel = OR over (x-1..x, y-1..y)), assembled from the raw threshold
// bits of rows y and y-1 that db_rawbits_kernel wrote (same lane-major layout); dil[0]/dil[1] carry the
// last raw bit of the previous group of row y / y-1.
template <int kEpl, bool kFull, bool kDilate>
__device__ __forceinline__ void db_scan_group(const DbParams& p, const float* v, int xg, int lane,
                                              unsigned long long* sc, uint32_t* out,
                                              unsigned long long& carry, unsigned& carry_bit, int& cnt,
                                              unsigned& first_bit, unsigned& worst,
                                              const uint32_t* raw_y, const uint32_t* raw_u, unsigned* dil) {
  constexpr int PPG = 32 * kEpl;
  const int x = xg + lane * kEpl;
  unsigned m[kEpl];    // ballot of pixel k of every lane
  unsigned pre[kEpl];  // lane-local inclusive prefix of the fixed-point values
  unsigned run = 0;
#pragma unroll
  for (int k = 0; k < kEpl; ++k) {
    const float f = v[k];
    const bool valid = kFull || x + k < p.W;
    if (!kDilate) m[k] = __ballot_sync(0xffffffffu, valid && f > p.thresh);
    // f in [0,1]: the mantissa of f + 1.0f is round(f * 2^23); anything else (negative, > 1, NaN, Inf)
    // gives u > 2^23 and is reported through `worst`
    unsigned u = __float_as_uint(f + 1.0f) - 0x3f800000u;
    if (!kFull) u = valid ? u : 0u;
    worst = max(worst, u);
    run += u;
    pre[k] = run;
  }
  if (kDilate) {
    const int w0 = (xg / PPG) * kEpl;
    unsigned ry[kEpl], ru[kEpl];
#pragma unroll
    for (int k = 0; k < kEpl; ++k) {
      ry[k] = raw_y[w0 + k];
      ru[k] = raw_u ? raw_u[w0 + k] : 0u;
    }
#pragma unroll
    for (int k = 0; k < kEpl; ++k) {   // pixel x-1: word k-1 at the same lane, or word kEpl-1 one lane down
      const unsigned py = k == 0 ? ((ry[kEpl - 1] << 1) | dil[0]) : ry[k - 1];
      const unsigned pu = k == 0 ? ((ru[kEpl - 1] << 1) | dil[1]) : ru[k - 1];
      m[k] = ry[k] | py | ru[k] | pu;
      if (!kFull) {
        const int nvalid = (p.W - xg - k + kEpl - 1) / kEpl;
        m[k] &= nvalid >= 32 ? 0xffffffffu : (nvalid <= 0 ? 0u : ((1u << nvalid) - 1u));
      }
    }
    dil[0] = ry[kEpl - 1] >> 31;
    dil[1] = ru[kEpl - 1] >> 31;
  }
  const unsigned total = __reduce_add_sync(0xffffffffu, run);  // <= 32 * kEpl * 2^23 <= 2^31
  if (xg == 0) first_bit = m[0] & 1u;
  // transitions: pixel k differs from pixel k-1 (the previous lane's last pixel for k == 0)
  unsigned tm[kEpl];
  unsigned lanes = 0;
#pragma unroll
  for (int k = 0; k < kEpl; ++k) {
    const unsigned prev = k == 0 ? ((m[kEpl - 1] << 1) | carry_bit) : m[k - 1];
    tm[k] = m[k] ^ prev;
    if (!kFull) {
      const int nvalid = (p.W - xg - k + kEpl - 1) / kEpl;  // lanes whose pixel k is inside the row
      tm[k] &= nvalid >= 32 ? 0xffffffffu : (nvalid <= 0 ? 0u : ((1u << nvalid) - 1u));
    }
    if (k == 0 && xg == 0) tm[k] &= ~1u;  // pixel 0 starts run 0, it is not a transition
    lanes |= tm[k];
  }
  // emission: every lane that holds a transition stores (x << 48 | cumulative sum before x) at its rank.
  // Only those (few) lanes need the sum of the lanes before them: one warp reduction per such lane
  // instead of a full prefix scan per group.
  if (lanes) {
    unsigned multi = 0;  // lanes holding more than one transition (rare: runs shorter than kEpl pixels)
#pragma unroll
    for (int k = 1; k < kEpl; ++k) {
      unsigned lower = tm[0];
#pragma unroll
      for (int q = 1; q < k; ++q) lower |= tm[q];
      multi |= tm[k] & lower;
    }
    unsigned excl = 0;
    for (unsigned rest = lanes; rest; rest &= rest - 1) {
      const int L = __ffs(rest) - 1;
      const unsigned e = __reduce_add_sync(0xffffffffu, lane < L ? run : 0u);
      if (lane == L) excl = e;
    }
    const unsigned long long base = carry + excl;
    if (multi == 0) {
      if ((lanes >> lane) & 1u) {
        const int j = cnt + __popc(lanes & ((1u << lane) - 1u));
        int k = 0;
        unsigned before = 0;
#pragma unroll
        for (int q = 1; q < kEpl; ++q)
          if ((tm[q] >> lane) & 1u) {
            k = q;
            before = pre[q - 1];
          }
        if (j < p.cap) sc[j] = ((unsigned long long)(x + k) << 48) | (base + before);
      }
      cnt += __popc(lanes);
    } else {
      const unsigned lt = (1u << lane) - 1u;
      int j = cnt, tot = 0;
#pragma unroll
      for (int k = 0; k < kEpl; ++k) {
        j += __popc(tm[k] & lt);
        tot += __popc(tm[k]);
      }
#pragma unroll
      for (int k = 0; k < kEpl; ++k) {
        if ((tm[k] >> lane) & 1u) {
          if (j < p.cap) sc[j] = ((unsigned long long)(x + k) << 48) | (base + (k > 0 ? pre[k - 1] : 0u));
          ++j;
        }
      }
      cnt += tot;
    }
  }
  carry += total;
  carry_bit = m[kEpl - 1] >> 31;
  if (lane == 0) {   // rows of the mask are 32-byte aligned: one vector store per group
    uint32_t* o = out + (xg / PPG) * kEpl;
    if (kEpl == 4) {
      *reinterpret_cast<uint4*>(o) = make_uint4(m[0], m[1 % kEpl], m[2 % kEpl], m[3 % kEpl]);
    } else if (kEpl == 8) {
      *reinterpret_cast<uint4*>(o) = make_uint4(m[0], m[1 % kEpl], m[2 % kEpl], m[3 % kEpl]);
      *reinterpret_cast<uint4*>(o + 4) = make_uint4(m[4 % kEpl], m[5 % kEpl], m[6 % kEpl], m[7 % kEpl]);
    } else {
      o[0] = m[0];
    }
  }
}

template <typename T, int kEpl, int kU = 4, bool kDilate = false>
__global__ void __launch_bounds__(kBinWarps * 32, kScanMinCtas) db_scan_kernel(DbParams p) {
  const int n = blockIdx.y + p.n0;
  const int y = blockIdx.x * kBinWarps + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (y >= p.H) return;
  const T* row = reinterpret_cast<const T*>(p.maps) + n * p.stride_n + y * p.stride_h;
  uint32_t* out = p.bits + ((size_t)n * p.H + y) * p.Wd;
  const size_t rowid = (size_t)n * p.H + y;
  unsigned long long* sc = p.scum + rowid * (p.cap + 1);
  constexpr int PPG = 32 * kEpl;  // pixels per group
  constexpr int U = kEpl == 1 ? 1 : kU;   // groups per main-loop iteration (all loads issued up front)
  unsigned worst = 0;
  unsigned long long carry = 0;   // sum of all pixels of the row before this group (warp-uniform)
  unsigned carry_bit = 0;         // bit of the pixel just before this group
  int cnt = 1;                    // run 0 starts at x = 0
  unsigned first_bit = 0;
  if (lane == 0) sc[0] = 0ull;   // run 0: x = 0, nothing before it
  const uint32_t* raw_y = kDilate ? p.rawbits + ((size_t)n * p.H + y) * p.Wd : nullptr;
  const uint32_t* raw_u = (kDilate && y > 0) ? raw_y - p.Wd : nullptr;
  unsigned dil[2] = {0u, 0u};
  // main loop: U full groups per iteration, all loads issued first, no per-group bounds checks
  int x0 = 0;
  for (; x0 + PPG * U <= p.W; x0 += PPG * U) {
    float v[U][kEpl];
#pragma unroll
    for (int u = 0; u < U; ++u) RowLoader<T, kEpl>::load(row, x0 + u * PPG + lane * kEpl, p.W, v[u]);
#pragma unroll
    for (int u = 0; u < U; ++u)
      db_scan_group<kEpl, true, kDilate>(p, v[u], x0 + u * PPG, lane, sc, out, carry, carry_bit, cnt, first_bit, worst,
                                         raw_y, raw_u, dil);
  }
  // tail: the remaining (fewer than U) groups, the last of them possibly partial
  for (; x0 < p.W; x0 += PPG) {
    float v[kEpl];
    RowLoader<T, kEpl>::load(row, x0 + lane * kEpl, p.W, v);
    if (x0 + PPG <= p.W)
      db_scan_group<kEpl, true, kDilate>(p, v, x0, lane, sc, out, carry, carry_bit, cnt, first_bit, worst, raw_y, raw_u, dil);
    else
      db_scan_group<kEpl, false, kDilate>(p, v, x0, lane, sc, out, carry, carry_bit, cnt, first_bit, worst, raw_y, raw_u, dil);
  }
  if (lane == 0) {
    p.srow_cnt[rowid] = cnt | (first_bit << 31);
    if (cnt <= p.cap) sc[cnt] = carry;
  }
  if (__any_sync(0xffffffffu, worst > 0x800000u) && lane == 0) atomicOr(&p.imgflags[n], OCRPP_IMG_VALUE_OUT_OF_RANGE);
}

// use_dilation, pass 1: raw threshold bits of every row in the lane-major layout (the dilated mask of row y
// needs rows y-1 and y, so the scan kernel cannot threshold on the fly)
template <typename T, int kEpl>
__global__ void __launch_bounds__(kBinWarps * 32) db_rawbits_kernel(DbParams p) {
  const int n = blockIdx.y + p.n0;
  const int y = blockIdx.x * kBinWarps + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (y >= p.H) return;
  const T* row = reinterpret_cast<const T*>(p.maps) + n * p.stride_n + y * p.stride_h;
  uint32_t* out = p.rawbits + ((size_t)n * p.H + y) * p.Wd;
  constexpr int PPG = 32 * kEpl;
  for (int x0 = 0; x0 < p.W; x0 += PPG) {
    float v[kEpl];
    RowLoader<T, kEpl>::load(row, x0 + lane * kEpl, p.W, v);
#pragma unroll
    for (int k = 0; k < kEpl; ++k) {
      const unsigned w = __ballot_sync(0xffffffffu, x0 + lane * kEpl + k < p.W && v[k] > p.thresh);
      if (lane == k) out[(x0 / PPG) * kEpl + k] = w;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K1': the same pass as db_scan_kernel with the per-run bookkeeping moved off the streaming loop (what the
// one-kernel-per-image stage takes: it does not need the bit mask in global memory).
//   phase 1 (streaming, per group of 32*kEpl pixels): load, fixed-point conversion, range check, kEpl ballots;
//     the pixel values, the sum of every lane's cell of kEpl pixels and the ballot words go to the warp's slice of
//     shared memory - no run logic, no warp reductions: ~26 instructions per group;
//   phase 2 (once per row): every lane takes C CONSECUTIVE cells (a contiguous chunk of the row). The bits of its
//     chunk are C-bit fields of the ballot words, so ALL transitions inside the chunk come from kEpl XORs; two warp
//     scans (transition counts, chunk sums) give every lane its first output slot and the row's cumulative sum
//     at its first pixel; only lanes that own a transition go to shared memory again for the values inside the cell.
// Same output as db_scan_kernel: per row the run starts `x << 48 | cumulative sum before x`, the run count and the
// polarity of the first run. About 450 instead of 915 warp instructions per 1280-pixel row.
// ------------------------------------------------------------------------------------------------
constexpr int kScan2Warps = 8;

__host__ __device__ inline int scan2_row_words(int C, int epl) { return (32 * C * epl + 32 * C + (C + 1) * epl + 3) & ~3; }

template <typename T, int kEpl>
__global__ void __launch_bounds__(kScan2Warps * 32) db_scan2_kernel(DbParams p, int C) {
  extern __shared__ __align__(16) uint32_t s_scan[];
  const int n = blockIdx.y + p.n0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int y = blockIdx.x * kScan2Warps + warp;
  if (y >= p.H) return;
  const int W = p.W;
  const int ncells = W / kEpl;               // the vector path requires W % kEpl == 0
  uint32_t* uval = s_scan + (size_t)warp * scan2_row_words(C, kEpl);   // [32*C*kEpl] pixel values, 2^-23 units
  uint32_t* csum = uval + 32 * C * kEpl;                               // [32*C] cell sums, later prefixes
  uint32_t* bits = csum + 32 * C;                                      // [(C+1)*kEpl] ballot words, g-major
  const T* row = reinterpret_cast<const T*>(p.maps) + n * p.stride_n + y * p.stride_h;
  const size_t rowid = (size_t)n * p.H + y;
  unsigned long long* sc = p.scum + rowid * (p.cap + 1);
  unsigned worst = 0;

  // ---- phase 1 ----
  constexpr int U = 5;   // groups per step, all loads issued up front
  if (lane < kEpl) bits[C * kEpl + lane] = 0u;
  for (int g0 = 0; g0 < C; g0 += U) {
    float v[U][kEpl];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int x = ((g0 + u) * 32 + lane) * kEpl;
      RowLoader<T, kEpl>::load(row, x, (g0 + u) < C ? W : 0, v[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int g = g0 + u;
      if (g < C) {
        const int c = g * 32 + lane;
        const bool valid = c < ncells;
        unsigned uu[kEpl], m[kEpl], run = 0;
#pragma unroll
        for (int k = 0; k < kEpl; ++k) {
          const float f = v[u][k];
          m[k] = __ballot_sync(0xffffffffu, valid && f > p.thresh);
          // f in [0,1]: the mantissa of f + 1.0f is round(f * 2^23); anything else (negative, > 1, NaN, Inf)
          // gives a value > 2^23 and is reported through `worst`
          unsigned q = __float_as_uint(f + 1.0f) - 0x3f800000u;
          q = valid ? q : 0u;
          worst = max(worst, q);
          run += q;
          uu[k] = q;
        }
        if (kEpl == 4) {
          *reinterpret_cast<uint4*>(uval + c * 4) = make_uint4(uu[0], uu[1 % kEpl], uu[2 % kEpl], uu[3 % kEpl]);
          if (lane == 0) *reinterpret_cast<uint4*>(bits + g * 4) = make_uint4(m[0], m[1 % kEpl], m[2 % kEpl], m[3 % kEpl]);
        } else {
          *reinterpret_cast<uint4*>(uval + c * 8) = make_uint4(uu[0], uu[1 % kEpl], uu[2 % kEpl], uu[3 % kEpl]);
          *reinterpret_cast<uint4*>(uval + c * 8 + 4) = make_uint4(uu[4 % kEpl], uu[5 % kEpl], uu[6 % kEpl], uu[7 % kEpl]);
          if (lane == 0) {
            *reinterpret_cast<uint4*>(bits + g * 8) = make_uint4(m[0], m[1 % kEpl], m[2 % kEpl], m[3 % kEpl]);
            *reinterpret_cast<uint4*>(bits + g * 8 + 4) = make_uint4(m[4 % kEpl], m[5 % kEpl], m[6 % kEpl], m[7 % kEpl]);
          }
        }
        csum[c] = run;
      }
    }
  }
  __syncwarp();

  // ---- phase 2: lane owns cells [c0, c0 + C), pixels [c0 * kEpl, ...) ----
  const int c0 = C * lane;
  const int nv = max(0, min(C, ncells - c0));                 // valid cells of the chunk
  const unsigned vmask = nv >= 32 ? 0xffffffffu : ((1u << nv) - 1u);
  unsigned F[kEpl], Tm[kEpl];
  {
    const int w0 = c0 >> 5, sh = c0 & 31;
#pragma unroll
    for (int k = 0; k < kEpl; ++k)
      F[k] = __funnelshift_r(bits[w0 * kEpl + k], bits[(w0 + 1) * kEpl + k], sh) & vmask;
  }
  // bit of the pixel just before the chunk (the previous lane's last valid pixel; pixel 0 has none)
  const unsigned mylast = nv > 0 ? (F[kEpl - 1] >> (nv - 1)) & 1u : 0u;
  unsigned prevbit = __shfl_up_sync(0xffffffffu, mylast, 1);
  if (lane == 0) prevbit = F[0] & 1u;
  int cnt_l = 0;
  unsigned cellmask = 0;
#pragma unroll
  for (int k = 0; k < kEpl; ++k) {
    const unsigned prev = k == 0 ? ((F[kEpl - 1] << 1) | prevbit) : F[k - 1];
    Tm[k] = (F[k] ^ prev) & vmask;
    cnt_l += __popc(Tm[k]);
    cellmask |= Tm[k];
  }
  // chunk-local exclusive prefix of the cell sums, in place
  unsigned acc = 0;
  for (int i = 0; i < C; ++i) {
    const unsigned t = csum[c0 + i];
    csum[c0 + i] = acc;
    acc += t;
  }
  // warp scans: transitions before this lane, sum of the row before this lane's first pixel
  int pos = cnt_l;
  unsigned long long base = acc;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int tp = __shfl_up_sync(0xffffffffu, pos, o);
    const unsigned long long tb = __shfl_up_sync(0xffffffffu, base, o);
    if (lane >= o) {
      pos += tp;
      base += tb;
    }
  }
  const int total_cnt = __shfl_sync(0xffffffffu, pos, 31) + 1;          // run 0 starts at x = 0
  const unsigned long long total_sum = __shfl_sync(0xffffffffu, base, 31);
  pos -= cnt_l;     // exclusive
  base -= acc;
  // emission: cells of this chunk that hold a transition
  for (unsigned rest = cellmask; rest; rest &= rest - 1) {
    const int i = __ffs(rest) - 1;
    const unsigned below = (1u << i) - 1u;
    int j = 1 + pos;
#pragma unroll
    for (int k = 0; k < kEpl; ++k) j += __popc(Tm[k] & below);
    const unsigned long long cb = base + csum[c0 + i];
    const uint32_t* uv = uval + (size_t)(c0 + i) * kEpl;
    unsigned within = 0;
#pragma unroll
    for (int k = 0; k < kEpl; ++k) {
      if ((Tm[k] >> i) & 1u) {
        if (j < p.cap) sc[j] = ((unsigned long long)((c0 + i) * kEpl + k) << 48) | (cb + within);
        ++j;
      }
      within += uv[k];
    }
  }
  if (lane == 0) {
    sc[0] = 0ull;   // run 0: x = 0, nothing before it
    p.srow_cnt[rowid] = total_cnt | ((F[0] & 1u) << 31);
    if (total_cnt <= p.cap) sc[total_cnt] = total_sum;
  }
  if (__any_sync(0xffffffffu, worst > 0x800000u) && lane == 0) atomicOr(&p.imgflags[n], OCRPP_IMG_VALUE_OUT_OF_RANGE);
}

// K1'': db_scan2_kernel specialised at compile time for rows of exactly kC full groups (W == 32 * kEpl * kC, e.g. 1280
// float32 pixels: kC = 10): no bounds predicates, constant shared-memory offsets, unrolled chunk sums.
template <typename T, int kEpl, int kC>
__global__ void __launch_bounds__(kScan2Warps * 32) db_scan3_kernel(DbParams p) {
  extern __shared__ __align__(16) uint32_t s_scan[];
  constexpr int kRowWords = (32 * kC * kEpl + 32 * kC + (kC + 1) * kEpl + 3) & ~3;
  const int n = blockIdx.y + p.n0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int y = blockIdx.x * kScan2Warps + warp;
  if (y >= p.H) return;
  uint32_t* uval = s_scan + (size_t)warp * kRowWords;   // [32*kC*kEpl] pixel values, 2^-23 units
  uint32_t* csum = uval + 32 * kC * kEpl;                // [32*kC] cell sums
  uint32_t* bits = csum + 32 * kC;                       // [(kC+1)*kEpl] ballot words, g-major
  const T* row = reinterpret_cast<const T*>(p.maps) + n * p.stride_n + y * p.stride_h;
  const size_t rowid = (size_t)n * p.H + y;
  unsigned long long* sc = p.scum + rowid * (p.cap + 1);
  const uint4* vp = reinterpret_cast<const uint4*>(row) + lane;
  const float thresh = p.thresh;
  unsigned worst = 0;

  // ---- phase 1: stream the row ----
  if (lane < kEpl) bits[kC * kEpl + lane] = 0u;
  constexpr int U = kC <= 10 ? kC : (kC % 5 == 0 ? 5 : (kC % 4 == 0 ? 4 : 1));   // loads in flight per lane: the whole row when it fits
#pragma unroll
  for (int g0 = 0; g0 < kC; g0 += U) {
    uint4 raw[U];
#pragma unroll
    for (int u = 0; u < U; ++u) raw[u] = ldg_stream_u4_ordered(vp + (g0 + u) * 32);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int g = g0 + u;
      float f[kEpl];
      if (kEpl == 4) {
        f[0] = __uint_as_float(raw[u].x); f[1 % kEpl] = __uint_as_float(raw[u].y);
        f[2 % kEpl] = __uint_as_float(raw[u].z); f[3 % kEpl] = __uint_as_float(raw[u].w);
      } else {
        f[0] = h2f_lo(raw[u].x); f[1 % kEpl] = h2f_hi(raw[u].x); f[2 % kEpl] = h2f_lo(raw[u].y); f[3 % kEpl] = h2f_hi(raw[u].y);
        f[4 % kEpl] = h2f_lo(raw[u].z); f[5 % kEpl] = h2f_hi(raw[u].z); f[6 % kEpl] = h2f_lo(raw[u].w); f[7 % kEpl] = h2f_hi(raw[u].w);
      }
      unsigned q[kEpl], m[kEpl], run = 0;
#pragma unroll
      for (int k = 0; k < kEpl; ++k) {
        m[k] = __ballot_sync(0xffffffffu, f[k] > thresh);
        q[k] = __float_as_uint(f[k] + 1.0f) - 0x3f800000u;   // round(f * 2^23) for f in [0,1], > 2^23 otherwise
        worst = max(worst, q[k]);
        run += q[k];
      }
      uint4* uv = reinterpret_cast<uint4*>(uval + (g * 32 + lane) * kEpl);
      uv[0] = make_uint4(q[0], q[1 % kEpl], q[2 % kEpl], q[3 % kEpl]);
      if (kEpl == 8) uv[1] = make_uint4(q[4 % kEpl], q[5 % kEpl], q[6 % kEpl], q[7 % kEpl]);
      csum[g * 32 + lane] = run;
      if (lane == 0) {
        uint4* bw = reinterpret_cast<uint4*>(bits + g * kEpl);
        bw[0] = make_uint4(m[0], m[1 % kEpl], m[2 % kEpl], m[3 % kEpl]);
        if (kEpl == 8) bw[1] = make_uint4(m[4 % kEpl], m[5 % kEpl], m[6 % kEpl], m[7 % kEpl]);
      }
    }
  }
  __syncwarp();

  // ---- phase 2: lane owns cells [c0, c0 + kC), pixels [c0 * kEpl, ...) ----
  const int c0 = kC * lane;
  constexpr unsigned vmask = kC >= 32 ? 0xffffffffu : ((1u << kC) - 1u);
  unsigned F[kEpl], Tm[kEpl];
  {
    const int w0 = c0 >> 5, sh = c0 & 31;
#pragma unroll
    for (int k = 0; k < kEpl; ++k)
      F[k] = __funnelshift_r(bits[w0 * kEpl + k], bits[(w0 + 1) * kEpl + k], sh) & vmask;
  }
  unsigned prevbit = __shfl_up_sync(0xffffffffu, (F[kEpl - 1] >> (kC - 1)) & 1u, 1);
  if (lane == 0) prevbit = F[0] & 1u;
  int cnt_l = 0;
  unsigned cellmask = 0;
#pragma unroll
  for (int k = 0; k < kEpl; ++k) {
    const unsigned prev = k == 0 ? ((F[kEpl - 1] << 1) | prevbit) : F[k - 1];
    Tm[k] = (F[k] ^ prev) & vmask;
    cnt_l += __popc(Tm[k]);
    cellmask |= Tm[k];
  }
  unsigned acc = 0;
#pragma unroll
  for (int i = 0; i < kC; ++i) acc += csum[c0 + i];
  int pos = cnt_l;
  unsigned long long base = acc;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int tp = __shfl_up_sync(0xffffffffu, pos, o);
    const unsigned long long tb = __shfl_up_sync(0xffffffffu, base, o);
    if (lane >= o) {
      pos += tp;
      base += tb;
    }
  }
  const int total_cnt = __shfl_sync(0xffffffffu, pos, 31) + 1;          // run 0 starts at x = 0
  const unsigned long long total_sum = __shfl_sync(0xffffffffu, base, 31);
  pos -= cnt_l;     // exclusive
  base -= acc;
  for (unsigned rest = cellmask; rest; rest &= rest - 1) {
    const int i = __ffs(rest) - 1;
    const unsigned below = (1u << i) - 1u;
    int j = 1 + pos;
#pragma unroll
    for (int k = 0; k < kEpl; ++k) j += __popc(Tm[k] & below);
    unsigned long long cb = base;
    for (int q = 0; q < i; ++q) cb += csum[c0 + q];
    const uint32_t* uv = uval + (size_t)(c0 + i) * kEpl;
    unsigned within = 0;
#pragma unroll
    for (int k = 0; k < kEpl; ++k) {
      if ((Tm[k] >> i) & 1u) {
        if (j < p.cap) sc[j] = ((unsigned long long)((c0 + i) * kEpl + k) << 48) | (cb + within);
        ++j;
      }
      within += uv[k];
    }
  }
  if (lane == 0) {
    sc[0] = 0ull;
    p.srow_cnt[rowid] = total_cnt | ((F[0] & 1u) << 31);
    if (total_cnt <= p.cap) sc[total_cnt] = total_sum;
  }
  if (__any_sync(0xffffffffu, worst > 0x800000u) && lane == 0) atomicOr(&p.imgflags[n], OCRPP_IMG_VALUE_OUT_OF_RANGE);
}

template <typename T, int kEpl, int kC>
int launch_scan3(const DbParams& p, dim3 grid, cudaStream_t s) {
  constexpr int kRowWords = (32 * kC * kEpl + 32 * kC + (kC + 1) * kEpl + 3) & ~3;
  constexpr int smem = kScan2Warps * kRowWords * (int)sizeof(uint32_t);
  OCRPP_CUDA(cudaFuncSetAttribute(db_scan3_kernel<T, kEpl, kC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  db_scan3_kernel<T, kEpl, kC><<<grid, kScan2Warps * 32, smem, s>>>(p);
  return OCRPP_OK;
}

// rows of exactly C full groups for which a specialisation exists (640 / 1024 / 1280 / 1920 / 2560 float32 pixels)
template <typename T, int kEpl>
int launch_scan3_any(const DbParams& p, int C, dim3 grid, cudaStream_t s, bool* done) {
  *done = true;
  switch (C) {
    case 4: return launch_scan3<T, kEpl, 4>(p, grid, s);
    case 5: return launch_scan3<T, kEpl, 5>(p, grid, s);
    case 8: return launch_scan3<T, kEpl, 8>(p, grid, s);
    case 10: return launch_scan3<T, kEpl, 10>(p, grid, s);
    case 15: return launch_scan3<T, kEpl, 15>(p, grid, s);
    case 20: return launch_scan3<T, kEpl, 20>(p, grid, s);
    default: *done = false; return OCRPP_OK;
  }
}

// pixel (x,y) of the lane-major bit mask written by db_scan_kernel
struct BitView {
  const uint32_t* bits;
  int Wd, epl;
  __device__ __forceinline__ bool at(int x, int y) const {
    const int ppg = 32 * epl;
    const int g = x / ppg, r = x - g * ppg;
    return (bits[(size_t)y * Wd + g * epl + (r % epl)] >> (r / epl)) & 1u;
  }
};

// ------------------------------------------------------------------------------------------------
// K2: per-row segments -> dense run table in raster order (both polarities) with per-run pixel sums.
// kImgCtas CTAs per image: each scans the row counts (cheap) and converts its share of the rows.
// ------------------------------------------------------------------------------------------------
constexpr int kRunThreads = 512;
constexpr int kImgCtas = 8;       // minimum CTAs per image for the run-parallel kernels (more for small batches)

__global__ void __launch_bounds__(kRunThreads) db_runs_kernel(DbParams p) {
  extern __shared__ int s_rowptr[];  // [H+1]
  const int n = blockIdx.y + p.n0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = kRunThreads / 32;
  const int32_t* rc = p.srow_cnt + (size_t)n * p.H;
  int32_t* rowptr = p.rowptr + (size_t)n * (p.H + 1);
  const int chunk = (p.H + kRunThreads - 1) / kRunThreads;
  const int y0 = threadIdx.x * chunk, y1 = min(p.H, y0 + chunk);
  int local = 0, over = 0;
  for (int y = y0; y < y1; ++y) {
    const int c = rc[y] & 0x7fffffff;
    over |= c > p.cap;
    local += c;
  }
  int total;
  int base = block_exclusive_scan(local, &total);
  for (int y = y0; y < y1; ++y) {
    s_rowptr[y] = base;
    base += rc[y] & 0x7fffffff;
  }
  if (threadIdx.x == 0) s_rowptr[p.H] = total;
  over = __syncthreads_or(over);
  if (over || total > p.R) {
    if (blockIdx.x == 0) {
      if (threadIdx.x == 0) {
        atomicOr(&p.imgflags[n], OCRPP_IMG_RUN_OVERFLOW);
        p.nruns[n] = 0;
      }
      for (int y = threadIdx.x; y <= p.H; y += kRunThreads) rowptr[y] = 0;
    }
    return;
  }
  if (blockIdx.x == 0) {
    for (int y = threadIdx.x; y <= p.H; y += kRunThreads) rowptr[y] = s_rowptr[y];
    if (threadIdx.x == 0) p.nruns[n] = total;
  }
  const size_t ro = (size_t)n * p.R;
  for (int y = blockIdx.x * nw + warp; y < p.H; y += gridDim.x * nw) {
    const int rbase = s_rowptr[y];
    const int c = rc[y] & 0x7fffffff;
    const int first = (unsigned)rc[y] >> 31;
    const size_t rowid = (size_t)n * p.H + y;
    const unsigned long long* sc = p.scum + rowid * (p.cap + 1);
    constexpr unsigned long long kCumMask = (1ull << 48) - 1ull;
    for (int j = lane; j < c; j += 32) {
      const size_t r = ro + rbase + j;
      const int fg = first ^ (j & 1);
      const unsigned long long e0 = sc[j], e1 = sc[j + 1];   // x << 48 | cumulative sum before x
      p.run_xs[r] = (uint16_t)(e0 >> 48);
      p.run_xe[r] = j + 1 < c ? (uint16_t)((e1 >> 48) - 1) : (uint16_t)(p.W - 1);
      p.run_yf[r] = (uint16_t)(y | (fg << 15));
      p.run_sum[r] = (long long)(((e1 & kCumMask) - (e0 & kCumMask)) << 9);   // 2^-23 units -> 32.32 fixed point
    }
  }
}

// per-component slots live at the index of the component's ROOT run; only roots are ever read
__device__ __forceinline__ void db_init_component(const DbParams& p, size_t r, int cflag) {
  p.area[r] = 0;
  p.xmin[r] = 0x7fffffff; p.xmax[r] = -1; p.ymax[r] = -1;
  p.dmin[r] = 0x7fffffff; p.dmax[r] = -0x7fffffff;
  p.smin[r] = 0x7fffffff; p.smax[r] = -0x7fffffff;
  p.sum[r] = 0; p.fcnt[r] = 0; p.fsum[r] = 0; p.xcnt[r] = 0; p.xsum[r] = 0;
  p.cpar[r] = -1; p.cflag[r] = cflag; p.rowoff[r] = -1;
}

// global-memory fallback of db_ccl_kernel, step 0: every run is its own set and owns fresh slots
__global__ void __launch_bounds__(256) db_slots_init_kernel(DbParams p) {
  const int n = blockIdx.y + p.n0;
  const int nr = p.nruns[n];
  const size_t ro = (size_t)n * p.R;
  for (int r = blockIdx.x * 256 + threadIdx.x; r < nr; r += gridDim.x * 256) {
    p.par[ro + r] = r;
    db_init_component(p, ro + r, 0);
  }
}

constexpr int kRunBlk = 256;

#define FOR_EACH_RUN(r, nr) for (int r = blockIdx.x * kRunBlk + threadIdx.x; r < (nr); r += gridDim.x * kRunBlk)

// K3: link every run with the overlapping same-polarity runs of the row above
// (foreground: 8-connectivity => overlap of [xs-1, xe+1]; background: 4-connectivity => [xs, xe]).
__global__ void __launch_bounds__(kRunBlk) db_link_kernel(DbParams p) {
  const int n = blockIdx.y + p.n0;
  const int nr = p.nruns[n];
  const size_t ro = (size_t)n * p.R;
  const int32_t* rowptr = p.rowptr + (size_t)n * (p.H + 1);
  const uint16_t *xs = p.run_xs + ro, *xe = p.run_xe + ro, *yf = p.run_yf + ro;
  int32_t* par = p.par + ro;
  FOR_EACH_RUN(r, nr) {
    const int y = yf[r] & 0x7fff, fg = yf[r] >> 15;
    if (y == 0) continue;
    const int lo = (int)xs[r] - fg, hi = (int)xe[r] + fg;
    int a = rowptr[y - 1], b = rowptr[y];
    // first run q in row y-1 with xe[q] >= lo
    int l = a, h = b;
    while (l < h) {
      const int m = (l + h) >> 1;
      if ((int)xe[m] < lo) l = m + 1; else h = m;
    }
    for (int q = l; q < b && (int)xs[q] <= hi; ++q)
      if ((yf[q] >> 15) == fg) uf_union(par, q, r);
  }
}

// K4a: flatten; background runs touching the frame mark their region as OUT.
__global__ void __launch_bounds__(kRunBlk) db_flatten_kernel(DbParams p) {
  const int n = blockIdx.y + p.n0;
  const int nr = p.nruns[n];
  const size_t ro = (size_t)n * p.R;
  const uint16_t *xs = p.run_xs + ro, *xe = p.run_xe + ro, *yf = p.run_yf + ro;
  int32_t* par = p.par + ro;
  FOR_EACH_RUN(r, nr) {
    const int root = uf_find(par, r);
    par[r] = root;
    const int y = yf[r] & 0x7fff, fg = yf[r] >> 15;
    if (!fg && (y == 0 || y == p.H - 1 || xs[r] == 0 || xe[r] == p.W - 1) &&
        !(__ldcg(&p.cflag[ro + root]) & kOutFlag))
      p.cflag[ro + root] = kOutFlag;   // test first: same-address stores serialise in L2
  }
}

// K3+K4a fused, shared-memory variant: ONE CTA per image keeps the union-find parent array of the
// whole image in shared memory (4 bytes per run; used when it fits), so the pointer chasing of link and
// flatten runs at shared-memory latency instead of L2 latency. Same result as db_link_kernel +
// db_flatten_kernel (the root of a set is its smallest run index in both).
constexpr int kCclThreads = 1024;
constexpr int kCclSmemMax = 200 * 1024;   // dynamic shared memory the kernel is always launched with

__device__ __forceinline__ int uf_find_s(volatile int* par, int x) {  // with path halving
  int q = par[x];
  while (q != x) {
    const int g = par[q];
    if (g != q) par[x] = g;
    x = q;
    q = g;
  }
  return x;
}

__device__ __forceinline__ void uf_union_s(int* par, int a, int b) {
  while (true) {
    a = uf_find_s(par, a);
    b = uf_find_s(par, b);
    if (a == b) return;
    if (a > b) {
      const int t = a;
      a = b;
      b = t;
    }
    const int old = atomicMin(par + b, a);
    if (old == b) return;
    b = old;
  }
}

__global__ void __launch_bounds__(kCclThreads) db_ccl_kernel(DbParams p) {
  extern __shared__ int s_par[];  // [R] parents, [R/32+1] words of per-run / per-root flags, then (when they
                                  // fit) copies of the run table: the overlap searches below are chains of
                                  // dependent loads, which cost ~30 cycles each from shared memory, ~600 from L2
  const int n = blockIdx.x + p.n0;
  const int nr = p.nruns[n];
  const size_t ro = (size_t)n * p.R;
  const int32_t* rowptr = p.rowptr + (size_t)n * (p.H + 1);
  const uint16_t *xs = p.run_xs + ro, *xe = p.run_xe + ro, *yf = p.run_yf + ro;
  unsigned* s_flag = reinterpret_cast<unsigned*>(s_par + p.R);
  const int nwords = (nr + 31) / 32;
  for (int w = threadIdx.x; w < nwords; w += kCclThreads) s_flag[w] = 0u;
  {
    uint16_t* stage = reinterpret_cast<uint16_t*>(s_flag + (p.R + 31) / 32 + 1);
    const size_t used = sizeof(int) * ((size_t)p.R + (p.R + 31) / 32 + 1);
    if (used + 6 * (size_t)nr + 16 <= (size_t)kCclSmemMax) {
      uint16_t *sx = stage, *se = stage + nr, *sy = stage + 2 * (size_t)nr;
      for (int r = threadIdx.x; r < nr; r += kCclThreads) {
        sx[r] = xs[r];
        se[r] = xe[r];
        sy[r] = yf[r];
      }
      xs = sx;
      xe = se;
      yf = sy;
    }
  }
  __syncthreads();
  // pass 1: every run points at the FIRST overlapping run of the same polarity in the row above (a
  // smaller index), or at itself. No find, no atomic: most runs of a text map overlap exactly one run.
  // Runs with further overlaps are flagged for pass 3.
  for (int r = threadIdx.x; r < nr; r += kCclThreads) {
    const int y = yf[r] & 0x7fff, fg = yf[r] >> 15;
    int first = r;
    if (y > 0) {
      const int lo = (int)xs[r] - fg, hi = (int)xe[r] + fg;
      int l = rowptr[y - 1], h = rowptr[y];
      const int b = h;
      while (l < h) {
        const int m = (l + h) >> 1;
        if ((int)xe[m] < lo) l = m + 1; else h = m;
      }
      bool more = false;
      for (int q = l; q < b && (int)xs[q] <= hi; ++q)
        if ((yf[q] >> 15) == fg) {
          if (first == r) first = q;
          else more = true;
        }
      if (more) atomicOr(&s_flag[r >> 5], 1u << (r & 31));
    }
    s_par[r] = first;
  }
  __syncthreads();
  // pass 2: pointer jumping until every run points at the top of its chain (the background region
  // forms chains as long as the image is high: log2(H) rounds, no divergence)
  while (true) {
    int changed = 0;
    for (int r = threadIdx.x; r < nr; r += kCclThreads) {
      const int q = s_par[r];
      const int g = s_par[q];
      if (g != q) {
        s_par[r] = g;
        changed = 1;
      }
    }
    if (!__syncthreads_or(changed)) break;
  }
  // pass 3: the remaining overlaps merge chains (atomicMin linking of roots, smallest index wins)
  for (int r = threadIdx.x; r < nr; r += kCclThreads) {
    if (!((s_flag[r >> 5] >> (r & 31)) & 1u)) continue;
    const int y = yf[r] & 0x7fff, fg = yf[r] >> 15;
    const int lo = (int)xs[r] - fg, hi = (int)xe[r] + fg;
    int l = rowptr[y - 1], h = rowptr[y];
    const int b = h;
    while (l < h) {
      const int m = (l + h) >> 1;
      if ((int)xe[m] < lo) l = m + 1; else h = m;
    }
    bool skipped = false;
    for (int q = l; q < b && (int)xs[q] <= hi; ++q)
      if ((yf[q] >> 15) == fg) {
        if (skipped) uf_union_s(s_par, q, r);
        skipped = true;
      }
  }
  __syncthreads();
  // pass 4: flatten; background regions that touch the frame are OUT. Thousands of runs share the root
  // of the outside region: the flag is collected in the shared bitmap (test before set) and written once
  // per root, because same-address global stores from a whole CTA serialise in L2.
  for (int w = threadIdx.x; w < nwords; w += kCclThreads) s_flag[w] = 0u;
  __syncthreads();
  for (int r = threadIdx.x; r < nr; r += kCclThreads) {
    const int root = uf_find_s(s_par, r);
    s_par[r] = root;
    p.par[ro + r] = root;
    const int y = yf[r] & 0x7fff, fg = yf[r] >> 15;
    if (!fg && (y == 0 || y == p.H - 1 || xs[r] == 0 || xe[r] == p.W - 1)) {
      const unsigned bit = 1u << (root & 31);
      if (!(reinterpret_cast<volatile unsigned*>(s_flag)[root >> 5] & bit)) atomicOr(&s_flag[root >> 5], bit);
    }
  }
  __syncthreads();
  for (int r = threadIdx.x; r < nr; r += kCclThreads)
    if (s_par[r] == r) db_init_component(p, ro + r, ((s_flag[r >> 5] >> (r & 31)) & 1u) ? kOutFlag : 0);
}

// K4b: per-component reductions from the per-run sums (no pixel is read again). One thread per run.
__global__ void __launch_bounds__(kRunBlk) db_stats_kernel(DbParams p) {
  const int n = blockIdx.y + p.n0;
  const int nr = p.nruns[n];
  const size_t ro = (size_t)n * p.R;
  const uint16_t *xs = p.run_xs + ro, *xe = p.run_xe + ro, *yf = p.run_yf + ro;
  FOR_EACH_RUN(r, nr) {
    const int root = p.par[ro + r];
    const int y = yf[r] & 0x7fff, fg = yf[r] >> 15;
    if (!fg && (p.cflag[ro + root] & kOutFlag)) continue;
    const int a = xs[r], b = xe[r];
    const size_t c = ro + root;
    atomicAdd(&p.area[c], b - a + 1);
    atomicAdd((unsigned long long*)&p.sum[c], (unsigned long long)p.run_sum[ro + r]);
    atomicMin(&p.xmin[c], a);
    atomicMax(&p.xmax[c], b);
    atomicMax(&p.ymax[c], y);
    // diagonal extents feed the "<= 2 contour points" rule only, which needs bbox_w == bbox_h == area, i.e.
    // one pixel per row: a run longer than one pixel rules it out, so only single-pixel runs report
    if (fg && a == b) {
      atomicMin(&p.dmin[c], a - y);
      atomicMax(&p.dmax[c], a - y);
      atomicMin(&p.smin[c], a + y);
      atomicMax(&p.smax[c], a + y);
    }
  }
}

// run of row y that contains pixel x
__device__ __forceinline__ int run_at(const int32_t* rowptr, const uint16_t* xs, int y, int x) {
  int l = rowptr[y], h = rowptr[y + 1];  // last run with xs <= x
  while (h - l > 1) {
    const int m = (l + h) >> 1;
    if ((int)xs[m] <= x) l = m; else h = m;
  }
  return l;
}

__device__ __forceinline__ bool is_candidate_root(const DbParams& p, size_t ro, int r) {
  if (p.par[ro + r] != r) return false;
  const int fg = p.run_yf[ro + r] >> 15;
  return fg || !(p.cflag[ro + r] & kOutFlag);
}

// K5: parent links of the component tree + row-extent slot allocation, one thread per candidate root.
__global__ void __launch_bounds__(kRunBlk) db_tree_kernel(DbParams p) {
  const int n = blockIdx.y + p.n0;
  const int nr = p.nruns[n];
  const size_t ro = (size_t)n * p.R;
  const int32_t* rowptr = p.rowptr + (size_t)n * (p.H + 1);
  const uint16_t *xs = p.run_xs + ro, *yf = p.run_yf + ro;
  FOR_EACH_RUN(r, nr) {
    if (!is_candidate_root(p, ro, r)) continue;
    const int y = yf[r] & 0x7fff, fg = yf[r] >> 15;
    int parent = -1, nrows;
    if (fg) {
      // region LEFT of the component's first pixel: previous run of the same row (background)
      if (xs[r] > 0) {
        const int h = p.par[ro + r - 1];
        if (!(p.cflag[ro + h] & kOutFlag)) parent = h;
      }
      nrows = p.ymax[ro + r] - y + 1;
    } else {
      // pixel ABOVE the hole's first pixel is foreground and belongs to the enclosing component
      parent = p.par[ro + run_at(rowptr, xs, y - 1, xs[r])];
      nrows = p.ymax[ro + r] - y + 3;  // ring rows ymin-1 .. ymax+1
    }
    p.cpar[ro + r] = parent;
    const int off = atomicAdd(&p.ext_alloc[n], nrows + 1);
    if (off + nrows + 1 <= p.E) {
      p.rowoff[ro + r] = off;
      for (int i = 0; i < nrows; ++i) {
        p.ext_l[(size_t)n * p.E + off + i] = 0x7fffffff;
        p.ext_r[(size_t)n * p.E + off + i] = -1;
      }
    }
  }
}

// K6: every candidate adds its own (count, sum) to all its ancestors: fill = own + descendants.
__global__ void __launch_bounds__(kRunBlk) db_fill_kernel(DbParams p) {
  const int n = blockIdx.y + p.n0;
  const int nr = p.nruns[n];
  const size_t ro = (size_t)n * p.R;
  FOR_EACH_RUN(r, nr) {
    if (!is_candidate_root(p, ro, r)) continue;
    const int cnt = p.area[ro + r];
    const long long s = p.sum[ro + r];
    int a = p.cpar[ro + r];
    while (a >= 0) {
      atomicAdd(&p.fcnt[ro + a], cnt);
      atomicAdd((unsigned long long*)&p.fsum[ro + a], (unsigned long long)s);
      a = p.cpar[ro + a];
    }
  }
}


// K7: row extents of every candidate's point set, stair pixels, hole rings. One thread per run.
template <typename T>
__global__ void __launch_bounds__(kRunBlk) db_extents_kernel(DbParams p) {
  const int n = blockIdx.y + p.n0;
  const int nr = p.nruns[n];
  const size_t ro = (size_t)n * p.R;
  const int32_t* rowptr = p.rowptr + (size_t)n * (p.H + 1);
  const uint16_t *xs = p.run_xs + ro, *xe = p.run_xe + ro, *yf = p.run_yf + ro;
  const int32_t* par = p.par + ro;
  const BitView bv{p.bits + (size_t)n * p.H * p.Wd, p.Wd, p.epl};
  int32_t* ext_l = p.ext_l + (size_t)n * p.E;
  int32_t* ext_r = p.ext_r + (size_t)n * p.E;
  const int H = p.H, W = p.W;
  const long long img = n * p.stride_n;

  auto px = [&](int x, int y) { return to_fixed(load_px<T>(p.maps, img + y * p.stride_h + x)); };
  // is pixel (x,y) a background pixel of hole region h ?
  auto in_hole = [&](int x, int y, int h) {
    if (x < 0 || y < 0 || x >= W || y >= H) return false;
    if (bv.at(x, y)) return false;
    return par[run_at(rowptr, xs, y, x)] == h;
  };

  FOR_EACH_RUN(r, nr) {
    const int y = yf[r] & 0x7fff, fg = yf[r] >> 15;
    const int a = xs[r], b = xe[r];
    const int root = par[r];
    if (fg) {
      const int off = p.rowoff[ro + root];
      if (off >= 0) {
        const int i = off + (y - (yf[root] & 0x7fff));
        atomicMin(&ext_l[i], a);
        atomicMax(&ext_r[i], b);
      }
      continue;
    }
    // ---- background run ----
    // (a) stair pixel of the OUTER contour of the component to the left: e = (a, y)
    if (a > 0 && p.stairs) {
      const int C = par[r - 1];
      const bool updn = (y > 0 && bv.at(a, y - 1)) || (y < H - 1 && bv.at(a, y + 1));
      if (updn) {
        const int pc = p.cpar[ro + C];
        const bool outer_region = pc < 0 ? (p.cflag[ro + root] & kOutFlag) != 0 : (root == pc);
        if (outer_region) {
          atomicAdd(&p.xcnt[ro + C], 1);
          atomicAdd((unsigned long long*)&p.xsum[ro + C], (unsigned long long)px(a, y));
        }
      }
    }
    if (p.cflag[ro + root] & kOutFlag) continue;
    // (b) hole run: ring pixels (each counted once: owner = first of up/left/right/down neighbour
    //     that lies in the hole) and their row extents; (c) stair pixels of the hole contour.
    const int h = root;
    const int C = p.cpar[ro + h];  // enclosing foreground component: the ring consists of ITS pixels only
    const int off = p.rowoff[ro + h];
    const int y0 = (yf[h] & 0x7fff) - 1;
    int cnt = 0;
    long long s = 0;
    // foreground pixel of the enclosing component (islands inside the hole are not ring pixels)
    auto in_C = [&](int x, int yy) { return bv.at(x, yy) && par[run_at(rowptr, xs, yy, x)] == C; };
    auto ring = [&](int x, int yy) {
      ++cnt;
      s += px(x, yy);
      if (off >= 0) {
        atomicMin(&ext_l[off + yy - y0], x);
        atomicMax(&ext_r[off + yy - y0], x);
      }
    };
    // hole runs never touch the frame: a-1, b+1, y-1, y+1 are inside the image
    for (int x = a; x <= b; ++x) {
      if (in_C(x, y + 1)) ring(x, y + 1);  // its UP neighbour is in h: always the owner
      if (in_C(x, y - 1)) {                // p = (x, y-1): down neighbour in h
        if (!in_hole(x, y - 2, h) && !in_hole(x - 1, y - 1, h) && !in_hole(x + 1, y - 1, h)) ring(x, y - 1);
      }
    }
    // p = (b+1, y): left neighbour in h; owner unless its up neighbour is in h
    if (in_C(b + 1, y) && !in_hole(b + 1, y - 1, h)) ring(b + 1, y);
    // p = (a-1, y): right neighbour in h; owner unless up or left neighbour is in h
    if (in_C(a - 1, y) && !in_hole(a - 1, y - 1, h) && !in_hole(a - 2, y, h)) ring(a - 1, y);
    // (c) o = (b, y) is the last pixel of a hole run, q = (b+1, y) is foreground; for dy in {-1,+1}:
    //     p = (b, y+dy) in C => the hole contour steps diagonally p <-> q and the 4-connected
    //     boundary also paints e = (b+1, y+dy)
    for (int dy = -1; dy <= 1 && p.stairs; dy += 2) {
      if (!in_C(b, y + dy)) continue;
      const int ex = b + 1, ey = y + dy;
      if (dy == -1) {
        // the same e is produced from o' = (b, y-2) with dy=+1 when that qualifies: count it there
        if (in_hole(b, y - 2, h) && bv.at(b + 1, y - 2)) continue;
      }
      if (bv.at(ex, ey)) {
        // foreground e already belongs to the ring when one of its other neighbours is in h
        if (in_hole(ex + 1, ey, h) || in_hole(ex, ey + dy, h)) continue;
      } else {
        if (par[run_at(rowptr, xs, ey, ex)] == h) continue;  // e itself is hole background
      }
      ++cnt;
      s += px(ex, ey);
    }
    if (cnt) {
      atomicAdd(&p.xcnt[ro + h], cnt);
      atomicAdd((unsigned long long*)&p.xsum[ro + h], (unsigned long long)s);
    }
  }
}

// K8: candidate ranking in cv2 order (reverse raster order of the first point), one CTA per image.
__global__ void __launch_bounds__(kRunThreads) db_rank_kernel(DbParams p) {
  const int n = blockIdx.x + p.n0;
  const int nr = p.nruns[n];
  const size_t ro = (size_t)n * p.R;
  const int chunk = (nr + kRunThreads - 1) / kRunThreads;
  // thread t owns runs [nr - (t+1)*chunk, nr - t*chunk) walked downwards: reverse order
  const int hi = nr - threadIdx.x * chunk, lo = max(0, hi - chunk);
  int local = 0;
  for (int r = hi - 1; r >= lo; --r) local += is_candidate_root(p, ro, r) ? 1 : 0;
  int total;
  int rank = block_exclusive_scan(local, &total);
  for (int r = hi - 1; r >= lo; --r) {
    if (!is_candidate_root(p, ro, r)) continue;
    if (rank < p.maxc) p.cand[(size_t)n * p.maxc + rank] = r;
    ++rank;
  }
  if (threadIdx.x == 0) {
    p.ncand[n] = min(total, p.maxc);
    if (total > p.maxc) atomicOr(&p.imgflags[n], OCRPP_IMG_CANDIDATES_TRUNCATED);
  }
}

// ------------------------------------------------------------------------------------------------
// K9b: generic per-candidate geometry, one warp per candidate: handles the (rare) candidates the
// fast kernel below defers (more than kSmallRows rows, or an unclip polygon above its capacity).
// ------------------------------------------------------------------------------------------------
// ---- the places where the reference's two branches differ (oracle/db_oracle.py lists them) ----
__device__ __forceinline__ float db_side(const DbParams& p, double w, double h) {
  return p.sem ? fminf((float)w, (float)h) : fmaxf((float)w, (float)h);   // py:176 min(w,h) | cpp:161 max(w,h)
}
__device__ __forceinline__ double db_distance(const DbParams& p, const float* mx, const float* my) {
  return p.sem ? geom::unclip_distance_py(mx, my, p.unclip_ratio_d)
               : (double)geom::unclip_distance(mx, my, p.unclip_ratio);
}
__device__ __forceinline__ bool db_low_score(const DbParams& p, double mean) {
  return p.sem ? mean < p.box_thresh_d : (float)mean < p.box_thresh;
}
__device__ __forceinline__ float db_px(const DbParams& p, int n, int x, int y) {
  const long long off = n * p.stride_n + y * p.stride_h + x;
  return p.dtype == OCRPP_F32 ? __ldg(reinterpret_cast<const float*>(p.maps) + off)
                              : __half2float(__ldg(reinterpret_cast<const __half*>(p.maps) + off));
}
// score_mode "box" (db_postprocess.py:109-110,178-194): mean of the map over cv2.fillPoly of the mini box. `lanes`
// threads (lane index li, all in `mask`) share the rows; L / R are scratch for bh row extents, filled by lane 0.
__device__ __forceinline__ double db_box_score(const DbParams& p, int n, const float* mx, const float* my, int* L, int* R,
                                               int li, int lanes, unsigned mask, int bxmin, int bymin, int bw, int bh,
                                               const int* qx, const int* qy) {
  if (li == 0) geom::fill_quad_rows(qx, qy, bw, bh, L, R);
  __syncwarp(mask);
  unsigned long long sum = 0;
  int cnt = 0;
  for (int y = li; y < bh; y += lanes) {
    const int l = L[y], r = R[y];
    for (int x = l; x <= r; ++x) sum += (unsigned long long)to_fixed(db_px(p, n, bxmin + x, bymin + y));
    cnt += r >= l ? r - l + 1 : 0;
  }
  for (int o = lanes >> 1; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(mask, sum, o);
    cnt += __shfl_xor_sync(mask, cnt, o);
  }
  __syncwarp(mask);
  return cnt ? ((double)(long long)sum / kFixScale) / (double)cnt : 0.0;   // cv2.mean of an empty mask is 0
}

// rescale + round + clamp of one corner -> int16 pair and the pre-rounding floats
__device__ __forceinline__ void db_store_corner(const DbParams& p, size_t ko, int q, float mxq, float myq, float sw, float sh) {
  if (p.sem) {
    double fx, fy;
    geom::db_rescale_py(mxq, myq, p.W, p.H, sw, sh, p.pad_resize, &fx, &fy);
    p.res_boxf[ko * 8 + 2 * q] = (float)fx;
    p.res_boxf[ko * 8 + 2 * q + 1] = (float)fy;
    p.res_box[ko * 8 + 2 * q] = (int16_t)(int)fmin(fmax(geom::round_half_even(fx), 0.0), (double)sw);
    p.res_box[ko * 8 + 2 * q + 1] = (int16_t)(int)fmin(fmax(geom::round_half_even(fy), 0.0), (double)sh);
  } else {
    float fx, fy;
    geom::db_rescale(mxq, myq, p.W, p.H, sw, sh, p.pad_resize, &fx, &fy);
    p.res_boxf[ko * 8 + 2 * q] = fx;
    p.res_boxf[ko * 8 + 2 * q + 1] = fy;
    p.res_box[ko * 8 + 2 * q] = (int16_t)(int)fminf(fmaxf(geom::roundf_half_away(fx), 0.f), sw);
    p.res_box[ko * 8 + 2 * q + 1] = (int16_t)(int)fminf(fmaxf(geom::roundf_half_away(fy), 0.f), sh);
  }
}

constexpr int kGeoWarps = 4;
constexpr int kSmallRows = 128;             // candidates up to this many rows build their hull in smem (taller ones: in the global
                                            // workspace, one lane walking it - ~100 us for a single candidate, which then is the
                                            // whole step's tail: the slow rank of the 4- and 8-GPU runs had one such page)
constexpr int kOffCap = 320;                // capacity of the unclip polygon (points)


// The candidates db_geometry_kernel deferred (point sets or unclip polygons that do not fit its fixed buffers; rare):
// one warp per candidate, generic buffers. Runs on the first kGeoWarps warps of db_compact_kernel's CTA, before the
// compaction (a launch of its own cost 6 us per step for a list that is almost always empty).
__device__ __forceinline__ void db_geometry_big_run(const DbParams& p, const int n) {
  __shared__ P2i s_pts[kGeoWarps][2 * kSmallRows];
  __shared__ P2i s_hull[kGeoWarps][2 * kSmallRows + 2];
  __shared__ P2i s_off[kGeoWarps][kOffCap];
  __shared__ P2i s_offh[kGeoWarps][kOffCap + 2];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nbig = min(p.nbig[n], p.maxc);
  for (int bi = wib; bi < nbig; bi += kGeoWarps) {
  const int k = p.big[(size_t)n * p.maxc + bi];
  const size_t ko = (size_t)n * p.maxc + k;
  // 2: deferred by the triage (db_image_kernel / db_hull_kernel; its row extents are in the global workspace),
  // 4: deferred by db_geometry_kernel (its hull is ready in p.hull, packed points, at most 2 * kFastRows + 2 of them)
  const int how = p.res_keep[ko];
  __syncwarp();
  if (lane == 0) p.res_keep[ko] = 0;
  // triaged by db_image_kernel / db_hull_kernel: <= 2-point rule and (score_mode poly) BoxScore passed
  const int off = p.cand_off[ko], nrows = p.cand_nrows[ko], y0 = p.cand_y0[ko];
  P2i* hull;
  int hn = 0;
  if (how == 4) {
    hull = s_hull[wib];
    hn = min(p.hull_n[ko], 2 * kSmallRows + 2);
    const int* gh = reinterpret_cast<const int*>(p.hull + ((size_t)n * p.E + off) * 4);
    for (int i = lane; i < hn; i += 32) hull[i] = P2i{pkx(gh[i]), pky(gh[i])};
    __syncwarp();
  } else {
    const int32_t* ext_l = p.ext_l + (size_t)n * p.E + off;
    const int32_t* ext_r = p.ext_r + (size_t)n * p.E + off;
    P2i* pts;
    if (nrows <= kSmallRows) {
      pts = s_pts[wib];
      hull = s_hull[wib];
    } else {
      pts = p.hull + ((size_t)n * p.E + off) * 4;
      hull = pts + 2 * nrows;
    }
    for (int i = lane; i < nrows; i += 32) {
      pts[2 * i] = P2i{ext_l[i], y0 + i};
      pts[2 * i + 1] = P2i{ext_r[i], y0 + i};
    }
    __syncwarp();
    if (lane == 0) hn = geom::hull_sorted(pts, 2 * nrows, hull);
    hn = __shfl_sync(0xffffffffu, hn, 0);
    __syncwarp();
  }

  geom::Rect rect;
  warp_min_area_rect(hull, hn, &rect, lane);
  // float32 RotatedRect semantics from here on (db_postprocess.cpp:159-192)
  float cx[4], cy[4], mx[4], my[4];
  for (int q = 0; q < 4; ++q) {
    cx[q] = (float)rect.cx[q];
    cy[q] = (float)rect.cy[q];
  }
  geom::mini_box(cx, cy, mx, my);
  const float ssid = db_side(p, rect.w, rect.h);
  if (ssid < 3.f) continue;
  float score = p.res_score[ko];
  if (p.score_box) {
    int bxmin, bymin, bw, bh, qx[4], qy[4];
    geom::box_score_quad(mx, my, p.W, p.H, &bxmin, &bymin, &bw, &bh, qx, qy);
    int* L = bh <= kSmallRows ? reinterpret_cast<int*>(s_pts[wib]) : reinterpret_cast<int*>(p.hull + ((size_t)n * p.E + off) * 4);
    if (bh > kSmallRows && 2 * bh > 8 * (nrows + 1)) {   // cannot happen: the mini box spans at most nrows + 3 rows
      if (lane == 0) atomicOr(&p.imgflags[n], OCRPP_IMG_RUN_OVERFLOW);
      continue;
    }
    __syncwarp();
    const double mean = db_box_score(p, n, mx, my, L, L + bh, lane, 32, 0xffffffffu, bxmin, bymin, bw, bh, qx, qy);
    score = (float)mean;
    if (db_low_score(p, mean)) continue;
  }

  // UnClip (db_postprocess.cpp:34-64)
  const double distance = db_distance(p, mx, my);
  P2i quad[4];
  for (int q = 0; q < 4; ++q) quad[q] = P2i{(int)mx[q], (int)my[q]};
  int m = 0;
  if (lane == 0) {
    m = geom::do_offset_quad(quad, distance, s_off[wib], kOffCap);
    if (m > 0) geom::sort_points_yx(s_off[wib], m);
  }
  m = __shfl_sync(0xffffffffu, m, 0);
  if (m < 0) {  // polygon larger than kOffCap points: unsupported size, fail loudly
    if (lane == 0) atomicOr(&p.imgflags[n], OCRPP_IMG_RUN_OVERFLOW);
    continue;
  }
  if (m == 0) continue;  // empty solution -> RotatedRect((0,0),(1,1),0) -> dropped by the 1.001 test
  int hm = 0;
  if (lane == 0) hm = geom::hull_sorted(s_off[wib], m, s_offh[wib]);
  hm = __shfl_sync(0xffffffffu, hm, 0);
  __syncwarp();
  geom::Rect rect2;
  warp_min_area_rect(s_offh[wib], hm, &rect2, lane);
  if ((float)rect2.h < 1.001 && (float)rect2.w < 1.001) continue;
  for (int q = 0; q < 4; ++q) {
    cx[q] = (float)rect2.cx[q];
    cy[q] = (float)rect2.cy[q];
  }
  geom::mini_box(cx, cy, mx, my);
  const float ssid2 = db_side(p, rect2.w, rect2.h);
  if (ssid2 < 5.f) continue;

  if (lane == 0) {
    const float sw = (float)p.src_wh[2 * n], sh = (float)p.src_wh[2 * n + 1];
    for (int q = 0; q < 4; ++q) db_store_corner(p, ko, q, mx[q], my[q], sw, sh);
    p.res_score[ko] = score;
    p.res_keep[ko] = 1;
  }
  }
}

// ------------------------------------------------------------------------------------------------
// K9: per-candidate geometry, FOUR candidates per warp (groups of 8 lanes), 32-bit integer
// projections, points packed as short2 in shared memory. Candidates that do not fit its fixed
// buffers are appended to the per-image `big` list and handled by db_geometry_big_run.
// ------------------------------------------------------------------------------------------------
constexpr int kGeoThreads = 16 * kGrp;              // 16 candidates per CTA
constexpr int kFastRows = 64;                 // max rows of a candidate's point set
constexpr int kFastPts = 64;                  // max vertices of its hull, and max rows of its mini box in score_mode box
constexpr int kFastOff = 48;                  // max points of its unclip polygon. A candidate of this path is at most kFastRows
                                              // rows high, so its unclip distance (area * ratio / perimeter < 0.85 * height at
                                              // ratio 1.7) stays below ~55 px and the four round joins make <= 40 points;
                                              // anything longer is deferred. The group's shared memory decides how many CTAs an
                                              // SM holds: 96 points of 8 bytes -> 64 -> 48 packed points of 4 bytes = 7 -> 8 -> 11


// K9a: candidate triage + convex hull of the row extents, ONE THREAD per candidate (the monotone chain is
// sequential: with a thread per candidate a warp advances 32 hulls at once instead of 4). The hull is
// written, packed, to the candidate's slice of the hull scratch; res_keep carries the verdict:
// 0 dropped, 2 deferred to db_geometry_big_run, 3 hull ready (hull_n vertices).
constexpr int kHullThreads = 64;

__global__ void __launch_bounds__(kHullThreads) db_hull_kernel(DbParams p) {
  const int n = blockIdx.y + p.n0;
  const int k = blockIdx.x * kHullThreads + threadIdx.x;
  if (k >= p.ncand[n]) return;
  const size_t ro = (size_t)n * p.R;
  const size_t ko = (size_t)n * p.maxc + k;
  const int c = p.cand[ko];
  const int yf = p.run_yf[ro + c];
  const int fg = yf >> 15, y_first = yf & 0x7fff;
  const int ymax = p.ymax[ro + c];
  const int area = p.area[ro + c];
  p.res_keep[ko] = 0;
  if (fg && p.skip2) {  // "contour has <= 2 points" (db_postprocess.cpp:255-257)
    const int bw = p.xmax[ro + c] - p.xmin[ro + c] + 1, bh = ymax - y_first + 1;
    const bool diag = (bw == bh && bw == area) &&
                      (p.dmin[ro + c] == p.dmax[ro + c] || p.smin[ro + c] == p.smax[ro + c]);
    if (area == 1 || (bh == 1 && area == bw) || (bw == 1 && area == bh) || diag) return;
  }
  const int nrows = fg ? (ymax - y_first + 1) : (ymax - y_first + 3);
  const int off = p.rowoff[ro + c];
  if (off < 0) {  // extent arena exhausted (cannot happen with E = 4R); fail loudly
    atomicOr(&p.imgflags[n], OCRPP_IMG_RUN_OVERFLOW);
    return;
  }
  // BoxScore first: a low score drops the candidate whatever its rectangle is
  const long long tot = p.sum[ro + c] + p.fsum[ro + c] + p.xsum[ro + c];
  const int cnt = area + p.fcnt[ro + c] + p.xcnt[ro + c];
  const double mean = ((double)tot / kFixScale) / (double)cnt;
  const float score = (float)mean;
  if (!p.score_box && db_low_score(p, mean)) return;
  p.res_score[ko] = score;
  p.cand_off[ko] = off;
  p.cand_y0[ko] = fg ? y_first : y_first - 1;
  p.cand_nrows[ko] = nrows;
  if (nrows > kFastRows || p.W >= 16384 || p.H >= 16384) {
    const int slot = atomicAdd(&p.nbig[n], 1);
    if (slot < p.maxc) p.big[(size_t)n * p.maxc + slot] = k;
    p.res_keep[ko] = 2;
    return;
  }

  const int y0 = fg ? y_first : y_first - 1;
  const int32_t* ext_l = p.ext_l + (size_t)n * p.E + off;
  const int32_t* ext_r = p.ext_r + (size_t)n * p.E + off;
  int* gout = reinterpret_cast<int*>(p.hull + ((size_t)n * p.E + off) * 4);   // >= 8 * (nrows + 1) ints
  int out[2 * kFastRows + 2];   // dynamically indexed => thread-local memory, which L1 caches write-back
  // monotone chain over the points (ext_l[i], y0+i), (ext_r[i], y0+i), already sorted by (y, x); same result as
  // hull_sorted32 on all of them, but every pass only visits the extents that can stay on its chain: with y
  // ascending a kept turn (cross > 0) bulges towards +x, so the first pass is the chain of RIGHT extents and a left
  // extent of a later row is always popped again by the right extent of its own row - and whatever it popped
  // before, that right extent pops too (cross(o, a, q) only decreases as q moves right along a row). Mirrored for
  // the second pass. The two top-of-stack points stay in registers.
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
    const int lo = kk + 1;
    for (int i = last; i >= 0; --i) {
      const int q = pk(ext_l[i], y0 + i);
      if (q == prev) continue;   // single-pixel last row
      prev = q;
      while (kk >= lo && cross32(a, b, q) <= 0) {
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

#include "db_image.cuh"

__global__ void __launch_bounds__(kGeoThreads) db_geometry_kernel(DbParams p) {
  constexpr int kGroups = kGeoThreads / kGrp;
  __shared__ int s_a[kGroups][kFastPts];             // sorted unclip polygon | fill rows of score_mode box
  __shared__ int s_b[kGroups][kFastPts + 2];         // hull
  __shared__ int s_off[kGroups][kFastOff];           // raw unclip polygon, points packed x | y << 16
  const int n = blockIdx.y + p.n0;
  const int g = threadIdx.x / kGrp, gl = threadIdx.x % kGrp;
  const unsigned gmask = ((1u << kGrp) - 1u) << ((threadIdx.x & 31) / kGrp * kGrp);
  const int k = blockIdx.x * kGroups + g;
  if (k >= p.ncand[n]) return;
  const size_t ko = (size_t)n * p.maxc + k;
  if (p.res_keep[ko] != 3) return;   // dropped or deferred by the triage (db_image_kernel / db_hull_kernel)
  __syncwarp(gmask);
  if (gl == 0) p.res_keep[ko] = 0;
  float score = p.res_score[ko];
  const int off = p.cand_off[ko];
  // hand the candidate to db_geometry_big_run (generic buffers). Mark 4 = "its hull is ready in p.hull": the row
  // extents it was built from may live only in db_image_kernel's shared memory, so they cannot be read again
  auto defer = [&]() {
    if (gl == 0) {
      p.res_keep[ko] = 4;
      const int slot = atomicAdd(&p.nbig[n], 1);
      if (slot < p.maxc) p.big[(size_t)n * p.maxc + slot] = k;
    }
  };
  int* A = s_a[g];
  int* B = s_b[g];
  const int hn = p.hull_n[ko];
  if (hn > kFastPts) {   // a hull of more points than the group's buffer (a near-ellipse 33+ rows high)
    defer();
    return;
  }
  {
    const int* hull = reinterpret_cast<const int*>(p.hull + ((size_t)n * p.E + off) * 4);
    for (int i = gl; i < hn; i += kGrp) B[i] = hull[i];
  }
  __syncwarp(gmask);
  geom::Rect rect;
  group_min_area_rect(B, hn, &rect, gl, gmask);
  float cx[4], cy[4], mx[4], my[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    cx[q] = (float)rect.cx[q];
    cy[q] = (float)rect.cy[q];
  }
  geom::mini_box(cx, cy, mx, my);
  if (db_side(p, rect.w, rect.h) < 3.f) return;
  if (p.score_box) {
    int bxmin, bymin, bw, bh, qx[4], qy[4];
    geom::box_score_quad(mx, my, p.W, p.H, &bxmin, &bymin, &bw, &bh, qx, qy);
    if (bh > kFastPts) {
      defer();
      return;
    }
    __syncwarp(gmask);
    const double mean = db_box_score(p, n, mx, my, A, B, gl, kGrp, gmask, bxmin, bymin, bw, bh, qx, qy);
    score = (float)mean;
    if (db_low_score(p, mean)) return;
  }

  // UnClip (db_postprocess.cpp:34-64)
  const double distance = db_distance(p, mx, my);
  P2i quad[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) quad[q] = P2i{(int)mx[q], (int)my[q]};
  const int m = group_do_offset_quad(quad, distance, s_off[g], kFastOff, gl, gmask);
  if (m < 0) {
    defer();
    return;
  }
  if (m == 0) return;  // empty solution -> RotatedRect((0,0),(1,1),0) -> dropped by the 1.001 test
  __syncwarp(gmask);
  // rank sort by (y, x, index) into A: every lane ranks its own points against all of them (one packed key per point:
  // coordinates are within +-16384 on this path)
  for (int i = gl; i < m; i += kGrp) {
    const int q = s_off[g][i];
    const unsigned kq = (unsigned)q ^ 0x80008000u;   // both signed halves biased: unsigned order = (y, x) order
    int rank = 0;
    for (int j = 0; j < m; ++j) {
      const unsigned ko = (unsigned)s_off[g][j] ^ 0x80008000u;
      rank += (ko < kq || (ko == kq && j < i)) ? 1 : 0;
    }
    A[rank] = q;
  }
  __syncwarp(gmask);
  // monotone chain (same result as hull_sorted32): its two passes are independent - the second starts from the last
  // point of the sorted list, which the first always ends on - so lanes 0 and 1 run them at the same time with ONE
  // instruction stream (direction = lane): base point + a stack of the points pushed after it, popped while the
  // turn is not strictly left. Lane 0's stack is B[1..], lane 1's goes to the (dead) raw polygon buffer.
  int hm = 0;
  {
    int* S = gl == 0 ? B + 1 : s_off[g];
    int cnt = 0;
    if (gl < 2) {
      const int dir = gl;
      const int base = A[dir ? m - 1 : 0];
      int prevq = base;
      for (int step = 1; step < m; ++step) {
        const int q = A[dir ? m - 1 - step : step];
        if (q == prevq) continue;
        prevq = q;
        while (cnt >= 1 && cross32(cnt >= 2 ? S[cnt - 2] : base, S[cnt - 1], q) <= 0) --cnt;
        S[cnt++] = q;
      }
      if (gl == 0) B[0] = base;
    }
    const int lane0 = (threadIdx.x & 31) & ~(kGrp - 1);
    const int c1 = __shfl_sync(gmask, cnt, lane0), c2 = __shfl_sync(gmask, cnt, lane0 + 1);
    const int k1 = 1 + c1;
    hm = k1 > 1 ? k1 + c2 - 1 : 1;
    __syncwarp(gmask);
    const int* S2 = s_off[g];
    for (int j = gl; j < c2 - 1; j += kGrp) B[k1 + j] = S2[j];
  }
  __syncwarp(gmask);
  geom::Rect rect2;
  group_min_area_rect(B, hm, &rect2, gl, gmask);
  if ((float)rect2.h < 1.001 && (float)rect2.w < 1.001) return;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    cx[q] = (float)rect2.cx[q];
    cy[q] = (float)rect2.cy[q];
  }
  geom::mini_box(cx, cy, mx, my);
  if (db_side(p, rect2.w, rect2.h) < 5.f) return;

  if (gl == 0) {
    const float sw = (float)p.src_wh[2 * n], sh = (float)p.src_wh[2 * n + 1];
#pragma unroll
    for (int q = 0; q < 4; ++q) db_store_corner(p, ko, q, mx[q], my[q], sw, sh);
    p.res_score[ko] = score;
    p.res_keep[ko] = 1;
  }
}

// K10: ordered compaction of the kept boxes, one CTA per image.
__global__ void __launch_bounds__(kRunThreads) db_compact_kernel(DbParams p) {
  const int n = blockIdx.x + p.n0;
  if (p.nbig[n] > 0) {   // uniform over the CTA
    if ((threadIdx.x >> 5) < kGeoWarps) db_geometry_big_run(p, n);
    __syncthreads();
  }
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

// debug: canonical 8-connected foreground label map (id = 1 + rank of the first raster pixel)
__global__ void __launch_bounds__(kRunThreads) db_labels_kernel(DbParams p) {
  const int n = blockIdx.x + p.n0;
  const int nr = p.nruns[n];
  const size_t ro = (size_t)n * p.R;
  const int chunk = (nr + kRunThreads - 1) / kRunThreads;
  const int lo = min(nr, (int)threadIdx.x * chunk), hi = min(nr, lo + chunk);
  int local = 0;
  for (int r = lo; r < hi; ++r) local += (p.par[ro + r] == r && (p.run_yf[ro + r] >> 15)) ? 1 : 0;
  int total;
  int id = block_exclusive_scan(local, &total);
  // publish each root's id through its `dmin` slot (dead after the geometry kernel)
  for (int r = lo; r < hi; ++r)
    if (p.par[ro + r] == r && (p.run_yf[ro + r] >> 15)) p.dmin[ro + r] = ++id;
  __syncthreads();
  int32_t* lab = p.labels_dbg + (size_t)n * p.H * p.W;
  for (int r = threadIdx.x; r < nr; r += kRunThreads) {
    if (!(p.run_yf[ro + r] >> 15)) continue;
    const int y = p.run_yf[ro + r] & 0x7fff;
    const int v = p.dmin[ro + p.par[ro + r]];
    for (int x = p.run_xs[ro + r]; x <= p.run_xe[ro + r]; ++x) lab[(size_t)y * p.W + x] = v;
  }
}

size_t carve(DbParams& p, void* ws) {
  Carver c{(char*)ws, 0};
  const size_t N = p.N, R = p.R, E = p.E;
  p.bits = c.take<uint32_t>(N * p.H * p.Wd);
  p.rawbits = c.take<uint32_t>(N * p.H * p.Wd);
  p.scum = c.take<unsigned long long>(N * p.H * (p.cap + 1));
  p.srow_cnt = c.take<int32_t>(N * p.H);
  p.run_sum = c.take<long long>(N * R);
  p.rowptr = c.take<int32_t>(N * (p.H + 1));
  p.run_xs = c.take<uint16_t>(N * R);
  p.run_xe = c.take<uint16_t>(N * R);
  p.run_yf = c.take<uint16_t>(N * R);
  p.par = c.take<int32_t>(N * R);
  p.area = c.take<int32_t>(N * R);
  p.xmin = c.take<int32_t>(N * R);
  p.xmax = c.take<int32_t>(N * R);
  p.ymax = c.take<int32_t>(N * R);
  p.dmin = c.take<int32_t>(N * R);
  p.dmax = c.take<int32_t>(N * R);
  p.smin = c.take<int32_t>(N * R);
  p.smax = c.take<int32_t>(N * R);
  p.sum = c.take<long long>(N * R);
  p.fcnt = c.take<int32_t>(N * R);
  p.fsum = c.take<long long>(N * R);
  p.xcnt = c.take<int32_t>(N * R);
  p.xsum = c.take<long long>(N * R);
  p.cpar = c.take<int32_t>(N * R);
  p.cflag = c.take<int32_t>(N * R);
  p.rowoff = c.take<int32_t>(N * R);
  p.ext_l = c.take<int32_t>(N * E);
  p.ext_r = c.take<int32_t>(N * E);
  p.hull = c.take<P2i>(N * E * 4);
  p.nruns = c.take<int32_t>(5 * N);  // nruns | ext_alloc | imgflags | ncand | nbig, cleared together
  p.ext_alloc = p.nruns ? p.nruns + N : nullptr;
  p.imgflags = p.nruns ? p.nruns + 2 * N : nullptr;
  p.ncand = p.nruns ? p.nruns + 3 * N : nullptr;
  p.nbig = p.nruns ? p.nruns + 4 * N : nullptr;
  p.big = c.take<int32_t>(N * p.maxc);
  p.cand = c.take<int32_t>(N * p.maxc);
  p.res_keep = c.take<int32_t>(N * p.maxc);
  p.hull_n = c.take<int32_t>(N * p.maxc);
  p.cand_off = c.take<int32_t>(N * p.maxc);
  p.cand_y0 = c.take<int32_t>(N * p.maxc);
  p.cand_nrows = c.take<int32_t>(N * p.maxc);
  p.hcnt = c.take<int32_t>(N * p.maxc * 2);
  p.cand_root = c.take<int32_t>(N * R);
  p.flagw = c.take<uint32_t>(N * (R / 32 + 1));
  p.res_box = c.take<int16_t>(N * p.maxc * 8);
  p.res_boxf = c.take<float>(N * p.maxc * 8);
  p.res_score = c.take<float>(N * p.maxc);
  return align_up(c.off, 256);
}

int db_words_per_row(int W) { return 8 * ((W + 255) / 256); }  // covers the 1-, 4- and 8-pixel-per-lane layouts

int db_row_cap(int H, int W, int R) {
  long long c = 4ll * R / H;
  if (c < 32) c = 32;
  if (c > W) c = W;
  return (int)c;
}

int resolve_max_runs(int H, int W, int max_runs) {
  const long long worst = (long long)H * (W + 1);
  if (max_runs <= 0 || max_runs > worst) return (int)worst;
  return max_runs;
}

constexpr int kImgSmemMax = 225 * 1024;   // dynamic shared memory of db_image_kernel (227 KB per CTA minus static)

// auxiliary streams / events of the current device (created once, never destroyed). Every sub-batch pipeline has a
// LOW-priority stream for its map scan and a HIGH-priority stream for everything after it: the block scheduler
// then places the (few, latency-bound) stage-2 / geometry CTAs of sub-batch i as soon as SM resources free up,
// instead of queueing them behind the thousands of scan CTAs of sub-batches i+1.. that are already pending.
constexpr int kDbMaxSplit = 8;
struct DbAux {
  cudaStream_t lo[kDbMaxSplit], hi[kDbMaxSplit];
  cudaEvent_t fork, scanned[kDbMaxSplit], join[kDbMaxSplit];
};
std::mutex g_db_enqueue_mu;   // the auxiliary streams/events are shared: one fork/join section at a time

DbAux* db_aux() {
  static std::mutex mu;   // first use may come from several host threads
  std::lock_guard<std::mutex> lock(mu);
  static DbAux aux[64];
  static bool ready[64] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (!ready[dev]) {
    int least = 0, greatest = 0;
    if (cudaDeviceGetStreamPriorityRange(&least, &greatest) != cudaSuccess) return nullptr;
    for (int i = 0; i < kDbMaxSplit; ++i) {
      if (cudaStreamCreateWithPriority(&aux[dev].lo[i], cudaStreamNonBlocking, least) != cudaSuccess) return nullptr;
      if (cudaStreamCreateWithPriority(&aux[dev].hi[i], cudaStreamNonBlocking, greatest) != cudaSuccess) return nullptr;
      if (cudaEventCreateWithFlags(&aux[dev].scanned[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
      if (cudaEventCreateWithFlags(&aux[dev].join[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
    }
    if (cudaEventCreateWithFlags(&aux[dev].fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    ready[dev] = true;
  }
  return &aux[dev];
}

// the whole chain for images [p.n0, p.n0 + N): the map scan on stream s_scan, everything after it on stream s
// (when they differ, `scanned` orders the two)
int db_pipeline(DbParams p, int N, int dtype, cudaStream_t s_scan, cudaStream_t s, cudaEvent_t scanned,
                ProfileScope* prof) {
  const int epl = dtype == OCRPP_F32 ? 4 : 8;
  const bool vec = ((uintptr_t)p.maps % 16 == 0) && (p.stride_n % epl == 0) && (p.stride_h % epl == 0) && (p.W % epl == 0);
  // Stage 2. Default: ONE kernel, one CTA per image, tables in shared memory (db_image.cuh). Large images
  // (many rows / pixels: a single CTA per image would be the bottleneck) and the test hook take the run-parallel
  // multi-kernel chain, which spreads every image over several CTAs with its tables in the global workspace.
  const int path = tuning(OCRPP_TUNE_DB_PATH);
  const bool fused = path != 3 && p.H <= 8191 && (long long)p.H * p.W <= (4ll << 20);
  // the two-phase scan (opt-in through the tuning hook: measured 5 % slower than the single-phase scan on B200, see
  // DESIGN.md): needs the vector layout, no dilation, a consumer that does not read the global bit mask (the
  // one-kernel stage 2) and a row that fits the warp's shared-memory slice
  const int scan2_C = (p.W / epl + 31) / 32;
  const bool scan2 = fused && vec && !p.dilate && scan2_C <= 32 && tuning(OCRPP_TUNE_DB_SCAN) >= 2 &&
                     kScan2Warps * scan2_row_words(scan2_C, epl) * sizeof(uint32_t) <= 100 * 1024;
  // the bulk-copy fed scan (default): same consumer and layout requirements as the two-phase scan; rows are copied as
  // whole 16-byte cells (vec: base, strides and width are multiples of 16 bytes)
  bool scan4_done = false;
  if (fused && vec && !p.dilate && tuning(OCRPP_TUNE_DB_SCAN) == 0) {
    Scan4Params q{};
    q.maps = p.maps; q.stride_n = p.stride_n; q.stride_h = p.stride_h;
    q.H = p.H; q.n0 = p.n0; q.nimg = N; q.cap = p.cap;
    q.ncells = p.W / epl;
    q.thresh = p.thresh;
    q.scum = p.scum; q.srow_cnt = p.srow_cnt; q.imgflags = p.imgflags;
    const int rc = dtype == OCRPP_F32 ? scan4_any<float>(q, s_scan) : scan4_any<__half>(q, s_scan);
    if (rc > 0) return rc;
    scan4_done = rc == OCRPP_OK;
  }
  {
    dim3 grid((p.H + kBinWarps - 1) / kBinWarps, N);
    p.epl = vec ? epl : 1;
    if (p.dilate) {
      if (dtype == OCRPP_F32) {
        if (vec) db_rawbits_kernel<float, 4><<<grid, kBinWarps * 32, 0, s_scan>>>(p);
        else db_rawbits_kernel<float, 1><<<grid, kBinWarps * 32, 0, s_scan>>>(p);
      } else {
        if (vec) db_rawbits_kernel<__half, 8><<<grid, kBinWarps * 32, 0, s_scan>>>(p);
        else db_rawbits_kernel<__half, 1><<<grid, kBinWarps * 32, 0, s_scan>>>(p);
      }
      OCRPP_LAUNCHED();
      if (dtype == OCRPP_F32) {
        if (vec) db_scan_kernel<float, 4, 4, true><<<grid, kBinWarps * 32, 0, s_scan>>>(p);
        else db_scan_kernel<float, 1, 4, true><<<grid, kBinWarps * 32, 0, s_scan>>>(p);
      } else {
        if (vec) db_scan_kernel<__half, 8, 4, true><<<grid, kBinWarps * 32, 0, s_scan>>>(p);
        else db_scan_kernel<__half, 1, 4, true><<<grid, kBinWarps * 32, 0, s_scan>>>(p);
      }
    } else if (scan4_done) {
    } else if (scan2) {
      const int smem = kScan2Warps * scan2_row_words(scan2_C, epl) * (int)sizeof(uint32_t);
      dim3 grid2((p.H + kScan2Warps - 1) / kScan2Warps, N);
      bool done = false;
      if (p.W == 32 * epl * scan2_C && tuning(OCRPP_TUNE_DB_SCAN) != 3) {   // compile-time specialisation for this width
        const int rc = dtype == OCRPP_F32 ? launch_scan3_any<float, 4>(p, scan2_C, grid2, s_scan, &done)
                                          : launch_scan3_any<__half, 8>(p, scan2_C, grid2, s_scan, &done);
        if (rc != OCRPP_OK) return rc;
      }
      if (done) {
      } else
      if (dtype == OCRPP_F32) {
        OCRPP_CUDA(cudaFuncSetAttribute(db_scan2_kernel<float, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        db_scan2_kernel<float, 4><<<grid2, kScan2Warps * 32, smem, s_scan>>>(p, scan2_C);
      } else {
        OCRPP_CUDA(cudaFuncSetAttribute(db_scan2_kernel<__half, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        db_scan2_kernel<__half, 8><<<grid2, kScan2Warps * 32, smem, s_scan>>>(p, scan2_C);
      }
    } else if (dtype == OCRPP_F32) {
      // the tail loop (one group per iteration, its load latency exposed) costs ~15 % on short rows
      const int ng = p.W % 128 == 0 ? p.W / 128 : 0;   // whole groups per row: pick U so that no tail loop is left
      if (vec && ng > 0 && ng % 5 == 0) db_scan_kernel<float, 4, 5><<<grid, kBinWarps * 32, 0, s_scan>>>(p);
      else if (vec && ng > 0 && ng % 4 != 0 && ng % 3 == 0) db_scan_kernel<float, 4, 3><<<grid, kBinWarps * 32, 0, s_scan>>>(p);
      else if (vec) db_scan_kernel<float, 4><<<grid, kBinWarps * 32, 0, s_scan>>>(p);
      else db_scan_kernel<float, 1><<<grid, kBinWarps * 32, 0, s_scan>>>(p);
    } else {
      if (vec) db_scan_kernel<__half, 8><<<grid, kBinWarps * 32, 0, s_scan>>>(p);
      else db_scan_kernel<__half, 1><<<grid, kBinWarps * 32, 0, s_scan>>>(p);
    }
    OCRPP_LAUNCHED();
    if (prof) prof->mark("db_scan");
    if (s_scan != s) {
      OCRPP_CUDA(cudaEventRecord(scanned, s_scan));
      OCRPP_CUDA(cudaStreamWaitEvent(s, scanned, 0));
    }
  }
  if (fused) {
    const size_t want = sizeof(int) * (p.H + 1) + 10 * (size_t)p.R + 68 * ((size_t)p.R / 4 + 64) +
                        8 * ((size_t)p.R + p.H) + 64;
    // 176 KB by default: what is left of the SM's 228 KB stays L1 for the kernel's global reads (row segments of the
    // scan, pixels under stair / ring positions): measured 0.189 -> 0.174 ms per 256 pages against the full 225 KB.
    // Pages whose tables need more take the global-table variant of the same kernel (db_image_kernel, mode 0).
    int smem_cap = 176 * 1024;
    if (tuning(OCRPP_TUNE_DB_IMG_SMEM_KB) > 0) smem_cap = min(kImgSmemMax, tuning(OCRPP_TUNE_DB_IMG_SMEM_KB) * 1024);
    const int smem = (int)(want < (size_t)smem_cap ? want : (size_t)smem_cap);
    const int mode = path == 2 ? 1 : 0;
    if (dtype == OCRPP_F32) {
      OCRPP_CUDA(cudaFuncSetAttribute(db_image_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, kImgSmemMax));
      db_image_kernel<float><<<N, kImgThreads, smem, s>>>(p, smem, mode);
    } else {
      OCRPP_CUDA(cudaFuncSetAttribute(db_image_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, kImgSmemMax));
      db_image_kernel<__half><<<N, kImgThreads, smem, s>>>(p, smem, mode);
    }
    OCRPP_LAUNCHED();
    if (prof) prof->mark("db_image");
  } else {
    const int ictas = max(kImgCtas, min(64, (kNumSMs * 4 + N - 1) / N));   // fill the GPU at small batch sizes too
    db_runs_kernel<<<dim3(ictas, N), kRunThreads, sizeof(int) * (p.H + 1), s>>>(p);
    OCRPP_LAUNCHED();
    if (prof) prof->mark("db_runs");
    dim3 rgrid(ictas, N);
    const size_t ccl_smem = sizeof(int) * ((size_t)p.R + (p.R + 31) / 32 + 1);
    if (ccl_smem <= (size_t)kCclSmemMax) {
      // opt in to > 48 KB of dynamic shared memory (a per-device function attribute: set on every call)
      OCRPP_CUDA(cudaFuncSetAttribute(db_ccl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kCclSmemMax));
      db_ccl_kernel<<<N, kCclThreads, kCclSmemMax, s>>>(p);
      OCRPP_LAUNCHED();
      if (prof) prof->mark("db_ccl");
    } else {
      db_slots_init_kernel<<<rgrid, 256, 0, s>>>(p);
      OCRPP_LAUNCHED();
      db_link_kernel<<<rgrid, kRunBlk, 0, s>>>(p);
      OCRPP_LAUNCHED();
      db_flatten_kernel<<<rgrid, kRunBlk, 0, s>>>(p);
      OCRPP_LAUNCHED();
      if (prof) prof->mark("db_ccl");
    }
    db_stats_kernel<<<rgrid, kRunBlk, 0, s>>>(p);
    OCRPP_LAUNCHED();
    if (prof) prof->mark("db_stats");
    db_tree_kernel<<<rgrid, kRunBlk, 0, s>>>(p);
    OCRPP_LAUNCHED();
    if (prof) prof->mark("db_tree");
    db_fill_kernel<<<rgrid, kRunBlk, 0, s>>>(p);
    OCRPP_LAUNCHED();
    if (prof) prof->mark("db_fill");
    if (dtype == OCRPP_F32) db_extents_kernel<float><<<rgrid, kRunBlk, 0, s>>>(p);
    else db_extents_kernel<__half><<<rgrid, kRunBlk, 0, s>>>(p);
    OCRPP_LAUNCHED();
    if (prof) prof->mark("db_extents");
    db_rank_kernel<<<N, kRunThreads, 0, s>>>(p);
    OCRPP_LAUNCHED();
    if (prof) prof->mark("db_rank");
    db_hull_kernel<<<dim3((p.maxc + kHullThreads - 1) / kHullThreads, N), kHullThreads, 0, s>>>(p);
    OCRPP_LAUNCHED();
    if (prof) prof->mark("db_hull");
    if (p.labels_dbg) {
      db_labels_kernel<<<N, kRunThreads, 0, s>>>(p);
      OCRPP_LAUNCHED();
    }
  }
  {
    constexpr int kGroups = kGeoThreads / kGrp;
    dim3 grid((p.maxc + kGroups - 1) / kGroups, N);
    db_geometry_kernel<<<grid, kGeoThreads, 0, s>>>(p);
    OCRPP_LAUNCHED();
    if (prof) prof->mark("db_geometry");
  }
  db_compact_kernel<<<N, kRunThreads, 0, s>>>(p);
  OCRPP_LAUNCHED();
  if (prof) prof->mark("db_compact");
  return OCRPP_OK;
}

}  // namespace
}  // namespace ocrpp

extern "C" size_t ocrpp_db_workspace_bytes(int N, int H, int W, int max_runs) {
  using namespace ocrpp;
  if (N <= 0 || H <= 0 || W <= 0) return 0;
  DbParams p{};
  p.N = N; p.H = H; p.W = W; p.Wd = db_words_per_row(W);
  p.R = resolve_max_runs(H, W, max_runs);
  p.cap = db_row_cap(H, W, p.R);
  p.E = 4 * p.R + 4;
  p.maxc = 1000;  // max_candidates is capped at the reference constant
  return carve(p, nullptr);
}

extern "C" int ocrpp_db_postprocess(const void* maps_dev, int dtype, int N, int H, int W,
                                    int64_t stride_n, int64_t stride_h, const int32_t* src_wh_dev,
                                    float thresh, float box_thresh, float unclip_ratio,
                                    int max_candidates, int max_runs, int use_dilation, int use_padding_resize,
                                    int16_t* boxes_out_dev,
                                    float* scores_out_dev, int32_t* counts_out_dev,
                                    int32_t* status_out_dev, float* boxes_f_out_dev,
                                    int32_t* labels_dbg_dev, void* workspace_dev,
                                    size_t workspace_bytes, void* stream) {
  return ocrpp_db_postprocess_ex(maps_dev, dtype, N, H, W, stride_n, stride_h, src_wh_dev, thresh, (double)box_thresh,
                                 (double)unclip_ratio, max_candidates, max_runs, use_dilation, use_padding_resize,
                                 OCRPP_DB_SEMANTICS_CPP, OCRPP_DB_SCORE_POLY, boxes_out_dev, scores_out_dev,
                                 counts_out_dev, status_out_dev, boxes_f_out_dev, labels_dbg_dev, workspace_dev,
                                 workspace_bytes, stream);
}

extern "C" int ocrpp_db_postprocess_ex(const void* maps_dev, int dtype, int N, int H, int W,
                                       int64_t stride_n, int64_t stride_h, const int32_t* src_wh_dev,
                                       float thresh, double box_thresh_d, double unclip_ratio_d,
                                       int max_candidates, int max_runs, int use_dilation, int use_padding_resize,
                                       int semantics, int score_mode,
                                       int16_t* boxes_out_dev,
                                       float* scores_out_dev, int32_t* counts_out_dev,
                                       int32_t* status_out_dev, float* boxes_f_out_dev,
                                       int32_t* labels_dbg_dev, void* workspace_dev,
                                       size_t workspace_bytes, void* stream) {
  using namespace ocrpp;
  const float box_thresh = (float)box_thresh_d, unclip_ratio = (float)unclip_ratio_d;
  OCRPP_CHECK_ARG(semantics == OCRPP_DB_SEMANTICS_CPP || semantics == OCRPP_DB_SEMANTICS_PYTHON, "db: bad semantics %d", semantics);
  OCRPP_CHECK_ARG(score_mode == OCRPP_DB_SCORE_POLY || (score_mode == OCRPP_DB_SCORE_BOX && semantics == OCRPP_DB_SEMANTICS_PYTHON),
                  "db: score_mode box exists in the Python branch's semantics only");
  OCRPP_CHECK_ARG(dtype == OCRPP_F32 || dtype == OCRPP_F16, "db: dtype must be OCRPP_F32 or OCRPP_F16");
  OCRPP_CHECK_ARG(N >= 0 && H > 0 && W > 0, "db: bad shape N=%d H=%d W=%d", N, H, W);
  OCRPP_CHECK_ARG(H < 32768 && W < 65536, "db: H must be < 32768 and W < 65536");
  OCRPP_CHECK_ARG(max_candidates > 0 && max_candidates <= 1000, "db: max_candidates must be in [1,1000]");
  if (N == 0) return OCRPP_OK;
  OCRPP_CHECK_ARG(maps_dev && src_wh_dev && boxes_out_dev && scores_out_dev && counts_out_dev && status_out_dev && workspace_dev,
                  "db: null pointer argument");
  DbParams p{};
  p.maps = maps_dev; p.stride_n = stride_n; p.stride_h = stride_h; p.src_wh = src_wh_dev;
  p.N = N; p.H = H; p.W = W; p.Wd = db_words_per_row(W);
  p.R = resolve_max_runs(H, W, max_runs);
  p.cap = db_row_cap(H, W, p.R);
  p.E = 4 * p.R + 4;
  p.maxc = max_candidates;
  p.thresh = thresh; p.box_thresh = box_thresh; p.unclip_ratio = unclip_ratio;
  p.pad_resize = use_padding_resize ? 1 : 0;
  p.dilate = use_dilation ? 1 : 0;
  p.dtype = dtype;
  p.sem = semantics;
  p.score_box = score_mode == OCRPP_DB_SCORE_BOX ? 1 : 0;
  p.box_thresh_d = box_thresh_d;
  p.unclip_ratio_d = unclip_ratio_d;
  p.stairs = semantics == OCRPP_DB_SEMANTICS_CPP ? 1 : 0;   // cpp:222 fillPoly(..., lineType = 1) | py:193 LINE_8
  p.skip2 = 1;   // cpp:252; in the Python branch such contours fail its min-side test (a side of length 0) anyway
  const size_t need = carve(p, workspace_dev);
  if (need > workspace_bytes)
    return set_error(OCRPP_ERR_WORKSPACE_TOO_SMALL, "db: workspace needs %zu bytes, got %zu", need, workspace_bytes);
  p.boxes_out = boxes_out_dev; p.scores_out = scores_out_dev; p.counts_out = counts_out_dev;
  p.status_out = status_out_dev; p.boxes_f_out = boxes_f_out_dev; p.labels_dbg = labels_dbg_dev;
  cudaStream_t s = (cudaStream_t)stream;

  OCRPP_CUDA(cudaMemsetAsync(p.nruns, 0, sizeof(int32_t) * 5 * N, s));
  if (labels_dbg_dev) OCRPP_CUDA(cudaMemsetAsync(labels_dbg_dev, 0, sizeof(int32_t) * (size_t)N * H * W, s));
  // Large batches run as independent sub-batch pipelines on separate streams: the issue-bound map scan of one
  // sub-batch overlaps the latency-bound stage 2 and the ALU-bound geometry of another. (Not while per-phase
  // profiling is on: the event marks describe one whole-batch chain.)
  int nsplit = N >= 64 ? 2 : 1;
  const int forced = tuning(OCRPP_TUNE_DB_SPLIT);
  if (forced > 0) nsplit = forced > kDbMaxSplit ? kDbMaxSplit : forced;
  DbAux* aux = (nsplit > 1 && nsplit <= N && !profile_on()) ? db_aux() : nullptr;
  if (aux) {
    const bool prio = tuning(OCRPP_TUNE_DB_PRIO) != 1;
    std::lock_guard<std::mutex> lock(g_db_enqueue_mu);
    OCRPP_CUDA(cudaEventRecord(aux->fork, s));
    int rc = OCRPP_OK;
    for (int i = 0; i < nsplit; ++i) {
      cudaStream_t ss = aux->lo[i], sp = prio ? aux->hi[i] : aux->lo[i];
      if (cudaStreamWaitEvent(ss, aux->fork, 0) != cudaSuccess) rc = set_error(OCRPP_ERR_CUDA, "db: cudaStreamWaitEvent failed");
      const int lo = (int)((long long)N * i / nsplit), hi = (int)((long long)N * (i + 1) / nsplit);
      p.n0 = lo;
      if (rc == OCRPP_OK) rc = db_pipeline(p, hi - lo, dtype, ss, sp, aux->scanned[i], nullptr);
      // always join, also after a failed launch: the auxiliary streams may still touch the caller's buffers
      cudaEventRecord(aux->join[i], sp);
      cudaStreamWaitEvent(s, aux->join[i], 0);
      if (sp != ss) {
        cudaEventRecord(aux->scanned[i], ss);
        cudaStreamWaitEvent(s, aux->scanned[i], 0);
      }
    }
    p.n0 = 0;
    if (rc != OCRPP_OK) return rc;
  } else {
    ProfileScope prof(s);
    p.n0 = 0;
    const int rc = db_pipeline(p, N, dtype, s, s, nullptr, &prof);
    if (rc != OCRPP_OK) return rc;
  }
  return OCRPP_OK;
}
