/* TEST INFRASTRUCTURE ONLY (oracle). Never linked, loaded or called by the product path
 * (pytorchocr_b200/): only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs
 * may use it, and there only as the checker / the thing timed as "CPU".
 *
 * Plain-C restatement of the integer parts of the reference's post-processing hot path.
 * Every function cites the reference lines it follows (R = /root/reference).
 * Parity pinning: the reference ships no tests/golden vectors for this path,
 * so these functions are pinned against the reference's own compiled Cython modules
 * (oracle/_ref/pse*.so, pa*.so, built unmodified by oracle/build_ref.py) in
 * tests/test_oracle_vs_reference.py and against fixtures in tests/golden/ that were produced
 * by running the reference classes (tests/golden/make_golden.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * 4-connected component labelling with ids in raster order of each component's first pixel.
 * Third-party behaviour being restated: cv2.connectedComponents(img, connectivity=4) as
 * called at R/pytocr/postprocess/pse_postprocess_fast/pse.pyx:68 and
 * R/pytocr/postprocess/pan_postprocess_fast/pa.pyx:101-102 (OpenCV numbers 4-connected
 * components in raster order of their first pixel; checked against cv2 4.13 in the tests).
 * Returns label_num (= number of components + 1, as cv2 does).
 * ------------------------------------------------------------------------------------------ */
static int32_t uf_find(int32_t* p, int32_t x) {
  while (p[x] != x) {
    p[x] = p[p[x]];
    x = p[x];
  }
  return x;
}

int oracle_ccl4(const uint8_t* img, int H, int W, int32_t* label) {
  const int64_t n = (int64_t)H * W;
  int32_t* parent = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      const int32_t i = y * W + x;
      if (!img[i]) {
        parent[i] = -1;
        continue;
      }
      parent[i] = i;
      if (x > 0 && img[i - 1]) {
        parent[i] = uf_find(parent, i - 1);
      }
      if (y > 0 && img[i - W]) {
        int32_t a = uf_find(parent, i - W), b = uf_find(parent, i);
        if (a < b) parent[b] = a; else if (b < a) parent[a] = b;
      }
    }
  /* roots are the minimum raster index of their component => numbering roots in raster order
   * numbers components by first pixel. */
  int32_t next = 1;
  for (int64_t i = 0; i < n; ++i) {
    if (parent[i] < 0) { label[i] = 0; continue; }
    int32_t r = uf_find(parent, (int32_t)i);
    if (r == i) label[i] = next++;
    else label[i] = label[r];
  }
  free(parent);
  return next;
}

/* simple FIFO of (row, col) int16 pairs == std::queue<pair<int16,int16>> (pse.pyx:25-28) */
typedef struct { int16_t* d; size_t cap, head, tail; } fifo_t;
static void fifo_init(fifo_t* q, size_t cap) { q->d = (int16_t*)malloc(cap * 4 + 4); q->cap = cap; q->head = q->tail = 0; }
static void fifo_push(fifo_t* q, int16_t a, int16_t b) {
  if (q->tail == q->cap) {
    /* compact or grow */
    size_t live = q->tail - q->head;
    if (q->head > 0) { memmove(q->d, q->d + 2 * q->head, live * 4); q->head = 0; q->tail = live; }
    if (q->tail == q->cap) { q->cap *= 2; q->d = (int16_t*)realloc(q->d, q->cap * 4 + 4); }
  }
  q->d[2 * q->tail] = a; q->d[2 * q->tail + 1] = b; q->tail++;
}
static int fifo_empty(const fifo_t* q) { return q->head == q->tail; }

/* ------------------------------------------------------------------------------------------
 * Progressive scale expansion. Restates `pse()` + `_pse()`:
 *   R/pytocr/postprocess/pse_postprocess_fast/pse.pyx:66-69  (wrapper: CCL of kernels[-1], passes
 *       kernels[:-1] but kernel_num = K, so the level loop K-1..0 reads the ORIGINAL array:
 *       index K-1 is the smallest kernel again)
 *   pse.pyx:21-23  drop labels whose area < min_area (float compare)
 *   pse.pyx:33-37  seed queue in raster order
 *   pse.pyx:41-62  per level: FIFO BFS, 4 neighbours in order (row-1),(row+1),(col-1),(col+1);
 *                  a popped pixel that claimed nothing goes to the next level's queue
 * kernels: uint8 [K,H,W]; pred_out: int32 [H,W].
 * ------------------------------------------------------------------------------------------ */
void oracle_pse(const uint8_t* kernels, int K, int H, int W, float min_area, int32_t* pred_out) {
  const int64_t n = (int64_t)H * W;
  int32_t* label = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
  const int label_num = oracle_ccl4(kernels + (int64_t)(K - 1) * n, H, W, label);
  int64_t* area = (int64_t*)calloc((size_t)label_num, sizeof(int64_t));
  for (int64_t i = 0; i < n; ++i) area[label[i]]++;
  for (int64_t i = 0; i < n; ++i)
    if (label[i] > 0 && (float)area[label[i]] < min_area) label[i] = 0; /* pse.pyx:21-23 */
  memset(pred_out, 0, sizeof(int32_t) * (size_t)n);

  fifo_t que, nxt;
  fifo_init(&que, 1024); fifo_init(&nxt, 1024);
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x)
      if (label[(int64_t)y * W + x] > 0) {
        fifo_push(&que, (int16_t)y, (int16_t)x);
        pred_out[(int64_t)y * W + x] = label[(int64_t)y * W + x];
      }
  static const int dx[4] = {-1, 1, 0, 0};  /* first index is the ROW (pse.pyx:29-30,49-50) */
  static const int dy[4] = {0, 0, -1, 1};
  for (int k = K - 1; k >= 0; --k) {
    const uint8_t* ker = kernels + (int64_t)k * n;
    while (!fifo_empty(&que)) {
      const int cy = que.d[2 * que.head], cx = que.d[2 * que.head + 1];
      que.head++;
      const int32_t cur_label = pred_out[(int64_t)cy * W + cx];
      int is_edge = 1;
      for (int j = 0; j < 4; ++j) {
        const int ty = cy + dx[j], tx = cx + dy[j];
        if (ty < 0 || ty >= H || tx < 0 || tx >= W) continue;
        const int64_t t = (int64_t)ty * W + tx;
        if (ker[t] == 0 || pred_out[t] > 0) continue;
        fifo_push(&que, (int16_t)ty, (int16_t)tx);
        pred_out[t] = cur_label;
        is_edge = 0;
      }
      if (is_edge) fifo_push(&nxt, (int16_t)cy, (int16_t)cx);
    }
    fifo_t t = que; que = nxt; nxt = t;
    nxt.head = nxt.tail = 0;
  }
  free(que.d); free(nxt.d); free(area); free(label);
}

/* ------------------------------------------------------------------------------------------
 * Pixel aggregation, BFS part. Restates `_pa()` after its label pre-pass:
 *   R/pytocr/postprocess/pan_postprocess_fast/pa.pyx:56-68  seeds = surviving kernel labels, raster order
 *   pa.pyx:72-95  ONE level (kernel_num-2 .. 0 with kernel_num = 2) over kernels[0] (text mask);
 *                 a claim by a flagged label is blocked when ||emb[:,p] - mean_emb[label]||_2 > 3
 * The pre-pass (areas, min-area drop, first pixels, ratio flags, float32 np.mean embeddings;
 * pa.pyx:28-54) is done by the caller in numpy (oracle/pan_oracle.py) so that the float32
 * pairwise mean is numpy's own; `label` arrives already filtered.
 * text: uint8 [H,W]; emb: float32 [4,H,W]; label: int32 [H,W]; flag: int32 [label_num];
 * mean_emb: float32 [label_num,4]; pred_out: int32 [H,W].
 * ------------------------------------------------------------------------------------------ */
void oracle_pa_expand(const uint8_t* text, const float* emb, const int32_t* label, int H, int W,
                      const int32_t* flag, const float* mean_emb, int32_t* pred_out) {
  const int64_t n = (int64_t)H * W;
  memset(pred_out, 0, sizeof(int32_t) * (size_t)n);
  fifo_t que;
  fifo_init(&que, 1024);
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x)
      if (label[(int64_t)y * W + x] > 0) {
        fifo_push(&que, (int16_t)y, (int16_t)x);
        pred_out[(int64_t)y * W + x] = label[(int64_t)y * W + x];
      }
  static const int dx[4] = {-1, 1, 0, 0};
  static const int dy[4] = {0, 0, -1, 1};
  while (!fifo_empty(&que)) {
    const int cy = que.d[2 * que.head], cx = que.d[2 * que.head + 1];
    que.head++;
    const int32_t cur_label = pred_out[(int64_t)cy * W + cx];
    for (int j = 0; j < 4; ++j) {
      const int ty = cy + dx[j], tx = cx + dy[j];
      if (ty < 0 || ty >= H || tx < 0 || tx >= W) continue;
      const int64_t t = (int64_t)ty * W + tx;
      if (text[t] == 0 || pred_out[t] > 0) continue;
      if (flag[cur_label] == 1) {
        float s = 0.f;
        for (int c = 0; c < 4; ++c) {
          const float d = emb[(int64_t)c * n + t] - mean_emb[cur_label * 4 + c];
          s += d * d;
        }
        if (sqrtf(s) > 3.f) continue; /* pa.pyx:86-87 */
      }
      fifo_push(&que, (int16_t)ty, (int16_t)tx);
      pred_out[t] = cur_label;
    }
  }
  free(que.d);
}

/* ------------------------------------------------------------------------------------------
 * CTC greedy decode, index part. Restates
 *   R/pytocr/postprocess/rec_postprocess.py:82-84 (argmax / max over classes, first maximum wins)
 *   rec_postprocess.py:40-50 (drop blank 0, drop repeats of the previous RAW index)
 * probs: float32, element (t,b,c) at probs[t*stride_t + b*stride_b + c]. Outputs per line b:
 * idx_out[b*T + 0..len) kept class ids, prob_out[b*T + ..] their max-probs, len_out[b].
 * The confidence mean (np.mean of float32 list, :58) is taken by the caller in numpy.
 * ------------------------------------------------------------------------------------------ */
void oracle_ctc_greedy(const float* probs, int T, int B, int C, int64_t stride_t, int64_t stride_b,
                       int32_t* idx_out, float* prob_out, int32_t* len_out,
                       int32_t* raw_idx_out /* optional [B,T] */) {
  for (int b = 0; b < B; ++b) {
    int n = 0;
    int32_t prev = -1;
    for (int t = 0; t < T; ++t) {
      const float* row = probs + t * stride_t + b * stride_b;
      int32_t best = 0;
      float bv = row[0];
      /* numpy's argmax / max (rec_postprocess.py:83-84): NaN is the maximum and the FIRST NaN wins */
      for (int c = 1; c < C && bv == bv; ++c)
        if (row[c] > bv || row[c] != row[c]) { bv = row[c]; best = c; }
      if (raw_idx_out) raw_idx_out[(int64_t)b * T + t] = best;
      if (best != 0 && !(t > 0 && prev == best)) {
        idx_out[(int64_t)b * T + n] = best;
        prob_out[(int64_t)b * T + n] = bv;
        ++n;
      }
      prev = best;
    }
    len_out[b] = n;
  }
}
