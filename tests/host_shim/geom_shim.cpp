// TEST ONLY: compiles pytorchocr_b200/csrc/geometry.cuh with g++ (host side of OCRPP_HD) so the
// exact code the CUDA kernels run per candidate can be checked against the oracle on a CPU box.
#include <vector>

#include "../../pytorchocr_b200/csrc/geometry.cuh"
#include "../../pytorchocr_b200/csrc/prep.cuh"

using namespace ocrpp::geom;

extern "C" {

// points int32 [n,2] (any order) -> corners[8] (x0,y0,...), wh[2]; returns hull size
int shim_min_area_rect(const int* xy, int n, double* corners, double* wh) {
  std::vector<P2i> p(n), h(n + 2);
  for (int i = 0; i < n; ++i) p[i] = P2i{xy[2 * i], xy[2 * i + 1]};
  sort_points_yx(p.data(), n);
  int hn = hull_sorted(p.data(), n, h.data());
  Rect r;
  min_area_rect(h.data(), hn, &r);
  for (int k = 0; k < 4; ++k) {
    corners[2 * k] = r.cx[k];
    corners[2 * k + 1] = r.cy[k];
  }
  wh[0] = r.w;
  wh[1] = r.h;
  return hn;
}

// points -> corners in cv2.boxPoints order, then order_points_clockwise (float32) -> oxy[8]
void shim_generate_box(const int* xy, int n, float* box_cv, float* box_ordered) {
  std::vector<P2i> p(n), h(n + 2);
  for (int i = 0; i < n; ++i) p[i] = P2i{xy[2 * i], xy[2 * i + 1]};
  sort_points_yx(p.data(), n);
  int hn = hull_sorted(p.data(), n, h.data());
  Rect r;
  min_area_rect(h.data(), hn, &r);
  double bx[4], by[4];
  cv_box_order(r, bx, by);
  float cx[4], cy[4], ox[4], oy[4];
  for (int k = 0; k < 4; ++k) {
    cx[k] = (float)bx[k];
    cy[k] = (float)by[k];
    box_cv[2 * k] = cx[k];
    box_cv[2 * k + 1] = cy[k];
  }
  order_points_clockwise(cx, cy, ox, oy);
  for (int k = 0; k < 4; ++k) {
    box_ordered[2 * k] = ox[k];
    box_ordered[2 * k + 1] = oy[k];
  }
}

int shim_do_offset(const int* quad_xy, double delta, int* out_xy, int cap) {
  P2i q[4];
  for (int i = 0; i < 4; ++i) q[i] = P2i{quad_xy[2 * i], quad_xy[2 * i + 1]};
  std::vector<P2i> out(cap);
  int m = do_offset_quad(q, delta, out.data(), cap);
  for (int i = 0; i < m; ++i) {
    out_xy[2 * i] = out[i].x;
    out_xy[2 * i + 1] = out[i].y;
  }
  return m;
}

void shim_mini_box(const float* cxy, float* oxy) {
  float cx[4], cy[4], ox[4], oy[4];
  for (int i = 0; i < 4; ++i) { cx[i] = cxy[2 * i]; cy[i] = cxy[2 * i + 1]; }
  mini_box(cx, cy, ox, oy);
  for (int i = 0; i < 4; ++i) { oxy[2 * i] = ox[i]; oxy[2 * i + 1] = oy[i]; }
}

void shim_order_clockwise(const float* cxy, float* oxy) {
  float cx[4], cy[4], ox[4], oy[4];
  for (int i = 0; i < 4; ++i) { cx[i] = cxy[2 * i]; cy[i] = cxy[2 * i + 1]; }
  order_points_clockwise(cx, cy, ox, oy);
  for (int i = 0; i < 4; ++i) { oxy[2 * i] = ox[i]; oxy[2 * i + 1] = oy[i]; }
}

float shim_unclip_distance(const float* bxy, float ratio) {
  float bx[4], by[4];
  for (int i = 0; i < 4; ++i) { bx[i] = bxy[2 * i]; by[i] = bxy[2 * i + 1]; }
  return unclip_distance(bx, by, ratio);
}

void shim_db_rescale(float mx, float my, int W, int H, float sw, float sh, int use_padding_resize, float* out) {
  db_rescale(mx, my, W, H, sw, sh, use_padding_resize, &out[0], &out[1]);
}

float shim_roundf(float v) { return roundf_half_away(v); }
double shim_round_half_even(double v) { return round_half_even(v); }

// score_mode "box": rectangle + integer quad of box_score, then the fillPoly row extents
void shim_box_score_rows(const float* bxy, int W, int H, int* rect, int* L, int* R) {
  float bx[4], by[4];
  for (int i = 0; i < 4; ++i) { bx[i] = bxy[2 * i]; by[i] = bxy[2 * i + 1]; }
  int qx[4], qy[4];
  box_score_quad(bx, by, W, H, &rect[0], &rect[1], &rect[2], &rect[3], qx, qy);
  fill_quad_rows(qx, qy, rect[2], rect[3], L, R);
}

void shim_fill_quad_rows(const int* qxy, int w, int h, int* L, int* R) {
  int qx[4], qy[4];
  for (int i = 0; i < 4; ++i) { qx[i] = qxy[2 * i]; qy[i] = qxy[2 * i + 1]; }
  fill_quad_rows(qx, qy, w, h, L, R);
}

// recogniser pre-processing of one crop [h,w,cin] -> float32 [cout, img_h, img_w] (zero padded), returns resized_w
int shim_rec_preprocess(const uint8_t* crop, int h, int w, int cin, int mode, int img_h, int img_w, float* out) {
  using namespace ocrpp::prep;
  const int cout = mode == 0 ? 1 : cin;
  const int rw = resized_width(h, w, img_h, img_w);
  for (int c = 0; c < cout; ++c)
    for (int y = 0; y < img_h; ++y)
      for (int x = 0; x < img_w; ++x)
        out[((size_t)c * img_h + y) * img_w + x] = x < rw ? normalise(resized_value(crop, h, w, cin, mode, rw, img_h, x, y, c)) : 0.f;
  return rw;
}

}  // extern "C"
