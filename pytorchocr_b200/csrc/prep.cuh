// Recogniser pre-processing of one text-line crop, per output pixel: what the reference's run_ocr.py does on the host
// between get_part_img and the recogniser forward (R/deploy/pytorch/run_ocr.py:212-220):
//   cv2.cvtColor(part_img, COLOR_BGR2GRAY | COLOR_BGR2RGB)   (rec_img_mode)
//   RecResizeImg -> resize_norm_img (R/pytocr/data/imaug/rec_img_aug.py:108-134): cv2.resize(img, (resized_w, imgH))
//   with resized_w = min(imgW, ceil(imgH * w / h)), astype(float32) / 255, (x - 0.5) / 0.5, zero padding to imgW.
// Third-party arithmetic restated from OpenCV's 8-bit code paths (bit exact against cv2 4.13, tests/test_geometry_host.py):
//   BGR2GRAY: (B * 3735 + G * 19235 + R * 9798 + 2^14) >> 15   (the 15-bit coefficients of OpenCV 4's RGB2Gray<uchar>)
//   resize INTER_LINEAR, 8U: source coordinate (float)((d + 0.5) * scale - 0.5) with scale = src / dst in double,
//     floor, 11-bit coefficients cvRound(f * 2048) (the pair sums to 2048 only up to rounding: both are rounded),
//     border taps replicated (x: coefficient reset to one tap; y: row index clamped), horizontal pass in int32,
//     vertical pass ((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2 >> 2;
//     an exact 2x2 decimation takes OpenCV's INTER_AREA shortcut (sum of the four + 2 >> 2);
//     equal sizes copy.
// OCRPP_HD: the same source is compiled for the host by tests/host_shim.
#pragma once
#include "geometry.cuh"

namespace ocrpp {
namespace prep {

OCRPP_HD int cv_round_half_even(double v) { return (int)geom::round_half_even(v); }

// one channel value of pixel (x, y) of the crop after the colour conversion; crop = dense [h, w, cin] uint8
// mode 0: GRAY (cin 3: BGR2GRAY; cin 1: as is), 1: RGB (channel order reversed), 2: BGR (as is)
OCRPP_HD int src_value(const uint8_t* crop, int w, int cin, int mode, int x, int y, int c) {
  const uint8_t* px = crop + ((size_t)y * w + x) * cin;
  if (mode == 0) {
    if (cin == 1) return px[0];
    return (px[0] * 3735 + px[1] * 19235 + px[2] * 9798 + (1 << 14)) >> 15;
  }
  return mode == 1 ? px[cin - 1 - c] : px[c];
}

// resize_norm_img's resized width (rec_img_aug.py:117-121), float64 like the Python code
OCRPP_HD int resized_width(int h, int w, int img_h, int img_w) {
  const double ratio = (double)w / (double)h;
  const double cw = ceil(geom::dmul((double)img_h, ratio));
  return cw > (double)img_w ? img_w : (int)cw;
}

struct Tap {
  int s0, s1;      // source indices
  int a0, a1;      // 11-bit coefficients
};

OCRPP_HD Tap linear_tap_x(int d, int dst, int src) {
  const double scale = (double)src / (double)dst;
  float f = (float)(geom::dmul((double)d + 0.5, scale) - 0.5);
  int s = (int)floorf(f);
  f -= (float)s;
  if (s < 0) { f = 0.f; s = 0; }
  if (s + 1 >= src) {
    if (s >= src - 1) { f = 0.f; s = src - 1; }
    // D[dx] = S[sx] * ONE for every dx from the first such column on
    Tap t;
    t.s0 = s; t.s1 = s; t.a0 = 2048; t.a1 = 0;
    return t;
  }
  Tap t;
  t.s0 = s; t.s1 = s + 1;
  t.a0 = cv_round_half_even((double)((1.f - f) * 2048.f));
  t.a1 = cv_round_half_even((double)(f * 2048.f));
  return t;
}

OCRPP_HD Tap linear_tap_y(int d, int dst, int src) {
  const double scale = (double)src / (double)dst;
  float f = (float)(geom::dmul((double)d + 0.5, scale) - 0.5);
  const int s = (int)floorf(f);
  f -= (float)s;
  Tap t;
  t.s0 = s < 0 ? 0 : (s > src - 1 ? src - 1 : s);
  t.s1 = s + 1 < 0 ? 0 : (s + 1 > src - 1 ? src - 1 : s + 1);
  t.a0 = cv_round_half_even((double)((1.f - f) * 2048.f));
  t.a1 = cv_round_half_even((double)(f * 2048.f));
  return t;
}

// uint8 value of channel c of pixel (x, y) of cv2.resize(convert(crop), (dw, dh))
OCRPP_HD int resized_value(const uint8_t* crop, int h, int w, int cin, int mode, int dw, int dh, int x, int y, int c) {
  if (dw == w && dh == h) return src_value(crop, w, cin, mode, x, y, c);
  if (w == 2 * dw && h == 2 * dh) {
    return (src_value(crop, w, cin, mode, 2 * x, 2 * y, c) + src_value(crop, w, cin, mode, 2 * x + 1, 2 * y, c) +
            src_value(crop, w, cin, mode, 2 * x, 2 * y + 1, c) + src_value(crop, w, cin, mode, 2 * x + 1, 2 * y + 1, c) + 2) >> 2;
  }
  const Tap tx = linear_tap_x(x, dw, w), ty = linear_tap_y(y, dh, h);
  const int r0 = src_value(crop, w, cin, mode, tx.s0, ty.s0, c) * tx.a0 + src_value(crop, w, cin, mode, tx.s1, ty.s0, c) * tx.a1;
  const int r1 = src_value(crop, w, cin, mode, tx.s0, ty.s1, c) * tx.a0 + src_value(crop, w, cin, mode, tx.s1, ty.s1, c) * tx.a1;
  return (((ty.a0 * (r0 >> 4)) >> 16) + ((ty.a1 * (r1 >> 4)) >> 16) + 2) >> 2;
}

// float32 network input: v / 255, - 0.5, / 0.5 in float32 as numpy evaluates them
OCRPP_HD float normalise(int v) {
  return geom::fdiv(geom::fsub(geom::fdiv((float)v, 255.f), 0.5f), 0.5f);
}

}  // namespace prep
}  // namespace ocrpp
