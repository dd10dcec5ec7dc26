// Recogniser pre-processing of a whole arena of text-line crops in one launch: the step
// between ocrpp_crop_boxes and ONE batched recogniser forward. Replaces, per crop, the host-side
//   cv2.cvtColor + RecResizeImg/resize_norm_img + .unsqueeze(0).to(device)
// of R/deploy/pytorch/run_ocr.py:212-220 and R/pytocr/data/imaug/rec_img_aug.py:108-134; the per-pixel arithmetic
// is prep.cuh (bit exact against cv2, host-checked). HBM-bound gather: every output pixel reads four source bytes
// per channel that its neighbours share through L1.
#include "common.cuh"
#include "prep.cuh"

namespace ocrpp {
namespace {

__global__ void __launch_bounds__(256) rec_preprocess_kernel(const uint8_t* crops, const long long* offsets, const int32_t* dims,
                                                              int cin, int mode, int img_h, int img_w, float* out) {
  const int k = blockIdx.y;
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= img_h * img_w) return;
  const int y = idx / img_w, x = idx - y * img_w;
  const int h = dims[2 * k], w = dims[2 * k + 1];
  const int cout = mode == 0 ? 1 : cin;
  float* o = out + (size_t)k * cout * img_h * img_w + idx;
  int rw = 0;
  if (h > 0 && w > 0) rw = prep::resized_width(h, w, img_h, img_w);
  const uint8_t* crop = crops + offsets[k];
  for (int c = 0; c < cout; ++c)
    o[(size_t)c * img_h * img_w] = x < rw ? prep::normalise(prep::resized_value(crop, h, w, cin, mode, rw, img_h, x, y, c)) : 0.f;
}

}  // namespace
}  // namespace ocrpp

extern "C" int ocrpp_rec_preprocess(const uint8_t* crops_dev, const int64_t* offsets_dev, const int32_t* dims_dev, int K,
                                    int channels, int img_mode, int img_h, int img_w, float* out_dev, void* stream) {
  using namespace ocrpp;
  OCRPP_CHECK_ARG(K >= 0 && img_h > 0 && img_w > 0, "rec_preprocess: bad shape K=%d img_h=%d img_w=%d", K, img_h, img_w);
  OCRPP_CHECK_ARG(channels == 1 || channels == 3, "rec_preprocess: crops must have 1 or 3 channels (got %d)", channels);
  OCRPP_CHECK_ARG(img_mode == OCRPP_IMG_MODE_GRAY || img_mode == OCRPP_IMG_MODE_RGB || img_mode == OCRPP_IMG_MODE_BGR,
                  "rec_preprocess: bad img_mode %d", img_mode);
  OCRPP_CHECK_ARG(K <= 65535, "rec_preprocess: at most 65535 crops per call");
  if (K == 0) return OCRPP_OK;
  OCRPP_CHECK_ARG(crops_dev && offsets_dev && dims_dev && out_dev, "rec_preprocess: null pointer");
  dim3 grid((img_h * img_w + 255) / 256, K);
  rec_preprocess_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(crops_dev, reinterpret_cast<const long long*>(offsets_dev), dims_dev,
                                                                channels, img_mode, img_h, img_w, out_dev);
  OCRPP_LAUNCHED();
  return OCRPP_OK;
}
