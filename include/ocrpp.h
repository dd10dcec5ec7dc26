/* ocrpp.h - C ABI of the B200-native OCR post-processing library (libocrpp.so).
 *
 * This is the drop-in boundary for the post-processing hot path of DYJNG/PyTorchOCR
 * (R = the reference tree). Each entry point replaces one native/numpy boundary of the
 * reference and is what a reference-side binding (ctypes stub) would bind:
 *
 *   ocrpp_ctc_greedy        <- R/pytocr/postprocess/rec_postprocess.py:77-89 (+ decode :35-59)
 *   ocrpp_db_postprocess    <- R/pytocr/postprocess/db_postprocess_fast/src/db_postprocess.cpp:319-325
 *                              (`db_postprocess(pred, bitmap, box_thresh, unclip_ratio, src_w, src_h,
 *                              use_padding_resize)`, pybind11 :362-370) plus the thresholding that
 *                              precedes it, R/pytocr/postprocess/db_postprocess.py:43-46
 *   ocrpp_pse_postprocess   <- R/pytocr/postprocess/pse_postprocess_fast/pse.pyx:66 `pse(kernels, min_area)`
 *                              plus R/pytocr/postprocess/pse_postprocess.py:34-45 (upsample/sigmoid/threshold)
 *                              and :65-105 (generate_box)
 *   ocrpp_pan_postprocess   <- R/pytocr/postprocess/pan_postprocess_fast/pa.pyx:99 `pa(kernels, emb, min_area)`
 *                              plus R/pytocr/postprocess/pan_postprocess.py:36-51 and :73-113
 *
 * Conventions
 *   - plain C: pointers, sizes, scalars. No torch / C++ types. Every function returns an int status
 *     (OCRPP_OK == 0); ocrpp_last_error() gives the message of the calling thread's last failure.
 *     No exceptions cross the boundary.
 *   - batch level and stream ordered: all pointers named *_dev are DEVICE pointers on the current
 *     CUDA device; work is enqueued on `stream` (a cudaStream_t passed as void*) and the call
 *     returns without synchronising. Results are valid after the stream is synchronised.
 *   - caller-owned buffers: outputs have fixed capacity plus per-item counts; scratch comes from a
 *     caller-provided workspace whose size the matching *_workspace_bytes() call returns.
 *   - there is NO CPU path: without a CUDA device every compute entry point fails.
 */
#ifndef OCRPP_H_
#define OCRPP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OCRPP_ABI_VERSION 2

#if defined(__GNUC__)
#define OCRPP_API __attribute__((visibility("default")))
#else
#define OCRPP_API
#endif

/* status codes */
#define OCRPP_OK 0
#define OCRPP_ERR_INVALID_ARGUMENT 1
#define OCRPP_ERR_CUDA 2
#define OCRPP_ERR_WORKSPACE_TOO_SMALL 3

/* element types of input maps */
#define OCRPP_F32 0
#define OCRPP_F16 1

/* per-image status bits written to `status_out_dev` by the detection entry points */
#define OCRPP_IMG_RUN_OVERFLOW 1        /* more runs than `max_runs` - result of that image is invalid; retry with a larger max_runs */
#define OCRPP_IMG_CANDIDATES_TRUNCATED 2 /* more candidates than `max_candidates`; the reference keeps cv2's first 1000, we keep ours */
#define OCRPP_IMG_BOX_DEGENERATE 8       /* crop: empty bounding rectangle, corner outside the page or collinear corners; its crop is empty */
#define OCRPP_IMG_CROPS_TRUNCATED 16     /* crop: the arena is too small; offsets/dims are complete, crops past the capacity are not written */
#define OCRPP_IMG_VALUE_OUT_OF_RANGE 4   /* DB: a map value was outside [0,1] (or NaN/Inf): not a probability map, the
                                          * fixed-point score accumulation is not valid for it */

OCRPP_API int ocrpp_abi_version(void);
OCRPP_API const char* ocrpp_last_error(void);

/* Number of kernel launches issued by this library since load / since the last reset
 * (bench.py reports it as gpu_launches). */
OCRPP_API int64_t ocrpp_launch_count(void);
OCRPP_API void ocrpp_reset_launch_count(void);

/* Test / tuning hook (process-wide, not part of the reference-facing surface): selects between code paths that
 * produce identical results, so that the parity tests can exercise each of them and bench.py can sweep them.
 * value 0 always restores the default. */
#define OCRPP_TUNE_DB_PATH 0   /* 0 auto | 1 one-kernel-per-image stage, shared-memory tables | 2 same, global tables |
                                * 3 run-parallel multi-kernel chain (what large images take) */
#define OCRPP_TUNE_DB_SPLIT 1  /* sub-batch pipelines of one DB call (0 = chosen from the batch size) */
#define OCRPP_TUNE_DB_PRIO 2   /* 1: sub-batch pipelines without stream priorities */
#define OCRPP_TUNE_DB_SCAN 3   /* map scan in front of the one-kernel stage 2. 0 auto: db_scan4_kernel (bulk-copy fed ring,
                                * lane-contiguous chunks) when the row layout allows it, else db_scan_kernel |
                                * 1: db_scan_kernel (warp per row, ballots) | 2: the two-phase scan (db_scan3_kernel where the
                                * width has a compile-time specialisation, else db_scan2_kernel) | 3: db_scan2_kernel only */
#define OCRPP_TUNE_DB_SCAN4_STAGES 4  /* ring slots per CTA of db_scan4_kernel (0 = fill 100 KB) */
#define OCRPP_TUNE_DB_SCAN4_CTAS 5    /* CTAs per SM of db_scan4_kernel's persistent grid (0 = 2) */
#define OCRPP_TUNE_DB_IMG_SMEM_KB 6    /* dynamic shared memory of db_image_kernel in KB (0 = default) */
#define OCRPP_TUNE_COUNT 8
OCRPP_API int ocrpp_set_tuning(int key, int value);

/* Per-kernel device timing for bench.py's roofline. While enabled, every detection entry point
 * brackets each of its kernels with cudaEventRecord on the caller's stream (events only, no
 * synchronisation). After the stream is synchronised, ocrpp_profile_read() returns, for phase i,
 * the SUM of that phase's durations in milliseconds over all calls since the last
 * ocrpp_profile_reset() (at most 64 calls are kept) and the number of calls in *calls_out.
 * Returns the number of phases written. Phase 0 is always the kernel that streams the input maps. */
OCRPP_API void ocrpp_profile_enable(int on);
OCRPP_API void ocrpp_profile_reset(void);
OCRPP_API int ocrpp_profile_read(float* ms_out, int cap, int* calls_out);
OCRPP_API const char* ocrpp_profile_phase_name(int phase);

/* ---------------------------------------------------------------------------------------------
 * CTC greedy decode. probs element (t,b,c) is at probs_dev[t*stride_t + b*stride_b + c]
 * (element strides; class stride is 1). For every line b:
 *   idx_out_dev[b*T + 0..len)  kept class ids  (argmax per step, first maximum wins; blank 0
 *                              dropped; a step equal to the previous RAW argmax dropped)
 *   prob_out_dev[b*T + 0..len) their max-probabilities (float32)
 *   len_out_dev[b]             number kept
 *   conf_out_dev[b]            float32 mean of the kept probabilities, NaN when none kept
 *   raw_idx_out_dev            optional [B*T] raw argmax per step (may be NULL)
 * ------------------------------------------------------------------------------------------- */
OCRPP_API int ocrpp_ctc_greedy(const void* probs_dev, int dtype, int T, int B, int C,
                     int64_t stride_t, int64_t stride_b,
                     int32_t* idx_out_dev, float* prob_out_dev, int32_t* len_out_dev,
                     float* conf_out_dev, int32_t* raw_idx_out_dev, void* stream);

/* ---------------------------------------------------------------------------------------------
 * DB / DB++ box extraction for a batch of probability maps.
 *   maps_dev: image n, pixel (y,x) at maps_dev[n*stride_n + y*stride_h + x] (element strides).
 *   src_wh_dev: int32 [N,2] = (src_w, src_h) of every image (shape_list columns 1,0).
 *   max_candidates: output capacity per image (reference constant: 1000).
 *   max_runs: capacity of the per-image run table (foreground + background runs);
 *             0 selects the worst case H*(W+1).
 *   use_dilation: extract the boxes from the 2x2 dilation of the thresholded map (db_postprocess.py:52-55,
 *             cv2.dilate with a [[1,1],[1,1]] kernel) instead of the thresholded map itself; BoxScore still
 *             averages the probabilities under the (dilated) region. Costs a second pass over the maps.
 *   use_padding_resize: map the boxes back through the inverse "pad to square + resize" affine transform
 *             (db_postprocess.cpp:293-302) instead of the plain src/map scaling (:303-311).
 * outputs (device):
 *   boxes_out_dev   int16 [N,max_candidates,4,2]  TL,TR,BR,BL, rescaled to (src_w,src_h), roundf, clamped
 *   scores_out_dev  float [N,max_candidates]      BoxScore of every kept box (the reference wrapper
 *                                                 discards it and reports 1.0)
 *   counts_out_dev  int32 [N]
 *   status_out_dev  int32 [N]                     OCRPP_IMG_* bits
 *   boxes_f_out_dev float [N,max_candidates,4,2]  optional (NULL ok): pre-rounding rescaled corners
 *   labels_dbg_dev  int32 [N,H,W]                 optional (NULL ok): 8-connected foreground labels,
 *                                                 id = 1 + rank of the component's first raster pixel
 * ------------------------------------------------------------------------------------------- */
OCRPP_API size_t ocrpp_db_workspace_bytes(int N, int H, int W, int max_runs);
/* The same with the reference's branch selectable (R/pytocr/postprocess/db_postprocess.py:56-71):
 *   semantics OCRPP_DB_SEMANTICS_CPP    = `cpp_speedup: True`, db_postprocess_fast/src/db_postprocess.cpp:231-317 (what
 *                                          ocrpp_db_postprocess computes)
 *             OCRPP_DB_SEMANTICS_PYTHON = `cpp_speedup: False`, DBPostProcess.boxes_from_bitmap (db_postprocess.py:76-194):
 *                                          short side = min(w,h) instead of max, BoxScore over the LINE_8 polygon fill
 *                                          (no 4-connected "stair" pixels), unclip distance = area * ratio / length in
 *                                          float64 (shapely), np.round (half to even) instead of roundf, box_thresh and
 *                                          unclip_ratio used as doubles, `max_candidates` = the number of contours (in
 *                                          cv2's order) that are looked at, the scores are part of the result.
 *   score_mode OCRPP_DB_SCORE_POLY: mean over the filled contour; OCRPP_DB_SCORE_BOX (Python semantics only,
 *             db_postprocess.py:109-110): mean over cv2.fillPoly of the contour's mini box (float corners shifted by the
 *             clipped floor of their minimum and truncated to int, LINE_8). */
#define OCRPP_DB_SEMANTICS_CPP 0
#define OCRPP_DB_SEMANTICS_PYTHON 1
#define OCRPP_DB_SCORE_POLY 0
#define OCRPP_DB_SCORE_BOX 1
OCRPP_API int ocrpp_db_postprocess_ex(const void* maps_dev, int dtype, int N, int H, int W,
                         int64_t stride_n, int64_t stride_h, const int32_t* src_wh_dev,
                         float thresh, double box_thresh, double unclip_ratio,
                         int max_candidates, int max_runs, int use_dilation, int use_padding_resize,
                         int semantics, int score_mode,
                         int16_t* boxes_out_dev, float* scores_out_dev, int32_t* counts_out_dev,
                         int32_t* status_out_dev, float* boxes_f_out_dev, int32_t* labels_dbg_dev,
                         void* workspace_dev, size_t workspace_bytes, void* stream);
OCRPP_API int ocrpp_db_postprocess(const void* maps_dev, int dtype, int N, int H, int W,
                         int64_t stride_n, int64_t stride_h, const int32_t* src_wh_dev,
                         float thresh, float box_thresh, float unclip_ratio,
                         int max_candidates, int max_runs, int use_dilation, int use_padding_resize,
                         int16_t* boxes_out_dev, float* scores_out_dev, int32_t* counts_out_dev,
                         int32_t* status_out_dev, float* boxes_f_out_dev, int32_t* labels_dbg_dev,
                         void* workspace_dev, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * PSENet progressive scale expansion + box generation for a batch of head outputs.
 *   maps_dev: logits; image n, channel k, pixel (y,x) at
 *             maps_dev[n*stride_n + k*stride_c + y*stride_h + x] (element strides), k < K <= 8,
 *             channel 0 = text, channel K-1 = smallest kernel, spatial size h x w.
 *   upsample_in:  nearest up-sampling applied BEFORE the expansion (the reference's
 *             F.interpolate(scale_factor=4//scale), pse_postprocess.py:34-36); processing
 *             resolution = (h*upsample_in) x (w*upsample_in). 1 = maps already at processing res.
 *   upsample_out: nearest up-sampling of the label/score maps AFTER the expansion
 *             (cv2.resize(INTER_NEAREST), pse_postprocess.py:58-62); boxes are in that resolution
 *             before the division by the ratios.
 *   shape_dev: float64 [N,4] = shape_list rows (src_h, src_w, ratio_h, ratio_w).
 *   min_area_seed: seed components with fewer pixels are dropped (pse.pyx:21-23; the operator
 *             passes min_area / scale^2); min_area_box / box_thresh: generate_box filters
 *             (pse_postprocess.py:75,80) on the up-sampled label.
 *   max_boxes: output capacity per image (the reference has no limit; OCRPP_IMG_CANDIDATES_TRUNCATED
 *             is set when it is exceeded). max_runs: capacity of the per-image run tables
 *             (0 = worst case). arena_elems: capacity (uint32 elements) of the batch-wide queue
 *             arena (0 = worst case 4*N*H*W); OCRPP_IMG_RUN_OVERFLOW reports either overflow.
 * outputs (device): boxes int16 [N,max_boxes,4,2] in label order, corners ordered by
 *   order_points_clockwise (utility.py:21-29), np.round + clip; scores float [N,max_boxes] =
 *   mean sigmoid(text logit) per label; counts/status int32 [N]; boxes_f optional float
 *   [N,max_boxes,4,2] pre-rounding; labels_dbg optional int32 [N,H,W] (processing resolution) = the
 *   label map pse() returns, ids of cv2.connectedComponents(connectivity=4).
 * ------------------------------------------------------------------------------------------- */
OCRPP_API size_t ocrpp_pse_workspace_bytes(int N, int K, int h, int w, int upsample_in, int max_boxes,
                                 int max_runs, int64_t arena_elems);
OCRPP_API int ocrpp_pse_postprocess(const void* maps_dev, int dtype, int N, int K, int h, int w,
                          int64_t stride_n, int64_t stride_c, int64_t stride_h,
                          int upsample_in, int upsample_out, const double* shape_dev,
                          float thresh, float box_thresh, float min_area_seed, float min_area_box,
                          int max_boxes, int max_runs, int64_t arena_elems,
                          int16_t* boxes_out_dev, float* scores_out_dev, int32_t* counts_out_dev,
                          int32_t* status_out_dev, float* boxes_f_out_dev, int32_t* labels_dbg_dev,
                          void* workspace_dev, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * PAN / PAN++ pixel aggregation + box generation. maps_dev has 6 channels: 0 text logit,
 * 1 kernel logit, 2..5 the 4-d embedding (pan_postprocess.py:40-47). Kernel components smaller
 * than min_kernel_area are dropped (pa.pyx:33-37; the operator passes min_kernel_area / scale^2);
 * kernels whose area ratio to another kernel of the same text component exceeds 1024 are
 * "flagged" and only claim pixels whose embedding lies within distance 3 of the kernel's mean
 * embedding (pa.pyx:42-54,86-87). Everything else as ocrpp_pse_postprocess.
 * ------------------------------------------------------------------------------------------- */
OCRPP_API size_t ocrpp_pan_workspace_bytes(int N, int h, int w, int upsample_in, int max_boxes,
                                 int max_runs, int64_t arena_elems);
OCRPP_API int ocrpp_pan_postprocess(const void* maps_dev, int dtype, int N, int h, int w,
                          int64_t stride_n, int64_t stride_c, int64_t stride_h,
                          int upsample_in, int upsample_out, const double* shape_dev,
                          float thresh, float box_thresh, float min_kernel_area, float min_area_box,
                          int max_boxes, int max_runs, int64_t arena_elems,
                          int16_t* boxes_out_dev, float* scores_out_dev, int32_t* counts_out_dev,
                          int32_t* status_out_dev, float* boxes_f_out_dev, int32_t* labels_dbg_dev,
                          void* workspace_dev, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Text-line crops of detected boxes for a batch of pages: the step between the detection operators and
 * the recogniser. Replaces sort_boxes (R/pytocr/utils/utility.py:32-50), get_part_img (:53-78:
 * bounding-rectangle crop, cv2.getPerspectiveTransform, cv2.warpPerspective with INTER_LINEAR and
 * BORDER_REPLICATE) and the rot90 rule of its caller (R/deploy/pytorch/run_ocr.py:188-191), with cv2's
 * 8-bit arithmetic reproduced bit for bit (oracle/crop_oracle.py).
 *   img_dev:   uint8 pages, page n pixel (y,x) channel c at img_dev[n*stride_n + y*stride_row + x*C + c]
 *              (byte strides; C = 1, 3 or 4; H, W <= 32767).
 *   boxes_dev: int16 [N,max_boxes,4,2] corners (x,y) - the layout ocrpp_db/pse/pan_postprocess write -
 *              and counts_dev int32 [N] (NULL: max_boxes boxes on every page); max_boxes <= 8192.
 *   sort_boxes:  reorder every page's boxes as the reference's sort_boxes does before cropping.
 *   rotate_tall: store crops with rows >= 1.5 * cols rotated counter-clockwise (np.rot90(crop, 1)).
 * outputs (device), entry e = n*max_boxes + r for the r-th box of page n in output order:
 *   crops_out_dev    uint8 arena of crops_capacity_bytes; crop e is the dense [rows,cols,C] array at byte
 *                    offsets_out_dev[e] (entries are packed back to back: offsets[e+1]-offsets[e] = its size)
 *   offsets_out_dev  int64 [N*max_boxes+1]   (the last entry = bytes needed for all crops)
 *   dims_out_dev     int32 [N*max_boxes,2]   rows, cols as stored (0,0 past the count / degenerate box)
 *   order_out_dev    int32 [N*max_boxes]     index of the box in the page's input list (-1 past the count)
 *   status_out_dev   int32 [N]               OCRPP_IMG_BOX_DEGENERATE / OCRPP_IMG_CROPS_TRUNCATED
 * ------------------------------------------------------------------------------------------- */
OCRPP_API size_t ocrpp_crop_workspace_bytes(int N, int max_boxes);
OCRPP_API int ocrpp_crop_boxes(const uint8_t* img_dev, int N, int H, int W, int C, int64_t stride_n, int64_t stride_row,
                     const int16_t* boxes_dev, const int32_t* counts_dev, int max_boxes, int sort_boxes,
                     int rotate_tall, uint8_t* crops_out_dev, size_t crops_capacity_bytes,
                     int64_t* offsets_out_dev, int32_t* dims_out_dev, int32_t* order_out_dev,
                     int32_t* status_out_dev, void* workspace_dev, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Recogniser pre-processing of the crop arena ocrpp_crop_boxes wrote: one float32 batch [K, Cout, img_h, img_w] for
 * ONE recogniser forward. Replaces, per crop, cv2.cvtColor + RecResizeImg (resize_norm_img: cv2.resize to
 * (min(img_w, ceil(img_h * w / h)), img_h), / 255, (x - 0.5) / 0.5, zero padding to img_w) + the upload of
 * R/deploy/pytorch/run_ocr.py:212-220 and R/pytocr/data/imaug/rec_img_aug.py:108-134, with cv2's 8-bit arithmetic
 * reproduced bit for bit.
 *   crops_dev / offsets_dev / dims_dev: the arena, byte offsets and (rows, cols) of K crops with `channels` (1 or 3,
 *   BGR) interleaved channels; a crop with rows == 0 gives an all-zero image.
 *   img_mode: OCRPP_IMG_MODE_GRAY (Cout = 1; BGR2GRAY for 3-channel crops), _RGB (channel order reversed), _BGR.
 * ------------------------------------------------------------------------------------------- */
#define OCRPP_IMG_MODE_GRAY 0
#define OCRPP_IMG_MODE_RGB 1
#define OCRPP_IMG_MODE_BGR 2
OCRPP_API int ocrpp_rec_preprocess(const uint8_t* crops_dev, const int64_t* offsets_dev, const int32_t* dims_dev, int K,
                         int channels, int img_mode, int img_h, int img_w, float* out_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OCRPP_H_ */
