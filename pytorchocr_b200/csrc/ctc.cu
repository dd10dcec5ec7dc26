// CTC greedy decode on sm_100a.
//
// Replaces R/pytocr/postprocess/rec_postprocess.py:77-89 (D2H of [T,B,C], numpy argmax/max over
// classes) and BaseRecLabelDecode.decode :35-59 (pure-Python B x T loop: drop blank, drop repeats
// of the previous raw index, mean confidence).
//
// Kernel 1 (HBM-bound, ~all of the time): one warp per (t,b) row of C probabilities. The row is
//   streamed once with 128-bit L1-bypassing loads (rows of C=6623 floats are only 4-byte aligned,
//   so up to 3 head and 3 tail elements are peeled), each lane keeps (max, first index), then a
//   shuffle reduction with "smaller index wins ties" reproduces numpy's first-maximum argmax.
//   Algorithmic bytes per text line: T*C*sizeof(elem), read exactly once.
// Kernel 2 (negligible): one warp per text line collapses the T raw indices in place with
//   ballot/popc compaction and computes the confidence with numpy's float32 pairwise-sum order.
#include "common.cuh"

namespace ocrpp {
namespace {

constexpr int kCtcThreads = 256;
constexpr int kCtcUnroll = 8;

struct Best {
  float v;
  int i;
};

__device__ __forceinline__ void upd(Best& b, float& acc, float v, int i) {
  acc += v;   // NaN canary
  if (v > b.v) {
    b.v = v;
    b.i = i;
  }
}

template <typename T>
struct VecOps;

template <>
struct VecOps<float> {
  static constexpr int kElems = 4;
  __device__ static __forceinline__ void consume(Best& b, float& acc, const uint4& u, int base) {
    const float x = __uint_as_float(u.x), y = __uint_as_float(u.y), z = __uint_as_float(u.z),
                w = __uint_as_float(u.w);
    acc += (x + y) + (z + w);   // NaN canary (fmaxf drops NaNs); see the end of ctc_argmax_kernel
    const float m = fmaxf(fmaxf(x, y), fmaxf(z, w));
    if (m > b.v) {  // rare after the first few vectors
      b.v = m;
      b.i = base + (x == m ? 0 : (y == m ? 1 : (z == m ? 2 : 3)));
    }
  }
};

template <>
struct VecOps<__half> {
  static constexpr int kElems = 8;
  __device__ static __forceinline__ void consume(Best& b, float& acc, const uint4& u, int base) {
    float f[8] = {h2f_lo(u.x), h2f_hi(u.x), h2f_lo(u.y), h2f_hi(u.y),
                  h2f_lo(u.z), h2f_hi(u.z), h2f_lo(u.w), h2f_hi(u.w)};
    float m = f[0];
    acc += ((f[0] + f[1]) + (f[2] + f[3])) + ((f[4] + f[5]) + (f[6] + f[7]));   // NaN canary
#pragma unroll
    for (int k = 1; k < 8; ++k) m = fmaxf(m, f[k]);
    if (m > b.v) {
      int k = 0;
#pragma unroll
      for (int q = 7; q >= 0; --q)
        if (f[q] == m) k = q;
      b.v = m;
      b.i = base + k;
    }
  }
};

template <typename T>
__global__ void __launch_bounds__(kCtcThreads)
ctc_argmax_kernel(const T* __restrict__ probs, int Tn, int B, int C, long long stride_t,
                  long long stride_b, int32_t* __restrict__ idx_out, float* __restrict__ prob_out,
                  int32_t* __restrict__ raw_out) {
  const long long row_id = ((long long)blockIdx.x * kCtcThreads + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row_id >= (long long)Tn * B) return;
  const int t = (int)(row_id / B), b = (int)(row_id % B);
  const T* row = probs + t * stride_t + b * stride_b;
  constexpr int EPV = VecOps<T>::kElems;

  int head = (int)(((16u - (unsigned)((uintptr_t)row & 15u)) & 15u) / sizeof(T));
  if (head > C) head = C;
  Best best{-INFINITY, 0x7fffffff};
  float acc = 0.f;   // becomes NaN when the lane has seen a NaN (or +inf and -inf): the row is then re-read below
  if (lane < head) upd(best, acc, load_scalar<T>(row + lane), lane);

  const uint4* vp = reinterpret_cast<const uint4*>(row + head);
  const int nvec = (C - head) / EPV;
  const uint32_t ninf = sizeof(T) == 4 ? 0xff800000u : 0xfc00fc00u;
  int j0 = 0;
  // full batches: kCtcUnroll independent 128-bit loads in flight per lane before any is consumed
  for (; j0 + 32 * kCtcUnroll <= nvec; j0 += 32 * kCtcUnroll) {
    uint4 v[kCtcUnroll];
#pragma unroll
    for (int u = 0; u < kCtcUnroll; ++u) v[u] = ldg_stream_u4(vp + j0 + u * 32 + lane);
#pragma unroll
    for (int u = 0; u < kCtcUnroll; ++u)
      VecOps<T>::consume(best, acc, v[u], head + (j0 + u * 32 + lane) * EPV);
  }
  {  // ragged last batch
    uint4 v[kCtcUnroll];
#pragma unroll
    for (int u = 0; u < kCtcUnroll; ++u) {
      const int j = j0 + u * 32 + lane;
      v[u] = j < nvec ? ldg_stream_u4(vp + j) : make_uint4(ninf, ninf, ninf, ninf);
    }
#pragma unroll
    for (int u = 0; u < kCtcUnroll; ++u)
      VecOps<T>::consume(best, acc, v[u], head + (j0 + u * 32 + lane) * EPV);
  }
  for (int i = head + nvec * EPV + lane; i < C; i += 32) upd(best, acc, load_scalar<T>(row + i), i);

#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best.v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, best.i, o);
    if (ov > best.v || (ov == best.v && oi < best.i)) {
      best.v = ov;
      best.i = oi;
    }
  }
  // numpy's argmax / max (rec_postprocess.py:83-84) treat NaN as the maximum: the FIRST NaN wins and the
  // probability is NaN. Rare: only rows whose canary tripped are read again.
  if (__any_sync(0xffffffffu, acc != acc)) {
    int first_nan = 0x7fffffff;
    for (int i = lane; i < C; i += 32) {
      const float v = load_scalar<T>(row + i);
      if (v != v && i < first_nan) first_nan = i;
    }
    first_nan = __reduce_min_sync(0xffffffffu, first_nan);
    if (first_nan != 0x7fffffff) {
      best.i = first_nan;
      best.v = __int_as_float(0x7fc00000);
    }
  }
  if (lane == 0) {
    const int bi = best.i == 0x7fffffff ? 0 : best.i;  // all -inf row: numpy says 0
    const long long o = (long long)b * Tn + t;
    idx_out[o] = bi;
    prob_out[o] = best.v;
    if (raw_out) raw_out[o] = bi;
  }
}

// numpy's pairwise float32 summation (numpy/_core/src/umath/loops_utils.h.src, pairwise_sum), the
// order np.mean uses on the reference's conf_list (rec_postprocess.py:58).
__device__ float np_pairwise_sum(const float* a, int n) {
  if (n < 8) {
    float res = -0.0f;
    for (int i = 0; i < n; ++i) res += a[i];
    return res;
  }
  if (n <= 128) {
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i;
    for (i = 8; i < n - (n % 8); i += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] += a[i + j];
    }
    float res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
  }
  int n2 = n / 2;
  n2 -= n2 % 8;
  return np_pairwise_sum(a, n2) + np_pairwise_sum(a + n2, n - n2);
}

__global__ void __launch_bounds__(128)
ctc_collapse_kernel(int Tn, int B, int32_t* __restrict__ idx, float* __restrict__ prob,
                    int32_t* __restrict__ len_out, float* __restrict__ conf_out) {
  const int b = (blockIdx.x * 128 + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  int32_t* li = idx + (long long)b * Tn;
  float* lp = prob + (long long)b * Tn;
  int n = 0;
  int prev_last = -1;
  for (int t0 = 0; t0 < Tn; t0 += 32) {
    const int t = t0 + lane;
    const int cur = t < Tn ? li[t] : 0;
    const float p = t < Tn ? lp[t] : 0.f;
    int prev = __shfl_up_sync(0xffffffffu, cur, 1);
    if (lane == 0) prev = prev_last;
    prev_last = __shfl_sync(0xffffffffu, cur, 31);
    const bool keep = t < Tn && cur != 0 && !(t > 0 && prev == cur);
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    __syncwarp();
    if (keep) {
      const int pos = n + __popc(m & ((1u << lane) - 1u));
      li[pos] = cur;  // pos <= t: compaction never overtakes unread input (see header comment)
      lp[pos] = p;
    }
    n += __popc(m);
  }
  __syncwarp();
  for (int t = n + lane; t < Tn; t += 32) li[t] = 0;  // zero tail: the host turns rows into strings
  if (lane == 0) {
    len_out[b] = n;
    conf_out[b] = n > 0 ? __fdiv_rn(np_pairwise_sum(lp, n), (float)n) : __int_as_float(0x7fc00000);
  }
}

}  // namespace
}  // namespace ocrpp

extern "C" int ocrpp_ctc_greedy(const void* probs_dev, int dtype, int T, int B, int C,
                                int64_t stride_t, int64_t stride_b, int32_t* idx_out_dev,
                                float* prob_out_dev, int32_t* len_out_dev, float* conf_out_dev,
                                int32_t* raw_idx_out_dev, void* stream) {
  using namespace ocrpp;
  OCRPP_CHECK_ARG(dtype == OCRPP_F32 || dtype == OCRPP_F16, "ctc: dtype must be OCRPP_F32 or OCRPP_F16");
  OCRPP_CHECK_ARG(T >= 0 && B >= 0 && C >= 1, "ctc: bad shape T=%d B=%d C=%d", T, B, C);
  OCRPP_CHECK_ARG(idx_out_dev && prob_out_dev && len_out_dev && conf_out_dev, "ctc: null output");
  if (T == 0 || B == 0) {
    if (B > 0) {
      // zero steps: every line is empty
      cudaStream_t s0 = (cudaStream_t)stream;
      OCRPP_CUDA(cudaMemsetAsync(len_out_dev, 0, sizeof(int32_t) * B, s0));
      OCRPP_CUDA(cudaMemsetAsync(conf_out_dev, 0xff, sizeof(float) * B, s0));  // NaN pattern
    }
    return OCRPP_OK;
  }
  OCRPP_CHECK_ARG(probs_dev != nullptr, "ctc: null input");
  cudaStream_t s = (cudaStream_t)stream;
  const long long rows = (long long)T * B;
  const long long blocks = (rows * 32 + kCtcThreads - 1) / kCtcThreads;
  OCRPP_CHECK_ARG(blocks < (1ll << 31), "ctc: too many rows");
  ProfileScope prof(s);
  if (dtype == OCRPP_F32)
    ctc_argmax_kernel<float><<<(unsigned)blocks, kCtcThreads, 0, s>>>(
        (const float*)probs_dev, T, B, C, stride_t, stride_b, idx_out_dev, prob_out_dev, raw_idx_out_dev);
  else
    ctc_argmax_kernel<__half><<<(unsigned)blocks, kCtcThreads, 0, s>>>(
        (const __half*)probs_dev, T, B, C, stride_t, stride_b, idx_out_dev, prob_out_dev, raw_idx_out_dev);
  OCRPP_LAUNCHED();
  prof.mark("ctc_argmax");
  ctc_collapse_kernel<<<(B * 32 + 127) / 128, 128, 0, s>>>(T, B, idx_out_dev, prob_out_dev, len_out_dev, conf_out_dev);
  OCRPP_LAUNCHED();
  prof.mark("ctc_collapse");
  return OCRPP_OK;
}
