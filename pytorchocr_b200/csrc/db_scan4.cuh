// K1 (default for rows that fit): the pass over the probability map fed by the bulk-copy engine.
//
// Replaces `pred > thresh` + the pixel traversal of cv2.findContours / BoxScore's mean
// (R/pytocr/postprocess/db_postprocess.py:43-46, db_postprocess_fast/src/db_postprocess.cpp:231-246,211-229) with the
// same per-row output as db_scan_kernel: run starts `x << 48 | cumulative row sum before x` (2^-23 fixed point), the
// run count and the polarity of the first run.
//
// Why another formulation. db_scan_kernel / db_scan3_kernel are bound by the ALU pipe (16 lanes per scheduler: one
// IADD3/LOP3/SHF/ISETP per two cycles), not by issue slots or HBM: ~700 of their ~900 / ~600 warp instructions per
// 1280-pixel row are integer bookkeeping around ballots. Here
//   * one elected thread per CTA streams whole rows into a shared-memory ring with cp.async.bulk (+ mbarrier
//     complete_tx): no load instructions, no registers and no L1 tags are spent on the stream, and the bytes in
//     flight (stages x row bytes per CTA) do not depend on how many warps compute;
//   * a consumer warp takes a row, every lane a CONTIGUOUS chunk of it (kCells 16-byte cells), so the row's mask is a
//     bit field per lane built without ballots: the sign of `thresh - f` (one FADD on the FMA pipe, exact) is shifted in
//     with one funnel shift per pixel; the fixed-point value is the mantissa of `f + 1.0f` (one FADD), summed and range
//     checked with three-input integer adds / max;
//   * transitions of a whole chunk come from one XOR of the (<= 64-bit) field; one packed 64-bit warp scan gives every
//     lane its first output slot and the row sum before its first pixel; only lanes that hold a transition go back to
//     shared memory for the sum inside their chunk (per-cell prefixes parked in a per-warp scratch by vector stores).
#pragma once

#include "common.cuh"

namespace ocrpp {
namespace {

struct Scan4Params {
  const void* maps;
  long long stride_n, stride_h;   // elements
  int H, n0, nimg, cap;
  int ncells;                     // 16-byte cells per row (W * sizeof(T) / 16)
  int stages;                     // ring slots
  float thresh;
  unsigned long long* scum;       // [N*H*(cap+1)]
  int32_t* srow_cnt;              // [N*H]
  int32_t* imgflags;              // [N]
};

constexpr int kS4Warps = 8;                         // consumer warps per CTA (+ 1 producer warp)
constexpr int kS4Threads = (kS4Warps + 1) * 32;

__device__ __forceinline__ uint32_t s4_smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void s4_mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void s4_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void s4_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void s4_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void s4_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

template <typename T>
struct S4Elem;
template <>
struct S4Elem<float> {
  static constexpr int kEpl = 4;
  __device__ static __forceinline__ void unpack(const uint4& u, float* f) {
    f[0] = __uint_as_float(u.x); f[1] = __uint_as_float(u.y); f[2] = __uint_as_float(u.z); f[3] = __uint_as_float(u.w);
  }
};
template <>
struct S4Elem<__half> {
  static constexpr int kEpl = 8;
  __device__ static __forceinline__ void unpack(const uint4& u, float* f) {
    f[0] = h2f_lo(u.x); f[1] = h2f_hi(u.x); f[2] = h2f_lo(u.y); f[3] = h2f_hi(u.y);
    f[4] = h2f_lo(u.z); f[5] = h2f_hi(u.z); f[6] = h2f_lo(u.w); f[7] = h2f_hi(u.w);
  }
};

// shared memory: [stages][ncells*16] rows | [kS4Warps][32*kScrWords] per-lane cell prefixes | full[stages] | empty[stages]
template <int kCells>
struct S4Layout {
  static constexpr int kScrWords = (kCells + 3) & ~3;   // words per lane (a whole number of 16-byte stores)
  __host__ __device__ static size_t bytes(int ncells, int stages) {
    return (size_t)stages * ncells * 16 + (size_t)kS4Warps * 32 * kScrWords * 4 + (size_t)stages * 16;
  }
};

template <typename T, int kCells, bool kExact>
__global__ void __launch_bounds__(kS4Threads) db_scan4_kernel(Scan4Params p) {
  extern __shared__ __align__(128) unsigned char s4_smem[];
  constexpr int kEpl = S4Elem<T>::kEpl;
  constexpr int kPx = kCells * kEpl;                       // pixels per lane, <= 64
  constexpr int kScrWords = S4Layout<kCells>::kScrWords;
  constexpr uint32_t kOne = 0x3f800000u;
  static_assert(kPx <= 64, "a lane's chunk is one 64-bit field");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int stages = p.stages;
  const uint32_t row_bytes = (uint32_t)p.ncells * 16u;
  unsigned char* ring = s4_smem;
  uint32_t* scratch = reinterpret_cast<uint32_t*>(s4_smem + (size_t)stages * row_bytes);
  const uint32_t bar0 = s4_smem_addr(scratch + kS4Warps * 32 * kScrWords);   // full[i] at bar0 + 8 i, empty[i] after them
  const uint32_t full0 = bar0, empty0 = bar0 + 8u * stages;
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) {
      s4_mbar_init(full0 + 8u * i, 1);
      s4_mbar_init(empty0 + 8u * i, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // this CTA's contiguous share of the rows of the sub-batch
  const long long rows = (long long)p.nimg * p.H;
  const int r0 = (int)(rows * blockIdx.x / gridDim.x), r1 = (int)(rows * (blockIdx.x + 1) / gridDim.x);
  const int cnt = r1 - r0;

  if (warp == kS4Warps) {   // ---- producer: one thread feeds the ring ----
    if (lane == 0) {
      int n = r0 / p.H, y = r0 - n * p.H;
      const T* base = reinterpret_cast<const T*>(p.maps);
      int slot = 0;
      uint32_t phase = 0;   // parity of the slot's previous use
      for (int i = 0; i < cnt; ++i) {
        if (i >= stages) s4_mbar_wait(empty0 + 8u * slot, phase ^ 1u);
        const T* src = base + (long long)(n + p.n0) * p.stride_n + (long long)y * p.stride_h;
        s4_mbar_expect_tx(full0 + 8u * slot, row_bytes);
        s4_bulk_g2s(s4_smem_addr(ring + (size_t)slot * row_bytes), src, row_bytes, full0 + 8u * slot);
        if (++y == p.H) { y = 0; ++n; }
        if (++slot == stages) { slot = 0; phase ^= 1u; }
      }
    }
    return;
  }

  // ---- consumers: warp w takes rows w, w + kS4Warps, ... of the CTA's share ----
  uint32_t* scr = scratch + (size_t)(warp * 32 + lane) * kScrWords;
  const float thresh = p.thresh;
  int slot = warp;          // stages is a multiple of kS4Warps: slot s is only ever used by warp s % kS4Warps
  uint32_t par = 0;
  for (int i = warp; i < cnt; i += kS4Warps) {
    const int r = r0 + i;
    s4_mbar_wait(full0 + 8u * slot, par);
    const uint4* rowp = reinterpret_cast<const uint4*>(ring + (size_t)slot * row_bytes) + lane * kCells;
#ifdef OCRPP_S4_DRAIN_ONLY   // tools/micro/scan4_bench.cu: the consumers only drain the ring (its own streaming rate)
    if (rowp[0].x == 0x7fc00001u) p.srow_cnt[r] = 1;
    __syncwarp();
    if (lane == 0) s4_mbar_arrive(empty0 + 8u * slot);
    slot += kS4Warps;
    if (slot >= stages) { slot -= stages; par ^= 1u; }
    continue;
#endif
    uint4 raw[kCells];
#pragma unroll
    for (int c = 0; c < kCells; ++c) {
      if (kExact || lane * kCells + c < p.ncells) raw[c] = rowp[c];
      else raw[c] = make_uint4(0u, 0u, 0u, 0u);
    }
    float f[kCells][kEpl];
#pragma unroll
    for (int c = 0; c < kCells; ++c) S4Elem<T>::unpack(raw[c], f[c]);
    // mask field (bit j <-> pixel j of the chunk), filled from the top of each 32-bit word down: one funnel shift
    // per pixel moves the sign of (thresh - f) in;  f > thresh  <=>  thresh - f < 0  (exact in IEEE arithmetic)
    uint32_t mlo = 0, mhi = 0;
#pragma unroll
    for (int j = kPx - 1; j >= 32; --j) mhi = __funnelshift_l(__float_as_uint(thresh - f[j / kEpl][j % kEpl]), mhi, 1);
#pragma unroll
    for (int j = (kPx < 32 ? kPx : 32) - 1; j >= 0; --j)
      mlo = __funnelshift_l(__float_as_uint(thresh - f[j / kEpl][j % kEpl]), mlo, 1);
    // f in [0,1]: the mantissa of f + 1.0f is round(f * 2^23), i.e. bits(f + 1.0f) - bits(1.0f); anything else
    // (negative, > 1, NaN, Inf) leaves [bits(1.0f), bits(2.0f)] and is caught by the running min / max. The bits are
    // summed as they are (mod 2^32) and the constant is taken off once per use. Cells past the end of the row hold 0.
    uint32_t pre[kCells];   // inclusive per-cell prefix of the raw bit patterns
    uint32_t run = 0, umax = 0u, umin = 0xffffffffu;
#pragma unroll
    for (int c = 0; c < kCells; ++c) {
#pragma unroll
      for (int k = 0; k < kEpl; ++k) {
        const uint32_t u = __float_as_uint(f[c][k] + 1.0f);
        umax = max(umax, u);
        umin = min(umin, u);
        run += u;
      }
      pre[c] = run;
    }
    const uint32_t lane_sum = run - (uint32_t)kPx * kOne;
    // park the cell prefixes for the emission (vector stores: off the ALU pipe)
#pragma unroll
    for (int c = 0; c < kScrWords; c += 4)
      *reinterpret_cast<uint4*>(scr + c) = make_uint4(pre[c < kCells ? c : kCells - 1], pre[c + 1 < kCells ? c + 1 : kCells - 1],
                                                      pre[c + 2 < kCells ? c + 2 : kCells - 1], pre[c + 3 < kCells ? c + 3 : kCells - 1]);
    unsigned long long m = ((unsigned long long)mhi << 32) | mlo;
    unsigned long long vmask = ~0ull >> (64 - kPx);   // valid pixels of the chunk
    if (!kExact) {
      const int nv = max(0, min(kPx, (p.ncells - lane * kCells) * kEpl));
      vmask = nv == 0 ? 0ull : (~0ull >> (64 - nv));
      m &= vmask;
    }
    // transitions: pixel j differs from pixel j-1 (the previous lane's last pixel for j == 0; pixel 0 of the row starts
    // run 0 and is not a transition)
    unsigned prevbit = __shfl_up_sync(0xffffffffu, (unsigned)(m >> (kPx - 1)) & 1u, 1);
    if (lane == 0) prevbit = (unsigned)m & 1u;
    unsigned long long tm = (m ^ ((m << 1) | prevbit)) & vmask;
    const unsigned cnt_l = __popcll(tm);
    // one packed warp scan: transitions in the high part, fixed-point sum (< 2^34 per row) in the low 40 bits
    const unsigned long long mine = ((unsigned long long)cnt_l << 40) | lane_sum;
    unsigned long long v = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    const unsigned long long tot = __shfl_sync(0xffffffffu, v, 31);
    v -= mine;   // exclusive
    unsigned j = 1u + (unsigned)(v >> 40);
    const unsigned long long base = v & ((1ull << 40) - 1);
    unsigned long long* sc = p.scum + ((size_t)p.n0 * p.H + r) * (size_t)(p.cap + 1);
    const unsigned x0 = lane * kPx;
    while (tm) {
      const unsigned k = (unsigned)__ffsll((long long)tm) - 1u;
      tm &= tm - 1;
      const unsigned cell = k / kEpl, within = k % kEpl;
      uint32_t before = cell ? scr[cell - 1] - cell * (kEpl * kOne) : 0u;
      float g[kEpl];
      S4Elem<T>::unpack(rowp[cell], g);
#pragma unroll
      for (int q = 0; q < kEpl - 1; ++q)
        if ((unsigned)q < within) before += __float_as_uint(g[q] + 1.0f) - kOne;
      if (j < (unsigned)p.cap) sc[j] = ((unsigned long long)(x0 + k) << 48) | (base + before);
      ++j;
    }
    if (lane == 0) {
      const int total_cnt = (int)(tot >> 40) + 1;
      sc[0] = 0ull;   // run 0: x = 0, nothing before it
      p.srow_cnt[(size_t)p.n0 * p.H + r] = total_cnt | (((unsigned)m & 1u) << 31);
      if (total_cnt <= p.cap) sc[total_cnt] = tot & ((1ull << 40) - 1);
    }
    __syncwarp();   // every lane is done with the slot
    if (lane == 0) s4_mbar_arrive(empty0 + 8u * slot);
    if (__any_sync(0xffffffffu, umax > 0x40000000u || umin < kOne) && lane == 0)
      atomicOr(&p.imgflags[p.n0 + r / p.H], OCRPP_IMG_VALUE_OUT_OF_RANGE);
    slot += kS4Warps;
    if (slot >= stages) { slot -= stages; par ^= 1u; }
  }
}

// ---- host side ----
constexpr int kS4SmemBudget = 100 * 1024;   // per CTA: two CTAs per SM

template <typename T, int kCells>
int scan4_launch(Scan4Params p, cudaStream_t s) {
  const size_t fixed = S4Layout<kCells>::bytes(p.ncells, 0);
  int stages = (int)((kS4SmemBudget - fixed) / ((size_t)p.ncells * 16 + 16));
  const int want = tuning(OCRPP_TUNE_DB_SCAN4_STAGES);
  if (want > 0) stages = want;
  // a slot must always be drained by the same warp: a warp that asked for use w of a slot while the copy of use w-1
  // (another warp's row) was still in flight would see the barrier's parity alias and pass
  stages = (stages > 64 ? 64 : stages) / kS4Warps * kS4Warps;
  if (stages < kS4Warps) return -1;
  p.stages = stages;
  const size_t smem = S4Layout<kCells>::bytes(p.ncells, stages);
  if (smem > 220 * 1024) return -1;
  const long long rows = (long long)p.nimg * p.H;
  int per_sm = tuning(OCRPP_TUNE_DB_SCAN4_CTAS);
  if (per_sm <= 0) per_sm = 2;
  long long grid = (long long)kNumSMs * per_sm;
  if (grid > (rows + kS4Warps - 1) / kS4Warps) grid = (rows + kS4Warps - 1) / kS4Warps;   // at least a row per consumer warp
  const bool exact = p.ncells == 32 * kCells;
  if (exact) {
    OCRPP_CUDA(cudaFuncSetAttribute(db_scan4_kernel<T, kCells, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    db_scan4_kernel<T, kCells, true><<<(int)grid, kS4Threads, smem, s>>>(p);
  } else {
    OCRPP_CUDA(cudaFuncSetAttribute(db_scan4_kernel<T, kCells, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    db_scan4_kernel<T, kCells, false><<<(int)grid, kS4Threads, smem, s>>>(p);
  }
  return OCRPP_OK;
}

// cells per lane: the smallest instantiated chunk that covers the row. Returns -1 when no variant fits (the caller
// then takes db_scan_kernel), OCRPP_OK after a launch, or an error code.
template <typename T>
int scan4_any(const Scan4Params& p, cudaStream_t s) {
  const int need = (p.ncells + 31) / 32;
  constexpr int kMaxCells = 64 / S4Elem<T>::kEpl;
  if (need > kMaxCells) return -1;
  if (need <= 2) return scan4_launch<T, 2>(p, s);
  if (need <= 3) return scan4_launch<T, 3>(p, s);
  if (need <= 4) return scan4_launch<T, 4>(p, s);
  if (need <= 5) return scan4_launch<T, 5>(p, s);
  if (need <= 6) return scan4_launch<T, 6>(p, s);
  if (need <= 8) return scan4_launch<T, 8>(p, s);
  if constexpr (kMaxCells >= 16) {
    if (need <= 10) return scan4_launch<T, 10>(p, s);
    if (need <= 12) return scan4_launch<T, 12>(p, s);
    if (need <= 16) return scan4_launch<T, 16>(p, s);
  }
  return -1;
}

}  // namespace
}  // namespace ocrpp
