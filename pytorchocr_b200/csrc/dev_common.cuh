// Device helpers shared by the detection pipelines (db.cu, expand.cu): block scan, run-transition
// bit tricks, lock-free union-find on runs, workspace carving. Internal linkage on purpose (each
// translation unit gets its own copy; the library is built without relocatable device code).
#pragma once
#include "common.cuh"

namespace ocrpp {
namespace {

// ------------------------------------------------------------------------------------------------
// block-wide exclusive scan of one int per thread; returns the exclusive prefix, total in *total
// ------------------------------------------------------------------------------------------------
__device__ int block_exclusive_scan(int v, int* total) {
  __shared__ int warp_sums[32];
  __shared__ int s_total;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();  // protect warp_sums reuse across calls
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int ws = lane < nw ? warp_sums[lane] : 0;
    int winc = ws;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    if (lane < nw) warp_sums[lane] = winc - ws;
    if (lane == 31) s_total = winc;
  }
  __syncthreads();
  *total = s_total;
  return warp_sums[warp] + inc - v;
}

__device__ __forceinline__ unsigned valid_mask(int k, int W) {
  const int rem = W - k * 32;
  return rem >= 32 ? 0xffffffffu : ((1u << rem) - 1u);
}

// transitions inside word k of a row: bit i set <=> pixel 32k+i differs from pixel 32k+i-1
// (bit 0 of word 0 is never a transition). `prev` = word k-1 (ignored for k == 0).
__device__ __forceinline__ unsigned transitions(unsigned w, unsigned prev, int k, int W) {
  const unsigned carry = k > 0 ? (prev >> 31) : (w & 1u);
  return (w ^ ((w << 1) | carry)) & valid_mask(k, W);
}

// ------------------------------------------------------------------------------------------------
// union-find on runs (atomicMin linking: the root of a set is its smallest run index = the run
// holding the component's first raster pixel)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int uf_find(const int32_t* par, int x) {
  int q;
  while ((q = __ldcg(par + x)) != x) x = q;
  return x;
}

// find with intermediate pointer jumping (path halving): every node visited is re-pointed to its
// grandparent. Parents only ever decrease towards the root, so the plain stores are benign under the
// concurrent atomicMin linking (as in ECL-CC); long chains (the background region spans every row of
// the image) collapse after a few finds.
__device__ __forceinline__ int uf_find_compress(int32_t* par, int x) {
  int p = __ldcg(par + x);
  while (p != x) {
    const int g = __ldcg(par + p);
    if (g != p) par[x] = g;
    x = p;
    p = g;
  }
  return x;
}

__device__ void uf_union(int32_t* par, int a, int b) {
  while (true) {
    a = uf_find_compress(par, a);
    b = uf_find_compress(par, b);
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

struct Carver {
  char* base;
  size_t off;
  template <typename T>
  T* take(size_t count) {
    off = align_up(off, 256);
    T* ptr = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += count * sizeof(T);
    return ptr;
  }
};

}  // namespace
}  // namespace ocrpp
