// Development aid: db_scan4_kernel alone on synthetic rows (ring slots x CTAs per SM), and the same ring with the
// consumers only draining it (-DOCRPP_S4_DRAIN_ONLY): the streaming ceiling of the bulk-copy ring itself.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I pytorchocr_b200/csrc tools/micro/scan4_bench.cu -o tools/micro/scan4_bench
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <vector>
namespace ocrpp {
int g_tune[8] = {0};
int tuning(int key) { return g_tune[key]; }
int set_error(int code, const char* fmt, ...) { fprintf(stderr, "error %d: %s\n", code, fmt); return code; }
std::atomic<long long> g_launch_count{0};
bool debug_sync() { return false; }
}
#include "db_scan4.cuh"
using namespace ocrpp;
int main(int argc, char** argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 128, H = 736, W = 1280, cap = 642;
  float* maps;
  size_t npx = (size_t)N * H * W;
  cudaMalloc(&maps, npx * 4);
  std::vector<float> h(H * W);
  srand(1);
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) h[y * W + x] = ((x / 37 + y / 23) % 5 == 0) ? 0.8f : 0.1f * (rand() % 1000) / 1000.f;
  for (int n = 0; n < N; ++n) cudaMemcpy(maps + (size_t)n * H * W, h.data(), H * W * 4, cudaMemcpyHostToDevice);
  Scan4Params p{};
  p.maps = maps; p.stride_n = (long long)H * W; p.stride_h = W; p.H = H; p.n0 = 0; p.nimg = N; p.cap = cap;
  p.ncells = W / 4; p.thresh = 0.3f;
  cudaMalloc(&p.scum, (size_t)N * H * (cap + 1) * 8);
  cudaMalloc(&p.srow_cnt, (size_t)N * H * 4);
  cudaMalloc(&p.imgflags, N * 4);
  cudaMemset(p.imgflags, 0, N * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int cfgs[][2] = {{8, 1}, {16, 1}, {32, 1}, {8, 2}, {16, 2}, {8, 3}, {8, 4}};
  for (auto& c : cfgs) {
    g_tune[OCRPP_TUNE_DB_SCAN4_STAGES] = c[0];
    g_tune[OCRPP_TUNE_DB_SCAN4_CTAS] = c[1];
    for (int i = 0; i < 3; ++i) scan4_any<float>(p, 0);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    const int K = 20;
    for (int i = 0; i < K; ++i) scan4_any<float>(p, 0);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= K;
    printf("stages %2d ctas/SM %d: %.4f ms  %.0f GB/s  (%s)\n", c[0], c[1], ms, npx * 4 / ms * 1e-6, cudaGetErrorString(cudaGetLastError()));
  }
  int flags = 0;
  cudaMemcpy(&flags, p.imgflags, 4, cudaMemcpyDeviceToHost);
  printf("flags %d\n", flags);
  return 0;
}
