// Experiment: how fast can the host expand uint8 image planes to float (the host half of a uint8 PCIe transport)?
//   g++ -O3 -pthread -o /tmp/exp_host_expand tools/exp_host_expand.cpp && /tmp/exp_host_expand
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

__attribute__((target_clones("avx512f", "avx2", "default")))
void expand(const uint8_t* __restrict src, float* __restrict dst, size_t n) {
  for (size_t i = 0; i < n; ++i) dst[i] = (float)src[i];
}

int main() {
  const size_t n = (size_t)64 * 6 * 512 * 384;  // one batch of frame pairs
  uint8_t* src = (uint8_t*)aligned_alloc(4096, n);
  float* dst = (float*)aligned_alloc(4096, n * sizeof(float));
  for (size_t i = 0; i < n; ++i) src[i] = (uint8_t)(i * 2654435761u >> 24);
  memset(dst, 0, n * sizeof(float));
  const unsigned hw = std::thread::hardware_concurrency();
  printf("hardware_concurrency %u\n", hw);
  for (unsigned T : {1u, 2u, 4u, 8u, 16u, 32u}) {
    if (T > hw && T > 16) break;
    double best = 1e9;
    for (int rep = 0; rep < 5; ++rep) {
      auto t0 = std::chrono::steady_clock::now();
      std::vector<std::thread> th;
      for (unsigned t = 0; t < T; ++t)
        th.emplace_back([&, t] { const size_t a = n * t / T, b = n * (t + 1) / T; expand(src + a, dst + a, b - a); });
      for (auto& x : th) x.join();
      const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
      if (ms < best) best = ms;
    }
    printf("threads %2u: %.2f ms per batch (%.1f GB/s written)\n", T, best, n * 4 / best * 1e-6);
  }
  return 0;
}
