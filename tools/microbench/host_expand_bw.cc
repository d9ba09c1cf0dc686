// Host-side bandwidth probe: how fast can the box's cores replicate compact channel rows [n_sc][L] over 14 symbols into
// the dense [n_sc][14][L] grid of the reference contract (1.47 MB per unit at config 2)?  Decides whether a host-buffer
// entry point that moves the FACTORS over PCIe and expands on the host can beat the PCIe-bound dense D2H (55 GB/s).
//   g++ -O3 -mavx2 -pthread -o host_expand_bw host_expand_bw.cc && ./host_expand_bw [threads]
#include <immintrin.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <thread>
#include <vector>

static void expand_rows(const float* src, float* dst, long rows, int L, bool stream) {
  const int row_f = 2 * L;              // floats per compact row
  for (long r = 0; r < rows; ++r) {
    const float* s = src + r * row_f;
    float* d = dst + r * 14 * row_f;
    if (L == 4 && stream) {
      const __m256 v = _mm256_loadu_ps(s);
      for (int t = 0; t < 14; ++t) _mm256_stream_ps(d + t * 8, v);
    } else {
      for (int t = 0; t < 14; ++t) memcpy(d + t * row_f, s, row_f * sizeof(float));
    }
  }
}

int main(int argc, char** argv) {
  const int threads = argc > 1 ? atoi(argv[1]) : (int)std::thread::hardware_concurrency();
  const long units = 1024, n_sc = 3276;
  const int L = 4;
  const long rows = units * n_sc;
  float* src = (float*)aligned_alloc(64, rows * 2 * L * sizeof(float));
  float* dst = (float*)aligned_alloc(64, rows * 14 * 2 * L * sizeof(float));
  for (long i = 0; i < rows * 2 * L; ++i) src[i] = float(i & 1023);
  memset(dst, 0, rows * 14 * 2 * L * sizeof(float));
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 3; ++rep) {
      auto t0 = std::chrono::steady_clock::now();
      std::vector<std::thread> th;
      for (int k = 0; k < threads; ++k) {
        const long r0 = rows * k / threads, r1 = rows * (k + 1) / threads;
        th.emplace_back(expand_rows, src + r0 * 2 * L, dst + r0 * 14 * 2 * L, r1 - r0, L, mode == 1);
      }
      for (auto& t : th) t.join();
      const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      printf("%s stores, %d threads: %.1f GB/s written, %.0f units/s = %.0f estimates/s\n", mode ? "streaming" : "memcpy", threads,
             rows * 14.0 * 2 * L * 4 / s / 1e9, units / s, 4 * units / s);
    }
  }
  return 0;
}
