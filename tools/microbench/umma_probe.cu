// Probe of the tcgen05 building blocks used by the tensor-core fill kernel (srsce_fill_tc.cu): one CTA,
// D[128 x N] (fp32, TMEM) = A[128 x K] * B[N x K]^T with bf16 operands in the canonical no-swizzle K-major
// shared-memory layout, K/16 MMAs, one commit, tcgen05.ld epilogue.  Checked against the host.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe umma_probe.cu && ./umma_probe
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

constexpr int M = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Shared-memory matrix descriptor, no swizzle, K-major: core matrix = 8 rows x 16 bytes stored as 128 contiguous
// bytes; lbo = byte distance between the two core matrices an MMA consumes along K, sbo = byte distance between
// consecutive 8-row groups.  (cute::UMMA::SmemDescriptor: start [0,14), lbo [16,30), sbo [32,46), version [46,48) = 1,
// layout type [61,64) = 0.)
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3ffff) >> 4);
  d |= (uint64_t)((lbo & 0x3ffff) >> 4) << 16;
  d |= (uint64_t)((sbo & 0x3ffff) >> 4) << 32;
  d |= 1ull << 46;
  return d;
}

// kind::f16 instruction descriptor: D fp32, A/B bf16, both K-major, M x N.
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a), "l"(b), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int N, int K>
__global__ void __launch_bounds__(128, 1) probe(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B,
                                                float* __restrict__ D) {
  extern __shared__ __align__(128) unsigned char smem[];
  __nv_bfloat16* sa = reinterpret_cast<__nv_bfloat16*>(smem);                 // [M/8][K/8][8][8]
  __nv_bfloat16* sb = sa + M * K;                                            // [N/8][K/8][8][8]
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int idx = tid; idx < M * K; idx += 128) {
    const int r = idx / K, k = idx % K;
    sa[((r >> 3) * (K >> 3) + (k >> 3)) * 64 + (r & 7) * 8 + (k & 7)] = A[idx];
  }
  for (int idx = tid; idx < N * K; idx += 128) {
    const int r = idx / K, k = idx % K;
    sb[((r >> 3) * (K >> 3) + (k >> 3)) * 64 + (r & 7) * 8 + (k & 7)] = B[idx];
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(N < 32 ? 32 : N));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy operand writes -> async proxy (MMA)
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;

  if (tid == 0) {
    constexpr uint32_t lbo = 128, sbo = (K / 8) * 128;
    constexpr uint32_t idesc = make_idesc(M, N);
#pragma unroll
    for (int kk = 0; kk < K / 16; ++kk) {
      const uint64_t da = make_desc(smem_u32(sa) + kk * 256, lbo, sbo);
      const uint64_t db = make_desc(smem_u32(sb) + kk * 256, lbo, sbo);
      mma_bf16(tmem, da, db, idesc, kk > 0);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // wait for the MMAs (phase 0)
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
          : "=r"(done)
          : "r"(smem_u32(&bar)), "r"(0u)
          : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // epilogue: warp w reads TMEM lanes 32w .. 32w+31 (= rows), 8 columns at a time
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t v[8];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) D[(warp * 32 + lane) * N + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(N < 32 ? 32 : N));
}

template <int N, int K>
int run() {
  std::vector<__nv_bfloat16> a(M * K), b(N * K);
  std::vector<float> af(M * K), bf(N * K), ref(M * N), got(M * N);
  srand(1234);
  for (int i = 0; i < M * K; ++i) { a[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); af[i] = __bfloat162float(a[i]); }
  for (int i = 0; i < N * K; ++i) { b[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); bf[i] = __bfloat162float(b[i]); }
  for (int r = 0; r < M; ++r)
    for (int c = 0; c < N; ++c) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)af[r * K + k] * bf[c * K + k];
      ref[r * N + c] = (float)s;
    }
  __nv_bfloat16 *da, *db;
  float* dd;
  cudaMalloc(&da, a.size() * 2);
  cudaMalloc(&db, b.size() * 2);
  cudaMalloc(&dd, got.size() * 4);
  cudaMemcpy(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, b.data(), b.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dd, 0xff, got.size() * 4);
  const int smem = (M + N) * K * 2;
  cudaFuncSetAttribute(probe<N, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<N, K><<<1, 128, smem>>>(da, db, dd);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("N=%d K=%d: CUDA error %s\n", N, K, cudaGetErrorString(e)); return 1; }
  cudaMemcpy(got.data(), dd, got.size() * 4, cudaMemcpyDeviceToHost);
  double worst = 0;
  int bad = 0;
  for (int i = 0; i < M * N; ++i) {
    const double d = fabs((double)got[i] - ref[i]);
    if (!(d <= 1e-4)) { if (bad < 5) printf("  mismatch row %d col %d: got %g want %g\n", i / N, i % N, got[i], ref[i]); ++bad; }
    if (d > worst) worst = d;
  }
  printf("umma_probe M=128 N=%d K=%d: max abs err %.3g, mismatches %d -> %s\n", N, K, worst, bad, bad ? "FAIL" : "OK");
  cudaFree(da); cudaFree(db); cudaFree(dd);
  return bad != 0;
}

int main() {
  int rc = 0;
  rc |= run<64, 32>();
  rc |= run<256, 96>();
  rc |= run<256, 192>();
  return rc;
}
