// Probe of the TMA pilot-RE gather used by the fused kernel: the received grid [unit][n_sc][14] complex64 (reference
// layout) is described as a rank-4 tensor [symbol 14][RE position 12][PRB][unit] of 8-byte elements; one
// cp.async.bulk.tensor op with box {2, 1, PB, 1} fetches the symbol PAIR holding a pilot symbol for one RE position of
// PB consecutive PRBs -- PB 16-byte rows, 1344 bytes apart in global memory, packed densely in shared memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_gather_probe tma_gather_probe.cu && ./tma_gather_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap map, float2* out, int pb, int n_ops, const int4* coords) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t bytes = uint32_t(n_ops) * pb * 16;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
    for (int k = 0; k < n_ops; ++k) {
      const int4 c = coords[k];
      asm volatile(
          "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
              smem_u32(smem + k * pb * 16)),
          "l"(&map), "r"(c.x), "r"(c.y), "r"(c.z), "r"(c.w), "r"(smem_u32(&bar))
          : "memory");
    }
  }
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
  const float2* s = reinterpret_cast<const float2*>(smem);
  for (int k = threadIdx.x; k < n_ops * pb * 2; k += blockDim.x) out[k] = s[k];
}

int main() {
  const int n_prb = 273, n_sc = 12 * n_prb, n_sym = 14, n_units = 3, pb = 144;   // pb * 16 bytes must be a multiple of 128: every op lands on a 128-byte aligned shared address
  std::vector<float2> h((size_t)n_units * n_sc * n_sym);
  for (size_t i = 0; i < h.size(); ++i) h[i] = make_float2(float(i % 1000003), float(i / 1000003));
  float2* d;
  cudaMalloc(&d, h.size() * 8);
  cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &q) != cudaSuccess || !encode) {
    printf("no cuTensorMapEncodeTiled\n");
    return 1;
  }
  alignas(64) CUtensorMap map;
  const cuuint64_t dims[4] = {(cuuint64_t)n_sym, 12, (cuuint64_t)n_prb, (cuuint64_t)n_units};
  const cuuint64_t strides[3] = {(cuuint64_t)n_sym * 8, (cuuint64_t)n_sym * 8 * 12, (cuuint64_t)n_sym * 8 * n_sc};
  const cuuint32_t box[4] = {2, 1, (cuuint32_t)pb, 1}, estr[4] = {1, 1, 1, 1};
  CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  // ops: unit 1, pilot symbol 13 (pair 12), RE positions 0, 4, 8 x two PRB boxes; unit 2, symbol 3 (pair 2), position 6
  std::vector<int4> c = {{12, 0, 0, 1}, {12, 0, pb, 1}, {12, 4, 0, 1}, {12, 4, pb, 1}, {12, 8, 0, 1}, {12, 8, pb, 1}, {2, 6, 0, 2}, {2, 6, pb, 2}};
  int4* dc;
  cudaMalloc(&dc, c.size() * sizeof(int4));
  cudaMemcpy(dc, c.data(), c.size() * sizeof(int4), cudaMemcpyHostToDevice);
  float2* dout;
  const int n_out = (int)c.size() * pb * 2;
  cudaMalloc(&dout, n_out * 8);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  probe<<<1, 128, (int)c.size() * pb * 16>>>(map, dout, pb, (int)c.size(), dc);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<float2> got(n_out);
  cudaMemcpy(got.data(), dout, n_out * 8, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (size_t k = 0; k < c.size(); ++k)
    for (int p = 0; p < pb; ++p)
      for (int e2 = 0; e2 < 2; ++e2) {
        const int prb = c[k].z + p;
        float2 want = make_float2(0.f, 0.f);                       // PRBs beyond the grid are zero filled
        if (prb < n_prb) want = h[((size_t)c[k].w * n_sc + 12 * prb + c[k].y) * n_sym + c[k].x + e2];
        const float2 g = got[(k * pb + p) * 2 + e2];
        if (g.x != want.x || g.y != want.y) {
          if (bad < 5) printf("  mismatch op %zu prb %d e %d: got (%g,%g) want (%g,%g)\n", k, prb, e2, g.x, g.y, want.x, want.y);
          ++bad;
        }
      }
  printf("tma_gather_probe: %zu ops x %d PRBs: mismatches %d -> %s\n", c.size(), pb, bad, bad ? "FAIL" : "OK");
  return bad != 0;
}
