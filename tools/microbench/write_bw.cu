// Write-bandwidth probe: what can a store-only stream reach on this GPU, by store flavour?
// (the estimator writes 75x more than it reads, so this -- not the copy figure -- is its ceiling)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

template <int MODE>
__device__ __forceinline__ void store16(float4* p, float4 v) {
  if (MODE == 0) *p = v;
  else if (MODE == 1) __stcs(p, v);
  else if (MODE == 2) __stcg(p, v);
  else if (MODE == 3) __stwt(p, v);
  else asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w));
}

// grid-stride, each warp writes 512 contiguous bytes per instruction
template <int MODE>
__global__ void fill_linear(float4* out, size_t n4) {
  const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) store16<MODE>(out + i, v);
}

// the estimator's pattern: a CTA owns a contiguous 1.4 MB unit, each warp writes rows of 448 B (28 lanes), 32 rows per tile
template <int MODE>
__global__ void fill_units(float4* out, int n_units, int rows, int cols) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
  for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
    float4* base = out + (size_t)u * rows * cols;
    if (lane < cols)
      for (int r0 = warp * 32; r0 < rows; r0 += nw * 32) {
        const int r1 = min(rows, r0 + 32);
#pragma unroll 4
        for (int r = r0; r < r1; ++r) store16<MODE>(base + (size_t)r * cols + lane, v);
      }
  }
}

// same bytes, but the unit is written as one linear 512-B-per-warp stream (what a re-blocked store loop could do)
template <int MODE>
__global__ void fill_units_linear(float4* out, int n_units, int n4_per_unit) {
  const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
  for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
    float4* base = out + (size_t)u * n4_per_unit;
#pragma unroll 4
    for (int i = threadIdx.x; i < n4_per_unit; i += blockDim.x) store16<MODE>(base + i, v);
  }
}

// block-cyclic chunks over the whole buffer: chunk c is written by CTA c % gridDim, each warp taking 512-B pieces;
// the write fronts of all CTAs form one contiguous moving window (what a global tile queue would produce)
template <int MODE>
__global__ void fill_chunked(float4* out, size_t n4, int chunk4) {
  const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
  const size_t n_chunks = n4 / chunk4;
  for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    float4* base = out + c * (size_t)chunk4;
#pragma unroll 4
    for (int i = threadIdx.x; i < chunk4; i += blockDim.x) store16<MODE>(base + i, v);
  }
}

// TMA bulk store: stage 16 KB in shared memory once, then stream it out with cp.async.bulk
__global__ void fill_bulk(char* out, int n_units, int bytes_per_unit, int chunk) {
  extern __shared__ __align__(128) char sm[];
  for (int i = threadIdx.x; i < chunk / 4; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = 1.f;
  asm volatile("fence.proxy.async.shared::cta;");
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
      char* base = out + (size_t)u * bytes_per_unit;
      for (int off = 0; off < bytes_per_unit; off += chunk) {
        const int n = min(chunk, bytes_per_unit - off);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(base + off),
                     "r"((unsigned)__cvta_generic_to_shared(sm)), "r"(n) : "memory");
        asm volatile("cp.async.bulk.commit_group;");
        asm volatile("cp.async.bulk.wait_group.read 8;" ::: "memory");
      }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

template <typename F>
float time_ms(F f, int reps = 5) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  f(); f();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int i = 0; i < reps; ++i) f();
  CK(cudaEventRecord(b));
  CK(cudaEventSynchronize(b));
  CK(cudaGetLastError());
  float ms; CK(cudaEventElapsedTime(&ms, a, b));
  return ms / reps;
}

int main() {
  const int n_units = 8192, rows = 3276, cols = 28;         // 12 GB
  const size_t n4 = (size_t)n_units * rows * cols;
  const double gb = n4 * 16.0 / 1e9;
  float4 *out, *in;
  CK(cudaMalloc(&out, n4 * 16));
  CK(cudaMalloc(&in, n4 * 16));
  CK(cudaMemset(in, 0, n4 * 16));
  int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  printf("SMs %d, %.2f GB per pass\n", sms, gb);
  float ms;
  ms = time_ms([&] { CK(cudaMemsetAsync(out, 0, n4 * 16)); });
  printf("cudaMemset                         %8.3f ms %8.1f GB/s\n", ms, gb / ms * 1e3);
  ms = time_ms([&] { CK(cudaMemcpyAsync(out, in, n4 * 16, cudaMemcpyDeviceToDevice)); });
  printf("cudaMemcpy D2D (r+w bytes)         %8.3f ms %8.1f GB/s\n", ms, 2 * gb / ms * 1e3);
  const char* names[5] = {"st", "st.cs", "st.cg", "st.wt", "st.L1::no_allocate"};
#define RUN_LIN(MODE, G, B) ms = time_ms([&] { fill_linear<MODE><<<G, B>>>(out, n4); }); \
  printf("linear   %-20s g=%5d b=%4d %8.3f ms %8.1f GB/s\n", names[MODE], G, B, ms, gb / ms * 1e3);
  RUN_LIN(0, sms * 8, 256) RUN_LIN(1, sms * 8, 256) RUN_LIN(2, sms * 8, 256) RUN_LIN(3, sms * 8, 256) RUN_LIN(4, sms * 8, 256)
  RUN_LIN(0, sms * 2, 1024) RUN_LIN(0, sms * 32, 256) RUN_LIN(1, sms * 32, 256)
#define RUN_UNI(MODE, G, B) ms = time_ms([&] { fill_units<MODE><<<G, B>>>(out, n_units, rows, cols); }); \
  printf("units448 %-20s g=%5d b=%4d %8.3f ms %8.1f GB/s\n", names[MODE], G, B, ms, gb / ms * 1e3);
  RUN_UNI(0, sms * 2, 256) RUN_UNI(1, sms * 2, 256) RUN_UNI(1, sms * 3, 256) RUN_UNI(1, sms * 4, 256) RUN_UNI(1, sms * 8, 256) RUN_UNI(4, sms * 3, 256) RUN_UNI(2, sms * 3, 256)
  RUN_UNI(1, sms * 1, 1024) RUN_UNI(1, sms * 2, 1024)
#define RUN_UL(MODE, G, B) ms = time_ms([&] { fill_units_linear<MODE><<<G, B>>>(out, n_units, rows * cols); }); \
  printf("unitslin %-20s g=%5d b=%4d %8.3f ms %8.1f GB/s\n", names[MODE], G, B, ms, gb / ms * 1e3);
  RUN_UL(0, sms * 3, 256) RUN_UL(1, sms * 3, 256) RUN_UL(1, sms * 8, 256) RUN_UL(1, sms * 2, 1024)
  for (int chunk4 : {896, 3584, 28 * 3276 / 4}) {            // 14 KB tile, 57 KB, a quarter unit
    for (int g : {sms * 2, sms * 3, sms * 8}) {
      ms = time_ms([&] { fill_chunked<1><<<g, 192>>>(out, n4, chunk4); });
      printf("chunked  st.cs chunk=%7d B g=%5d b= 192 %8.3f ms %8.1f GB/s\n", chunk4 * 16, g, ms, gb / ms * 1e3);
    }
  }
  for (int chunk : {4096, 16384, 65536}) {
    CK(cudaFuncSetAttribute(fill_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, chunk));
    for (int per : {1, 2, 4}) {
      if (per * chunk > 200 * 1024) continue;
      ms = time_ms([&] { fill_bulk<<<sms * per, 128, chunk>>>((char*)out, n_units, rows * cols * 16, chunk); });
      printf("tma bulk chunk=%6d ctas/sm=%d        %8.3f ms %8.1f GB/s\n", chunk, per, ms, gb / ms * 1e3);
    }
  }
  return 0;
}
