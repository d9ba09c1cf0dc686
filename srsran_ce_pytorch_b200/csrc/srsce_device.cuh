// Device helpers shared by the fused estimator kernel (srsce_kernels.cu) and the tensor-core fill kernel
// (srsce_fill_tc.cu): complex arithmetic and the writer warps' streaming loops.  Include inside namespace srsce.
#pragma once

#include <cuda_runtime.h>

#include "srsce_internal.h"

namespace srsce {
namespace {

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cmul_conj(float2 a, float2 b) {  // a * conj(b)
  return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cscale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }

// Writer hot loop for a full 32-row tile of a 14-symbol grid with LL (2 or 4) layers: every lane owns one 16-byte
// store slot (symbol, layer pair) and walks the rows with compile-time strides, so each row costs one LDS.128 and
// one STG.128 with the offsets folded into the instructions.  (32 / cols) rows are covered per warp pass.
// The shared loads are issued kStreamBatch rows at a time ahead of their stores: the writer warps share the SM's
// in-order load/store pipe with the estimator warps, and while the stores drain at HBM speed a shared load waits
// ~1 k cycles in that pipe -- the bytes a writer warp has in flight per such round trip set its store rate.
constexpr int kStreamBatch = 16;
template <int LL, bool ROT>
__device__ __forceinline__ void stream_full_tile(const float4* __restrict__ src, float4* __restrict__ dst, bool on,
                                                 float2 rot0, float2 rot1) {
  constexpr int kCols = 14 * LL / 2, kRowsPerPass = 32 / kCols, kPasses = (kTileRows + kRowsPerPass - 1) / kRowsPerPass;
  static_assert(kTileRows % kRowsPerPass == 0, "tile rows must split evenly over the warp passes");
  constexpr int kB = kPasses < kStreamBatch ? kPasses : kStreamBatch;
  static_assert(kPasses % kB == 0, "passes must split evenly into batches");
#pragma unroll
  for (int p0 = 0; p0 < kPasses; p0 += kB) {
    float4 v[kB];
#pragma unroll
    for (int k = 0; k < kB; ++k) v[k] = on ? src[(p0 + k) * (kRowsPerPass * LL / 2)] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < kB; ++k) {
      float4 w = v[k];
      if (ROT) {
        const float2 a = cmul(make_float2(w.x, w.y), rot0), b = cmul(make_float2(w.z, w.w), rot1);
        w = make_float4(a.x, a.y, b.x, b.y);
      }
      __stcs(dst + (p0 + k) * (kRowsPerPass * kCols), w);
    }
  }
}

// Odd layer counts (1 or 3): the two complex64 elements of a 16-byte slot belong to different (symbol, layer)
// pairs, so each is fetched (8 bytes) and gated on its own; strides are still compile-time constants.
template <int LL, bool ROT>
__device__ __forceinline__ void stream_full_tile_odd(const float2* __restrict__ src0, const float2* __restrict__ src1,
                                                     float4* __restrict__ dst, bool on0, bool on1, float2 rot0,
                                                     float2 rot1) {
  constexpr int kCols = 14 * LL / 2, kRowsPerPass = 32 / kCols, kPasses = kTileRows / kRowsPerPass;
  static_assert(kTileRows % kRowsPerPass == 0, "tile rows must split evenly over the warp passes");
  if (LL == 1) {
    // one layer: both elements of a slot are the SAME row value on two symbols -- one 8-byte shared load per store,
    // all loads of the tile issued ahead of the stores (see kStreamBatch above)
    float2 v[kPasses];
#pragma unroll
    for (int p = 0; p < kPasses; ++p) v[p] = (on0 || on1) ? src0[p * kRowsPerPass] : make_float2(0.f, 0.f);
#pragma unroll
    for (int p = 0; p < kPasses; ++p) {
      float2 a = on0 ? v[p] : make_float2(0.f, 0.f), b = on1 ? v[p] : make_float2(0.f, 0.f);
      if (ROT) {
        a = cmul(a, rot0);
        b = cmul(b, rot1);
      }
      __stcs(dst + p * (kRowsPerPass * kCols), make_float4(a.x, a.y, b.x, b.y));
    }
    return;
  }
#pragma unroll
  for (int p = 0; p < kPasses; ++p) {
    float2 a = on0 ? src0[p * (kRowsPerPass * LL)] : make_float2(0.f, 0.f);
    float2 b = on1 ? src1[p * (kRowsPerPass * LL)] : make_float2(0.f, 0.f);
    if (ROT) {
      a = cmul(a, rot0);
      b = cmul(b, rot1);
    }
    __stcs(dst + p * (kRowsPerPass * kCols), make_float4(a.x, a.y, b.x, b.y));
  }
}

}  // namespace
}  // namespace srsce
