// Tensor-core fill for sm_100a: the frequency interpolation of fill_ch_est_cdm -- the fixed-weight "CNN" inpainting
// operator of ce_dl_cnn.py:233-322 (:433-508) or the linear interpolation of ce_rule_tensorized.py:237-354 -- applied
// to a batch of smoothed pilot vectors as a banded GEMM on tcgen05, followed by the same dense [n_sc, 14, L] grid
// write as the fused estimator.
//
//   per CDM group g and 96-row band tile m (8 PRBs):
//     D_g[96 x (unit, layer in g, re/im)]  =  A_{kind(m), g}[96 x K]  *  X_g[K x (unit, layer, re/im)]
//   A: operator rows of the tile over its pilot window [k0(m), k0(m) + K), built by the host planner (float64
//      recurrence of the reference, srsce_plan.cc) and stored as bf16 terms a0 (+ a1 + a2 when not exact in bf16);
//   X: the pilots of 32 / 16 / 8 units, float32 values split on the fly into three bf16 terms x0 + x1 + x2;
//   the products a_i * x_j with i + j <= 2 are accumulated in float32 in tensor memory, which reproduces the float32
//   operator result to ~1e-7 (the direct path computes the same sum with float32 FMAs).
//
// One persistent CTA per SM, warp specialised:
//   warps 0-7   epilogue / writers: tcgen05.ld of the accumulator rows (lane = band row), 32-row tile in shared
//               memory, then the streaming 448-byte row stores of the fused kernel (st.global.cs.v4)
//   warps 8-11  loaders: pilots of the next work item -> bf16 terms -> canonical K-major core-matrix layout
//   warp 12     one thread issues the tcgen05.mma (M = 128, N = units * layers * 2, K = 16) and the commits
// Pipelines: X stages (2) between loaders and MMA, accumulator stages (2 x 256 tensor-memory columns) between MMA
// and epilogue, all on mbarriers.  Arithmetic intensity is ~0.4 FLOP/B, so the kernel is bound by the grid write
// exactly like the direct path; the tensor pipe is busy for ~1 % of the time by construction.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "srsce_device.cuh"
#include "srsce_internal.h"

namespace srsce {

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}

// ---- tcgen05
// Shared-memory matrix descriptor, no swizzle, K-major (cute::UMMA::SmemDescriptor): core matrix = 8 rows x 16 bytes
// stored as 128 contiguous bytes; lbo = distance between the two core matrices one MMA consumes along K, sbo = distance
// between consecutive 8-row groups.  Start [0,14), lbo [16,30), sbo [32,46) in 16-byte units, version [46,48) = 1.
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr & 0x3ffff) >> 4) | ((uint64_t)((lbo & 0x3ffff) >> 4) << 16) |
         ((uint64_t)((sbo & 0x3ffff) >> 4) << 32) | (1ull << 46);
}
// kind::f16 instruction descriptor: D float32 (bit 4), A and B bf16 (bits 7, 10), both K-major, N >> 3 at [17,23),
// M >> 4 at [24,29)  (cute::UMMA::InstrDescriptor)
__device__ __forceinline__ uint32_t umma_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a), "l"(b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {      // arrives on bar when all MMAs issued so far are done
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&v)[4]) {
  uint32_t r0, r1, r2, r3;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(taddr));
  v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
}
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, float (&v)[4]) {
  uint32_t r0, r1;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(taddr));
  v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = 0.f; v[3] = 0.f;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// float32 -> three bf16 terms (x = x0 + x1 + x2 up to 2^-24 relative)
__device__ __forceinline__ void split3(float x, unsigned short (&t)[3]) {
  const __nv_bfloat16 b0 = __float2bfloat16_rn(x);
  const float r1 = x - __bfloat162float(b0);
  const __nv_bfloat16 b1 = __float2bfloat16_rn(r1);
  const float r2 = r1 - __bfloat162float(b1);
  const __nv_bfloat16 b2 = __float2bfloat16_rn(r2);
  t[0] = __bfloat16_as_ushort(b0);
  t[1] = __bfloat16_as_ushort(b1);
  t[2] = __bfloat16_as_ushort(b2);
}

struct FillBars {            // at smem + off_bar
  uint64_t x_full[2], x_empty[2], d_full[2], d_empty[2];
  uint32_t tmem_base;
};

// Stream `rows_here` rows (<= 32) of a [row][L] tile in shared memory into the dense grid: generic form of
// stream_full_tile for partial tiles; slot c of a row = elements 2c, 2c+1 of its 14*L elements.
template <int LL>
__device__ __forceinline__ void stream_rows(const float2* __restrict__ tile, float2* __restrict__ out_row0, int rows_here,
                                            int lane, int sym0, int sym1) {
  constexpr int kCols = 14 * LL / 2, kRowsPerPass = 32 / kCols;
  const int my_rs = lane / kCols, my_c = lane - my_rs * kCols;
  if (my_rs >= kRowsPerPass) return;
  const int e0 = 2 * my_c, e1 = e0 + 1;
  const int s0 = e0 / LL, l0 = e0 - s0 * LL, s1 = e1 / LL, l1 = e1 - s1 * LL;
  const bool on0 = s0 >= sym0 && s0 < sym1, on1 = s1 >= sym0 && s1 < sym1;
  for (int r = my_rs; r < rows_here; r += kRowsPerPass) {
    const float2 a = on0 ? tile[r * LL + l0] : make_float2(0.f, 0.f);
    const float2 b = on1 ? tile[r * LL + l1] : make_float2(0.f, 0.f);
    __stcs(reinterpret_cast<float4*>(out_row0 + (long long)r * (14 * LL) + e0), make_float4(a.x, a.y, b.x, b.y));
  }
}

template <int LL>
__global__ void __launch_bounds__(kFillThreads, 1)
fill_tc_kernel(const __grid_constant__ FillTcDev F, const float2* __restrict__ h, float2* __restrict__ out,
               long long n_units) {
  extern __shared__ __align__(1024) unsigned char smem[];
  FillBars& bars = *reinterpret_cast<FillBars*>(smem + F.off_bar);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int G = (LL + 1) / 2;
  const int KT = F.k_tile, U = F.units, N = F.n_re;
  const long long n_batches = (n_units + U - 1) / U;
  const long long n_items = n_batches * F.n_tiles;

  // ---- one-time setup: operator blocks -> shared memory, barriers, tensor memory
  {
    const int a_bytes = F.n_kinds * G * F.n_a * F.a_block_bytes;
    const uint4* src = reinterpret_cast<const uint4*>(F.a_tiles);
    uint4* dst = reinterpret_cast<uint4*>(smem + F.off_a);
    for (int k = tid; k < a_bytes / 16; k += kFillThreads) dst[k] = __ldg(src + k);
    // (phantom rows 96..127 of the last block read the X stages behind it: finite or not, those accumulator rows are never read)
  }
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars.x_full[s], 32 * kFillLoadWarps);
      mbar_init(&bars.x_empty[s], 1);
      mbar_init(&bars.d_full[s], 1);
      mbar_init(&bars.d_empty[s], 6);             // lane 0 of the six epilogue warps that own rows 0..95
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kFillEpiWarps + kFillLoadWarps) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars.tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // operator blocks were written by the generic proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars.tmem_base;

  if (warp < kFillEpiWarps) {
    // ===================================== epilogue / writer warps ===========================================
    const int quad = warp & 3, half = warp >> 2;
    float2* const tile = reinterpret_cast<float2*>(smem + F.off_tile) + warp * (kTileRows * LL);
    // rows of the grid outside the hop band are zero: each warp takes 32-row chunks of (unit, chunk) pairs
    {
      const int below = F.band_sc0, above = F.n_sc - (F.band_sc0 + F.band_rows);
      const int chunks_b = (below + 31) / 32, chunks_a = (above + 31) / 32, chunks = chunks_b + chunks_a;
      const long long total = n_units * chunks;
      for (long long w = (long long)blockIdx.x * kFillEpiWarps + warp; w < total; w += (long long)gridDim.x * kFillEpiWarps) {
        const long long u = w / chunks;
        const int c = int(w - u * chunks);
        const int r0 = c < chunks_b ? c * 32 : F.band_sc0 + F.band_rows + (c - chunks_b) * 32;
        const int r1 = c < chunks_b ? min(below, r0 + 32) : min(F.n_sc, r0 + 32);
        float4* dst = reinterpret_cast<float4*>(out + (u * F.n_sc + r0) * (14 * LL));
        const int n16 = (r1 - r0) * 14 * LL / 2;
        for (int k = lane; k < n16; k += 32) __stcs(dst + k, make_float4(0.f, 0.f, 0.f, 0.f));
      }
    }
    if (quad == 3) return;                                   // MMA rows 96..127 do not exist in a 96-row tile
    // this lane's store slot for the unrolled full-tile loops
    constexpr int kCols = 14 * LL / 2, kRowsPerPass = 32 / kCols;
    const int my_rs = lane / kCols, my_c = lane - my_rs * kCols;
    const bool lane_on = my_rs < kRowsPerPass;
    const int e0 = 2 * my_c, e1 = e0 + 1;
    const int s0 = e0 / LL, l0 = e0 - s0 * LL, s1 = e1 / LL, l1 = e1 - s1 * LL;
    const bool on0 = s0 >= F.sym0 && s0 < F.sym1, on1 = s1 >= F.sym0 && s1 < F.sym1;
    const float2 one = make_float2(1.f, 0.f);
    int it = 0;
    for (long long item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int stage = it & 1, par = (it >> 1) & 1;
      const long long batch = item / F.n_tiles;
      const int m = int(item - batch * F.n_tiles);
      const int row0 = m * kFillTileRows + quad * 32;                 // band row of this warp's lane 0
      const int rows_here = min(32, F.band_rows - row0);
      const long long u0 = batch * U;
      const int u_here = int(min((long long)U, n_units - u0));
      mbar_wait(&bars.d_full[stage], par);
      tc_fence_after();
      if (rows_here > 0) {
        // software pipeline over this warp's units: the accumulator row of unit u + 2 is fetched from tensor memory
        // while the rows of unit u stream out; the 32-row tile is double buffered so that one __syncwarp per unit suffices
        auto fetch = [&](int u, float (&v)[G][4]) {
#pragma unroll
          for (int g = 0; g < G; ++g) {
            const int nl = (2 * g + 1 < LL) ? 2 : 1;
            const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + stage * 256 + F.col0[g] + u * nl * 2;
            if (2 * g + 1 < LL) tmem_ld4(taddr, v[g]);
            else tmem_ld2(taddr, v[g]);
          }
        };
        float v[G][4];
        if (half < u_here) fetch(half, v);
        int tb = 0;
        for (int u = half; u < u_here; u += 2, tb ^= 1) {
          float2* const tl = tile + tb * (kFillEpiWarps * kTileRows * LL);
          tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < G; ++g) {
            tl[lane * LL + 2 * g] = make_float2(v[g][0], v[g][1]);
            if (2 * g + 1 < LL) tl[lane * LL + 2 * g + 1] = make_float2(v[g][2], v[g][3]);
          }
          if (u + 2 < u_here) fetch(u + 2, v);
          __syncwarp();                       // tile complete; also: every lane is done streaming the tile of unit u - 2 (same buffer)
          float2* const orow = out + ((u0 + u) * F.n_sc + F.band_sc0 + row0) * (14 * LL);
          if (rows_here == 32) {
            if (lane_on) {
              float4* dst = reinterpret_cast<float4*>(orow + (long long)my_rs * (14 * LL) + e0);
              if constexpr (LL == 4 || LL == 2)
                stream_full_tile<LL, false>(reinterpret_cast<const float4*>(tl + my_rs * LL + l0), dst, on0, one, one);
              else
                stream_full_tile_odd<LL, false>(tl + my_rs * LL + l0, tl + my_rs * LL + l1, dst, on0, on1, one, one);
            }
          } else {
            stream_rows<LL>(tl, orow, rows_here, lane, F.sym0, F.sym1);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.d_empty[stage]);
    }
  } else if (warp < kFillEpiWarps + kFillLoadWarps) {
    // ============================================ loader warps ================================================
    const int lt = tid - 32 * kFillEpiWarps;                           // 0 .. 127
    const int kchunks = KT >> 3;
    int it = 0;
    for (long long item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int stage = it & 1, par = (it >> 1) & 1;
      const long long batch = item / F.n_tiles;
      const int m = int(item - batch * F.n_tiles);
      const int k0 = F.tile_k0[m];
      const long long u0 = batch * U;
      unsigned char* const xs = smem + F.off_b + stage * F.b_stage_bytes;
      mbar_wait(&bars.x_empty[stage], par ^ 1);
      // task = (unit, 8-pilot chunk, CDM group): 8 x (layers of the group) complex64 loads, 3 bf16 terms each,
      // one 16-byte core-matrix row per (layer, re/im, term)
      const int tasks = U * kchunks * G;
      for (int t = lt; t < tasks; t += 32 * kFillLoadWarps) {
        const int kc = t % kchunks, rest = t / kchunks, g = rest % G, u = rest / G;
        const int nl = (2 * g + 1 < LL) ? 2 : 1;
        const bool unit_ok = u0 + u < n_units;
        const float2* src = h + ((u0 + u) * N + k0 + kc * 8) * LL + 2 * g;
        uint32_t term[2][2][3][4];                                     // [layer][re/im][term][k / 2]: two bf16 per word
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const bool ok = unit_ok && (k0 + kc * 8 + k) < N;
#pragma unroll
          for (int l = 0; l < 2; ++l) {
            if (l < nl) {
              const float2 x = ok ? __ldg(src + k * LL + l) : make_float2(0.f, 0.f);
              unsigned short tr[3], ti[3];
              split3(x.x, tr);
              split3(x.y, ti);
#pragma unroll
              for (int j = 0; j < 3; ++j) {
                if (k & 1) {
                  term[l][0][j][k >> 1] |= (uint32_t)tr[j] << 16;
                  term[l][1][j][k >> 1] |= (uint32_t)ti[j] << 16;
                } else {
                  term[l][0][j][k >> 1] = tr[j];
                  term[l][1][j][k >> 1] = ti[j];
                }
              }
            }
          }
        }
        unsigned char* const xg = xs + F.b_group_off[g];
#pragma unroll
        for (int l = 0; l < 2; ++l) {
          if (l < nl) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const int n = (u * nl + l) * 2 + c;                       // row of X^T (the MMA's N index)
              const int off = ((n >> 3) * kchunks + kc) * 128 + (n & 7) * 16;
#pragma unroll
              for (int j = 0; j < 3; ++j) {
                const uint32_t* q = term[l][c][j];
                *reinterpret_cast<uint4*>(xg + j * F.b_block_bytes[g] + off) = make_uint4(q[0], q[1], q[2], q[3]);
              }
            }
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // X written by the generic proxy, read by the MMA
      mbar_arrive(&bars.x_full[stage]);
    }
  } else {
    // ============================================ MMA issuer ==================================================
    if (lane == 0) {
      const uint32_t lbo = 128, sbo = (uint32_t)(KT >> 3) * 128;
      uint32_t idesc[G];
#pragma unroll
      for (int g = 0; g < G; ++g) idesc[g] = umma_idesc(128, F.n_cols[g]);
      const uint32_t a_base = smem_u32(smem + F.off_a), x_base = smem_u32(smem + F.off_b);
      int it = 0;
      for (long long item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const int stage = it & 1, par = (it >> 1) & 1;
        const int m = int(item % F.n_tiles);
        const int kind = F.tile_kind[m];
        mbar_wait(&bars.d_empty[stage], par ^ 1);                        // the epilogue has drained this accumulator stage
        mbar_wait(&bars.x_full[stage], par);
        tc_fence_after();
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const uint32_t d = tmem + stage * 256 + F.col0[g];
          uint32_t acc = 0;
          for (int pr = 0; pr < F.n_prod; ++pr) {
            const uint32_t a_blk = a_base + ((kind * G + g) * F.n_a + F.prod_a[pr]) * F.a_block_bytes;
            const uint32_t x_blk = x_base + stage * F.b_stage_bytes + F.b_group_off[g] + F.prod_x[pr] * F.b_block_bytes[g];
            for (int kk = 0; kk < (KT >> 4); ++kk) {
              umma_bf16(d, umma_desc(a_blk + kk * 256, lbo, sbo), umma_desc(x_blk + kk * 256, lbo, sbo), idesc[g], acc);
              acc = 1;
            }
          }
        }
        umma_commit(&bars.x_empty[stage]);                               // X stage may be refilled ...
        umma_commit(&bars.d_full[stage]);                                // ... and the accumulators are complete
      }
    }
  }

  // ---- teardown (the quad-3 epilogue warps left early; everyone else meets here)
  tc_fence_before();
  asm volatile("bar.sync 1, %0;" ::"r"(kFillThreads - 64) : "memory");
  if (warp == kFillEpiWarps + kFillLoadWarps) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
  }
}

typedef void (*FillFn)(const FillTcDev, const float2*, float2*, long long);
FillFn pick_fill(int layers) {
  switch (layers) {
    case 1: return fill_tc_kernel<1>;
    case 2: return fill_tc_kernel<2>;
    case 3: return fill_tc_kernel<3>;
    default: return fill_tc_kernel<4>;
  }
}

}  // namespace

int prepare_fill_tc(Plan& p, std::string& err) {
  if (!p.has_fill_tc) return SRSCE_OK;
  cudaError_t e;
  const size_t bytes = p.fill_tc_a.size() * sizeof(uint16_t);
  if ((e = cudaMalloc(&p.d_fill_tc_a, bytes)) != cudaSuccess ||
      (e = cudaMemcpy(p.d_fill_tc_a, p.fill_tc_a.data(), bytes, cudaMemcpyHostToDevice)) != cudaSuccess) {
    err = std::string("tensor-core fill tables: ") + cudaGetErrorString(e);
    return SRSCE_ERR_CUDA;
  }
  p.fill_tc.a_tiles = static_cast<const uint16_t*>(p.d_fill_tc_a);
  if ((e = cudaFuncSetAttribute(reinterpret_cast<const void*>(pick_fill(p.fill_tc.n_layers)),
                                cudaFuncAttributeMaxDynamicSharedMemorySize, max_dynamic_smem())) != cudaSuccess) {
    err = std::string("cudaFuncSetAttribute(fill_tc): ") + cudaGetErrorString(e);
    return SRSCE_ERR_CUDA;
  }
  return SRSCE_OK;
}

int launch_fill_tc(const Plan& p, const void* h_pilots, void* out, int64_t n_units, void* stream, int* launches,
                   std::string& err) {
  *launches = 0;
  if (p.device < 0) { err = "plan has no CUDA device (created without a GPU); there is no CPU compute path"; return SRSCE_ERR_NO_DEVICE; }
  if (!p.has_fill_tc) { err = "tensor-core fill unavailable for this plan: " + p.fill_tc_why; return SRSCE_ERR_UNSUPPORTED; }
  int dev = -1;
  cudaGetDevice(&dev);
  if (dev != p.device) { err = "plan was created on another CUDA device"; return SRSCE_ERR_INVALID; }
  if (n_units == 0) return SRSCE_OK;
  const FillTcDev& f = p.fill_tc;
  const long long items = ((n_units + f.units - 1) / f.units) * f.n_tiles;
  // the tensor-memory allocation (all 512 columns) assumes one CTA per SM: shared memory is padded so that two never fit
  const int smem = f.smem_bytes > 116 * 1024 ? f.smem_bytes : 116 * 1024;
  const unsigned grid = unsigned(items < p.n_sms ? items : p.n_sms);
  pick_fill(f.n_layers)<<<grid, kFillThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      f, static_cast<const float2*>(h_pilots), static_cast<float2*>(out), (long long)n_units);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { err = std::string("fill_tc launch: ") + cudaGetErrorString(e); return SRSCE_ERR_CUDA; }
  *launches = 1;
  return SRSCE_OK;
}

}  // namespace srsce
