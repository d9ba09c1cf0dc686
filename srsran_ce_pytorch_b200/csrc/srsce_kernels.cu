// Fused per-slot port channel estimator for sm_100a.
//
// A persistent, warp-specialised CTA walks units (= reference calls: one received grid, one UE's pilots)
// through the whole reference path without leaving the SM; its estimator warps run phase A of unit k+1
// while its writer warps stream phase B of unit k (see the comment above estimate_kernel):
//
//   A. gather pilot REs + EPRE + LS de-spreading (y * conj(p))          ce_rule_tensorized.py:571-593
//      CFO from the first two pilot symbols, de-rotation                :357-451, :596-609
//      symbol average, CDM (OCC) pair average                           :613-628
//      smoothing: mean | raised-cosine FIR with virtual pilots | none   :633-668, :69-140, :459-493
//      time alignment: argmax of |IFFT_4096|^2 in +-144 samples         :674-698
//      reconstruction, noise, RSRP                                      :701-730
//      final normalisation                                              :899-919
//   B. frequency interpolation (lerp / CNN operator), replication over the allocated symbols,
//      final CFO rotation, and ONE streaming write of the dense [n_sc, n_sym, L] grid
//                                                                       :237-354, :922-929
//
// HBM traffic per unit is therefore the algorithmic minimum: pilot REs and pilots read once
// (the second touch for the noise estimate hits L1/L2), the grid written once.  Phase B is the
// bandwidth phase (>= 98 % of the bytes); phase A runs concurrently on the estimator warps.
//
// Precision islands follow the reference: complex64 data, float32 |.|^2 partial sums combined in
// float64, float32 angle; float64 for the virtual-pilot regression, phase unwrap, FIR
// accumulation, CFO phase tables and every scalar accumulator.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "srsce_internal.h"
#include "srsce_device.cuh"

namespace srsce {

namespace {

struct ColInfo {           // one store slot of an output row (VEC consecutive complex64 elements); 24 bytes
  float2 rot[2];           // final CFO rotation of each element's symbol
  uint8_t layer[2];
  uint8_t flags[2];        // bit0: symbol inside hop-1 allocation, bit1: inside hop-2 allocation
  uint8_t pad[4];
};

struct Misc {
  double red[kMaxWarpsA][8];
  double tot[8];
  double acc[8];           // 0 epre, 1 noise, 2 rsrp, 3 ta, 4 cfo, 5 cfo defined
  double ta_hop[2];
  double taps[SRSCE_MAX_TAPS + 1];
  LerpEntry lerp[2][kNRE];
  int32_t re_pos[2][kNRE];
  unsigned int r_magic, r_magic_pad;
  float cnn_coef[2][kNRE][kMaxFillTerms];
  int32_t cnn_idx0[2][kNRE];
  ColInfo col[2][SRSCE_MAX_SYMBOLS * SRSCE_MAX_LAYERS];   // one store-slot table per hall buffer
  int32_t re_inv[kNRE];                                   // RE position inside a PRB -> pilot ordinal of the LAST CDM group (-1: none)
  float2 w8[8];                                           // e^{+j 2 pi k / 8}: last (fused) stage of the TA transform
};
struct HopScratch {        // per hop, after Misc
  float2 rot_neg[SRSCE_MAX_SYMBOLS];
  float2 rot_pos[SRSCE_MAX_SYMBOLS];
  float prof[2 * SRSCE_TA_WINDOW];
  float vp_mod[SRSCE_MAX_LAYERS][2][12];   // modulus / phase of the edge pilots (virtual-pilot regression)
  float vp_ph[SRSCE_MAX_LAYERS][2][12];
};
static_assert(sizeof(HopScratch) <= kHopScratchBytes, "HopScratch does not fit its shared-memory slot");
static_assert(sizeof(Misc) <= kMiscBytes, "Misc does not fit its shared-memory slot");

// Profiling build only (make prof -> libsrsce_prof.so, used by tools/): ablation switches and per-section cycle
// counters.  The production library has neither: `dbg` is a compile-time 0 and PROF_MARK expands to nothing.
#ifdef SRSCE_PROFILING
constexpr int kProfSlots = 16;
__device__ long long* g_prof_buf = nullptr;       // [gridDim.x][kProfSlots] cycles, set by srsce_prof_set_buffer
// Attribute the cycles since the previous mark of this thread to section k.  The shared-memory read makes the
// thread really wait for a preceding bar.sync (BAR.SYNC.DEFER_BLOCKING lets the clock read run ahead otherwise).
#define PROF_DECL long long prof_acc[kProfSlots] = {0}; long long prof_t = clock64();
#define PROF_MARK(cond, k)                                                                          \
  do {                                                                                              \
    if (cond) {                                                                                     \
      unsigned prof_dummy;                                                                          \
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(prof_dummy) : "r"((unsigned)__cvta_generic_to_shared(smem)) : "memory"); \
      const long long prof_now = clock64();                                                         \
      prof_acc[k] += prof_now - prof_t + (prof_dummy & 0);                                          \
      prof_t = prof_now;                                                                            \
    }                                                                                               \
  } while (0)
#define PROF_FLUSH(cond, base)                                                                      \
  do {                                                                                              \
    if ((cond) && g_prof_buf != nullptr)                                                            \
      for (int k_ = 0; k_ < kProfSlots; ++k_)                                                       \
        if (prof_acc[k_]) g_prof_buf[(long long)blockIdx.x * kProfSlots + k_] = prof_acc[k_];       \
  } while (0)
#else
#define PROF_DECL
#define PROF_MARK(cond, k) do { } while (0)
#define PROF_FLUSH(cond, base) do { } while (0)
#endif

// Named barriers: 0 is __syncthreads (whole CTA); 1 joins the estimator warps only; 2 ("full") hands
// hall[b] + store-slot table b from the estimator warps to the writer warps, 4+b ("empty") hands them back.
constexpr int kBarA = 1, kBarFull = 2, kBarEmpty = 4;   // + buffer index (0/1)
constexpr int kModeUnstaged = 0, kModeStaged = 1, kModeGlobalH = 2;   // estimate_kernel's MODE
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
template <int NA>
__device__ __forceinline__ void sync_a() { bar_sync(kBarA, NA); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum K per-thread values over the CTA; every thread gets the totals in m.tot[0..K).
template <int NA, int K>
__device__ __forceinline__ void block_sum(Misc& m, double (&v)[K]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const double s = warp_sum(v[k]);
    if (lane == 0) m.red[warp][k] = s;
  }
  sync_a<NA>();
  if (threadIdx.x < K) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < NA / 32; ++w) s += m.red[w][threadIdx.x];
    m.tot[threadIdx.x] = s;
  }
  sync_a<NA>();
}

// The pilots of the two layers of a CDM group are adjacent in memory ([pilot][symbol][layer]).  A thread owns a pilot, so
// its loads are 128 bytes apart from its neighbours': every load instruction touches 32 cache lines whatever its width,
// and with an even layer count one 16-byte load (both layers) halves the L1 wavefronts of the pilot passes.
__device__ __forceinline__ void ld_pilot_pair(const float2* __restrict__ pp, bool vec, bool two, float2& a, float2& b) {
  if (vec) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(pp));
    a = make_float2(v.x, v.y);
    b = make_float2(v.z, v.w);
  } else {
    a = __ldg(pp);
    b = two ? __ldg(pp + 1) : make_float2(0.f, 0.f);
  }
}

// Subcarrier of pilot i of CDM group g inside a hop band.
// (i / R as one IMAD.HI: r_magic = ceil(2^32 / R) is exact for i < 2^16, 2 <= R <= 12; R == 1 has no 32-bit magic.
//  Written with __umulhi on purpose: the 64-bit form `(u64(i) * magic) >> 32` followed by `q * R` was mis-folded by
//  ptxas 12.9 in one instantiation -- it reassociated to i * hi32(magic * R) -- and faulted)
__device__ __forceinline__ int pilot_sc(const Misc& m, int band_sc0, int R, int g, int i) {
  const int q = (R == 1) ? i : int(__umulhi(static_cast<unsigned>(i), m.r_magic));
  return band_sc0 + q * kNRE + m.re_pos[g][i - q * R];
}

// create_virtual_pilots (:69-140), second half: given modulus and (float32) phase of the nv pilots at one
// edge of one layer (k counted from the band edge inwards), return the virtual pilot at abscissa k_out - nv.
// Straight-line fits of modulus and unwrapped phase in float64 (:105-138); unwrap as _unwrap_1d (:35-66).
__device__ float2 virtual_pilot(int nv, double mx, double den, const float* mod, const float* ph, int k_out) {
  if (nv == 1) {                                    // :95-101 constant extrapolation
    float s, c;
    sincosf(ph[0], &s, &c);
    return make_float2(mod[0] * c, mod[0] * s);
  }
  double sy_m = 0.0, sxy_m = 0.0, sy_p = 0.0, sxy_p = 0.0, prev = 0.0, corr = 0.0;
  const double two_pi = 2.0 * M_PI;
  for (int k = 0; k < nv; ++k) {
    const double a = double(mod[k]), p = double(ph[k]);
    if (k > 0) {
      const double d = p - prev;
      if (fabs(d) >= M_PI) {
        double dm = (d + M_PI) - two_pi * floor((d + M_PI) / two_pi) - M_PI;   // floored modulo, as torch.remainder
        if (dm == -M_PI && d > 0.0) dm += two_pi;
        corr += dm - d;
      }
    }
    prev = p;
    const double pu = p + corr;
    sy_m += a;
    sxy_m += double(k) * a;
    sy_p += pu;
    sxy_p += double(k) * pu;
  }
  const double n = double(nv);
  const double my_m = sy_m / n, my_p = sy_p / n;
  const double a_m = (sxy_m - n * mx * my_m) / den, b_m = my_m - a_m * mx;
  const double a_p = (sxy_p - n * mx * my_p) / den, b_p = my_p - a_p * mx;
  const double x = double(k_out - nv);
  double sn, cs;
  sincos(a_p * x + b_p, &sn, &cs);
  const double amp = a_m * x + b_m;
  return make_float2(float(amp * cs), float(amp * sn));
}

// ---------------------------------------------------------------------------------------------------------------
// Time-alignment transform (:674-698): inverse (e^{+j...}) decimation-in-frequency FFT of `nl` vectors of length
// M = 512 * RF (RF = 2, 4, 8 <=> M = 1024, 2048, 4096) as three register-blocked radix-8 passes at strides M/8, M/64
// and RF; the last stage (radix RF on RF consecutive elements) is not run as a pass: only 288 bins are ever looked
// at, so each of them is evaluated directly from its RF inputs when the delay profile is formed.
//   pass 1 reads the pilots straight from `hall` (zero padded / zero stuffed on the fly): no separate fill pass
//   twiddles are computed (one sincospif of an exact argument per task, powers by multiplication), not loaded
// Every element therefore makes three shared-memory round trips (the radix-4 version made 6-7) with 8-point
// butterflies in registers.  Vectors live in `ext` with one pad element per 8 (index i at i + i / 8) so that pass 3's
// short strides fall on distinct banks.  After the passes, bin k = d1 + 8 d2 + 64 d3 + 512 d4 is
//   sum_{m < RF} x[d1 * M/8 + d2 * M/64 + d3 * RF + m] * e^{+j 2 pi m d4 / RF}.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int fft_pad(int i) { return i + (i >> 3); }

// 4-point inverse DFT in place: y[r] = sum_m a[m] * (+j)^(m r)
__device__ __forceinline__ void bf4(float2& a0, float2& a1, float2& a2, float2& a3) {
  const float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3);
  const float2 d = csub(a1, a3);
  const float2 t3 = make_float2(-d.y, d.x);              // +j * (a1 - a3)
  a0 = cadd(t0, t2);
  a1 = cadd(t1, t3);
  a2 = csub(t0, t2);
  a3 = csub(t1, t3);
}
// 8-point inverse DFT in registers, natural order: radix 2 x radix 4.
__device__ __forceinline__ void fft8(float2 (&x)[8]) {
  constexpr float h = 0.70710678118654752f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 u = cadd(x[j], x[j + 4]), v = csub(x[j], x[j + 4]);
    x[j] = u;
    x[j + 4] = v;
  }
  x[5] = cmul(x[5], make_float2(h, h));        // w8^1
  x[6] = make_float2(-x[6].y, x[6].x);         // w8^2 = +j
  x[7] = cmul(x[7], make_float2(-h, h));       // w8^3
  bf4(x[0], x[1], x[2], x[3]);                 // -> y[0], y[2], y[4], y[6]
  bf4(x[4], x[5], x[6], x[7]);                 // -> y[1], y[3], y[5], y[7]
  const float2 y1 = x[4], y2 = x[1], y3 = x[5], y4 = x[2], y5 = x[6], y6 = x[3];
  x[1] = y1; x[2] = y2; x[3] = y3; x[4] = y4; x[5] = y5; x[6] = y6;
}


// One radix-8 task: x[m] (already loaded, natural order) -> butterfly -> twiddle by w^(r), w = e^{+j pi two_j_over_n}
__device__ __forceinline__ void fft8_twiddled(float2 (&x)[8], float two_j_over_n) {
  fft8(x);
  float sn, cs;
  sincospif(two_j_over_n, &sn, &cs);
  const float2 w1 = make_float2(cs, sn), w2 = cmul(w1, w1), w3 = cmul(w2, w1), w4 = cmul(w2, w2);
  x[1] = cmul(x[1], w1);
  x[2] = cmul(x[2], w2);
  x[3] = cmul(x[3], w3);
  x[4] = cmul(x[4], w4);
  x[5] = cmul(cmul(x[5], w4), w1);
  x[6] = cmul(cmul(x[6], w4), w2);
  x[7] = cmul(cmul(x[7], w4), w3);
}

// Passes 2 and 3 over `nl` padded vectors of length M (Mp = M + M / 8 elements apart) in s[].
template <int NA>
__device__ __forceinline__ void fft_pass23(float2* s, int nl, int log2M) {
  const int tid = threadIdx.x, Mp = (1 << log2M) + (1 << (log2M - 3));
  const int log2t = log2M - 3;                                    // tasks per vector in every pass = M / 8
#pragma unroll 1
  for (int pass = 2; pass <= 3; ++pass) {
    const int log2q = log2M - 3 * pass;                           // stride of the pass: M / 64, then RF = M / 512
    for (int t = tid; t < (nl << log2t); t += NA) {
      const int l = t >> log2t, rem = t & ((1 << log2t) - 1);
      const int blk = rem >> log2q, j = rem & ((1 << log2q) - 1);
      float2* const v = s + l * Mp;
      const int i0 = (blk << (log2q + 3)) + j;
      float2 x[8];
#pragma unroll
      for (int m = 0; m < 8; ++m) x[m] = v[fft_pad(i0 + (m << log2q))];
      fft8_twiddled(x, float(2 * j) / float(8 << log2q));         // w = e^{+j 2 pi j / n}, n = 8 q (exact in float)
#pragma unroll
      for (int r = 0; r < 8; ++r) v[fft_pad(i0 + (r << log2q))] = x[r];
    }
    sync_a<NA>();
  }
}

// ---- the radix-4 transform of round 1, kept for the plans whose phase A is light (see PlanDev::ta_radix4)
// Where output bin k of the in-place decimation-in-frequency transform of length M ends up.
__device__ __forceinline__ int dif_position(int k, int M) {
  int pos = 0, n = M;
  while (n >= 4) {
    const int q = n >> 2;
    pos += (k & 3) * q;
    k >>= 2;
    n = q;
  }
  if (n == 2) pos += k;
  return pos;
}

// In-place inverse (e^{+j...}) DIF FFT of `nl` length-M vectors stored back to back in s[]:
// radix-4 passes, one radix-2 pass if M is an odd power of two.  Output in digit-reversed order.
template <int NA>
__device__ void fft_dif_inverse(float2* s, int nl, int M, int log2M, const float2* __restrict__ tw) {
  const int tid = threadIdx.x;
  int n = M, log2n = log2M;
  while (n >= 4) {
    const int q = n >> 2, log2q = log2n - 2;
    const int step = M >> log2n;                    // twiddle stride: w_n^j = tw[j * M / n], tw[k] = exp(+j 2 pi k / M)
    const int per_vec = M >> 2;
    for (int b = tid; b < nl * per_vec; b += NA) {
      const int l = b >> (log2M - 2), r = b & (per_vec - 1);
      const int blk = r >> log2q, j = r & (q - 1);
      float2* p = s + l * M + (blk << log2n) + j;
      const float2 a0 = p[0], a1 = p[q], a2 = p[2 * q], a3 = p[3 * q];
      const float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3);
      const float2 d = csub(a1, a3);
      const float2 t3 = make_float2(-d.y, d.x);     // +j * (a1 - a3)
      const float2 w1 = __ldg(tw + j * step), w2 = __ldg(tw + 2 * j * step), w3 = __ldg(tw + 3 * j * step);
      p[0] = cadd(t0, t2);
      p[q] = cmul(cadd(t1, t3), w1);
      p[2 * q] = cmul(csub(t0, t2), w2);
      p[3 * q] = cmul(csub(t1, t3), w3);
    }
    sync_a<NA>();
    n = q;
    log2n = log2q;
  }
  if (n == 2) {
    for (int b = tid; b < nl * (M >> 1); b += NA) {
      float2* p = s + 2 * b;
      const float2 a0 = p[0], a1 = p[1];
      p[0] = cadd(a0, a1);
      p[1] = csub(a0, a1);
    }
    sync_a<NA>();
  }
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
  // .L2::64B: each pilot RE is alone in its 128-byte line (reference layout), so ask L2 for the smallest fill it
  // offers -- measured on B200: DRAM read traffic of the gather halves (225 -> 120 KB per unit at config 2)
  asm volatile("cp.async.ca.shared.global.L2::64B [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// The same for two hops: rows inside hop 2's band take hop 2's value on the symbols of hop 2's allocation
// (it wins where the hops overlap), otherwise hop 1's on hop 1's allocated symbols, otherwise zero.
template <int LL, bool ROT>
__device__ __forceinline__ void stream_full_tile_2hop(const float4* __restrict__ src1, const float4* __restrict__ src2,
                                                      float4* __restrict__ dst, int flags, unsigned in1, unsigned in2,
                                                      int first_row, float2 rot0, float2 rot1) {
  constexpr int kCols = 14 * LL / 2, kRowsPerPass = 32 / kCols, kPasses = kTileRows / kRowsPerPass;
  const unsigned use1 = (flags & 1) ? (in1 >> first_row) : 0u, use2 = (flags & 2) ? (in2 >> first_row) : 0u;
#pragma unroll
  for (int p = 0; p < kPasses; ++p) {
    const int r = p * kRowsPerPass;                      // row of this lane = first_row + r
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if ((use2 >> r) & 1u) v = src2[p * (kRowsPerPass * LL / 2)];
    else if ((use1 >> r) & 1u) v = src1[p * (kRowsPerPass * LL / 2)];
    if (ROT) {
      const float2 a = cmul(make_float2(v.x, v.y), rot0), b = cmul(make_float2(v.z, v.w), rot1);
      v = make_float4(a.x, a.y, b.x, b.y);
    }
    __stcs(dst + p * (kRowsPerPass * kCols), v);
  }
}

// Warp-specialised, software-pipelined CTA.  Per unit:
//   loader   (estimator threads, one unit ahead)  cp.async gather of the pilot REs of the received grid
//            into ystage[next]; the pilots themselves are warmed in L2
//   estimator warps (256 threads)                 phase A out of shared memory -> hall + store-slot table
//   writer warps (128 threads)                    phase B: interpolate + stream the dense grid to HBM
// so that the store stream of unit k overlaps the arithmetic of unit k+1 and the loads of unit k+2.
// Shared-memory layouts (all complex64, layer major so that every per-layer pass is unit stride):
//   ystage[2][G*N*S_total]     received pilot REs, [hop][group][pilot][symbol], double buffered
//   hall[hop][layer][N]        smoothed pilots of each hop: written by the FIR, read by the writer warps
//   ext[layer][EF]             raw pilots with nv virtual pilots and `pad` zeros either side (FIR input),
//                              EF = N + 2*nv + 2*pad; later the FFT scratch
//   tiles[writer warp][hop][32][L]  interpolated rows being streamed out
template <int NA, int VEC, bool TWO_HOPS, bool HAS_ROT, int MODE>
__global__ void __launch_bounds__(cta_threads(NA, TWO_HOPS), NA == 256 ? 2 : 1)
estimate_kernel(const __grid_constant__ PlanDev P, const float2* __restrict__ rx, long long su, long long ssc,
                long long ssym, const float2* __restrict__ pilots, int pilot_share, double beta_d,
                float2* __restrict__ out, double* __restrict__ scalars, long long n_units, int dbg_arg,
                float2* __restrict__ out_c, float2* __restrict__ rot_c, const long long* __restrict__ rx_index,
                const float2* __restrict__ h_in, float2* h_scratch) {
  // MODE 1: the received pilot REs are staged in shared memory; 0: read where they are used (staging buffer does not
  // fit); 2: as 0, and the smoothed pilots live in a per-CTA global scratch (L2 resident) instead of shared memory --
  // the largest two-hop geometries (dense mask on two full-band hops), where 2 * N * L * 8 bytes alone fill the SM
  constexpr bool STAGED = MODE == kModeStaged, H_GLOBAL = MODE == kModeGlobalH;
#ifdef SRSCE_PROFILING
  const int dbg = dbg_arg;                                  // ablation bits (profiling build only)
#else
  constexpr int dbg = 0;
#endif
  constexpr int kThreads = NA;                              // estimator threads
  constexpr int kWarpsB = (NA == 256) ? (TWO_HOPS ? 6 : 4) : (TWO_HOPS ? 8 : 6);   // = writer_warps(NA, TWO_HOPS)
  constexpr int kCtaThreads = NA + 32 * kWarpsB;
  extern __shared__ __align__(16) unsigned char smem[];
  float2* const hall0 = H_GLOBAL ? h_scratch + (long long)blockIdx.x * (P.hall_bufs * (TWO_HOPS ? 2 : 1) * P.n_layers * P.n_re)
                                 : reinterpret_cast<float2*>(smem + P.off_h);   // hall_bufs buffers of n_hops*L*N
  float2* const ystage = reinterpret_cast<float2*>(smem + P.off_y);
  float2* const tiles = reinterpret_cast<float2*>(smem + P.off_tile);
  float2* const ext = reinterpret_cast<float2*>(smem + P.off_ext);
  Misc& m = *reinterpret_cast<Misc*>(smem + P.off_misc);
  HopScratch* const hsx = reinterpret_cast<HopScratch*>(smem + P.off_misc + kMiscBytes);   // [n_hops]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int L = P.n_layers, G = P.n_cdm, R = P.re_per_prb, N = P.n_re, ST = P.n_dmrs_total;
  const int nv = P.n_virtual, K = P.n_taps, pad = K >> 1;
  const int EF = N + 2 * nv + 2 * pad;
  const int n_hops = TWO_HOPS ? 2 : 1;
  const float beta = float(beta_d);
  const float inv_beta = 1.0f / beta;
  const int n_sc = P.n_sc, row_elems = P.row_elems, cols = P.cols;
  const int ylen = G * N * ST;
  const bool filt = P.smoothing == SRSCE_SMOOTH_FILTER;
  const bool pvec = (L & 1) == 0;                          // 16-byte pilot loads (the ABI asks for 16-byte aligned pilots then)

  // plan tables that are indexed per lane go to shared memory once per CTA
  for (int k = tid; k < SRSCE_MAX_TAPS; k += kThreads) m.taps[k] = P.taps[k];
  for (int k = tid; k < 2 * kNRE; k += kThreads) {
    m.lerp[k / kNRE][k % kNRE] = P.lerp[k / kNRE][k % kNRE];
    m.re_pos[k / kNRE][k % kNRE] = P.re_pos[k / kNRE][k % kNRE];
    m.cnn_idx0[k / kNRE][k % kNRE] = P.cnn_idx0[k / kNRE][k % kNRE];
  }
  if (tid < 8) {
    float sn, cs;
    sincospif(0.25f * float(tid), &sn, &cs);
    m.w8[tid] = make_float2(cs, sn);
  }
  if (tid < kNRE) {
    int r = -1;
    for (int k = 0; k < R; ++k)
      if (P.re_pos[G - 1][k] == tid) r = k;
    m.re_inv[tid] = r;
  }
  if (tid == 0) m.r_magic = (R == 1) ? 0u : static_cast<unsigned>(((1ull << 32) + static_cast<unsigned>(R) - 1) / static_cast<unsigned>(R));
  for (int k = tid; k < 2 * kNRE * kMaxFillTerms; k += kThreads)
    (&m.cnn_coef[0][0][0])[k] = (&P.cnn_coef[0][0][0])[k];
  __syncthreads();

  if (tid < kThreads) {
  // ================================ estimator warps ==========================================
  PROF_DECL
  // gather the pilot REs of unit `un` into a stage buffer (8-byte async copies; nothing waits here)
  auto stage_unit = [&](long long un, float2* dst) {
    if (!STAGED || un >= n_units || (dbg & 4) || h_in != nullptr) return;   // (unstaged fallback / ablation bit 2 / fill-only launch)
    const float2* rxn = rx + (rx_index ? __ldg(rx_index + un) : un) * su;     // several units may share one received grid
    for (int hh = 0; hh < n_hops; ++hh) {
      const HopDev& H = P.hop[hh];
      float2* const d0 = dst + G * N * H.pilot_col0;
      // neighbouring threads fetch neighbouring subcarriers (the CDM groups interleave in frequency), so that the
      // 64-byte fills of one warp walk the grid in address order
      for (int k = tid; k < G * N; k += kThreads) {
        const int g = (G == 2) ? (k & 1) : 0, i = (G == 2) ? (k >> 1) : k;
        const float2* src = rxn + (long long)pilot_sc(m, H.band_sc0, R, g, i) * ssc;
        float2* const d = d0 + (g * N + i) * H.n_dmrs;
        for (int s = 0; s < H.n_dmrs; ++s) cp_async8(d + s, src + (long long)H.sym_idx[s] * ssym);
      }
    }
    const char* pn = reinterpret_cast<const char*>(pilots + (un / pilot_share) * ((long long)N * ST * L));
    const int pil_lines = (N * ST * L * 8 + 127) >> 7;
    for (int k = tid; k < pil_lines; k += kThreads) prefetch_l2(pn + ((long long)k << 7));
  };
  // store-slot table of the writer warps: layer / allocation flags / final CFO rotation of every slot
  auto build_store_table = [&](ColInfo* colw) {
    for (int c = tid; c < cols; c += kThreads) {
      ColInfo ci;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int el = min(c * VEC + e, row_elems - 1);
        const int sym = el / L;
        ci.layer[e] = uint8_t(el - sym * L);
        int f = (sym >= P.hop[0].start_symbol && sym < P.hop[0].end_symbol) ? 1 : 0;
        if (TWO_HOPS && sym >= P.hop[1].start_symbol && sym < P.hop[1].end_symbol) f |= 2;
        ci.flags[e] = uint8_t(f);
        double sn = 0.0, cs = 1.0;
        if (HAS_ROT) sincos(2.0 * M_PI * P.t_sym[sym] * m.acc[4], &sn, &cs);   // :922-929
        ci.rot[e] = make_float2(float(cs), float(sn));
      }
      colw[c] = ci;
    }
  };
  // compact output: the per-symbol final CFO rotation goes out as a [n_sym] vector instead of being multiplied in
  auto write_symbol_rotation = [&](long long unit) {
    if (rot_c != nullptr && tid < P.n_sym) {
      double sn = 0.0, cs = 1.0;
      if (HAS_ROT) sincos(2.0 * M_PI * P.t_sym[tid] * m.acc[4], &sn, &cs);
      rot_c[unit * P.n_sym + tid] = make_float2(float(cs), float(sn));
    }
  };
  if (h_in != nullptr) {
    // fill-only launch (srsce_fill, direct implementation): the smoothed pilots [hop][N][L] of every unit are an
    // input; transpose them into hall ([hop][layer][N]) and hand them to the writer warps -- phase B alone, i.e.
    // fill_ch_est_cdm (:237-354 / ce_dl_cnn.py:233-322) for every CDM group of every hop
    int itf = 0;
    for (long long u = blockIdx.x; u < n_units; u += gridDim.x, ++itf) {
      const int hb = (P.hall_bufs == 2) ? (itf & 1) : 0;
      float2* const hdst = hall0 + hb * (n_hops * L * N);
      if (itf >= P.hall_bufs) bar_sync(kBarEmpty + hb, kCtaThreads);
      const float2* __restrict__ hu = h_in + u * ((long long)n_hops * N * L);
      for (int idx = tid; idx < n_hops * N * L; idx += kThreads) {
        const int hh = idx / (N * L), rem = idx - hh * (N * L), i = rem / L, l = rem - i * L;
        hdst[hh * L * N + l * N + i] = __ldg(hu + idx);
      }
      build_store_table(m.col[hb]);
      bar_arrive(kBarFull + hb, kCtaThreads);
    }
  } else {
  stage_unit(blockIdx.x, ystage);
  cp_async_commit();

  int it = 0;
  for (long long u = blockIdx.x; u < n_units; u += gridDim.x, ++it) {
    const float2* __restrict__ pu = pilots + (u / pilot_share) * ((long long)N * ST * L);
    cp_async_wait_all();                                      // this unit's pilot REs have landed ...
    sync_a<NA>();                                                 // ... for every estimator thread
    PROF_MARK(tid == 0, 0);
    const float2* const ys = ystage + ((P.y_bufs == 2) ? (it & 1) * ylen : 0);
    // received pilot RE (hop H, group g, pilot i, pilot symbol s): from the staged copy, or -- the fallback for
    // geometries whose staging buffer does not fit beside the rest (STAGED = false, a separate instantiation so
    // that the staged variants do not carry its address arithmetic) -- straight from the received grid
    const float2* const rxu = STAGED ? rx : rx + (rx_index ? __ldg(rx_index + u) : u) * su;
    auto ld_y = [&](const HopDev& H, int g, int i, int s) -> float2 {
      if (STAGED) return ys[G * N * H.pilot_col0 + (g * N + i) * H.n_dmrs + s];
      return __ldg(rxu + (long long)pilot_sc(m, H.band_sc0, R, g, i) * ssc + (long long)H.sym_idx[s] * ssym);
    };
    const int hb_idx = (P.hall_bufs == 2) ? (it & 1) : 0;
    float2* const hall = hall0 + hb_idx * (n_hops * L * N);
    const bool wait_empty = it >= P.hall_bufs;                 // the writers last used this buffer hall_bufs units ago
    if (P.y_bufs == 2) {                                       // double buffered: fetch the next unit a whole unit ahead
      stage_unit(u + gridDim.x, ystage + ((it + 1) & 1) * ylen);
      cp_async_commit();
    }
    if (tid < 8) m.acc[tid] = 0.0;
    // (visibility of m.acc is ordered by the barriers below)

    // ===================================== phase A ==========================================
    // Step major: every step handles both hops before the next barrier, so that a two-hop unit pays the
    // barrier / latency chain of phase A once, not twice.
    if (!(dbg & 1)) {
      // ---- CFO: correlate the first two pilot symbols of each hop (compensate_cfo :357-451)
#pragma unroll
      for (int hh = 0; hh < n_hops; ++hh) {
        const HopDev& H = P.hop[hh];
        const int S = H.n_dmrs, col0 = H.pilot_col0;
        if (S < 2) continue;
        float2 c[2][2] = {{{0.f, 0.f}, {0.f, 0.f}}, {{0.f, 0.f}, {0.f, 0.f}}};
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (g < G) {
            for (int i = tid; i < N; i += kThreads) {
              const float2 y0 = ld_y(H, g, i, 0), y1 = ld_y(H, g, i, 1);
              const float2* pp = pu + ((long long)i * ST + col0) * L + 2 * g;
              float2 p0[2], p1[2];
              ld_pilot_pair(pp, pvec, 2 * g + 1 < L, p0[0], p0[1]);
              ld_pilot_pair(pp + L, pvec, 2 * g + 1 < L, p1[0], p1[1]);
#pragma unroll
              for (int k = 0; k < 2; ++k) {
                if (2 * g + k < L) {
                  const float2 x0 = cmul_conj(y0, p0[k]), x1 = cmul_conj(y1, p1[k]);
                  c[g][k] = cadd(c[g][k], cmul_conj(x1, x0));      // conj(x0) * x1
                }
              }
            }
          }
        }
        double v[8] = {c[0][0].x, c[0][0].y, c[0][1].x, c[0][1].y, c[1][0].x, c[1][0].y, c[1][1].x, c[1][1].y};
        block_sum<NA, 8>(m, v);
        if (tid == 0) {
          double tot = 0.0;
          for (int k = 0; k + 1 < L; k += 2) {
            const float re = float(m.tot[2 * k]) + float(m.tot[2 * k + 2]);
            const float im = float(m.tot[2 * k + 1]) + float(m.tot[2 * k + 3]);
            tot += double(atan2f(im, re));
          }
          if (L & 1) tot += double(atan2f(float(m.tot[2 * (L - 1) + 1]), float(m.tot[2 * (L - 1)])));
          const double cfo_hop = tot / H.cfo_den;
          m.acc[4] = (m.acc[5] != 0.0) ? 0.5 * (m.acc[4] + cfo_hop) : cfo_hop;   // :605-609
          m.acc[5] = 1.0;
          const bool derot = P.cfo_compensate != 0;
          for (int s = 0; s < S; ++s) {
            double sn = 0.0, cs = 1.0;
            if (derot) sincos(2.0 * M_PI * H.t_dmrs[s] * cfo_hop, &sn, &cs);
            hsx[hh].rot_pos[s] = make_float2(float(cs), float(sn));
            hsx[hh].rot_neg[s] = make_float2(float(cs), float(-sn));
          }
        }
      }
      if (P.cfo_defined) sync_a<NA>();
      PROF_MARK(tid == 0, 1);

      // ---- LS estimate, symbol average, CDM pair average (:593, :613-628) + EPRE (:584), then smoothing
      //      (:633-668).  With the raised-cosine filter the raw pilots go through ext (FIR input, one region
      //      per hop); when the planner asked for it, one CDM group (<= 2 layers) at a time so that ext stays small.
      float epre_part[2] = {0.f, 0.f};
      if (!filt && wait_empty) bar_sync(kBarEmpty + hb_idx, kCtaThreads);   // LS writes hall: writers must be done with it
      PROF_MARK(tid == 0 && !filt, 4);
      const int n_batches = (filt && P.fir_per_group) ? G : 1;
      for (int bt = 0; bt < n_batches; ++bt) {
        const int g_lo = (n_batches == 1) ? 0 : bt, g_hi = (n_batches == 1) ? G : bt + 1;
        const int l_lo = 2 * g_lo, LB = min(L, 2 * g_hi) - l_lo;          // layers of this batch: l_lo .. l_lo + LB - 1
#pragma unroll
        for (int hh = 0; hh < n_hops; ++hh) {
          const HopDev& H = P.hop[hh];
          const int S = H.n_dmrs, col0 = H.pilot_col0;
          const bool derot = S >= 2 && P.cfo_compensate;
          float2* const exth = ext + hh * P.ext_hop;
          float2* const dst = filt ? exth + pad + nv : hall + hh * L * N;   // filt: layer index relative to the batch
          const int dstride = filt ? EF : N;
          if (filt) {                                  // zero margins of the FIR input
            for (int k = tid; k < 2 * pad * LB; k += kThreads) {
              const int l = k / (2 * pad), j = k - l * 2 * pad;
              exth[l * EF + (j < pad ? j : EF - 2 * pad + j)] = make_float2(0.f, 0.f);
            }
          }
          const float inv_s = 1.0f / float(S);
          const int n_pairs = (N + 1) >> 1;
          for (int g = g_lo; g < g_hi; ++g) {
            const bool two = 2 * g + 1 < L;
            const int lrel = filt ? 2 * g - l_lo : 2 * g;
            for (int pr = tid; pr < n_pairs; pr += kThreads) {
              float2 h[2][2];
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int i = 2 * pr + e;
                float2 a0 = make_float2(0.f, 0.f), a1 = a0;
                if (i < N) {
                  const float2* pp = pu + ((long long)i * ST + col0) * L + 2 * g;
                  for (int s = 0; s < S; ++s) {
                    const float2 y = ld_y(H, g, i, s);
                    float2 pa, pb;
                    ld_pilot_pair(pp + s * L, pvec, two, pa, pb);
                    epre_part[hh] += y.x * y.x + y.y * y.y;
                    float2 x0 = cmul_conj(y, pa);
                    if (derot) x0 = cmul(x0, hsx[hh].rot_neg[s]);
                    a0 = cadd(a0, x0);
                    if (two) {
                      float2 x1 = cmul_conj(y, pb);
                      if (derot) x1 = cmul(x1, hsx[hh].rot_neg[s]);
                      a1 = cadd(a1, x1);
                    }
                  }
                }
                h[e][0] = cscale(cscale(a0, inv_beta), inv_s);
                h[e][1] = cscale(cscale(a1, inv_beta), inv_s);
              }
              const int i0 = 2 * pr;
              if (L >= 2 && i0 + 1 < N) {
                h[0][0] = h[1][0] = cscale(cadd(h[0][0], h[1][0]), 0.5f);
                h[0][1] = h[1][1] = cscale(cadd(h[0][1], h[1][1]), 0.5f);
              }
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int i = i0 + e;
                if (i < N) {
                  dst[lrel * dstride + i] = h[e][0];
                  if (two) dst[(lrel + 1) * dstride + i] = h[e][1];
                }
              }
            }
          }
        }
        sync_a<NA>();
        PROF_MARK(tid == 0, 2);

        if (filt) {
          if (nv > 0) {
            // one thread per (hop, layer, edge, pilot): modulus / phase first, then every thread fits the two
            // lines of its edge and evaluates its own virtual pilot
            const int per_hop = 2 * LB * nv;
            const int hh = tid / per_hop, t = tid - hh * per_hop;
            const int l = t / (2 * nv), r = t - l * 2 * nv, end = r / nv, k = r - end * nv;
            const bool mine_vp = tid < n_hops * per_hop;
            float2* const x = ext + (mine_vp ? hh : 0) * P.ext_hop + l * EF + pad;   // x[0..nv) virtual | x[nv..nv+N) real | virtual
            HopScratch& hs = hsx[mine_vp ? hh : 0];
            if (mine_vp) {
              const float2 e = end ? x[nv + N - 1 - k] : x[nv + k];
              hs.vp_mod[l][end][k] = sqrtf(e.x * e.x + e.y * e.y);
              hs.vp_ph[l][end][k] = atan2f(e.y, e.x);
            }
            sync_a<NA>();
            if (mine_vp) {
              const float2 val = virtual_pilot(nv, P.vp_mx, P.vp_den, hs.vp_mod[l][end], hs.vp_ph[l][end], k);
              if (end) x[nv + N + (nv - 1 - k)] = val;
              else x[k] = val;
            }
          }
          sync_a<NA>();
          PROF_MARK(tid == 0, 3);
          if (bt == 0 && wait_empty) bar_sync(kBarEmpty + hb_idx, kCtaThreads);   // the FIR overwrites hall: writers must be done with it
          PROF_MARK(tid == 0, 4);
          // conv(ext, taps, "same") cropped to the real pilots, float64 accumulation (:459-493);
          // two adjacent outputs per thread share all but one of their input samples
#pragma unroll
          for (int hh = 0; hh < n_hops; ++hh) {
            for (int l = 0; l < LB; ++l) {
              const float2* const xl = ext + hh * P.ext_hop + l * EF + nv + 2 * pad;    // tap k of output i multiplies xl[i - k]
              float2* const hl = hall + hh * L * N + (l_lo + l) * N;
              for (int i = 2 * tid; i < N; i += 2 * kThreads) {
                double ar0 = 0.0, ai0 = 0.0, ar1 = 0.0, ai1 = 0.0;
                const float2 t1 = (i + 1 < N) ? xl[i + 1] : make_float2(0.f, 0.f);
                double tx1 = double(t1.x), ty1 = double(t1.y);
                for (int k = 0; k < K; ++k) {
                  const float2 t0 = xl[i - k];
                  const double tx0 = double(t0.x), ty0 = double(t0.y), w = m.taps[k];
                  ar0 += w * tx0;
                  ai0 += w * ty0;
                  ar1 += w * tx1;
                  ai1 += w * ty1;
                  tx1 = tx0;
                  ty1 = ty0;
                }
                hl[i] = make_float2(float(ar0), float(ai0));
                if (i + 1 < N) hl[i + 1] = make_float2(float(ar1), float(ai1));
              }
            }
          }
          sync_a<NA>();
          if (P.blend && N > 2) {                      // ce_dl_cnn.py:712-717
#pragma unroll
            for (int hh = 0; hh < n_hops; ++hh) {
              float2* const hb0 = hall + hh * L * N + l_lo * N;
              float2* const sc0 = ext + hh * P.ext_hop;
              for (int idx = tid; idx < N * LB; idx += kThreads) {
                const int l = idx / N, i = idx - l * N;
                const int il = (i == 0) ? 1 : i - 1, ir = (i == N - 1) ? N - 2 : i + 1;
                const float2 a = hb0[l * N + il], b = hb0[idx], c = hb0[l * N + ir];
                sc0[idx] = make_float2(float(0.25 * double(a.x) + 0.5 * double(b.x) + 0.25 * double(c.x)),
                                       float(0.25 * double(a.y) + 0.5 * double(b.y) + 0.25 * double(c.y)));
              }
            }
            sync_a<NA>();
#pragma unroll
            for (int hh = 0; hh < n_hops; ++hh) {
              float2* const hb0 = hall + hh * L * N + l_lo * N;
              const float2* const sc0 = ext + hh * P.ext_hop;
              for (int idx = tid; idx < N * LB; idx += kThreads) {
                const float2 s0 = hb0[idx], lp = sc0[idx];
                hb0[idx] = make_float2(s0.x + P.cnn_alpha * (lp.x - s0.x), s0.y + P.cnn_alpha * (lp.y - s0.y));
              }
            }
            sync_a<NA>();
          }
          PROF_MARK(tid == 0, 5);
        }
      }  // group batches

      if (P.smoothing == SRSCE_SMOOTH_MEAN) {
#pragma unroll
        for (int hh = 0; hh < n_hops; ++hh) {
          float2* const hbuf = hall + hh * L * N;
          double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
          for (int l = 0; l < SRSCE_MAX_LAYERS; ++l) {
            if (l < L) {
              float sx = 0.f, sy = 0.f;
              for (int i = tid; i < N; i += kThreads) {
                const float2 t = hbuf[l * N + i];
                sx += t.x;
                sy += t.y;
              }
              v[2 * l] = sx;
              v[2 * l + 1] = sy;
            }
          }
          block_sum<NA, 8>(m, v);
          const float inv_n = 1.0f / float(N);
          for (int idx = tid; idx < N * L; idx += kThreads) {
            const int l = idx / N;
            hbuf[idx] = make_float2(float(m.tot[2 * l]) * inv_n, float(m.tot[2 * l + 1]) * inv_n);
          }
          sync_a<NA>();
        }
      }

      // ---- the smoothed pilots of every hop are final: build the store-slot table (final CFO
      //      rotation :922-929, allocation flags) and release the writer warps; the noise estimate and
      //      the time alignment below only read hall, concurrently with them
      PROF_MARK(tid == 0, 6);
      build_store_table(m.col[hb_idx]);
      write_symbol_rotation(u);
      bar_arrive(kBarFull + hb_idx, kCtaThreads);
      PROF_MARK(tid == 0, 7);

      // ---- reconstruction, noise, RSRP (:701-730)
      {
        double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};                    // per hop: epre, noise, rsrp
#pragma unroll
        for (int hh = 0; hh < n_hops; ++hh) {
          const HopDev& H = P.hop[hh];
          const int S = H.n_dmrs, col0 = H.pilot_col0;
          const bool derot = S >= 2 && P.cfo_compensate;
          const float2* const hbuf = hall + hh * L * N;
          float noise_part = 0.f, rsrp_part = 0.f;
          if (pvec && S >= 2) {
            // Several pilot symbols, even layer count: one thread per 16-byte chunk (pilot, symbol, CDM group) in memory
            // order, so that a warp's loads are 512 contiguous bytes (4 cache lines) instead of 32 lines -- the residual
            // is elementwise, only the order of the float32 partial sums changes.  (The same mapping for the CFO
            // correlation and the LS step needs shuffles to pair / sum the symbols of a pilot and measured slower.)
            const int per_pilot = S * G, total = N * per_pilot;
            for (int c = tid; c < total; c += kThreads) {
              const int i = c / per_pilot, r = c - i * per_pilot, sx = (G == 2) ? (r >> 1) : r, g = (G == 2) ? (r & 1) : 0;
              const float4 pv = __ldg(reinterpret_cast<const float4*>(pu + ((long long)i * ST + col0 + sx) * L + 2 * g));
              float2 h0 = hbuf[2 * g * N + i], h1 = hbuf[(2 * g + 1) * N + i];
              if (derot) {
                h0 = cmul(h0, hsx[hh].rot_pos[sx]);
                h1 = cmul(h1, hsx[hh].rot_pos[sx]);
              }
              const float2 e = cadd(cmul(make_float2(pv.x, pv.y), h0), cmul(make_float2(pv.z, pv.w), h1));
              const float2 d = csub(ld_y(H, g, i, sx), cscale(e, beta));
              noise_part += d.x * d.x + d.y * d.y;
            }
          } else
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            if (g < G) {
              const bool two = 2 * g + 1 < L;
              for (int i = tid; i < N; i += kThreads) {
                const float2* pp = pu + ((long long)i * ST + col0) * L + 2 * g;
                const float2 h0 = hbuf[2 * g * N + i];
                const float2 h1 = two ? hbuf[(2 * g + 1) * N + i] : make_float2(0.f, 0.f);
                for (int s = 0; s < S; ++s) {
                  const float2 y = ld_y(H, g, i, s);
                  float2 pa, pb;
                  ld_pilot_pair(pp + s * L, pvec, two, pa, pb);
                  float2 e = cmul(pa, derot ? cmul(h0, hsx[hh].rot_pos[s]) : h0);
                  if (two) e = cadd(e, cmul(pb, derot ? cmul(h1, hsx[hh].rot_pos[s]) : h1));
                  const float2 d = csub(y, cscale(e, beta));
                  noise_part += d.x * d.x + d.y * d.y;
                }
              }
            }
          }
          for (int idx = tid; idx < N * L; idx += kThreads) {
            const float2 t = hbuf[idx];
            rsrp_part += t.x * t.x + t.y * t.y;
          }
          v[3 * hh] = double(epre_part[hh]);
          v[3 * hh + 1] = double(noise_part);
          v[3 * hh + 2] = double(rsrp_part);
        }
        block_sum<NA, 8>(m, v);
        if (tid == 0) {
          // float32 totals per hop and per quantity, accumulated in float64 like the reference's scalar accumulators
#pragma unroll
          for (int hh = 0; hh < n_hops; ++hh) {
            m.acc[0] += double(float(m.tot[3 * hh]));
            m.acc[1] += double(float(m.tot[3 * hh + 1]));
            m.acc[2] += beta_d * beta_d * double(float(m.tot[3 * hh + 2])) * double(P.hop[hh].n_dmrs);
          }
        }
      }
      sync_a<NA>();
      PROF_MARK(tid == 0, 8);
      if (P.y_bufs == 1) {                                      // single staging buffer: this unit's pilot REs are dead now;
        stage_unit(u + gridDim.x, ystage);                     // the next unit's arrive while the IFFT below runs
        cp_async_commit();
      }
      PROF_MARK(tid == 0, 10);

      // ---- time alignment (:674-698): IFFT of every (hop, layer) vector, fft_layers at a time; every thread keeps
      //      the delay-profile bins it owns in registers across the rounds (layers summed in order, per hop)
      {
        if (P.ta_radix4) {
          // radix-4 passes over a dense twiddle table (round 1): slower by itself, but measured gentler on the writer warps
          // when phase A is light -- the planner's choice (srsce_plan.cc)
        const int M = P.fft_size, log2M = 31 - __clz(M);
        const int* posl = m.re_pos[G - 1];
        const int n_vec = n_hops * L;
        for (int vb = 0; vb < n_vec; vb += P.fft_layers) {
          const int nl = min(P.fft_layers, n_vec - vb);
          if (P.fft_regular) {
            for (int idx = tid; idx < nl * M; idx += kThreads) {
              const int v = vb + (idx >> log2M), n = idx & (M - 1);     // vector v = hop * L + layer: hall is [hop][layer][N]
              ext[idx] = (n < N) ? hall[v * N + n] : make_float2(0.f, 0.f);
            }
          } else {
            for (int idx = tid; idx < nl * M; idx += kThreads) ext[idx] = make_float2(0.f, 0.f);
            sync_a<NA>();
            for (int idx = tid; idx < nl * N; idx += kThreads) {
              const int l = idx / N, i = idx - l * N;
              const int q = i / R;
              ext[l * M + q * kNRE + posl[i - q * R] - posl[0]] = hall[(vb + l) * N + i];
            }
          }
          sync_a<NA>();
          PROF_MARK(tid == 0, 13);
          fft_dif_inverse<NA>(ext, nl, M, log2M, P.twiddle);
          PROF_MARK(tid == 0, 14);
          for (int t = tid; t < 2 * SRSCE_TA_WINDOW; t += kThreads) {
            const int bin = (t < SRSCE_TA_WINDOW) ? t : SRSCE_TA_FFT - 2 * SRSCE_TA_WINDOW + t;
            const int pos = dif_position(bin & (M - 1), M);
            for (int l = 0; l < nl; ++l) {                       // layers are summed in order within their hop's profile
              const int v = vb + l, hh = (TWO_HOPS && v >= L) ? 1 : 0;
              const float2 z = ext[l * M + pos];
              const float pw = z.x * z.x + z.y * z.y;
              hsx[hh].prof[t] = (v == hh * L) ? pw : hsx[hh].prof[t] + pw;
            }
          }
          sync_a<NA>();
          PROF_MARK(tid == 0, 15);
        }
        } else {
        const int M = P.fft_size, log2M = 31 - __clz(M), log2rf = log2M - 9, Mp = M + (M >> 3);
        const int n_vec = n_hops * L;
        constexpr int kBins = (2 * SRSCE_TA_WINDOW + NA - 1) / NA;
        float ta_acc[kBins][2];
#pragma unroll
        for (int k = 0; k < kBins; ++k) ta_acc[k][0] = ta_acc[k][1] = 0.f;
        const int pos0 = m.re_pos[G - 1][0];
        for (int vb = 0; vb < n_vec; vb += P.fft_layers) {
          const int nl = min(P.fft_layers, n_vec - vb);
          // pass 1 (stride M / 8): one (vector, j) task per thread and step, inputs straight from hall
          for (int t = tid; t < (nl << (log2M - 3)); t += kThreads) {
            const int l = t >> (log2M - 3), j = t & ((M >> 3) - 1);
            const float2* const hv = hall + (vb + l) * N;              // vector v = hop * L + layer: hall is [hop][layer][N]
            float2 x[8];
            if (P.fft_regular) {
#pragma unroll
              for (int mm = 0; mm < 8; ++mm) {
                const int n = j + mm * (M >> 3);
                x[mm] = (n < N) ? hv[n] : make_float2(0.f, 0.f);
              }
            } else {
              // zero-stuffed 4096-point input: sample n sits pos0 subcarriers above the first pilot's PRB origin
#pragma unroll
              for (int mm = 0; mm < 8; ++mm) {
                const int sc = j + mm * (M >> 3) + pos0, q = int(__umulhi(unsigned(sc), 357913942u));   // sc / 12
                const int r = m.re_inv[sc - q * kNRE], i = q * R + r;
                x[mm] = (r >= 0 && i < N) ? hv[i] : make_float2(0.f, 0.f);
              }
            }
            fft8_twiddled(x, float(2 * j) / float(M));
            float2* const dst = ext + l * Mp;
#pragma unroll
            for (int r = 0; r < 8; ++r) dst[fft_pad(j + r * (M >> 3))] = x[r];
          }
          sync_a<NA>();
          PROF_MARK(tid == 0, 13);
          fft_pass23<NA>(ext, nl, log2M);
          PROF_MARK(tid == 0, 14);
          // delay profile: the bins of the two windows, each from the RF inputs of its last butterfly
#pragma unroll
          for (int k = 0; k < kBins; ++k) {
            const int t = tid + k * kThreads;
            if (t < 2 * SRSCE_TA_WINDOW) {
              const int bin = ((t < SRSCE_TA_WINDOW) ? t : SRSCE_TA_FFT - 2 * SRSCE_TA_WINDOW + t) & (M - 1);
              const int d1 = bin & 7, d2 = (bin >> 3) & 7, d3 = (bin >> 6) & 7, d4 = bin >> 9;
              const int base = fft_pad(d1 * (M >> 3) + d2 * (M >> 6) + (d3 << log2rf));   // RF <= 8 elements: one pad group
              const int wstep = d4 << (3 - log2rf);                 // e^{+j 2 pi m d4 / RF} = w8[(m d4 8 / RF) mod 8]
              for (int l = 0; l < nl; ++l) {
                const float2* const e = ext + l * Mp + base;
                float2 z = e[0];
                for (int mm = 1; mm < (1 << log2rf); ++mm) z = cadd(z, cmul(e[mm], m.w8[(mm * wstep) & 7]));
                const float pw = z.x * z.x + z.y * z.y;
                if (TWO_HOPS && vb + l >= L) ta_acc[k][1] += pw;
                else ta_acc[k][0] += pw;
              }
            }
          }
          if (vb + P.fft_layers < n_vec) sync_a<NA>();                   // the next round overwrites ext
        }
#pragma unroll
        for (int k = 0; k < kBins; ++k) {
          const int t = tid + k * kThreads;
          if (t < 2 * SRSCE_TA_WINDOW) {
            hsx[0].prof[t] = ta_acc[k][0];
            if (TWO_HOPS) hsx[1].prof[t] = ta_acc[k][1];
          }
        }
        sync_a<NA>();
        PROF_MARK(tid == 0, 15);
        }
        if (warp < n_hops) {
          // first maximum of each window; ties keep the lower index, as torch.max / np.argmax
          const float* prof = hsx[warp].prof;
          float bd = -1.f, ba = -1.f;
          int id = 0, ia = 0;
          for (int t = lane; t < SRSCE_TA_WINDOW; t += 32) {
            const float d = prof[t], a = prof[SRSCE_TA_WINDOW + t];
            if (d > bd) { bd = d; id = t; }
            if (a > ba) { ba = a; ia = t; }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const float od = __shfl_xor_sync(0xffffffffu, bd, o), oa = __shfl_xor_sync(0xffffffffu, ba, o);
            const int oid = __shfl_xor_sync(0xffffffffu, id, o), oia = __shfl_xor_sync(0xffffffffu, ia, o);
            if (od > bd || (od == bd && oid < id)) { bd = od; id = oid; }
            if (oa > ba || (oa == ba && oia < ia)) { ba = oa; ia = oia; }
          }
          if (lane == 0) {
            const int imax = (bd >= ba) ? id : -(SRSCE_TA_WINDOW - ia);
            m.ta_hop[warp] = double(imax) / double(SRSCE_TA_FFT) / P.scs_hz;
          }
        }
      }
      sync_a<NA>();   // ext (IFFT scratch), the profiles and ta_hop are complete
      PROF_MARK(tid == 0, 9);
      if (tid == 0) {
#pragma unroll
        for (int hh = 0; hh < n_hops; ++hh) m.acc[3] += m.ta_hop[hh];
      }
    }

    if (dbg & 1) {                                            // profiling ablation: writers run on stale pilots
      if (wait_empty) bar_sync(kBarEmpty + hb_idx, kCtaThreads);
      build_store_table(m.col[hb_idx]);
      bar_arrive(kBarFull + hb_idx, kCtaThreads);
    }
    // ---- per-unit scalars (:899-919, :932-937)
    if (tid == 0) {
      double* o = scalars + u * 6;
      o[0] = m.acc[1] * P.norm_noise;
      o[1] = m.acc[2] * P.norm_rsrp;
      o[2] = m.acc[0] * P.norm_epre;
      o[3] = m.acc[3] / P.ta_hop_div;
      o[4] = (m.acc[5] != 0.0) ? m.acc[4] * P.scs_hz : 0.0;
      o[5] = m.acc[5];
    }
    // (the barrier at the top of the next unit orders these reads of m.acc before it is cleared)
  }
  }  // estimator loop (not a fill-only launch)
  PROF_FLUSH(tid == 0, 0);
  } else {
  // ================================== writer warps ============================================
  const int wb = warp - NA / 32;
  const long long n_iter = (n_units - blockIdx.x + gridDim.x - 1) / gridDim.x;
  int it = 0;
  PROF_DECL
  for (long long u = blockIdx.x; u < n_units; u += gridDim.x, ++it) {
    const int hb_idx = (P.hall_bufs == 2) ? (it & 1) : 0;
    const float2* const hall = hall0 + hb_idx * (n_hops * L * N);
    const ColInfo* const colb = m.col[hb_idx];
    bar_sync(kBarFull + hb_idx, kCtaThreads);
    PROF_MARK(wb == 0 && lane == 0, 11);
    // ===================================== phase B ==========================================
    // Each warp takes 32 consecutive subcarriers: lane r interpolates row r into the warp's tile
    // ([hop][row][layer]), then the warp streams the rows out, every lane owning a fixed store
    // slot (symbol, layers) of the row.
    {
      float2* const tile0 = tiles + wb * (kTileRows * L * n_hops);
      float2* const outu = out + u * ((long long)n_sc * row_elems);
      const bool fast = cols <= 32;
      const int rpi = fast ? 32 / cols : 1;
      const int my_rs = fast ? lane / cols : 0;
      const int my_c = fast ? lane - my_rs * cols : 0;
      const bool lane_on = fast && (my_rs < rpi);
      const ColInfo mine = colb[my_c];
      const bool pair = (VEC == 2) && ((L & 1) == 0);      // both elements of a slot are adjacent layers of one symbol

      // interpolate the 32 rows starting at row0 into a tile ([hop][row][layer]); returns the per-hop in-band row masks
      auto interp_tile = [&](int row0, float2* tile0, unsigned& in1, unsigned& in2) {
        const int row = row0 + lane;
        in1 = in2 = 0;
#pragma unroll
        for (int hh = 0; hh < n_hops; ++hh) {
          const HopDev& H = P.hop[hh];
          const float2* const hb = hall + hh * L * N;
          float2* const trow = tile0 + (hh * kTileRows + lane) * L;
          const bool inband = row >= H.band_sc0 && row < H.band_sc1;
          const unsigned bal = __ballot_sync(0xffffffffu, inband);
          if (hh == 0) in1 = bal;
          else in2 = bal;
          if (bal == 0) continue;                    // whole tile outside this hop's band
          if (!inband) {
            for (int l = 0; l < L; ++l) trow[l] = make_float2(0.f, 0.f);
          } else {
            const int br = row - H.band_sc0;
            const int q = br / kNRE, pos = br - q * kNRE;
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              if (g < G) {
                float2 v0, v1 = make_float2(0.f, 0.f);
                const bool two = 2 * g + 1 < L;
                const float2* const ha = hb + 2 * g * N;
                if (P.fill == SRSCE_FILL_LINEAR) {
                  const LerpEntry e = m.lerp[g][pos];
                  const int il = max(0, min(N - 1, q * R + e.jl)), ir = max(0, min(N - 1, q * R + e.jr));
                  const float2 l0 = ha[il], r0 = ha[ir];
                  v0 = make_float2(l0.x + e.alpha * (r0.x - l0.x), l0.y + e.alpha * (r0.y - l0.y));
                  if (two) {
                    const float2 l1 = ha[N + il], r1 = ha[N + ir];
                    v1 = make_float2(l1.x + e.alpha * (r1.x - l1.x), l1.y + e.alpha * (r1.y - l1.y));
                  }
                } else {
                  const int T = P.fill_terms;
                  const int bands = H.band_sc1 - H.band_sc0;
                  const bool interior = P.fill_periodic && br >= kCnnEdgeRows && br < bands - kCnnEdgeRows;
                  const int k0 = interior ? m.cnn_idx0[g][pos] + q * R : __ldg(P.fill_idx + g * bands + br);
                  const float* cf = interior ? m.cnn_coef[g][pos] : P.fill_coef + ((long long)g * bands + br) * T;
                  v0 = make_float2(0.f, 0.f);
                  for (int t = 0; t < T; ++t) {
                    const float w = cf[t];
                    const float2 a = ha[k0 + t];
                    v0.x += w * a.x;
                    v0.y += w * a.y;
                    if (two) {
                      const float2 b = ha[N + k0 + t];
                      v1.x += w * b.x;
                      v1.y += w * b.y;
                    }
                  }
                }
                trow[2 * g] = v0;
                if (two) trow[2 * g + 1] = v1;
              }
            }
          }
        }
      };
      // stream the rows of a finished tile to the grid (or the compact output)
      auto emit_tile = [&](int row0, const float2* tile0, unsigned in1, unsigned in2) {
        const int row = row0 + lane;
        const int rows_here = min(kTileRows, n_sc - row0);
        if (out_c != nullptr) {
          // compact output [unit][hop][subcarrier][layer]: every lane writes the row it interpolated
          if (row < n_sc) {
#pragma unroll
            for (int hh = 0; hh < n_hops; ++hh) {
              float2* d = out_c + (((long long)u * n_hops + hh) * n_sc + row) * L;
              const bool have = (((hh == 0) ? in1 : in2) >> lane) & 1u;
              const float2* t = tile0 + (hh * kTileRows + lane) * L;
              for (int l = 0; l < L; ++l) __stcs(d + l, have ? t[l] : make_float2(0.f, 0.f));
            }
          }
          return;
        }

        if (fast) {
          if (lane_on) {
            float2* dst = outu + (long long)(row0 + my_rs) * row_elems + my_c * VEC;
            const long long dstep = (long long)rpi * row_elems;
            if ((in1 | in2) == 0) {                  // tile entirely outside every band: zeros
              for (int r = my_rs; r < rows_here; r += rpi, dst += dstep) {
                if (VEC == 2) __stcs(reinterpret_cast<float4*>(dst), make_float4(0.f, 0.f, 0.f, 0.f));
                else __stcs(dst, make_float2(0.f, 0.f));
              }
            } else if (pair && !TWO_HOPS && rows_here == kTileRows && row_elems == 14 * L && (L == 4 || L == 2)) {
              // the common shapes (14 symbols x 4 or 2 layers, full tile): fully unrolled with compile-time strides
              const float4* src = reinterpret_cast<const float4*>(tile0 + my_rs * L + mine.layer[0]);
              const bool on = (mine.flags[0] & 1) != 0;
              if (L == 4) stream_full_tile<4, HAS_ROT>(src, reinterpret_cast<float4*>(dst), on, mine.rot[0], mine.rot[1]);
              else stream_full_tile<2, HAS_ROT>(src, reinterpret_cast<float4*>(dst), on, mine.rot[0], mine.rot[1]);
            } else if (pair && TWO_HOPS && rows_here == kTileRows && row_elems == 14 * L && (L == 4 || L == 2)) {
              const float4* s1 = reinterpret_cast<const float4*>(tile0 + my_rs * L + mine.layer[0]);
              const float4* s2 = s1 + kTileRows * L / 2;
              if (L == 4)
                stream_full_tile_2hop<4, HAS_ROT>(s1, s2, reinterpret_cast<float4*>(dst), mine.flags[0], in1, in2, my_rs,
                                                  mine.rot[0], mine.rot[1]);
              else
                stream_full_tile_2hop<2, HAS_ROT>(s1, s2, reinterpret_cast<float4*>(dst), mine.flags[0], in1, in2, my_rs,
                                                  mine.rot[0], mine.rot[1]);
            } else if (VEC == 2 && !TWO_HOPS && rows_here == kTileRows && row_elems == 14 * L && (L == 1 || L == 3)) {
              const float2* s0 = tile0 + my_rs * L + mine.layer[0];
              const float2* s1 = tile0 + my_rs * L + mine.layer[1];
              const bool on0 = (mine.flags[0] & 1) != 0, on1 = (mine.flags[1] & 1) != 0;
              if (L == 1) stream_full_tile_odd<1, HAS_ROT>(s0, s1, reinterpret_cast<float4*>(dst), on0, on1, mine.rot[0], mine.rot[1]);
              else stream_full_tile_odd<3, HAS_ROT>(s0, s1, reinterpret_cast<float4*>(dst), on0, on1, mine.rot[0], mine.rot[1]);
            } else if (pair && !TWO_HOPS) {
              // hot loop of the headline configuration: one 16-byte shared load, one 16-byte streaming store
              const float4* src = reinterpret_cast<const float4*>(tile0 + my_rs * L + mine.layer[0]);
              const int sstep = rpi * L / 2;
              const bool on = (mine.flags[0] & 1) != 0;
#pragma unroll 4
              for (int r = my_rs; r < rows_here; r += rpi, dst += dstep, src += sstep) {
                float4 v = on ? *src : make_float4(0.f, 0.f, 0.f, 0.f);
                if (HAS_ROT) {
                  const float2 a = cmul(make_float2(v.x, v.y), mine.rot[0]), b = cmul(make_float2(v.z, v.w), mine.rot[1]);
                  v = make_float4(a.x, a.y, b.x, b.y);
                }
                __stcs(reinterpret_cast<float4*>(dst), v);
              }
            } else {
              for (int r = my_rs; r < rows_here; r += rpi, dst += dstep) {
                float2 val[2];
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                  float2 t = make_float2(0.f, 0.f);
                  if (TWO_HOPS && (mine.flags[e] & 2) && ((in2 >> r) & 1u)) t = tile0[(kTileRows + r) * L + mine.layer[e]];
                  else if ((mine.flags[e] & 1) && ((in1 >> r) & 1u)) t = tile0[r * L + mine.layer[e]];
                  val[e] = HAS_ROT ? cmul(t, mine.rot[e]) : t;
                }
                if (VEC == 2) __stcs(reinterpret_cast<float4*>(dst), make_float4(val[0].x, val[0].y, val[1].x, val[1].y));
                else __stcs(dst, val[0]);
              }
            }
          }
        } else {
          for (int r = 0; r < rows_here; ++r) {
            for (int c = lane; c < cols; c += 32) {
              const ColInfo ci = colb[c];
              float2 val[2];
#pragma unroll
              for (int e = 0; e < VEC; ++e) {
                float2 t = make_float2(0.f, 0.f);
                if (TWO_HOPS && (ci.flags[e] & 2) && ((in2 >> r) & 1u)) t = tile0[(kTileRows + r) * L + ci.layer[e]];
                else if ((ci.flags[e] & 1) && ((in1 >> r) & 1u)) t = tile0[r * L + ci.layer[e]];
                val[e] = HAS_ROT ? cmul(t, ci.rot[e]) : t;
              }
              float2* dst = outu + (long long)(row0 + r) * row_elems + c * VEC;
              if (VEC == 2) __stcs(reinterpret_cast<float4*>(dst), make_float4(val[0].x, val[0].y, val[1].x, val[1].y));
              else __stcs(dst, val[0]);
            }
          }
        }
      };
      const int row_end = (dbg & 2) ? 0 : n_sc, row_step = kWarpsB * kTileRows;
      // (Tried: a software-pipelined variant that fetches the next tile's pilots before streaming this one, into a second
      //  tile buffer -- no gain on B200)
      for (int row0 = wb * kTileRows; row0 < row_end; row0 += row_step) {
        unsigned in1, in2;
        __syncwarp();
        interp_tile(row0, tile0, in1, in2);
        __syncwarp();
        emit_tile(row0, tile0, in1, in2);
      }
    }
    PROF_MARK(wb == 0 && lane == 0, 12);
    if (it + P.hall_bufs < n_iter) bar_arrive(kBarEmpty + hb_idx, kCtaThreads);   // this buffer may be rewritten
  }
  PROF_FLUSH(wb == 0 && lane == 0, 0);
  }
}

typedef void (*KernelFn)(const PlanDev, const float2*, long long, long long, long long, const float2*, int,
                         double, float2*, double*, long long, int, float2*, float2*, const long long*, const float2*,
                         float2*);

template <int NA, int STAGED>
KernelFn pick_kernel_na(int vec, bool two_hops, bool has_rot) {
  if (vec == 2) {
    if (two_hops) return has_rot ? estimate_kernel<NA, 2, true, true, STAGED> : estimate_kernel<NA, 2, true, false, STAGED>;
    return has_rot ? estimate_kernel<NA, 2, false, true, STAGED> : estimate_kernel<NA, 2, false, false, STAGED>;
  }
  if (two_hops) return has_rot ? estimate_kernel<NA, 1, true, true, STAGED> : estimate_kernel<NA, 1, true, false, STAGED>;
  return has_rot ? estimate_kernel<NA, 1, false, true, STAGED> : estimate_kernel<NA, 1, false, false, STAGED>;
}

KernelFn pick_kernel(const PlanDev& v, bool fill_only = false) {
  const bool rot = v.cfo_compensate && v.cfo_defined && !fill_only, two = v.n_hops == 2;   // (the fill has no CFO rotation)
  if (v.h_global) {                                                            // two-hop plans only (srsce_plan.cc, layout F)
    if (v.vec == 2) return rot ? estimate_kernel<512, 2, true, true, kModeGlobalH> : estimate_kernel<512, 2, true, false, kModeGlobalH>;
    return rot ? estimate_kernel<512, 1, true, true, kModeGlobalH> : estimate_kernel<512, 1, true, false, kModeGlobalH>;
  }
  if (v.y_bufs == 0) return pick_kernel_na<512, kModeUnstaged>(v.vec, two, rot);    // the planner pairs the unstaged layout with NA = 512
  return v.est_threads == 512 ? pick_kernel_na<512, kModeStaged>(v.vec, two, rot) : pick_kernel_na<256, kModeStaged>(v.vec, two, rot);
}

// Ablation switch, PROFILING BUILD ONLY (SRSCE_DEBUG_ABLATE: bit 0 skips phase A, bit 1 skips phase B, bit 2 skips
// the staging loads): results are garbage when set.  The production library never reads the environment.
#ifdef SRSCE_PROFILING
int debug_mode() {
  static const int mode = [] {
    const char* e = getenv("SRSCE_DEBUG_ABLATE");
    const int v = e ? atoi(e) : 0;
    if (v) fprintf(stderr, "libsrsce_prof: SRSCE_DEBUG_ABLATE=%d -- profiling ablation active, RESULTS ARE INVALID\n", v);
    return v;
  }();
  return mode;
}
#else
constexpr int debug_mode() { return 0; }
#endif

std::string cuda_msg(const char* what, cudaError_t e) {
  return std::string(what) + ": " + cudaGetErrorString(e);
}

}  // namespace

int max_dynamic_smem() { return 227 * 1024; }   // architectural per-CTA limit on sm_100

#ifdef SRSCE_PROFILING
// [grid][16] int64 cycle counters per CTA (device memory, zeroed by the caller); nullptr switches them off.
extern "C" int srsce_prof_set_buffer(void* device_ptr) {
  long long* p = static_cast<long long*>(device_ptr);
  return cudaMemcpyToSymbol(g_prof_buf, &p, sizeof p) == cudaSuccess ? 0 : -1;
}
#endif

int upload_plan(Plan& p, std::string& err) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    (void)cudaGetLastError();
    p.device = -1;
    return SRSCE_OK;                       // host-only plan (CPU-side tests); compute calls will refuse it
  }
  int dev = 0;
  if ((e = cudaGetDevice(&dev)) != cudaSuccess) { err = cuda_msg("cudaGetDevice", e); return SRSCE_ERR_CUDA; }

  // twiddles of the transform this plan runs, dense (M entries), so that the first radix-4 pass walks
  // them with unit stride and the whole table occupies M * 8 bytes of L1
  std::vector<float2> tw(p.dev.fft_size);
  for (int k = 0; k < p.dev.fft_size; ++k) {
    const double a = 2.0 * M_PI * double(k) / double(p.dev.fft_size);
    tw[k] = make_float2(float(cos(a)), float(sin(a)));
  }
  if ((e = cudaMalloc(&p.d_twiddle, tw.size() * sizeof(float2))) != cudaSuccess) { err = cuda_msg("cudaMalloc", e); return SRSCE_ERR_CUDA; }
  if ((e = cudaMemcpy(p.d_twiddle, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice)) != cudaSuccess) {
    err = cuda_msg("cudaMemcpy", e);
    return SRSCE_ERR_CUDA;
  }
  p.dev.twiddle = static_cast<const float2*>(p.d_twiddle);
  if (p.desc.fill == SRSCE_FILL_CNN) {
    const size_t bi = p.fill_idx.size() * sizeof(int32_t), bc = p.fill_coef.size() * sizeof(float);
    if ((e = cudaMalloc(&p.d_fill_idx, bi)) != cudaSuccess || (e = cudaMalloc(&p.d_fill_coef, bc)) != cudaSuccess) {
      err = cuda_msg("cudaMalloc", e);
      return SRSCE_ERR_CUDA;
    }
    if ((e = cudaMemcpy(p.d_fill_idx, p.fill_idx.data(), bi, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(p.d_fill_coef, p.fill_coef.data(), bc, cudaMemcpyHostToDevice)) != cudaSuccess) {
      err = cuda_msg("cudaMemcpy", e);
      return SRSCE_ERR_CUDA;
    }
    p.dev.fill_idx = static_cast<const int32_t*>(p.d_fill_idx);
    p.dev.fill_coef = static_cast<const float*>(p.d_fill_coef);
  }
  int sms = 0;
  if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) {
    err = cuda_msg("cudaDeviceGetAttribute", e);
    return SRSCE_ERR_CUDA;
  }
  auto prepare = [&](const PlanDev& v, int* per_sm) -> int {
    KernelFn fill_fn = pick_kernel(v, true);            // srsce_fill (direct) launches the variant without the final rotation
    if ((e = cudaFuncSetAttribute(reinterpret_cast<const void*>(fill_fn), cudaFuncAttributeMaxDynamicSharedMemorySize, max_dynamic_smem())) != cudaSuccess) {
      err = cuda_msg("cudaFuncSetAttribute(smem)", e);
      return SRSCE_ERR_CUDA;
    }
    KernelFn fn = pick_kernel(v);
    // the limit is per kernel variant, not per plan: always raise it to the architectural maximum so that
    // plans with different footprints can share a variant
    if ((e = cudaFuncSetAttribute(reinterpret_cast<const void*>(fn), cudaFuncAttributeMaxDynamicSharedMemorySize, max_dynamic_smem())) != cudaSuccess) {
      err = cuda_msg("cudaFuncSetAttribute(smem)", e);
      return SRSCE_ERR_CUDA;
    }
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, reinterpret_cast<const void*>(fn), cta_threads(v.est_threads, v.n_hops == 2), v.smem_bytes)) != cudaSuccess) {
      err = cuda_msg("occupancy query", e);
      return SRSCE_ERR_CUDA;
    }
    if (*per_sm < 1) { err = "fused kernel does not fit on an SM"; return SRSCE_ERR_UNSUPPORTED; }
    return SRSCE_OK;
  };
  if (int rc = prepare(p.dev, &p.ctas_per_sm)) return rc;
  if (p.has_pair) {
    p.dev_pair.twiddle = p.dev.twiddle;
    p.dev_pair.fill_idx = p.dev.fill_idx;
    p.dev_pair.fill_coef = p.dev.fill_coef;
    if (int rc = prepare(p.dev_pair, &p.ctas_per_sm_pair)) return rc;
  }
  p.n_sms = sms;
  p.device = dev;
  return prepare_fill_tc(p, err);
}

void release_plan(Plan& p) {
  if (p.d_twiddle) cudaFree(p.d_twiddle);
  if (p.d_fill_idx) cudaFree(p.d_fill_idx);
  if (p.d_fill_coef) cudaFree(p.d_fill_coef);
  if (p.d_fill_tc_a) cudaFree(p.d_fill_tc_a);
  p.d_twiddle = p.d_fill_idx = p.d_fill_coef = p.d_fill_tc_a = nullptr;
}

// Plans of layout F keep the smoothed pilots of each resident CTA in global memory: a stream-ordered allocation per
// launch (the plan stays immutable and shareable between streams; the pool serves repeat calls without a driver call).
static int alloc_h_scratch(const PlanDev& v, unsigned grid, void* stream, float2** out, std::string& err) {
  *out = nullptr;
  if (!v.h_global) return SRSCE_OK;
  const size_t bytes = size_t(grid) * v.hall_bufs * v.n_hops * v.n_layers * v.n_re * sizeof(float2);
  void* ptr = nullptr;
  const cudaError_t e = cudaMallocAsync(&ptr, bytes, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) { err = cuda_msg("cudaMallocAsync(smoothed-pilot scratch)", e); return SRSCE_ERR_CUDA; }
  *out = static_cast<float2*>(ptr);
  return SRSCE_OK;
}

int launch_estimate(const Plan& p, const void* rx, const int64_t strides[3], const void* pilots, int pilot_share,
                    double beta, void* out, double* scalars, int64_t n_units, void* stream, int* launches,
                    std::string& err, void* out_compact, void* rot_compact, const int64_t* rx_index) {
  *launches = 0;
  if (p.device < 0) { err = "plan has no CUDA device (created without a GPU); there is no CPU compute path"; return SRSCE_ERR_NO_DEVICE; }
  int dev = -1;
  cudaGetDevice(&dev);
  if (dev != p.device) { err = "plan was created on another CUDA device"; return SRSCE_ERR_INVALID; }
  if (n_units == 0) return SRSCE_OK;
  // the reference call shape (dense output, one [n_sc, n_sym] grid per unit) runs the plan's primary layout;
  // shared grids, other strides and the compact output run the CTA pair when the plan has one
  const bool reference_shape = rx_index == nullptr && out_compact == nullptr && strides[2] == 1;
  const bool pair = p.has_pair && !reference_shape;
  const PlanDev& v = pair ? p.dev_pair : p.dev;
  KernelFn fn = pick_kernel(v);
  const long long cap = (long long)p.n_sms * (pair ? p.ctas_per_sm_pair : p.ctas_per_sm);
  const unsigned grid = unsigned(n_units < cap ? n_units : cap);
  float2* h_scratch = nullptr;
  if (int rc = alloc_h_scratch(v, grid, stream, &h_scratch, err)) return rc;
  fn<<<grid, cta_threads(v.est_threads, v.n_hops == 2), v.smem_bytes, static_cast<cudaStream_t>(stream)>>>(
      v, static_cast<const float2*>(rx), (long long)strides[0], (long long)strides[1], (long long)strides[2],
      static_cast<const float2*>(pilots), pilot_share, beta, static_cast<float2*>(out), scalars, (long long)n_units,
      debug_mode(), static_cast<float2*>(out_compact), static_cast<float2*>(rot_compact),
      reinterpret_cast<const long long*>(rx_index), nullptr, h_scratch);
  cudaError_t e = cudaGetLastError();
  if (h_scratch) cudaFreeAsync(h_scratch, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) { err = cuda_msg("kernel launch", e); return SRSCE_ERR_CUDA; }
  *launches = 1;
  return SRSCE_OK;
}

// srsce_fill, direct implementation: the fused kernel's writer warps fed with given smoothed pilots (no phase A).
int launch_fill_direct(const Plan& p, const void* h_pilots, void* out, int64_t n_units, void* stream, int* launches,
                       std::string& err, void* out_compact) {
  *launches = 0;
  if (p.device < 0) { err = "plan has no CUDA device (created without a GPU); there is no CPU compute path"; return SRSCE_ERR_NO_DEVICE; }
  int dev = -1;
  cudaGetDevice(&dev);
  if (dev != p.device) { err = "plan was created on another CUDA device"; return SRSCE_ERR_INVALID; }
  if (n_units == 0) return SRSCE_OK;
  const bool pair = p.has_pair;                       // no gather, light estimator side: the CTA pair when the plan has one
  const PlanDev& v = pair ? p.dev_pair : p.dev;
  KernelFn fn = pick_kernel(v, true);
  const long long cap = (long long)p.n_sms * (pair ? p.ctas_per_sm_pair : p.ctas_per_sm);
  const unsigned grid = unsigned(n_units < cap ? n_units : cap);
  float2* h_scratch = nullptr;
  if (int rc = alloc_h_scratch(v, grid, stream, &h_scratch, err)) return rc;
  fn<<<grid, cta_threads(v.est_threads, v.n_hops == 2), v.smem_bytes, static_cast<cudaStream_t>(stream)>>>(
      v, nullptr, 0, 0, 0, nullptr, 1, 1.0, static_cast<float2*>(out), nullptr, (long long)n_units, 0,
      static_cast<float2*>(out_compact), nullptr, nullptr, static_cast<const float2*>(h_pilots), h_scratch);
  cudaError_t e = cudaGetLastError();
  if (h_scratch) cudaFreeAsync(h_scratch, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) { err = cuda_msg("fill launch", e); return SRSCE_ERR_CUDA; }
  *launches = 1;
  return SRSCE_OK;
}

}  // namespace srsce
