// Fused per-slot port channel estimator for sm_100a.
//
// One CTA owns one unit (= one reference call: one received grid, one UE's pilots) at a time and
// walks the whole reference path without leaving the SM:
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
// bandwidth phase (>= 98 % of the bytes); phase A of other resident CTAs overlaps it.
//
// Precision islands follow the reference: complex64 data, float32 |.|^2 partial sums combined in
// float64, float32 angle; float64 for the virtual-pilot regression, phase unwrap, FIR
// accumulation, CFO phase tables and every scalar accumulator.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>

#include <mutex>

#include "srsce_internal.h"

namespace srsce {

namespace {

struct ColInfo {           // one store slot of an output row (VEC consecutive complex64 elements)
  float2 rot[2];           // final CFO rotation of each element's symbol
  int32_t layer[2];
  int32_t flags[2];        // bit0: symbol inside hop-1 allocation, bit1: inside hop-2 allocation
};

struct Misc {
  double red[kWarps][8];
  double tot[8];
  double acc[8];           // 0 epre, 1 noise, 2 rsrp, 3 ta, 4 cfo, 5 cfo defined
  double taps[SRSCE_MAX_TAPS + 1];
  float2 rot_neg[SRSCE_MAX_SYMBOLS];
  float2 rot_pos[SRSCE_MAX_SYMBOLS];
  float prof[2 * SRSCE_TA_WINDOW];
  LerpEntry lerp[2][kNRE];
  int32_t re_pos[2][kNRE];
  ColInfo col[SRSCE_MAX_SYMBOLS * SRSCE_MAX_LAYERS];
};
static_assert(sizeof(Misc) <= kMiscBytes, "Misc does not fit its shared-memory slot");

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cmul_conj(float2 a, float2 b) {  // a * conj(b)
  return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cscale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum K per-thread values over the CTA; every thread gets the totals in m.tot[0..K).
template <int K>
__device__ __forceinline__ void block_sum(Misc& m, double (&v)[K]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const double s = warp_sum(v[k]);
    if (lane == 0) m.red[warp][k] = s;
  }
  __syncthreads();
  if (threadIdx.x < K) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s += m.red[w][threadIdx.x];
    m.tot[threadIdx.x] = s;
  }
  __syncthreads();
}

// Subcarrier of pilot i of CDM group g inside a hop band.
__device__ __forceinline__ int pilot_sc(const Misc& m, int band_sc0, int R, int g, int i) {
  const int q = i / R;
  return band_sc0 + q * kNRE + m.re_pos[g][i - q * R];
}

// create_virtual_pilots (:69-140) for one edge of one layer.  edge(k) is the k-th pilot counted
// from the band edge inwards; out(k), k = 0..nv-1, receives the virtual pilot at abscissa k - nv.
template <typename EdgeFn, typename OutFn>
__device__ void virtual_pilots(int nv, double mx, double den, EdgeFn edge, OutFn out) {
  if (nv <= 0) return;
  if (nv == 1) {                                    // :95-101 constant extrapolation
    const float2 e = edge(0);
    const float amp = sqrtf(e.x * e.x + e.y * e.y);
    const float ph = atan2f(e.y, e.x);
    float s, c;
    sincosf(ph, &s, &c);
    out(0, make_float2(amp * c, amp * s));
    return;
  }
  // straight-line fits of modulus and unwrapped phase, float64 (:105-138); unwrap as _unwrap_1d (:35-66)
  double sy_m = 0.0, sxy_m = 0.0, sy_p = 0.0, sxy_p = 0.0;
  double prev = 0.0, corr = 0.0;
  const double two_pi = 2.0 * M_PI;
  for (int k = 0; k < nv; ++k) {
    const float2 e = edge(k);
    const double mod = double(sqrtf(e.x * e.x + e.y * e.y));
    const double ph = double(atan2f(e.y, e.x));
    if (k > 0) {
      const double d = ph - prev;
      double dm = fmod(d + M_PI, two_pi);
      if (dm < 0.0) dm += two_pi;                   // floored modulo, as torch.remainder
      dm -= M_PI;
      if (dm == -M_PI && d > 0.0) dm += two_pi;
      if (fabs(d) >= M_PI) corr += dm - d;
    }
    prev = ph;
    const double phu = ph + corr;
    sy_m += mod;
    sxy_m += double(k) * mod;
    sy_p += phu;
    sxy_p += double(k) * phu;
  }
  const double n = double(nv);
  const double my_m = sy_m / n, my_p = sy_p / n;
  const double a_m = (sxy_m - n * mx * my_m) / den, b_m = my_m - a_m * mx;
  const double a_p = (sxy_p - n * mx * my_p) / den, b_p = my_p - a_p * mx;
  for (int k = 0; k < nv; ++k) {
    const double x = double(k - nv);
    const double mod = a_m * x + b_m, arg = a_p * x + b_p;
    double s, c;
    sincos(arg, &s, &c);
    out(k, make_float2(float(mod * c), float(mod * s)));
  }
}

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
__device__ void fft_dif_inverse(float2* s, int nl, int M, int log2M, const float2* __restrict__ tw) {
  const int tid = threadIdx.x;
  int n = M, log2n = log2M;
  while (n >= 4) {
    const int q = n >> 2, log2q = log2n - 2;
    const int step = SRSCE_TA_FFT >> log2n;         // twiddle stride: w_n^j = tw[j * 4096 / n]
    const int per_vec = M >> 2;
    for (int b = tid; b < nl * per_vec; b += kThreads) {
      const int l = b / per_vec, r = b - l * per_vec;
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
    __syncthreads();
    n = q;
    log2n = log2q;
  }
  if (n == 2) {
    for (int b = tid; b < nl * (M >> 1); b += kThreads) {
      float2* p = s + 2 * b;
      const float2 a0 = p[0], a1 = p[1];
      p[0] = cadd(a0, a1);
      p[1] = csub(a0, a1);
    }
    __syncthreads();
  }
}

template <int VEC, bool TWO_HOPS, bool HAS_ROT>
__global__ void __launch_bounds__(kThreads, 2)
estimate_kernel(const __grid_constant__ PlanDev P, const float2* __restrict__ rx, long long su, long long ssc,
                long long ssym, const float2* __restrict__ pilots, int pilot_share, double beta_d,
                float2* __restrict__ out, double* __restrict__ scalars, long long n_units) {
  extern __shared__ __align__(16) unsigned char smem[];
  float2* const hall = reinterpret_cast<float2*>(smem + P.off_h);
  float2* const ext = reinterpret_cast<float2*>(smem + P.off_ext);
  Misc& m = *reinterpret_cast<Misc*>(smem + P.off_misc);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int L = P.n_layers, G = P.n_cdm, R = P.re_per_prb, N = P.n_re, ST = P.n_dmrs_total;
  const int nv = P.n_virtual, K = P.n_taps;
  const int n_hops = TWO_HOPS ? 2 : 1;
  const float beta = float(beta_d);
  const float inv_beta = 1.0f / beta;

  // plan tables that are indexed per lane go to shared memory once per CTA
  for (int k = tid; k < SRSCE_MAX_TAPS; k += kThreads) m.taps[k] = P.taps[k];
  for (int k = tid; k < 2 * kNRE; k += kThreads) {
    m.lerp[k / kNRE][k % kNRE] = P.lerp[k / kNRE][k % kNRE];
    m.re_pos[k / kNRE][k % kNRE] = P.re_pos[k / kNRE][k % kNRE];
  }
  __syncthreads();

  for (long long u = blockIdx.x; u < n_units; u += gridDim.x) {
    const float2* __restrict__ rxu = rx + u * su;
    const float2* __restrict__ pu = pilots + (u / pilot_share) * ((long long)N * ST * L);
    if (tid < 8) m.acc[tid] = 0.0;
    // (visibility of m.acc is ordered by the barriers inside block_sum below)

    // ===================================== phase A ==========================================
    for (int hh = 0; hh < n_hops; ++hh) {
      const HopDev& H = P.hop[hh];
      const int S = H.n_dmrs, col0 = H.pilot_col0, sc0 = H.band_sc0;
      float2* const hbuf = hall + hh * N * L;
      const bool est_cfo = S >= 2;
      const bool derot = est_cfo && P.cfo_compensate;

      // ---- CFO: correlate the first two pilot symbols (compensate_cfo :357-451)
      if (est_cfo) {
        float2 c[2][2] = {{{0.f, 0.f}, {0.f, 0.f}}, {{0.f, 0.f}, {0.f, 0.f}}};
        const long long o0 = (long long)H.sym_idx[0] * ssym, o1 = (long long)H.sym_idx[1] * ssym;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (g < G) {
            for (int i = tid; i < N; i += kThreads) {
              const long long rs = (long long)pilot_sc(m, sc0, R, g, i) * ssc;
              const float2 y0 = __ldg(rxu + rs + o0), y1 = __ldg(rxu + rs + o1);
              const float2* pp = pu + ((long long)i * ST + col0) * L + 2 * g;
#pragma unroll
              for (int k = 0; k < 2; ++k) {
                if (2 * g + k < L) {
                  const float2 x0 = cmul_conj(y0, __ldg(pp + k)), x1 = cmul_conj(y1, __ldg(pp + L + k));
                  c[g][k] = cadd(c[g][k], cmul_conj(x1, x0));      // conj(x0) * x1
                }
              }
            }
          }
        }
        double v[8] = {c[0][0].x, c[0][0].y, c[0][1].x, c[0][1].y, c[1][0].x, c[1][0].y, c[1][1].x, c[1][1].y};
        block_sum<8>(m, v);
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
          for (int s = 0; s < S; ++s) {
            double sn = 0.0, cs = 1.0;
            if (derot) sincos(2.0 * M_PI * H.t_dmrs[s] * cfo_hop, &sn, &cs);
            m.rot_pos[s] = make_float2(float(cs), float(sn));
            m.rot_neg[s] = make_float2(float(cs), float(-sn));
          }
        }
        __syncthreads();
      }

      // ---- LS estimate, symbol average, CDM pair average (:593, :613-628) + EPRE (:584)
      float epre_part = 0.f;
      {
        float2* const dst = (P.smoothing == SRSCE_SMOOTH_FILTER) ? ext + nv * L : hbuf;
        const float inv_s = 1.0f / float(S);
        const int n_pairs = (N + 1) >> 1;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (g < G) {
            for (int pr = tid; pr < n_pairs; pr += kThreads) {
              float2 h[2][2];
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int i = 2 * pr + e;
                float2 a0 = make_float2(0.f, 0.f), a1 = a0;
                if (i < N) {
                  const long long rs = (long long)pilot_sc(m, sc0, R, g, i) * ssc;
                  const float2* pp = pu + ((long long)i * ST + col0) * L + 2 * g;
                  for (int s = 0; s < S; ++s) {
                    const float2 y = __ldg(rxu + rs + (long long)H.sym_idx[s] * ssym);
                    epre_part += y.x * y.x + y.y * y.y;
                    float2 x0 = cmul_conj(y, __ldg(pp + s * L));
                    if (derot) x0 = cmul(x0, m.rot_neg[s]);
                    a0 = cadd(a0, x0);
                    if (2 * g + 1 < L) {
                      float2 x1 = cmul_conj(y, __ldg(pp + s * L + 1));
                      if (derot) x1 = cmul(x1, m.rot_neg[s]);
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
                  dst[i * L + 2 * g] = h[e][0];
                  if (2 * g + 1 < L) dst[i * L + 2 * g + 1] = h[e][1];
                }
              }
            }
          }
        }
      }
      __syncthreads();

      // ---- smoothing (:633-668)
      if (P.smoothing == SRSCE_SMOOTH_MEAN) {
        double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int i = tid; i < N; i += kThreads)
          for (int l = 0; l < L; ++l) {
            const float2 t = hbuf[i * L + l];
            if (l == 0) { v[0] += t.x; v[1] += t.y; }
            else if (l == 1) { v[2] += t.x; v[3] += t.y; }
            else if (l == 2) { v[4] += t.x; v[5] += t.y; }
            else { v[6] += t.x; v[7] += t.y; }
          }
        block_sum<8>(m, v);
        const float inv_n = 1.0f / float(N);
        for (int idx = tid; idx < N * L; idx += kThreads) {
          const int l = idx % L;
          hbuf[idx] = make_float2(float(m.tot[2 * l]) * inv_n, float(m.tot[2 * l + 1]) * inv_n);
        }
        __syncthreads();
      } else if (P.smoothing == SRSCE_SMOOTH_FILTER) {
        if (tid < 2 * L && nv > 0) {
          const int l = tid >> 1;
          if ((tid & 1) == 0) {
            virtual_pilots(nv, P.vp_mx, P.vp_den,
                           [&](int k) { return ext[(nv + k) * L + l]; },
                           [&](int k, float2 val) { ext[k * L + l] = val; });
          } else {
            virtual_pilots(nv, P.vp_mx, P.vp_den,
                           [&](int k) { return ext[(nv + N - 1 - k) * L + l]; },
                           [&](int k, float2 val) { ext[(nv + N + (nv - 1 - k)) * L + l] = val; });
          }
        }
        __syncthreads();
        // conv(ext, taps, "same") cropped to the real pilots, float64 accumulation (:459-493)
        const int pad = K >> 1, E = N + 2 * nv;
        for (int idx = tid; idx < N * L; idx += kThreads) {
          const int i = idx / L, l = idx - i * L;
          double ar = 0.0, ai = 0.0;
          for (int k = 0; k < K; ++k) {
            const int q = i + nv + pad - k;
            if (q >= 0 && q < E) {
              const float2 t = ext[q * L + l];
              const double w = m.taps[k];
              ar += w * double(t.x);
              ai += w * double(t.y);
            }
          }
          hbuf[idx] = make_float2(float(ar), float(ai));
        }
        __syncthreads();
        if (P.blend && N > 2) {                      // ce_dl_cnn.py:712-717
          for (int idx = tid; idx < N * L; idx += kThreads) {
            const int i = idx / L, l = idx - i * L;
            const int il = (i == 0) ? 1 : i - 1, ir = (i == N - 1) ? N - 2 : i + 1;
            const float2 a = hbuf[il * L + l], b = hbuf[idx], c = hbuf[ir * L + l];
            ext[idx] = make_float2(float(0.25 * double(a.x) + 0.5 * double(b.x) + 0.25 * double(c.x)),
                                   float(0.25 * double(a.y) + 0.5 * double(b.y) + 0.25 * double(c.y)));
          }
          __syncthreads();
          for (int idx = tid; idx < N * L; idx += kThreads) {
            const float2 s0 = hbuf[idx], lp = ext[idx];
            hbuf[idx] = make_float2(s0.x + P.cnn_alpha * (lp.x - s0.x), s0.y + P.cnn_alpha * (lp.y - s0.y));
          }
          __syncthreads();
        }
      }

      // ---- time alignment (:674-698)
      {
        const int M = P.fft_size, log2M = 31 - __clz(M);
        const int* posl = m.re_pos[G - 1];
        for (int lb = 0; lb < L; lb += P.fft_layers) {
          const int nl = min(P.fft_layers, L - lb);
          if (P.fft_regular) {
            for (int idx = tid; idx < nl * M; idx += kThreads) {
              const int l = idx >> log2M, n = idx & (M - 1);
              ext[idx] = (n < N) ? hbuf[n * L + lb + l] : make_float2(0.f, 0.f);
            }
          } else {
            for (int idx = tid; idx < nl * M; idx += kThreads) ext[idx] = make_float2(0.f, 0.f);
            __syncthreads();
            for (int idx = tid; idx < nl * N; idx += kThreads) {
              const int l = idx / N, i = idx - l * N;
              const int q = i / R;
              ext[l * M + q * kNRE + posl[i - q * R] - posl[0]] = hbuf[i * L + lb + l];
            }
          }
          __syncthreads();
          fft_dif_inverse(ext, nl, M, log2M, P.twiddle);
          for (int t = tid; t < 2 * SRSCE_TA_WINDOW; t += kThreads) {
            const int bin = (t < SRSCE_TA_WINDOW) ? t : SRSCE_TA_FFT - 2 * SRSCE_TA_WINDOW + t;
            const int pos = dif_position(bin & (M - 1), M);
            float acc = (lb == 0) ? 0.f : m.prof[t];
            for (int l = 0; l < nl; ++l) {
              const float2 z = ext[l * M + pos];
              acc += z.x * z.x + z.y * z.y;
            }
            m.prof[t] = acc;
          }
          __syncthreads();
        }
        if (warp == 0) {
          // first maximum of each window; ties keep the lower index, as torch.max / np.argmax
          float bd = -1.f, ba = -1.f;
          int id = 0, ia = 0;
          for (int t = lane; t < SRSCE_TA_WINDOW; t += 32) {
            const float d = m.prof[t], a = m.prof[SRSCE_TA_WINDOW + t];
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
            m.acc[3] += double(imax) / double(SRSCE_TA_FFT) / P.scs_hz;
          }
        }
      }

      // ---- reconstruction, noise, RSRP (:701-730)
      {
        float noise_part = 0.f, rsrp_part = 0.f;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (g < G) {
            for (int i = tid; i < N; i += kThreads) {
              const long long rs = (long long)pilot_sc(m, sc0, R, g, i) * ssc;
              const float2* pp = pu + ((long long)i * ST + col0) * L + 2 * g;
              const float2 h0 = hbuf[i * L + 2 * g];
              const float2 h1 = (2 * g + 1 < L) ? hbuf[i * L + 2 * g + 1] : make_float2(0.f, 0.f);
              for (int s = 0; s < S; ++s) {
                const float2 y = __ldg(rxu + rs + (long long)H.sym_idx[s] * ssym);
                float2 e = cmul(__ldg(pp + s * L), derot ? cmul(h0, m.rot_pos[s]) : h0);
                if (2 * g + 1 < L) e = cadd(e, cmul(__ldg(pp + s * L + 1), derot ? cmul(h1, m.rot_pos[s]) : h1));
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
        double v[3] = {double(epre_part), double(noise_part), double(rsrp_part)};
        block_sum<3>(m, v);
        if (tid == 0) {
          m.acc[0] += double(float(m.tot[0]));
          m.acc[1] += double(float(m.tot[1]));
          m.acc[2] += beta_d * beta_d * double(float(m.tot[2])) * double(S);
        }
      }
      __syncthreads();
    }  // hops

    // ---- per-unit scalars (:899-919, :932-937) and the store-slot table of phase B
    if (tid == 0) {
      double* o = scalars + u * 6;
      o[0] = m.acc[1] * P.norm_noise;
      o[1] = m.acc[2] * P.norm_rsrp;
      o[2] = m.acc[0] * P.norm_epre;
      o[3] = m.acc[3] / P.ta_hop_div;
      o[4] = (m.acc[5] != 0.0) ? m.acc[4] * P.scs_hz : 0.0;
      o[5] = m.acc[5];
    }
    for (int c = tid; c < P.cols; c += kThreads) {
      ColInfo ci;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int el = min(c * VEC + e, P.row_elems - 1);
        const int sym = el / L;
        ci.layer[e] = el - sym * L;
        int f = (sym >= P.hop[0].start_symbol && sym < P.hop[0].end_symbol) ? 1 : 0;
        if (TWO_HOPS && sym >= P.hop[1].start_symbol && sym < P.hop[1].end_symbol) f |= 2;
        ci.flags[e] = f;
        double sn = 0.0, cs = 1.0;
        if (HAS_ROT) sincos(2.0 * M_PI * P.t_sym[sym] * m.acc[4], &sn, &cs);   // :922-929
        ci.rot[e] = make_float2(float(cs), float(sn));
      }
      m.col[c] = ci;
    }
    __syncthreads();

    // ===================================== phase B ==========================================
    // Each warp takes 32 consecutive subcarriers: lane r interpolates row r into the warp's tile,
    // then the warp streams the rows out, every lane owning a fixed store slot (symbol, layers).
    {
      float2* const tile0 = ext + warp * (kTileRows * L * n_hops);
      float2* const outu = out + u * ((long long)P.n_sc * P.row_elems);
      const int cols = P.cols;
      const bool fast = cols <= 32;
      const int rpi = fast ? 32 / cols : 1;
      const int my_rs = fast ? lane / cols : 0;
      const int my_c = fast ? lane - my_rs * cols : 0;
      const bool lane_on = fast ? (my_rs < rpi) : true;
      ColInfo mine = m.col[my_c];

      for (int row0 = warp * kTileRows; row0 < P.n_sc; row0 += kWarps * kTileRows) {
        const int row = row0 + lane;
        unsigned in2 = 0;
        __syncwarp();
#pragma unroll
        for (int hh = 0; hh < n_hops; ++hh) {
          const HopDev& H = P.hop[hh];
          float2* const hb = hall + hh * N * L;
          float2* const trow = tile0 + (hh * kTileRows + lane) * L;
          const bool inband = row >= H.band_sc0 && row < H.band_sc1;
          if (hh == 1) in2 = __ballot_sync(0xffffffffu, inband);
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
                if (P.fill == SRSCE_FILL_LINEAR) {
                  const LerpEntry e = m.lerp[g][pos];
                  const int il = max(0, min(N - 1, q * R + e.jl)), ir = max(0, min(N - 1, q * R + e.jr));
                  const float2 l0 = hb[il * L + 2 * g], r0 = hb[ir * L + 2 * g];
                  v0 = make_float2(l0.x + e.alpha * (r0.x - l0.x), l0.y + e.alpha * (r0.y - l0.y));
                  if (two) {
                    const float2 l1 = hb[il * L + 2 * g + 1], r1 = hb[ir * L + 2 * g + 1];
                    v1 = make_float2(l1.x + e.alpha * (r1.x - l1.x), l1.y + e.alpha * (r1.y - l1.y));
                  }
                } else {
                  const int T = P.fill_terms;
                  const int bands = H.band_sc1 - H.band_sc0;
                  const int k0 = __ldg(P.fill_idx + g * bands + br);
                  const float* cf = P.fill_coef + ((long long)g * bands + br) * T;
                  v0 = make_float2(0.f, 0.f);
                  for (int t = 0; t < T; ++t) {
                    const float w = __ldg(cf + t);
                    const float2 a = hb[(k0 + t) * L + 2 * g];
                    v0.x += w * a.x;
                    v0.y += w * a.y;
                    if (two) {
                      const float2 b = hb[(k0 + t) * L + 2 * g + 1];
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
        __syncwarp();

        const int rows_here = min(kTileRows, P.n_sc - row0);
        if (fast) {
          if (lane_on) {
            for (int r = my_rs; r < rows_here; r += rpi) {
              float2 val[2];
#pragma unroll
              for (int e = 0; e < VEC; ++e) {
                float2 t = make_float2(0.f, 0.f);
                if (TWO_HOPS && (mine.flags[e] & 2) && ((in2 >> r) & 1u)) t = tile0[(kTileRows + r) * L + mine.layer[e]];
                else if (mine.flags[e] & 1) t = tile0[r * L + mine.layer[e]];
                val[e] = HAS_ROT ? cmul(t, mine.rot[e]) : t;
              }
              float2* dst = outu + (long long)(row0 + r) * P.row_elems + my_c * VEC;
              if (VEC == 2) __stcs(reinterpret_cast<float4*>(dst), make_float4(val[0].x, val[0].y, val[1].x, val[1].y));
              else __stcs(dst, val[0]);
            }
          }
        } else {
          for (int r = 0; r < rows_here; ++r) {
            for (int c = lane; c < cols; c += 32) {
              const ColInfo ci = m.col[c];
              float2 val[2];
#pragma unroll
              for (int e = 0; e < VEC; ++e) {
                float2 t = make_float2(0.f, 0.f);
                if (TWO_HOPS && (ci.flags[e] & 2) && ((in2 >> r) & 1u)) t = tile0[(kTileRows + r) * L + ci.layer[e]];
                else if (ci.flags[e] & 1) t = tile0[r * L + ci.layer[e]];
                val[e] = HAS_ROT ? cmul(t, ci.rot[e]) : t;
              }
              float2* dst = outu + (long long)(row0 + r) * P.row_elems + c * VEC;
              if (VEC == 2) __stcs(reinterpret_cast<float4*>(dst), make_float4(val[0].x, val[0].y, val[1].x, val[1].y));
              else __stcs(dst, val[0]);
            }
          }
        }
      }
    }
    __syncthreads();   // tiles alias ext, m.col / m.acc are rewritten by the next unit
  }
}

typedef void (*KernelFn)(const PlanDev, const float2*, long long, long long, long long, const float2*, int,
                         double, float2*, double*, long long);

KernelFn pick_kernel(int vec, bool two_hops, bool has_rot) {
  if (vec == 2) {
    if (two_hops) return has_rot ? estimate_kernel<2, true, true> : estimate_kernel<2, true, false>;
    return has_rot ? estimate_kernel<2, false, true> : estimate_kernel<2, false, false>;
  }
  if (two_hops) return has_rot ? estimate_kernel<1, true, true> : estimate_kernel<1, true, false>;
  return has_rot ? estimate_kernel<1, false, true> : estimate_kernel<1, false, false>;
}

std::string cuda_msg(const char* what, cudaError_t e) {
  return std::string(what) + ": " + cudaGetErrorString(e);
}

}  // namespace

int max_dynamic_smem() { return 227 * 1024; }

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

  std::vector<float2> tw(SRSCE_TA_FFT);
  for (int k = 0; k < SRSCE_TA_FFT; ++k) {
    const double a = 2.0 * M_PI * double(k) / double(SRSCE_TA_FFT);
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
  KernelFn fn = pick_kernel(p.dev.vec, p.dev.n_hops == 2, p.dev.cfo_compensate && p.dev.cfo_defined);
  // the limit is per kernel variant, not per plan: always raise it to the architectural maximum so that
  // plans with different footprints can share a variant
  if ((e = cudaFuncSetAttribute(reinterpret_cast<const void*>(fn), cudaFuncAttributeMaxDynamicSharedMemorySize, max_dynamic_smem())) != cudaSuccess) {
    err = cuda_msg("cudaFuncSetAttribute(smem)", e);
    return SRSCE_ERR_CUDA;
  }
  int per_sm = 0, sms = 0;
  if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, reinterpret_cast<const void*>(fn), kThreads, p.dev.smem_bytes)) != cudaSuccess ||
      (e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) {
    err = cuda_msg("occupancy query", e);
    return SRSCE_ERR_CUDA;
  }
  if (per_sm < 1) { err = "fused kernel does not fit on an SM"; return SRSCE_ERR_UNSUPPORTED; }
  p.ctas_per_sm = per_sm;
  p.n_sms = sms;
  p.device = dev;
  return SRSCE_OK;
}

void release_plan(Plan& p) {
  if (p.d_twiddle) cudaFree(p.d_twiddle);
  if (p.d_fill_idx) cudaFree(p.d_fill_idx);
  if (p.d_fill_coef) cudaFree(p.d_fill_coef);
  p.d_twiddle = p.d_fill_idx = p.d_fill_coef = nullptr;
}

int launch_estimate(const Plan& p, const void* rx, const int64_t strides[3], const void* pilots, int pilot_share,
                    double beta, void* out, double* scalars, int64_t n_units, void* stream, int* launches,
                    std::string& err) {
  *launches = 0;
  if (p.device < 0) { err = "plan has no CUDA device (created without a GPU); there is no CPU compute path"; return SRSCE_ERR_NO_DEVICE; }
  int dev = -1;
  cudaGetDevice(&dev);
  if (dev != p.device) { err = "plan was created on another CUDA device"; return SRSCE_ERR_INVALID; }
  if (n_units == 0) return SRSCE_OK;
  KernelFn fn = pick_kernel(p.dev.vec, p.dev.n_hops == 2, p.dev.cfo_compensate && p.dev.cfo_defined);
  const long long cap = (long long)p.n_sms * p.ctas_per_sm;
  const unsigned grid = unsigned(n_units < cap ? n_units : cap);
  fn<<<grid, kThreads, p.dev.smem_bytes, static_cast<cudaStream_t>(stream)>>>(
      p.dev, static_cast<const float2*>(rx), (long long)strides[0], (long long)strides[1], (long long)strides[2],
      static_cast<const float2*>(pilots), pilot_share, beta, static_cast<float2*>(out), scalars, (long long)n_units);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { err = cuda_msg("kernel launch", e); return SRSCE_ERR_CUDA; }
  *launches = 1;
  return SRSCE_OK;
}

}  // namespace srsce
