// Pilot sequence generation on the device -- an EXTENSION: the reference takes the pilots as an input
// tensor (ce_rule_tensorized.py:746-747) and contains no sequence generator, so there is nothing to be bit-compatible
// with ("parity unpinned").  What is generated is the low-PAPR sequence of 3GPP TS 38.211 section 5.2.2 as the SRS of
// section 6.4.1.4.2 uses it, in the layout the estimator reads ([UE][pilot][pilot symbol][layer]):
//
//   r_{u,v}^{(alpha)}(n) = e^{j alpha n} * rbar_{u,v}(n),   0 <= n < M
//   M >= 36:  rbar(n) = x_q(n mod N_zc),  x_q(m) = exp(-j pi q m (m + 1) / N_zc),  N_zc = largest prime < M,
//             qbar = N_zc (u + 1) / 31,  q = floor(qbar + 1/2) + v (-1)^floor(2 qbar)
//   M == 30:  rbar(n) = exp(-j pi (u + 1)(n + 1)(n + 2) / 31)
//   (M in {6, 12, 18, 24} are table defined in the specification; the tables are not reproduced here -> unsupported.)
//
// Layer l of a UE sits on comb (CDM group) l / 2 with cyclic shift alpha = pi (l % 2): two ports per comb, half of
// n_cs_max apart -- exactly the OCC-2 pairs the estimator separates (:619-628).  The phase index q m (m + 1) mod 2 N_zc
// is formed in 64-bit integers, so the argument handed to sincospi is exact.
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "srsce_internal.h"

namespace srsce {
namespace {

__global__ void gen_low_papr_kernel(float2* __restrict__ out, long long n_ue, int m_zc, int n_zc, int n_dmrs, int n_layers,
                                    const int32_t* __restrict__ u_of, const int32_t* __restrict__ v_of) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;       // (ue, n)
  if (idx >= n_ue * m_zc) return;
  const long long ue = idx / m_zc;
  const int n = int(idx - ue * m_zc);
  const int u = u_of[ue], v = v_of ? v_of[ue] : 0;
  double sn, cs;
  if (m_zc == 30) {
    const long long k = ((long long)(u + 1) * (n + 1) * (n + 2)) % 62;           // phase = -pi k / 31, period 2 * 31
    sincospi(-double(k) / 31.0, &sn, &cs);
  } else {
    const double qbar = double(n_zc) * double(u + 1) / 31.0;
    const long long q = (long long)floor(qbar + 0.5) + (((long long)floor(2.0 * qbar) & 1) ? -v : v);
    const long long m = n % n_zc;
    const long long k = (q % (2LL * n_zc)) * ((m * (m + 1)) % (2LL * n_zc)) % (2LL * n_zc);
    sincospi(-double(k) / double(n_zc), &sn, &cs);
  }
  const float2 base = make_float2(float(cs), float(sn));
  const float2 odd = (n & 1) ? make_float2(-base.x, -base.y) : base;           // e^{j pi n} * rbar(n)
  float2* o = out + idx * ((long long)n_dmrs * n_layers);
  for (int s = 0; s < n_dmrs; ++s)
    for (int l = 0; l < n_layers; ++l) o[s * n_layers + l] = (l & 1) ? odd : base;
}

}  // namespace

int largest_prime_below(int m) {
  for (int p = m - 1; p >= 2; --p) {
    bool prime = true;
    for (int d = 2; d * d <= p && prime; ++d) prime = (p % d) != 0;
    if (prime) return p;
  }
  return 0;
}

int launch_gen_pilots(void* pilots, int64_t n_ue, int32_t m_zc, int32_t n_dmrs, int32_t n_layers, const int32_t* u_of,
                      const int32_t* v_of, void* stream, std::string& err) {
  if (n_ue == 0) return SRSCE_OK;
  const int n_zc = (m_zc == 30) ? 31 : largest_prime_below(m_zc);
  const long long total = n_ue * m_zc;
  const unsigned blocks = unsigned((total + 255) / 256);
  gen_low_papr_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<float2*>(pilots), n_ue, m_zc, n_zc,
                                                                              n_dmrs, n_layers, u_of, v_of);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { err = std::string("gen_pilots launch: ") + cudaGetErrorString(e); return SRSCE_ERR_CUDA; }
  return SRSCE_OK;
}

}  // namespace srsce
