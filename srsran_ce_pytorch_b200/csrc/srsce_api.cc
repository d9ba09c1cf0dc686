// C ABI glue (include/srsce.h): argument checks, error strings, plan lifetime.  No C++ exception
// leaves this file.
#include <stdio.h>
#include <string.h>

#include <new>
#include <string>

#include "srsce_internal.h"

struct srsce_plan {
  srsce::Plan impl;
};

namespace {
thread_local std::string g_error;
thread_local int g_launches = 0;

int fail(int code, const std::string& msg) {
  g_error = msg;
  return code;
}
}  // namespace

extern "C" {

int srsce_abi_version(void) { return SRSCE_ABI_VERSION; }

const char* srsce_last_error(void) { return g_error.c_str(); }

int srsce_plan_create(const srsce_plan_desc* desc, srsce_plan** out) {
  if (!desc || !out) return fail(SRSCE_ERR_INVALID, "null argument");
  *out = nullptr;
  try {
    srsce_plan* p = new srsce_plan();
    std::string err;
    int rc = srsce::build_plan(*desc, p->impl, err);
    if (rc == SRSCE_OK) rc = srsce::upload_plan(p->impl, err);
    if (rc != SRSCE_OK) {
      srsce::release_plan(p->impl);
      delete p;
      return fail(rc, err);
    }
    *out = p;
    return SRSCE_OK;
  } catch (const std::exception& e) {
    return fail(SRSCE_ERR_INVALID, e.what());
  }
}

void srsce_plan_destroy(srsce_plan* plan) {
  if (!plan) return;
  srsce::release_plan(plan->impl);
  delete plan;
}

int srsce_plan_query(const srsce_plan* plan, srsce_plan_info* info) {
  if (!plan || !info) return fail(SRSCE_ERR_INVALID, "null argument");
  const srsce::PlanDev& v = plan->impl.dev;
  memset(info, 0, sizeof *info);
  info->n_sc = v.n_sc;
  info->n_sym = v.n_sym;
  info->n_layers = v.n_layers;
  info->n_cdm = v.n_cdm;
  info->n_re = v.n_re;
  info->n_dmrs_total = v.n_dmrs_total;
  info->n_taps = v.n_taps;
  info->n_virtual = v.n_virtual;
  info->fft_size = v.fft_size;
  info->cfo_defined = v.cfo_defined;
  info->smem_bytes = v.smem_bytes;
  info->threads = srsce::cta_threads(v.est_threads, v.n_hops == 2);
  info->fill_terms = plan->impl.fill_terms_host;
  info->has_device = plan->impl.device >= 0;
  info->bytes_per_unit = plan->impl.bytes_per_unit;
  info->fill_tensor = plan->impl.has_fill_tc ? 1 : 0;
  info->fill_tensor_k = plan->impl.fill_tc.k_tile;
  info->fill_tensor_terms = plan->impl.fill_tc.n_prod;
  info->fill_tensor_units = plan->impl.fill_tc.units;
  info->h_in_global = v.h_global;
  info->fill_tensor_tiles = plan->impl.fill_tc.n_tiles;
  info->fill_tensor_kinds = plan->impl.fill_tc.n_kinds;
  info->fill_tensor_a_terms = plan->impl.fill_tc.n_a;
  return SRSCE_OK;
}

int srsce_plan_tables(const srsce_plan* plan, double* taps, int32_t* fill_idx, float* fill_coef) {
  if (!plan) return fail(SRSCE_ERR_INVALID, "null argument");
  const srsce::Plan& p = plan->impl;
  if (taps) memcpy(taps, p.taps.data(), p.taps.size() * sizeof(double));
  if (fill_idx) memcpy(fill_idx, p.fill_idx.data(), p.fill_idx.size() * sizeof(int32_t));
  if (fill_coef) memcpy(fill_coef, p.fill_coef.data(), p.fill_coef.size() * sizeof(float));
  return SRSCE_OK;
}

int srsce_plan_fill_tensor_tables(const srsce_plan* plan, uint16_t* a_tiles, int32_t* tile_kind, int32_t* tile_k0) {
  if (!plan) return fail(SRSCE_ERR_INVALID, "null argument");
  const srsce::Plan& p = plan->impl;
  if (!p.has_fill_tc) return fail(SRSCE_ERR_UNSUPPORTED, "tensor-core fill unavailable for this plan: " + p.fill_tc_why);
  if (a_tiles) memcpy(a_tiles, p.fill_tc_a.data(), p.fill_tc_a.size() * sizeof(uint16_t));
  for (int m = 0; m < p.fill_tc.n_tiles; ++m) {
    if (tile_kind) tile_kind[m] = p.fill_tc.tile_kind[m];
    if (tile_k0) tile_k0[m] = p.fill_tc.tile_k0[m];
  }
  return SRSCE_OK;
}

int srsce_estimate(const srsce_plan* plan, const void* rx_grid, const int64_t rx_strides[3], const void* pilots,
                   int32_t pilot_share, double beta, void* ch_est, double* scalars, int64_t n_units, void* stream) {
  g_launches = 0;
  if (!plan || !rx_strides) return fail(SRSCE_ERR_INVALID, "null argument");
  if (n_units < 0) return fail(SRSCE_ERR_INVALID, "negative unit count");
  if (n_units > 0 && (!rx_grid || !pilots || !ch_est || !scalars)) return fail(SRSCE_ERR_INVALID, "null buffer");
  if (pilot_share < 1 || (n_units % pilot_share) != 0)
    return fail(SRSCE_ERR_INVALID, "n_units must be a multiple of pilot_share (>= 1)");
  if (!(beta != 0.0)) return fail(SRSCE_ERR_INVALID, "beta_dmrs must be non-zero");
  if ((reinterpret_cast<uintptr_t>(ch_est) & 15u) || (reinterpret_cast<uintptr_t>(rx_grid) & 7u) ||
      (reinterpret_cast<uintptr_t>(pilots) & 7u) || (reinterpret_cast<uintptr_t>(scalars) & 7u))
    return fail(SRSCE_ERR_INVALID, "buffers must be aligned (ch_est 16 B, others 8 B)");
  try {
    if ((plan->impl.dev.n_layers & 1) == 0 && (reinterpret_cast<uintptr_t>(pilots) & 15u))
      return fail(SRSCE_ERR_INVALID, "pilots must be 16-byte aligned when the layer count is even (the kernel loads layer pairs)");
    std::string err;
    const int rc = srsce::launch_estimate(plan->impl, rx_grid, rx_strides, pilots, pilot_share, beta, ch_est, scalars,
                                          n_units, stream, &g_launches, err);
    return rc == SRSCE_OK ? rc : fail(rc, err);
  } catch (const std::exception& e) {
    return fail(SRSCE_ERR_INVALID, e.what());
  }
}

int srsce_estimate_compact(const srsce_plan* plan, const void* rx_grid, const int64_t rx_strides[3], const void* pilots,
                           int32_t pilot_share, double beta, void* ch_sc, void* sym_rot, double* scalars,
                           int64_t n_units, void* stream) {
  g_launches = 0;
  if (!plan || !rx_strides) return fail(SRSCE_ERR_INVALID, "null argument");
  if (n_units < 0) return fail(SRSCE_ERR_INVALID, "negative unit count");
  if (n_units > 0 && (!rx_grid || !pilots || !ch_sc || !sym_rot || !scalars)) return fail(SRSCE_ERR_INVALID, "null buffer");
  if (pilot_share < 1 || (n_units % pilot_share) != 0)
    return fail(SRSCE_ERR_INVALID, "n_units must be a multiple of pilot_share (>= 1)");
  if (!(beta != 0.0)) return fail(SRSCE_ERR_INVALID, "beta_dmrs must be non-zero");
  if ((reinterpret_cast<uintptr_t>(ch_sc) & 7u) || (reinterpret_cast<uintptr_t>(sym_rot) & 7u) ||
      (reinterpret_cast<uintptr_t>(rx_grid) & 7u) || (reinterpret_cast<uintptr_t>(pilots) & 7u) ||
      (reinterpret_cast<uintptr_t>(scalars) & 7u))
    return fail(SRSCE_ERR_INVALID, "buffers must be 8-byte aligned");
  try {
    if ((plan->impl.dev.n_layers & 1) == 0 && (reinterpret_cast<uintptr_t>(pilots) & 15u))
      return fail(SRSCE_ERR_INVALID, "pilots must be 16-byte aligned when the layer count is even (the kernel loads layer pairs)");
    std::string err;
    const int rc = srsce::launch_estimate(plan->impl, rx_grid, rx_strides, pilots, pilot_share, beta, nullptr, scalars,
                                          n_units, stream, &g_launches, err, ch_sc, sym_rot);
    return rc == SRSCE_OK ? rc : fail(rc, err);
  } catch (const std::exception& e) {
    return fail(SRSCE_ERR_INVALID, e.what());
  }
}

int srsce_estimate_indexed(const srsce_plan* plan, const void* rx_grids, const int64_t rx_strides[3],
                           const int64_t* rx_index, const void* pilots, int32_t pilot_share, double beta, void* ch_est,
                           double* scalars, int64_t n_units, void* stream) {
  g_launches = 0;
  if (!plan || !rx_strides) return fail(SRSCE_ERR_INVALID, "null argument");
  if (n_units < 0) return fail(SRSCE_ERR_INVALID, "negative unit count");
  if (n_units > 0 && (!rx_grids || !rx_index || !pilots || !ch_est || !scalars)) return fail(SRSCE_ERR_INVALID, "null buffer");
  if (pilot_share < 1 || (n_units % pilot_share) != 0)
    return fail(SRSCE_ERR_INVALID, "n_units must be a multiple of pilot_share (>= 1)");
  if (!(beta != 0.0)) return fail(SRSCE_ERR_INVALID, "beta_dmrs must be non-zero");
  if ((reinterpret_cast<uintptr_t>(ch_est) & 15u) || (reinterpret_cast<uintptr_t>(rx_grids) & 7u) ||
      (reinterpret_cast<uintptr_t>(pilots) & 7u) || (reinterpret_cast<uintptr_t>(scalars) & 7u) ||
      (reinterpret_cast<uintptr_t>(rx_index) & 7u))
    return fail(SRSCE_ERR_INVALID, "buffers must be aligned (ch_est 16 B, others 8 B)");
  try {
    if ((plan->impl.dev.n_layers & 1) == 0 && (reinterpret_cast<uintptr_t>(pilots) & 15u))
      return fail(SRSCE_ERR_INVALID, "pilots must be 16-byte aligned when the layer count is even (the kernel loads layer pairs)");
    std::string err;
    const int rc = srsce::launch_estimate(plan->impl, rx_grids, rx_strides, pilots, pilot_share, beta, ch_est, scalars,
                                          n_units, stream, &g_launches, err, nullptr, nullptr, rx_index);
    return rc == SRSCE_OK ? rc : fail(rc, err);
  } catch (const std::exception& e) {
    return fail(SRSCE_ERR_INVALID, e.what());
  }
}

int srsce_fill(const srsce_plan* plan, const void* h_pilots, void* ch_est, int64_t n_units, int32_t impl, void* stream) {
  g_launches = 0;
  if (!plan) return fail(SRSCE_ERR_INVALID, "null argument");
  if (n_units < 0) return fail(SRSCE_ERR_INVALID, "negative unit count");
  if (n_units > 0 && (!h_pilots || !ch_est)) return fail(SRSCE_ERR_INVALID, "null buffer");
  if (impl != SRSCE_FILL_IMPL_DIRECT && impl != SRSCE_FILL_IMPL_TENSOR) return fail(SRSCE_ERR_INVALID, "unknown fill implementation");
  if ((reinterpret_cast<uintptr_t>(ch_est) & 15u) || (reinterpret_cast<uintptr_t>(h_pilots) & 7u))
    return fail(SRSCE_ERR_INVALID, "buffers must be aligned (ch_est 16 B, h_pilots 8 B)");
  try {
    std::string err;
    const int rc = impl == SRSCE_FILL_IMPL_TENSOR
                       ? srsce::launch_fill_tc(plan->impl, h_pilots, ch_est, n_units, stream, &g_launches, err)
                       : srsce::launch_fill_direct(plan->impl, h_pilots, ch_est, n_units, stream, &g_launches, err);
    return rc == SRSCE_OK ? rc : fail(rc, err);
  } catch (const std::exception& e) {
    return fail(SRSCE_ERR_INVALID, e.what());
  }
}

int srsce_gen_pilots(void* pilots, int64_t n_ue, int32_t m_zc, int32_t n_dmrs, int32_t n_layers, const int32_t* u_of,
                     const int32_t* v_of, void* stream) {
  g_launches = 0;
  if (n_ue < 0) return fail(SRSCE_ERR_INVALID, "negative UE count");
  if (n_ue > 0 && (!pilots || !u_of)) return fail(SRSCE_ERR_INVALID, "null buffer");
  if (n_dmrs < 1 || n_dmrs > SRSCE_MAX_SYMBOLS || n_layers < 1 || n_layers > SRSCE_MAX_LAYERS)
    return fail(SRSCE_ERR_INVALID, "n_dmrs must be 1..14 and n_layers 1..4");
  if (m_zc != 30 && m_zc < 36)
    return fail(SRSCE_ERR_UNSUPPORTED, "sequence lengths 6, 12, 18, 24 are table defined in TS 38.211 5.2.2.2 (tables not included); supported: 30 and >= 36");
  if (m_zc > SRSCE_TA_FFT) return fail(SRSCE_ERR_UNSUPPORTED, "sequence longer than 4096");
  if (reinterpret_cast<uintptr_t>(pilots) & 7u) return fail(SRSCE_ERR_INVALID, "pilots must be 8-byte aligned");
  std::string err;
  const int rc = srsce::launch_gen_pilots(pilots, n_ue, m_zc, n_dmrs, n_layers, u_of, v_of, stream, err);
  if (rc == SRSCE_OK && n_ue > 0) g_launches = 1;
  return rc == SRSCE_OK ? rc : fail(rc, err);
}

int srsce_fill_compact(const srsce_plan* plan, const void* h_pilots, void* ch_sc, int64_t n_units, void* stream) {
  g_launches = 0;
  if (!plan) return fail(SRSCE_ERR_INVALID, "null argument");
  if (n_units < 0) return fail(SRSCE_ERR_INVALID, "negative unit count");
  if (n_units > 0 && (!h_pilots || !ch_sc)) return fail(SRSCE_ERR_INVALID, "null buffer");
  if ((reinterpret_cast<uintptr_t>(ch_sc) & 7u) || (reinterpret_cast<uintptr_t>(h_pilots) & 7u))
    return fail(SRSCE_ERR_INVALID, "buffers must be 8-byte aligned");
  try {
    std::string err;
    const int rc = srsce::launch_fill_direct(plan->impl, h_pilots, nullptr, n_units, stream, &g_launches, err, ch_sc);
    return rc == SRSCE_OK ? rc : fail(rc, err);
  } catch (const std::exception& e) {
    return fail(SRSCE_ERR_INVALID, e.what());
  }
}

int srsce_replicate_host(const srsce_plan* plan, const void* ch_sc, void* ch_est, int64_t n_units, int32_t n_threads,
                         int32_t flags) {
  if (!plan) return fail(SRSCE_ERR_INVALID, "null argument");
  if (n_units < 0) return fail(SRSCE_ERR_INVALID, "negative unit count");
  if (n_units > 0 && (!ch_sc || !ch_est)) return fail(SRSCE_ERR_INVALID, "null buffer");
  if ((reinterpret_cast<uintptr_t>(ch_sc) & 7u) || (reinterpret_cast<uintptr_t>(ch_est) & 7u))
    return fail(SRSCE_ERR_INVALID, "buffers must be 8-byte aligned");
  try {
    std::string err;
    const int rc = srsce::replicate_host(plan->impl, ch_sc, ch_est, n_units, n_threads, flags, err);
    return rc == SRSCE_OK ? rc : fail(rc, err);
  } catch (const std::exception& e) {
    return fail(SRSCE_ERR_INVALID, e.what());
  }
}

int srsce_last_launch_count(void) { return g_launches; }

}  // extern "C"
