// Host-side planner: everything about one estimator call that does not depend on the received
// samples is computed once, here, in float64 -- raised-cosine taps, virtual-pilot regression
// constants, CFO time tables, interpolation tables, the fixed-stencil "CNN" inpainting operator,
// FFT geometry and the shared-memory carve-up of the fused kernel.
//
// Reference behaviour restated (all citations into /root/reference/src):
//   get_rc_filter / _rcosdesign_normal      ce_rule_tensorized.py:184-234, :143-181
//   virtual-pilot count                     :644-647
//   create_virtual_pilots constants         :105-118
//   CFO sample count, symbol start times    :410-426, :809-820
//   fill_ch_est_cdm (linear)                :311-340
//   ce_dl_cnn inpainting + low-pass         ce_dl_cnn.py:433-508, :274-295
//   final normalisation                     ce_rule_tensorized.py:899-919
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>

#include "srsce_internal.h"

namespace srsce {

static int popcount12(uint32_t m) {
  int c = 0;
  for (int k = 0; k < kNRE; ++k) c += (m >> k) & 1u;
  return c;
}

// Raised cosine, roll-off 0.2, span n_rbs symbols at 10 samples/symbol; the centre tap and every
// stride-th sample either side are kept and the result is normalised to unit sum.
void rc_taps(int stride, int n_rbs, std::vector<double>& taps) {
  const int sps = 10;
  const double beta = 0.2;
  const int half = (n_rbs * sps) / 2;
  const int len = 2 * half + 1;
  std::vector<double> h(len);
  const double t0 = 1.0 / (2.0 * beta);
  for (int k = 0; k < len; ++k) {
    const double t = double(k - half) / double(sps);
    double v;
    if (fabs(fabs(t) - t0) < (1.0 / sps) * 1e-6) {
      v = (M_PI * beta / 2.0) * sin(1.0 / (2.0 * beta));
    } else {
      const double sinc = (t == 0.0) ? 1.0 : sin(M_PI * t) / (M_PI * t);
      v = sinc * cos(M_PI * beta * t) / (1.0 - (2.0 * beta * t) * (2.0 * beta * t));
    }
    h[k] = v;
  }
  const int reach = ((len / 2) / stride) * stride;
  taps.clear();
  double sum = 0.0;
  for (int k = half - reach; k <= half + reach; k += stride) {
    taps.push_back(h[k]);
    sum += h[k];
  }
  for (double& v : taps) v /= sum;
}

// ---------------------------------------------------------------------------------------------
// The ce_dl_cnn "CNN": partial-convolution inpainting with the fixed stencil [.25 .5 .25],
// known samples re-imposed every iteration, then two low-pass passes kept at the unknown
// positions only.  It is linear in the known samples and -- because known samples are clamped
// every iteration -- strictly local: an unknown run only ever sees the known sample(s) that
// bound it.  So the operator is built exactly, run by run, by iterating the reference recurrence
// in float64 on the run alone, and then composing with the two (reflect-padded) low-pass passes.
// ---------------------------------------------------------------------------------------------
typedef std::map<int, double> SparseRow;   // known-sample ordinal -> coefficient

static void stencil_reflect(const std::vector<double>& v, std::vector<double>& out) {
  const int n = int(v.size());
  out.resize(n);
  for (int i = 0; i < n; ++i) {
    double l, r;
    if (n == 1) {
      l = r = v[0];
    } else {
      l = (i == 0) ? v[1] : v[i - 1];
      r = (i == n - 1) ? v[n - 2] : v[i + 1];
    }
    out[i] = 0.25 * l + 0.5 * v[i] + 0.25 * r;
  }
}

// Response of the unknown samples of a local segment to a unit value at local index `src`
// (a known position of the segment), after n_iters iterations.
static void inpaint_segment(const std::vector<char>& known, int src, int n_iters,
                            std::vector<double>& resp) {
  const int n = int(known.size());
  std::vector<double> sparse(n, 0.0), cur, m(n), den, num, tmp(n);
  sparse[src] = 1.0;
  cur = sparse;
  for (int i = 0; i < n; ++i) m[i] = known[i] ? 1.0 : 0.0;
  for (int it = 0; it < std::max(1, n_iters); ++it) {
    stencil_reflect(m, den);
    for (int i = 0; i < n; ++i) tmp[i] = cur[i] * m[i];
    stencil_reflect(tmp, num);
    for (int i = 0; i < n; ++i) {
      const double prop = num[i] / (den[i] + 1e-12);
      if (den[i] > 1e-12 && m[i] < 1.0) m[i] = 1.0;
      cur[i] = known[i] ? sparse[i] : prop;
    }
  }
  resp = cur;
}

static void lowpass_rows(std::vector<SparseRow>& rows) {
  const int n = int(rows.size());
  if (n <= 2) return;
  std::vector<SparseRow> out(n);
  for (int i = 0; i < n; ++i) {
    const int l = (i == 0) ? 1 : i - 1;
    const int r = (i == n - 1) ? n - 2 : i + 1;
    SparseRow& o = out[i];
    for (const auto& kv : rows[l]) o[kv.first] += 0.25 * kv.second;
    for (const auto& kv : rows[i]) o[kv.first] += 0.5 * kv.second;
    for (const auto& kv : rows[r]) o[kv.first] += 0.25 * kv.second;
  }
  rows.swap(out);
}

static void cnn_operator_rows(const std::vector<char>& known, int n_iters, std::vector<SparseRow>& rows) {
  const int n = int(known.size());
  std::vector<int> ord(n, -1);
  int n_known = 0;
  for (int i = 0; i < n; ++i)
    if (known[i]) ord[i] = n_known++;
  rows.assign(n, SparseRow());
  for (int i = 0; i < n; ++i)
    if (known[i]) rows[i][ord[i]] = 1.0;
  if (n_known == n) {                     // ce_dl_cnn.py:488-489: low-pass only, nothing re-imposed
    lowpass_rows(rows);
    lowpass_rows(rows);
    return;
  }
  // unknown runs
  int i = 0;
  while (i < n) {
    if (known[i]) { ++i; continue; }
    int a = i;                            // first unknown of the run
    int b = i;
    while (b < n && !known[b]) ++b;       // one past the last unknown
    const int lo = (a > 0) ? a - 1 : a;   // include the bounding known samples
    const int hi = (b < n) ? b : b - 1;
    std::vector<char> seg(known.begin() + lo, known.begin() + hi + 1);
    std::vector<double> resp;
    if (a > 0) {
      inpaint_segment(seg, 0, n_iters, resp);
      for (int k = a; k < b; ++k) rows[k][ord[a - 1]] += resp[k - lo];
    }
    if (b < n) {
      inpaint_segment(seg, hi - lo, n_iters, resp);
      for (int k = a; k < b; ++k) rows[k][ord[b]] += resp[k - lo];
    }
    i = b;
  }
  std::vector<SparseRow> lp = rows;
  lowpass_rows(lp);
  lowpass_rows(lp);
  for (int k = 0; k < n; ++k)
    if (!known[k]) rows[k].swap(lp[k]);
}

// Rows are packed as (first ordinal, T dense coefficients): T = widest row span of the operator ...
static int row_span(const std::vector<SparseRow>& rows) {
  int t_max = 1;
  for (const auto& r : rows) {
    int lo = 1 << 30, hi = -1;
    for (const auto& kv : r)
      if (kv.second != 0.0) { lo = std::min(lo, kv.first); hi = std::max(hi, kv.first); }
    if (hi >= 0) t_max = std::max(t_max, hi - lo + 1);
  }
  return t_max;
}

// ... and the first ordinal is pulled back at the band end so that first + T <= n_known and the kernel never clamps.
static void emit_rows(const std::vector<SparseRow>& rows, int n_known, int terms,
                      int32_t* idx, float* coef) {
  for (size_t r = 0; r < rows.size(); ++r) {
    int lo = 1 << 30;
    for (const auto& kv : rows[r])
      if (kv.second != 0.0) lo = std::min(lo, kv.first);
    if (lo == (1 << 30)) lo = 0;
    lo = std::max(0, std::min(lo, n_known - terms));
    idx[r] = lo;
    for (int t = 0; t < terms; ++t) coef[r * terms + t] = 0.f;
    for (const auto& kv : rows[r])
      if (kv.second != 0.0) coef[r * terms + (kv.first - lo)] = float(kv.second);
  }
}

// ---------------------------------------------------------------------------------------------
// Tensor-core fill: pack the operator rows (lerp or CNN inpainting) into bf16 tiles for tcgen05.mma.
// ---------------------------------------------------------------------------------------------
static uint16_t bf16_rn(float f) {                      // round to nearest even, finite inputs only
  uint32_t u;
  memcpy(&u, &f, 4);
  u += 0x7fffu + ((u >> 16) & 1u);
  return uint16_t(u >> 16);
}
static float bf16_to_float(uint16_t h) {
  const uint32_t u = uint32_t(h) << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

// Fills p.fill_tc / p.fill_tc_a from the packed operator rows (p.fill_idx / p.fill_coef, fill_terms_host terms per
// row).  Unsupported geometries leave has_fill_tc = false with the reason in fill_tc_why; the direct fill remains.
static void build_fill_tc(Plan& p, int n_hops) {
  const srsce_plan_desc& d = p.desc;
  FillTcDev& f = p.fill_tc;
  memset(&f, 0, sizeof f);
  p.has_fill_tc = false;
  p.fill_tc_a.clear();
  auto no = [&](const char* why) { p.fill_tc_why = why; };
  if (n_hops != 1) return no("two hops");
  if (d.n_sym != SRSCE_MAX_SYMBOLS) return no("grid is not 14 symbols");
  const int L = d.n_layers, G = (L + 1) / 2, T = p.fill_terms_host, N = p.dev.n_re, rows = p.band_rows;
  const int n_tiles = (rows + kFillTileRows - 1) / kFillTileRows;
  if (n_tiles > kFillMaxTiles) return no("too many tiles");
  // pilot window of every tile
  std::vector<int> k0(n_tiles), k1(n_tiles);
  int span = 1;
  for (int m = 0; m < n_tiles; ++m) {
    int lo = 1 << 30, hi = -1;
    for (int g = 0; g < G; ++g)
      for (int r = m * kFillTileRows; r < std::min(rows, (m + 1) * kFillTileRows); ++r)
        for (int t = 0; t < T; ++t)
          if (p.fill_coef[(size_t(g) * rows + r) * T + t] != 0.f) {
            lo = std::min(lo, p.fill_idx[size_t(g) * rows + r] + t);
            hi = std::max(hi, p.fill_idx[size_t(g) * rows + r] + t);
          }
    if (hi < 0) lo = hi = 0;
    k0[m] = lo;
    k1[m] = hi;
    span = std::max(span, hi - lo + 1);
  }
  const int kt = ((span + 15) / 16) * 16;
  if (kt > 128) return no("operator window wider than 128 pilots per tile");
  // dense float tiles [m][g][96][kt], then their bf16 terms
  auto tile_of = [&](int m, int g, std::vector<float>& a) {
    a.assign(size_t(kFillTileRows) * kt, 0.f);
    for (int r = m * kFillTileRows; r < std::min(rows, (m + 1) * kFillTileRows); ++r)
      for (int t = 0; t < T; ++t) {
        const float c = p.fill_coef[(size_t(g) * rows + r) * T + t];
        if (c != 0.f) a[size_t(r - m * kFillTileRows) * kt + (p.fill_idx[size_t(g) * rows + r] + t - k0[m])] = c;
      }
  };
  std::vector<std::vector<float>> kinds;              // [kind] = concatenation over groups
  std::vector<float> a, cat;
  for (int m = 0; m < n_tiles; ++m) {
    cat.clear();
    for (int g = 0; g < G; ++g) {
      tile_of(m, g, a);
      cat.insert(cat.end(), a.begin(), a.end());
    }
    int kind = -1;
    for (size_t q = 0; q < kinds.size() && kind < 0; ++q)
      if (kinds[q] == cat) kind = int(q);
    if (kind < 0) {
      if (int(kinds.size()) == kFillMaxKinds) return no("operator tiles do not repeat (more than 6 distinct tiles)");
      kind = int(kinds.size());
      kinds.push_back(cat);
    }
    f.tile_kind[m] = uint8_t(kind);
    f.tile_k0[m] = k0[m];
  }
  // bf16 terms: a = a0 + a1 + a2 (each the bf16 rounding of what is left); keep as many as are non-zero anywhere
  int n_a = 1;
  for (const auto& kd : kinds)
    for (float v : kd) {
      const float r1 = v - bf16_to_float(bf16_rn(v));
      if (r1 != 0.f) {
        n_a = std::max(n_a, 2);
        if (r1 - bf16_to_float(bf16_rn(r1)) != 0.f) n_a = 3;
      }
    }
  f.n_a = n_a;
  f.n_prod = 0;
  for (int i = 0; i < n_a; ++i)
    for (int j = 0; i + j <= 2; ++j) {
      f.prod_a[f.n_prod] = i;
      f.prod_x[f.n_prod] = j;
      ++f.n_prod;
    }
  f.k_tile = kt;
  f.units = kt <= 32 ? 32 : (kt <= 64 ? 16 : 8);
  f.n_kinds = int(kinds.size());
  f.a_block_bytes = kFillTileRows * kt * 2;
  p.fill_tc_a.assign(size_t(f.n_kinds) * G * n_a * kFillTileRows * kt, 0);
  for (int q = 0; q < f.n_kinds; ++q)
    for (int g = 0; g < G; ++g)
      for (int r = 0; r < kFillTileRows; ++r)
        for (int k = 0; k < kt; ++k) {
          float v = kinds[q][(size_t(g) * kFillTileRows + r) * kt + k];
          for (int i = 0; i < n_a; ++i) {
            const uint16_t h = bf16_rn(v);
            v -= bf16_to_float(h);
            // canonical K-major core-matrix layout: [row / 8][k / 8][row % 8][k % 8]
            const size_t blk = (size_t(q) * G + g) * n_a + i;
            p.fill_tc_a[blk * kFillTileRows * kt + (size_t(r >> 3) * (kt >> 3) + (k >> 3)) * 64 + (r & 7) * 8 + (k & 7)] = h;
          }
        }
  f.n_sc = p.dev.n_sc;
  f.n_layers = L;
  f.n_cdm = G;
  f.n_re = N;
  f.band_sc0 = p.dev.hop[0].band_sc0;
  f.band_rows = rows;
  f.n_tiles = n_tiles;
  f.sym0 = p.dev.hop[0].start_symbol;
  f.sym1 = p.dev.hop[0].end_symbol;
  int col = 0, boff = 0;
  for (int g = 0; g < G; ++g) {
    const int nl = std::min(2, L - 2 * g);
    f.n_cols[g] = f.units * nl * 2;
    f.col0[g] = col;
    col += f.n_cols[g];
    f.b_block_bytes[g] = f.n_cols[g] * kt * 2;
    f.b_group_off[g] = boff;
    boff += 3 * f.b_block_bytes[g];
  }
  f.b_stage_bytes = boff;
  auto up128 = [](int b) { return (b + 127) / 128 * 128; };
  f.off_a = 0;
  f.off_b = up128(f.n_kinds * G * n_a * f.a_block_bytes);
  f.off_tile = f.off_b + up128(2 * f.b_stage_bytes);
  f.off_bar = f.off_tile + 2 * kFillEpiWarps * kTileRows * SRSCE_MAX_LAYERS * 8;   // double-buffered 32-row tile per epilogue warp
  f.smem_bytes = f.off_bar + 128;
  // (the MMA reads 128 operator rows where a block holds 96: the 32 phantom rows of the last block fall into the X
  //  stages behind it, 64 * k_tile bytes at most, always inside the allocation)
  if (f.smem_bytes > 227 * 1024) return no("operator tiles + staging exceed 227 KB of shared memory");
  if (col > 256) return no("accumulators exceed half of tensor memory");
  p.has_fill_tc = true;
  p.fill_tc_why.clear();
}

#define FAIL(code, ...)                                  \
  do {                                                   \
    char buf_[256];                                      \
    snprintf(buf_, sizeof buf_, __VA_ARGS__);            \
    err = buf_;                                          \
    return code;                                         \
  } while (0)

int build_plan(const srsce_plan_desc& d, Plan& p, std::string& err) {
  p.desc = d;
  PlanDev& v = p.dev;
  memset(&v, 0, sizeof v);

  if (d.n_layers < 1 || d.n_layers > SRSCE_MAX_LAYERS)
    FAIL(SRSCE_ERR_UNSUPPORTED, "n_layers=%d outside 1..%d", d.n_layers, SRSCE_MAX_LAYERS);
  if (d.n_prb_grid < 1 || d.n_prb_grid * kNRE > SRSCE_TA_FFT)
    FAIL(SRSCE_ERR_UNSUPPORTED, "grid of %d PRBs: n_sc must be in 12..4096 (time-alignment IFFT is 4096 points)", d.n_prb_grid);
  if (d.n_sym < 1 || d.n_sym > SRSCE_MAX_SYMBOLS)
    FAIL(SRSCE_ERR_UNSUPPORTED, "n_sym=%d outside 1..14", d.n_sym);
  if (d.smoothing != SRSCE_SMOOTH_NONE && d.smoothing != SRSCE_SMOOTH_MEAN && d.smoothing != SRSCE_SMOOTH_FILTER)
    FAIL(SRSCE_ERR_INVALID, "Unknown smoothing strategy %d.", d.smoothing);
  if (d.fill != SRSCE_FILL_LINEAR && d.fill != SRSCE_FILL_CNN) FAIL(SRSCE_ERR_INVALID, "unknown fill mode %d", d.fill);
  if (!(d.scs_hz > 0.0)) FAIL(SRSCE_ERR_INVALID, "scs must be positive");
  if (d.cfo_compensate && d.n_cp < 14)
    FAIL(SRSCE_ERR_INVALID, "config.CyclicPrefixDurations must have length >= 14 to match MATLAB code.");

  const int L = d.n_layers, G = (L + 1) / 2;
  const uint32_t sym_all = (d.n_sym >= 32) ? 0xffffffffu : ((1u << d.n_sym) - 1u);
  const int n_hops = d.hop[1].dmrs_symbol_mask ? 2 : 1;
  if (d.hop[0].dmrs_symbol_mask == 0) FAIL(SRSCE_ERR_INVALID, "hop 1 has no pilot symbol");
  if (n_hops == 2) {
    if (d.hop[0].dmrs_symbol_mask & d.hop[1].dmrs_symbol_mask) FAIL(SRSCE_ERR_ASSERT, "Hops should not overlap.");
    if (d.hop[1].n_prbs != d.hop[0].n_prbs)
      FAIL(SRSCE_ERR_INVALID, "both hops must span the same number of PRBs (one pilots tensor is split by symbol)");
  }
  const int R = popcount12(d.re_mask[0]);
  if (R < 1) FAIL(SRSCE_ERR_INVALID, "DMRSREmask column 0 is empty");
  for (int g = 0; g < G; ++g) {
    if ((d.re_mask[g] & ~0xfffu) != 0) FAIL(SRSCE_ERR_INVALID, "DMRSREmask has more than 12 rows");
    if (popcount12(d.re_mask[g]) != R)
      FAIL(SRSCE_ERR_INVALID, "DMRSREmask columns must select the same number of REs per PRB");
  }
  int s_total = 0;
  for (int h = 0; h < n_hops; ++h) {
    const srsce_hop_desc& hd = d.hop[h];
    if (hd.dmrs_symbol_mask & ~sym_all) FAIL(SRSCE_ERR_INVALID, "hop %d: pilot symbol outside the grid", h + 1);
    if (hd.n_prbs < 1 || hd.prb_start < 0 || hd.prb_start + hd.n_prbs > d.n_prb_grid)
      FAIL(SRSCE_ERR_INVALID, "hop %d: PRB range [%d, %d) outside the %d-PRB grid", h + 1, hd.prb_start,
           hd.prb_start + hd.n_prbs, d.n_prb_grid);
    if (hd.start_symbol < 0 || hd.n_alloc_symbols < 0 || hd.start_symbol + hd.n_alloc_symbols > d.n_sym)
      FAIL(SRSCE_ERR_INVALID, "hop %d: allocated symbols [%d, %d) outside the grid", h + 1, hd.start_symbol,
           hd.start_symbol + hd.n_alloc_symbols);
    s_total += __builtin_popcount(hd.dmrs_symbol_mask);
  }

  const int n_prbs = d.hop[0].n_prbs;
  const int N = n_prbs * R;
  v.n_sc = d.n_prb_grid * kNRE;
  v.n_sym = d.n_sym;
  v.n_layers = L;
  v.n_cdm = G;
  v.re_per_prb = R;
  v.n_re = N;
  v.n_dmrs_total = s_total;
  v.n_hops = n_hops;
  v.smoothing = d.smoothing;
  v.cfo_compensate = d.cfo_compensate ? 1 : 0;
  v.fill = d.fill;
  v.scs_hz = d.scs_hz;

  // symbol start times in symbol units: cumsum([CP0, CP1 + 1, ... CP13 + 1])  (:809-820)
  double cpd[SRSCE_MAX_SYMBOLS] = {0};
  for (int t = 0; t < SRSCE_MAX_SYMBOLS && t < d.n_cp; ++t) cpd[t] = d.cp_ms[t] * d.scs_hz / 1000.0;
  double t_sym[SRSCE_MAX_SYMBOLS] = {0};
  if (d.n_cp >= 14) {
    double acc = 0.0;
    for (int t = 0; t < 14; ++t) {
      acc += (t == 0) ? cpd[0] : cpd[t] + 1.0;
      t_sym[t] = acc;
    }
  }
  for (int t = 0; t < SRSCE_MAX_SYMBOLS; ++t) v.t_sym[t] = t_sym[t];

  int col0 = 0;
  for (int h = 0; h < n_hops; ++h) {
    const srsce_hop_desc& hd = d.hop[h];
    HopDev& hv = v.hop[h];
    hv.band_sc0 = kNRE * hd.prb_start;
    hv.band_sc1 = kNRE * (hd.prb_start + hd.n_prbs);
    hv.start_symbol = hd.start_symbol;
    hv.end_symbol = hd.start_symbol + hd.n_alloc_symbols;
    hv.pilot_col0 = col0;
    int s = 0;
    for (int t = 0; t < d.n_sym; ++t)
      if ((hd.dmrs_symbol_mask >> t) & 1u) {
        hv.sym_idx[s] = t;
        hv.t_dmrs[s] = t_sym[t < 14 ? t : 13];
        ++s;
      }
    hv.n_dmrs = s;
    col0 += s;
    if (s >= 2) {
      v.cfo_defined = 1;
      // n_samples between the first two pilot symbols, CPs included (:410-421); a short CP
      // vector is summed as far as it goes, like the reference's slice
      double n_samp = double(hv.sym_idx[1] - hv.sym_idx[0]);
      for (int t = hv.sym_idx[0] + 1; t <= hv.sym_idx[1] && t < d.n_cp && t < SRSCE_MAX_SYMBOLS; ++t) n_samp += cpd[t];
      hv.cfo_den = 2.0 * M_PI * n_samp * double(G);
    } else {
      hv.cfo_den = 1.0;
    }
  }
  if (v.cfo_compensate && v.cfo_defined && d.n_sym != 14)
    FAIL(SRSCE_ERR_INVALID, "CFO compensation with >= 2 pilot symbols needs a 14-symbol grid (got %d)", d.n_sym);

  // RE positions, interpolation table
  for (int g = 0; g < G; ++g) {
    int r = 0;
    for (int k = 0; k < kNRE; ++k)
      if ((d.re_mask[g] >> k) & 1u) v.re_pos[g][r++] = k;
    for (int pos = 0; pos < kNRE; ++pos) {
      int jr = 0;
      while (jr < R && v.re_pos[g][jr] < pos) ++jr;          // searchsorted(..., side="left")
      const int jl = jr - 1;
      const int pr = (jr < R) ? v.re_pos[g][jr] : v.re_pos[g][0] + kNRE;
      const int pl = (jl >= 0) ? v.re_pos[g][jl] : v.re_pos[g][R - 1] - kNRE;
      LerpEntry& e = v.lerp[g][pos];
      e.jl = jl;
      e.jr = jr;
      e.alpha = (float(pos) - float(pl)) / (float(pr) - float(pl));
    }
  }

  // smoothing filter
  v.n_taps = 0;
  v.n_virtual = 0;
  p.taps.clear();
  if (d.smoothing == SRSCE_SMOOTH_FILTER) {
    const int stride = kNRE / R;
    if (stride < 1) FAIL(SRSCE_ERR_INVALID, "stride must be >= 1");
    rc_taps(stride, std::min(3, n_prbs), p.taps);
    if (int(p.taps.size()) > SRSCE_MAX_TAPS) FAIL(SRSCE_ERR_UNSUPPORTED, "filter longer than %d taps", SRSCE_MAX_TAPS);
    v.n_taps = int(p.taps.size());
    v.n_virtual = (n_prbs > 1) ? std::min(12, v.n_taps / 2) : R;
    if (v.n_virtual > N) FAIL(SRSCE_ERR_UNSUPPORTED, "more virtual pilots than pilots");
    for (int k = 0; k < v.n_taps; ++k) v.taps[k] = p.taps[k];
    const int n = v.n_virtual;
    if (n >= 2) {
      double sx = 0.0, sxx = 0.0;
      for (int k = 0; k < n; ++k) { sx += k; sxx += double(k) * k; }
      const double nrm = sqrt(sxx);
      v.vp_mx = sx / n;
      v.vp_den = nrm * nrm - n * v.vp_mx * v.vp_mx;
    }
  }
  v.blend = (d.fill == SRSCE_FILL_CNN && d.smoothing == SRSCE_SMOOTH_FILTER && d.cnn_alpha > 0.0) ? 1 : 0;
  v.cnn_alpha = float(std::max(0.0, std::min(1.0, d.cnn_alpha)));

  // time alignment: the positions of the LAST CDM group are used for every layer (:674-675).
  // If that comb is regular with a stride s that divides 4096, |IFFT_4096(zero-stuffed)| equals the
  // tiled |IFFT_{4096/s}(pilots)| (position offset only adds a linear phase), so a short transform suffices.
  {
    const int* pos = v.re_pos[G - 1];
    const int s = kNRE / R;
    bool regular = (R * s == kNRE) && (SRSCE_TA_FFT % s == 0);
    for (int r = 1; r < R && regular; ++r) regular = (pos[r] - pos[r - 1] == s);
    if (regular && N <= SRSCE_TA_FFT / s) {
      v.fft_regular = 1;
      v.fft_size = SRSCE_TA_FFT / s;
    } else {
      v.fft_regular = 0;
      v.fft_size = SRSCE_TA_FFT;
    }
    v.fft_layers = std::max(1, std::min(L, 4096 / v.fft_size));
  }

  // normalisation (:899-919)
  {
    const double n_pil = double(n_prbs) * R * s_total;
    v.norm_rsrp = 1.0 / (n_pil * L);
    v.norm_epre = 1.0 / n_pil;
    v.norm_noise = 1.0 / (double(G) * n_pil - 1.0);
    v.ta_scale = 1.0 / double(SRSCE_TA_FFT) / d.scs_hz;
    v.ta_hop_div = (n_hops == 2) ? 2.0 : 1.0;
  }

  // output row geometry
  v.row_elems = d.n_sym * L;
  v.vec = (v.row_elems % 2 == 0) ? 2 : 1;
  v.cols = v.row_elems / v.vec;

  // fill operator.  Host copy always (tests); the device copy is only used by the CNN fill.
  p.band_rows = kNRE * n_prbs;
  {
    std::vector<std::vector<SparseRow>> all(G);
    int terms = 1;
    for (int g = 0; g < G; ++g) {
      std::vector<SparseRow>& rows = all[g];
      if (d.fill == SRSCE_FILL_CNN) {
        std::vector<char> known(p.band_rows);
        for (int r = 0; r < p.band_rows; ++r) known[r] = (d.re_mask[g] >> (r % kNRE)) & 1u;
        cnn_operator_rows(known, std::max(6, p.band_rows / 8), rows);
      } else {
        rows.assign(p.band_rows, SparseRow());
        for (int r = 0; r < p.band_rows; ++r) {
          const LerpEntry& e = v.lerp[g][r % kNRE];
          const int q = r / kNRE;
          const int il = std::max(0, std::min(N - 1, q * R + e.jl));
          const int ir = std::max(0, std::min(N - 1, q * R + e.jr));
          rows[r][il] += 1.0 - double(e.alpha);
          rows[r][ir] += double(e.alpha);
        }
      }
      terms = std::max(terms, std::min(row_span(rows), N));
    }
    if (d.fill == SRSCE_FILL_CNN && terms > kMaxFillTerms)
      FAIL(SRSCE_ERR_UNSUPPORTED, "CNN operator rows span %d pilots (max %d)", terms, kMaxFillTerms);
    p.fill_terms_host = terms;
    p.fill_idx.assign(size_t(G) * p.band_rows, 0);
    p.fill_coef.assign(size_t(G) * p.band_rows * terms, 0.f);
    for (int g = 0; g < G; ++g)
      emit_rows(all[g], N, terms, &p.fill_idx[size_t(g) * p.band_rows], &p.fill_coef[size_t(g) * p.band_rows * terms]);
    v.fill_terms = (d.fill == SRSCE_FILL_CNN) ? terms : 2;
    // Away from the band edges every PRB sees the same neighbourhood, so the operator rows repeat with
    // the PRB period (pilot index advancing by R per PRB).  Verify that bit for bit; if it holds the
    // kernel keeps one PRB's rows in shared memory and only looks the edge rows up in global memory.
    v.fill_periodic = 0;
    if (d.fill == SRSCE_FILL_CNN && p.band_rows >= 2 * kCnnEdgeRows + kNRE) {
      bool ok = true;
      for (int g = 0; g < G && ok; ++g) {
        const int32_t* idx = &p.fill_idx[size_t(g) * p.band_rows];
        const float* coef = &p.fill_coef[size_t(g) * p.band_rows * terms];
        const int q0 = kCnnEdgeRows / kNRE;
        for (int pos = 0; pos < kNRE; ++pos) {
          v.cnn_idx0[g][pos] = idx[kCnnEdgeRows + pos] - q0 * R;
          for (int t = 0; t < kMaxFillTerms; ++t)
            v.cnn_coef[g][pos][t] = (t < terms) ? coef[size_t(kCnnEdgeRows + pos) * terms + t] : 0.f;
        }
        for (int r = kCnnEdgeRows; r < p.band_rows - kCnnEdgeRows && ok; ++r) {
          const int q = r / kNRE, pos = r - q * kNRE;
          ok = idx[r] == v.cnn_idx0[g][pos] + q * R;
          for (int t = 0; t < terms && ok; ++t) ok = coef[size_t(r) * terms + t] == v.cnn_coef[g][pos][t];
        }
      }
      v.fill_periodic = ok ? 1 : 0;
    }
  }

  // shared memory: smoothed pilots of every hop (x hall_bufs) | staged pilot REs (x y_bufs) | raw pilots +
  // virtual pilots, later FFT scratch | the writer warps' expansion tiles | small tables.
  // The planner walks a short list of layouts and takes the first that fits its budget:
  //   A  two CTAs of 256 estimator threads (+4 writer warps) per SM, pilot REs fetched a whole unit ahead.  The
  //      two-CTA footprint is held under the 196 KB carve-out (1 KB per CTA is reserved by the system): one
  //      step higher leaves 28 KB instead of 60 KB of L1 for phase A's twiddle / pilot loads, measured to cost
  //      11 % of the headline throughput.
  //   B  the same CTA shape with a slimmer footprint: one staging buffer (the next unit's pilot REs are fetched
  //      while the IFFT runs), FIR input one CDM group at a time, IFFT two layers at a time.
  //   C  one CTA of 512 estimator threads (+6 writer warps), one staging buffer, FIR input for all layers (the
  //      stretch between "writers done" and "next FIR done" is shortest when both CDM groups are de-spread ahead).
  //   D  the same with the slim footprint (the largest geometries).
  //   E  D without a staging buffer: the pilot REs are read from the received grid where they are used (dense
  //      masks with several pilot symbols).
  //   F  E with the smoothed pilots in a per-CTA global scratch (L2 resident) instead of shared memory: two-hop plans
  //      whose smoothed pilots of both hops (2 * N * L * 8 bytes) leave no room for the rest -- in practice the dense
  //      mask (12 pilots per PRB) on two hops at 273 PRB with >= 3 layers or 341 PRB with >= 2 layers (host sweep of
  //      1 120 geometries: those 20 and nothing else).  Slower (every pass over the pilots goes to L2) but complete.
  //   Measured on B200 (comb-2, 273 PRB): single staging buffer 0.60 (C) / 0.59 (D), two staging buffers 0.58 / 0.57,
  //   smoothed pilots double buffered 0.50-0.54 of the HBM peak -- phase A re-reads the pilots through L1, the
  //   gather of the next unit pollutes L1 unless it is issued after the noise estimate, and L1 is what shared
  //   memory leaves over.
  {
    const int h_elems = n_hops * N * L;
    const int y_unit = G * N * s_total;
    const int fir_len = N + 2 * v.n_virtual + 2 * (v.n_taps / 2);
    const int fft_default = std::max(1, std::min(L, 4096 / v.fft_size));
    auto up16 = [](int bytes) { return ((bytes + 15) / 16) * 16; };
    auto layout = [&](int na, int hall_bufs, int y_bufs, bool slim, bool h_global = false) {
      v.h_global = h_global ? 1 : 0;
      const int ext_layers = slim ? std::min(L, 2) : L;
      v.fir_per_group = (slim && L > 2) ? 1 : 0;
      v.fft_layers = slim ? std::min(fft_default, 2) : fft_default;
      v.ext_hop = fir_len * ext_layers;                              // FIR input incl. virtual pilots and zero margins, per hop
      int ext = n_hops * v.ext_hop;                                 // (also the scratch of the CNN blend: N * layers <= ext_hop)
      ext = std::max(ext, v.fft_layers * (v.fft_size + v.fft_size / 8));   // FFT scratch: one pad element per 8 (bank spread)
      const int tile_elems = writer_warps(na, n_hops == 2) * kTileRows * L * n_hops;
      v.ext_elems = ext;
      v.est_threads = na;
      v.hall_bufs = hall_bufs;
      v.y_bufs = y_bufs;
      v.off_h = 0;
      v.off_y = v.off_h + (h_global ? 0 : up16(hall_bufs * h_elems * 8));
      v.off_ext = v.off_y + up16(y_bufs * y_unit * 8);
      v.off_tile = v.off_ext + up16(ext * 8);
      v.off_misc = v.off_tile + up16(tile_elems * 8);
      v.smem_bytes = v.off_misc + kMiscBytes + n_hops * kHopScratchBytes;
      return v.smem_bytes;
    };
    const int two_cta_budget = 196 * 1024 / 2 - 1024, one_cta_l1_budget = 196 * 1024 - 1024;
    // Light phase A (one hop, one pilot symbol, a regular comb) with the vectorised writer (even layer count): one
    // CTA of 512 + 6 warps beats two of 256 + 4 (C before A; config 2: 0.880 vs 0.852, 2 layers 0.86 vs 0.81, narrow
    // hops +3 %).  With more pilot symbols, two hops, odd layer counts or a 4096-point IFFT the two-CTA shapes win
    // (0.75 vs 0.73, 0.69 vs 0.64, 0.80 vs 0.78, 0.52 vs 0.50): their phase A is long enough that a single CTA's
    // writers starve between hand-overs.
    const bool light = n_hops == 1 && s_total == 1 && L % 2 == 0 && v.fft_regular;
    // Time-alignment transform.  The register-blocked radix-8 version (3 shared-memory round trips per element, no
    // twiddle loads) takes a third of the radix-4 version's time and lifts every phase-A-bound geometry (same box,
    // fraction of the HBM peak: comb-2 0.65 -> 0.75, type-2 mask 0.53 -> 0.77, S = 4 0.59 -> 0.63, ce_dl_cnn fill
    // 0.83 -> 0.86, two-hop shapes 0.52 -> 0.72).  The writer-bound headline family (one hop, one pilot symbol, four
    // ports on a comb-4: config 2 and its narrow-band hops) measures 2.5-3.7 % SLOWER with it (config 2 at bench
    // size: 0.888 -> 0.856) although its phase A is shorter, so those plans keep the radix-4 passes; with two ports
    // or one the radix-8 passes are level or ahead.  A measured rule like the CTA-shape rule above, not a model.
    v.ta_radix4 = (light && G == 2 && v.fft_size == 1024 && d.fill == SRSCE_FILL_LINEAR) ? 1 : 0;
    if (!(light && layout(512, 1, 1, false) <= one_cta_l1_budget) &&
        layout(256, 1, 2, false) > two_cta_budget && layout(256, 1, 1, true) > two_cta_budget &&
        layout(512, 1, 1, false) > 227 * 1024 && layout(512, 1, 1, true) > 227 * 1024) {
      if (layout(512, 1, 0, true) > 227 * 1024 && n_hops == 2) layout(512, 1, 0, true, true);   // E, else F
    }
    // A plan that runs the single CTA also keeps the CTA-pair layout: without the scattered gather (received grids
    // shared by the UEs of a slot, symbol-major grids) or without the dense write (compact output) the pair is the
    // faster shape (same box: shared grids 0.948 vs 0.885, compact output 28.9 vs 25.2 M estimates/s).
    p.has_pair = false;
    if (light && v.est_threads == 512) {
      const PlanDev primary = v;
      if (layout(256, 1, 2, false) <= two_cta_budget || layout(256, 1, 1, true) <= two_cta_budget) {
        p.dev_pair = v;
        p.has_pair = true;
      }
      v = primary;
    }
#ifdef SRSCE_PROFILING
    if (const char* e = getenv("SRSCE_DEBUG_LAYOUT")) {
      p.has_pair = false;       // experiments only: "na,hall_bufs,y_bufs,slim"
      int na = 256, hb = 1, yb = 2, sl = 0;
      if (sscanf(e, "%d,%d,%d,%d", &na, &hb, &yb, &sl) == 4 && (na == 256 || na == 512))
        layout(yb == 0 ? 512 : na, hb, yb, sl != 0);              // (the unstaged kernels exist for NA = 512 only)
    }
#endif
    if (v.smem_bytes > 227 * 1024)
      FAIL(SRSCE_ERR_UNSUPPORTED, "geometry needs %d bytes of shared memory per unit (limit %d)", v.smem_bytes, 227 * 1024);
  }

  build_fill_tc(p, n_hops);
  p.bytes_per_unit = 8LL * (int64_t(G) * N * s_total + int64_t(N) * s_total * L + int64_t(v.n_sc) * d.n_sym * L) + 48;
  return SRSCE_OK;
}

}  // namespace srsce
