// Internal definitions shared by the host planner (srsce_plan.cc), the C ABI glue (srsce_api.cc)
// and the CUDA kernels (srsce_kernels.cu).  Not part of the public ABI (include/srsce.h is).
#pragma once

#include <stdint.h>

#include <string>
#include <vector>

#include <vector_types.h>

#include "srsce.h"

namespace srsce {

constexpr int kNRE = 12;              // resource elements per PRB
// A CTA = NA estimator threads (phase A: pilots -> smoothed channel + scalars) + writer warps (phase B:
// interpolate + stream the dense grid out).  Geometries whose shared-memory footprint allows two CTAs
// per SM run NA = 256 (+4 writer warps, 6 with two hops); heavier ones run one CTA of NA = 512 (+6 / 8 writer warps).
constexpr int kMaxWarpsA = 16;
constexpr int kMaxWarpsB = 8;
// Writer warps per CTA, measured on B200 at BASELINE config 2 (fraction of the HBM peak, same box):
//   two CTAs of 256 estimator threads: 4 warps 0.852 | 5 warps 0.832 | 6 warps 0.834
//   one CTA of 512 estimator threads:  3 warps 0.80 | 4-5 warps 0.85 | 6 warps 0.880 | 8 warps 0.870 | 10 warps 0.84 | 12 warps 0.82
// More writers than the store stream needs only add contention for phase A's loads.  Two-hop units do more work per
// stored row (two sources, per-symbol select) and keep the larger teams: 6 / 8 warps measure 4-12 % faster there.
constexpr int writer_warps(int na, bool two_hops) { return na == 256 ? (two_hops ? 6 : 4) : (two_hops ? 8 : 6); }
constexpr int cta_threads(int na, bool two_hops) { return na + 32 * writer_warps(na, two_hops); }
constexpr int kTileRows = 32;         // subcarrier rows a warp expands per tile (one per lane)
constexpr int kMaxFillTerms = 8;      // widest CNN operator row the kernel accepts
constexpr int kCnnEdgeRows = 2 * kNRE; // band rows at either end whose CNN operator rows are looked up individually
constexpr int kMiscBytes = 5616;      // shared-memory slot of the kernel's small tables / reduction scratch ...
constexpr int kHopScratchBytes = 2176; // ... followed by one slot per hop (rotation tables, delay profile, edge-pilot scratch)

// One entry of the per-PRB interpolation table (fill_ch_est_cdm, ce_rule_tensorized.py:311-340):
// a subcarrier at position p of PRB q takes pilots q*R+jl and q*R+jr of its CDM group
// (jl = -1 / jr = R reach into the neighbouring PRB; the kernel clamps to [0, N-1], which
// reproduces the flat extrapolation at both band edges) and alpha = (p - pos_l) / (pos_r - pos_l).
struct LerpEntry {
  int32_t jl;
  int32_t jr;
  float alpha;
  int32_t pad;
};

struct HopDev {
  int32_t band_sc0;        // 12 * PRBstart
  int32_t band_sc1;        // 12 * (PRBstart + nPRBs)
  int32_t start_symbol;
  int32_t end_symbol;      // startSymbol + nAllocatedSymbols
  int32_t n_dmrs;          // pilot symbols S of this hop
  int32_t pilot_col0;      // first column of this hop in pilots[:, :, :]
  int32_t pad0, pad1;
  int32_t sym_idx[SRSCE_MAX_SYMBOLS];   // grid symbol of each pilot symbol
  double t_dmrs[SRSCE_MAX_SYMBOLS];     // symbol_start_time (symbol units) at those symbols (:809-820)
  double cfo_den;                       // 2*pi*n_samples*ceil(L/2)  (:410-426)
};

// Plain-old-data view of a plan passed to the kernel as a __grid_constant__ parameter.
struct PlanDev {
  int32_t n_sc, n_sym, n_layers, n_cdm;
  int32_t re_per_prb;      // R
  int32_t n_re;            // N = nPRBs * R pilots per CDM group
  int32_t n_dmrs_total;    // S_total
  int32_t n_hops;
  int32_t smoothing, cfo_compensate, fill, cfo_defined;
  int32_t n_taps, n_virtual;
  int32_t fft_size;        // M
  int32_t fft_regular;     // 1: M-point transform of the pilot vector (comb stride divides 4096); 0: zero-stuffed 4096
  int32_t fft_layers;      // layers transformed together in the scratch buffer
  int32_t fill_terms;      // T of the CNN operator
  int32_t fill_periodic;   // interior rows of the CNN operator repeat with the PRB period: use cnn_coef / cnn_idx0
  int32_t row_elems;       // n_sym * n_layers complex64 per subcarrier row of the output
  int32_t vec;             // 2: float4 stores (row_elems even), 1: float2 stores
  int32_t cols;            // row_elems / vec store slots per row
  int32_t blend;           // CNN blend of the smoothed pilots active (ce_dl_cnn.py:712-717)
  int32_t est_threads;     // NA: 256 (two CTAs per SM) or 512 (one CTA per SM)
  int32_t hall_bufs;       // 1: writers and the next unit's FIR alternate on one pilot buffer; 2: double buffered
  int32_t y_bufs;          // staging buffers of the received pilot REs: 2 = fetched a whole unit ahead, 1 = during the IFFT
  int32_t fir_per_group;   // FIR input (ext) holds one CDM group (<= 2 layers) at a time instead of all layers
  int32_t h_global;        // 1: the smoothed pilots live in a per-CTA global scratch, not in shared memory (layout F)
  int32_t ta_radix4;       // time-alignment transform: 1 = radix-4 passes + twiddle table (light phase A), 0 = register-blocked radix 8
  // shared-memory carve-up (byte offsets from the dynamic smem base)
  int32_t off_h, off_y, off_ext, off_tile, off_misc, smem_bytes, ext_elems;
  int32_t ext_hop;         // complex64 elements of one hop's FIR-input region inside ext
  float cnn_alpha;
  float pad_f;
  double scs_hz;
  double vp_mx, vp_den;                   // create_virtual_pilots regression constants (:105-118)
  double t_sym[SRSCE_MAX_SYMBOLS];        // symbol_start_time for the final CFO rotation (:922-929)
  double taps[SRSCE_MAX_TAPS];
  double norm_rsrp, norm_epre, norm_noise, ta_scale, ta_hop_div;  // :899-919
  int32_t re_pos[2][kNRE];                // positions of the set bits of each CDM group's RE mask
  LerpEntry lerp[2][kNRE];
  float cnn_coef[2][kNRE][kMaxFillTerms];   // CNN operator of an interior PRB: row pos takes pilots cnn_idx0 + q*R + t
  int32_t cnn_idx0[2][kNRE];
  HopDev hop[2];
  // device tables (owned by the plan)
  const float2* twiddle;                  // exp(+j*2*pi*k/M), k < M = fft_size
  const int32_t* fill_idx;                // [n_cdm][band rows]           (CNN fill only)
  const float* fill_coef;                 // [n_cdm][band rows][fill_terms]
};

// ---- tensor-core fill (srsce_fill_tc.cu): the band interpolation operator of fill_ch_est_cdm as a banded GEMM
//      D[96 band rows x (unit, layer, re/im)] = A_tile[96 x K] * X[K x (unit, layer, re/im)]   per CDM group,
// bf16 operands (float32 values split into bf16 terms, so the float32 result is reproduced), float32 accumulators
// in tensor memory.  A tile = 8 PRBs: interior tiles share one operator block because the rows repeat per PRB.
constexpr int kFillTileRows = 96;
constexpr int kFillMaxTiles = 48;      // ceil(4096 / 96) + slack
constexpr int kFillMaxKinds = 6;       // distinct operator tiles kept in shared memory (first / interior / last ...)
constexpr int kFillEpiWarps = 8;       // two warpgroups; the warps of lane quadrant 3 (rows 96..127 of the MMA) idle
constexpr int kFillLoadWarps = 4;
constexpr int kFillThreads = 32 * (kFillEpiWarps + kFillLoadWarps + 1);
struct FillTcDev {
  int32_t n_sc, n_layers, n_cdm, n_re;
  int32_t band_sc0, band_rows, n_tiles;
  int32_t sym0, sym1;                  // allocated symbols [sym0, sym1)
  int32_t k_tile;                      // pilots per tile window (multiple of 16)
  int32_t units;                       // units per work item (UMMA N = units * layers in group * 2)
  int32_t n_a;                         // bf16 terms of the operator coefficients (1 when they are exact in bf16)
  int32_t n_kinds;
  int32_t n_prod;                      // bf16 products a_i * x_j kept (i + j <= 2)
  int32_t prod_a[6], prod_x[6];
  int32_t n_cols[2];                   // UMMA N of each CDM group
  int32_t col0[2];                     // first tensor-memory column of each group inside an accumulator stage
  int32_t a_block_bytes;               // one (kind, group, term) operator block: 96 x k_tile bf16, canonical K-major
  int32_t b_block_bytes[2];            // one (group, term) block of X: n_cols[g] x k_tile bf16
  int32_t b_group_off[2];              // byte offset of group g inside a B stage
  int32_t b_stage_bytes;
  int32_t off_a, off_b, off_tile, off_bar, smem_bytes;
  uint8_t tile_kind[kFillMaxTiles];
  int32_t tile_k0[kFillMaxTiles];      // first pilot of each tile's window
  const uint16_t* a_tiles;             // device copy: [kind][group][term] blocks
};

struct Plan {
  srsce_plan_desc desc;
  PlanDev dev;                        // layout for the reference call shape (dense output, one received grid per unit)
  PlanDev dev_pair;                   // CTA-pair layout of the same plan (valid when has_pair): launches whose pilot
  bool has_pair = false;              //   gather is cheap -- shared or symbol-major grids -- and the compact output
  std::vector<double> taps;
  std::vector<int32_t> fill_idx;      // host copy of the operator (lerp rows too, for the tests)
  std::vector<float> fill_coef;
  int fill_terms_host = 0;
  int band_rows = 0;
  int device = -1;                    // CUDA device the tables live on, -1: host-only plan
  int ctas_per_sm = 0;                // occupancy of the selected kernel variant
  int ctas_per_sm_pair = 0;
  int n_sms = 0;
  void* d_twiddle = nullptr;
  void* d_fill_idx = nullptr;
  void* d_fill_coef = nullptr;
  int64_t bytes_per_unit = 0;
  // tensor-core fill
  bool has_fill_tc = false;
  std::string fill_tc_why;            // why the tensor-core fill is unavailable for this plan
  FillTcDev fill_tc;
  std::vector<uint16_t> fill_tc_a;    // host copy of the operator blocks (bf16 bit patterns)
  void* d_fill_tc_a = nullptr;
};

// srsce_plan.cc
int build_plan(const srsce_plan_desc& d, Plan& p, std::string& err);
void rc_taps(int stride, int n_rbs, std::vector<double>& taps);

// srsce_kernels.cu
int upload_plan(Plan& p, std::string& err);
void release_plan(Plan& p);
int launch_estimate(const Plan& p, const void* rx, const int64_t strides[3], const void* pilots,
                    int pilot_share, double beta, void* out, double* scalars, int64_t n_units,
                    void* stream, int* launches, std::string& err, void* out_compact = nullptr,
                    void* rot_compact = nullptr, const int64_t* rx_index = nullptr);
int max_dynamic_smem();
int launch_fill_direct(const Plan& p, const void* h_pilots, void* out, int64_t n_units, void* stream, int* launches,
                       std::string& err, void* out_compact = nullptr);

// srsce_seq.cu
// srsce_hostio.cc: replication of the compact per-hop columns over the allocated symbols, in host memory, on host threads
int replicate_host(const Plan& plan, const void* ch_sc, void* ch_est, int64_t n_units, int n_threads, int flags,
                   std::string& err);
int launch_gen_pilots(void* pilots, int64_t n_ue, int32_t m_zc, int32_t n_dmrs, int32_t n_layers, const int32_t* u_of,
                      const int32_t* v_of, void* stream, std::string& err);
int largest_prime_below(int m);

// srsce_fill_tc.cu
int prepare_fill_tc(Plan& p, std::string& err);
int launch_fill_tc(const Plan& p, const void* h_pilots, void* out, int64_t n_units, void* stream, int* launches,
                   std::string& err);

}  // namespace srsce
