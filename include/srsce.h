/* srsce.h -- C ABI of the B200-native srsRAN port channel estimator.
 *
 * This is the drop-in boundary for the ONE hot path of pjookim/srsran-ce-pytorch:
 *
 *   srs_channel_estimator(received_rg, pilots, beta_dmrs, hop1, hop2, config)
 *       reference: src/ce_rule_tensorized.py:745-937  (same in ce_rule_baseline.py:761, ce_dl_cnn.py:802)
 *
 * The reference has no FFI of its own (it is pure Python on torch eager ops); these entry points
 * are what a ctypes binding of that function binds (see INTEGRATION.md for the stub).  Plain
 * pointers and sizes only, no torch types, no C++ exceptions across the boundary.
 *
 * Ownership / threading contract
 *   - The caller owns every device buffer passed to srsce_estimate / srsce_estimate_compact and keeps it
 *     alive until the stream reaches the kernel.  The library allocates nothing per call (one exception: plans with
 *     srsce_plan_info.h_in_global take a stream-ordered scratch from the device's memory pool per launch), never
 *     synchronises and never touches the default stream unless it is the stream passed in.
 *   - A plan owns its (small) host tables and their device copies; it is immutable after creation
 *     and may be shared by threads and streams.  A plan is bound to the CUDA device that was
 *     current in srsce_plan_create.  Without a CUDA device, plan creation still succeeds
 *     (host tables only, used by the CPU-side tests); compute calls on it fail with
 *     SRSCE_ERR_NO_DEVICE.  There is no CPU compute path.
 *   - Every function returns 0 on success or a negative SRSCE_ERR_* code; the message is in
 *     srsce_last_error() (thread local).
 */
#ifndef SRSCE_H_
#define SRSCE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SRSCE_ABI_VERSION 3

#define SRSCE_MAX_LAYERS 4
#define SRSCE_MAX_SYMBOLS 14   /* the reference's CFO tables assume a 14-symbol slot (:809-820) */
#define SRSCE_MAX_TAPS 31      /* get_rc_filter(1, 3): longest raised-cosine filter (:184-234) */
#define SRSCE_TA_FFT 4096      /* ce_rule_tensorized.py:677 */
#define SRSCE_TA_WINDOW 144    /* floor((144/2)*4096/2048), :682 */

enum {
  SRSCE_OK = 0,
  SRSCE_ERR_INVALID = -1,     /* ValueError in the reference (unknown smoothing :668, CP vector < 14 :815, ...) */
  SRSCE_ERR_ASSERT = -2,      /* AssertionError in the reference (hop overlap :862, differing RE masks :869) */
  SRSCE_ERR_UNSUPPORTED = -3, /* outside the documented limits (n_sc <= 4096, <= 4 layers, <= 14 symbols) */
  SRSCE_ERR_NO_DEVICE = -4,
  SRSCE_ERR_CUDA = -5
};

enum { SRSCE_SMOOTH_NONE = 0, SRSCE_SMOOTH_MEAN = 1, SRSCE_SMOOTH_FILTER = 2 };  /* :633-668 */
enum { SRSCE_FILL_LINEAR = 0, SRSCE_FILL_CNN = 1 };  /* fill_ch_est_cdm of ce_rule_* (:237-354) / ce_dl_cnn.py:233-322 */

/* One frequency hop: mirrors HopConfig (src/ce_rule_tensorized.py:13-21). */
typedef struct srsce_hop_desc {
  uint32_t dmrs_symbol_mask;  /* bit t set <=> DMRSsymbols[t]; 0 => the hop is absent */
  int32_t prb_start;          /* PRBstart (0-based) */
  int32_t n_prbs;             /* nPRBs; maskPRBs must be exactly [prb_start, prb_start+n_prbs) */
  int32_t start_symbol;       /* startSymbol */
  int32_t n_alloc_symbols;    /* nAllocatedSymbols */
} srsce_hop_desc;

/* Everything data independent about a call: HopConfig x2 + EstimatorConfig (:24-29) + tensor shapes. */
typedef struct srsce_plan_desc {
  int32_t n_prb_grid;         /* len(maskPRBs); n_sc = 12 * n_prb_grid <= 4096 */
  int32_t n_sym;              /* received_rg.shape[1] (<= 14) */
  int32_t n_layers;           /* pilots.shape[2], 1..4 */
  int32_t n_cp;               /* entries of cp_ms that are valid (reference needs >= 14 with CFO compensation) */
  uint16_t re_mask[2];        /* DMRSREmask[:, g] as 12 bits (bit k <=> row k), g < ceil(n_layers/2) */
  int32_t smoothing;          /* SRSCE_SMOOTH_* */
  int32_t cfo_compensate;     /* EstimatorConfig.CFOCompensate */
  int32_t fill;               /* SRSCE_FILL_* */
  double cnn_alpha;           /* CNNSmoothingAlpha (ce_dl_cnn.py:864-867); used only with SRSCE_FILL_CNN */
  double scs_hz;              /* EstimatorConfig.scs */
  double cp_ms[SRSCE_MAX_SYMBOLS]; /* EstimatorConfig.CyclicPrefixDurations (ms) */
  srsce_hop_desc hop[2];      /* hop[1].dmrs_symbol_mask == 0 <=> single hop */
} srsce_plan_desc;

typedef struct srsce_plan srsce_plan;

/* Shapes derived from a plan (what the caller must allocate). */
typedef struct srsce_plan_info {
  int32_t n_sc;               /* 12 * n_prb_grid */
  int32_t n_sym;
  int32_t n_layers;
  int32_t n_cdm;              /* ceil(n_layers / 2) */
  int32_t n_re;               /* pilots.shape[0] = n_prbs * REs per PRB */
  int32_t n_dmrs_total;       /* pilots.shape[1] = pilot symbols of hop 1 + hop 2 */
  int32_t n_taps;             /* raised-cosine taps (0 unless smoothing == filter) */
  int32_t n_virtual;          /* virtual pilots per edge */
  int32_t fft_size;           /* TA transform actually run (4096 / comb stride, or 4096 zero-stuffed) */
  int32_t cfo_defined;        /* 1 if some hop has >= 2 pilot symbols (cfo output is a scalar, not empty) */
  int32_t smem_bytes;         /* dynamic shared memory of the fused kernel */
  int32_t threads;            /* CTA size */
  int32_t fill_terms;         /* operator terms per subcarrier (2 = lerp; CNN operator width otherwise) */
  int32_t has_device;         /* device tables were uploaded */
  int64_t bytes_per_unit;     /* algorithmic HBM bytes of one unit with pilots read once per unit */
  int32_t fill_tensor;        /* 1: srsce_fill(..., SRSCE_FILL_IMPL_TENSOR) is available for this plan */
  int32_t fill_tensor_k;      /* pilots per 96-row tile window (K of the banded GEMM before the bf16 term split) */
  int32_t fill_tensor_terms;  /* bf16 products a_i * x_j accumulated per tile (3 when the operator is exact in bf16) */
  int32_t fill_tensor_units;  /* units per MMA (N = units * layers in the CDM group * 2) */
  int32_t fill_tensor_tiles;  /* 96-row tiles of the hop band */
  int32_t fill_tensor_kinds;  /* distinct operator tiles (first / interior / last ...) kept in shared memory */
  int32_t fill_tensor_a_terms;/* bf16 terms per operator coefficient (1 when the coefficients are exact in bf16) */
  int32_t h_in_global;        /* 1: the smoothed pilots of a unit do not fit in shared memory beside the rest and live in a
                               * per-CTA global scratch (largest two-hop geometries; compute calls then make one
                               * stream-ordered allocation per launch, cudaMallocAsync / cudaFreeAsync on the caller's stream) */
} srsce_plan_info;

int srsce_abi_version(void);
const char* srsce_last_error(void);

int srsce_plan_create(const srsce_plan_desc* desc, srsce_plan** out);
void srsce_plan_destroy(srsce_plan* plan);
int srsce_plan_query(const srsce_plan* plan, srsce_plan_info* info);

/* Host-side tables (for tests of the host logic; no device needed).
 *   taps:      n_taps doubles                       get_rc_filter (:184-234)
 *   fill_idx:  n_hop_sc * n_cdm int32               first pilot index of each band subcarrier's operator row
 *   fill_coef: n_hop_sc * n_cdm * fill_terms float  operator coefficients (linear fill: the row {1-a, a} of the lerp)
 * Pass NULL to skip a table. */
int srsce_plan_tables(const srsce_plan* plan, double* taps, int32_t* fill_idx, float* fill_coef);

/* Host-side tables of the tensor-core fill (tests of the host logic; no device needed):
 *   a_tiles    fill_tensor_kinds * n_cdm * fill_tensor_a_terms blocks of 96 x fill_tensor_k bf16 bit patterns in the
 *              canonical K-major core-matrix layout [row / 8][k / 8][row % 8][k % 8]
 *   tile_kind  fill_tensor_tiles int32: operator tile used by each band tile
 *   tile_k0    fill_tensor_tiles int32: first pilot of each band tile's window
 * Fails with SRSCE_ERR_UNSUPPORTED (reason in srsce_last_error) when the plan has no tensor-core fill. */
int srsce_plan_fill_tensor_tables(const srsce_plan* plan, uint16_t* a_tiles, int32_t* tile_kind, int32_t* tile_k0);

/* The fused per-slot path, batched over independent units (one unit = one reference call =
 * one (slot, UE, rx antenna)):
 *   rx_grid   complex64, unit u at rx_grid + u*rx_strides[0], element (sc, sym) at
 *             + sc*rx_strides[1] + sym*rx_strides[2]   (strides in complex64 elements; the
 *             reference layout [n_sc, n_sym] is {n_sc*n_sym, n_sym, 1})
 *   pilots    complex64 [n_units / pilot_share][n_re][n_dmrs_total][n_layers] contiguous; unit u
 *             reads block u / pilot_share (the rx antennas of one UE share its pilots); 16-byte aligned
 *             when n_layers is even (8-byte otherwise)
 *   beta      beta_dmrs
 *   ch_est    complex64 [n_units][n_sc][n_sym][n_layers] contiguous, fully overwritten
 *   scalars   double [n_units][6]: noise_est, rsrp, epre, time_alignment (s), cfo_hz, cfo_valid
 *             (cfo_valid == 0 <=> the reference returns an empty cfo tensor)
 *   stream    cudaStream_t
 */
int srsce_estimate(const srsce_plan* plan, const void* rx_grid, const int64_t rx_strides[3],
                   const void* pilots, int32_t pilot_share, double beta, void* ch_est,
                   double* scalars, int64_t n_units, void* stream);

/* Compact-output variant (NOT part of the reference contract).  The reference spends
 * 13/14 of its output bytes replicating one column over the allocated symbols
 * (ce_rule_tensorized.py:346-347) and rotating it per symbol (:922-929).  This entry point returns the
 * factors instead:
 *   ch_sc    complex64 [n_units][n_hops][n_sc][n_layers]  interpolated channel of each hop (0 outside its band)
 *   sym_rot  complex64 [n_units][n_sym]                   final CFO rotation of each symbol (1 when none applies)
 * so that   channel_est_rg[sc][t][l] = sym_rot[t] * (hop-2 value if t is in hop 2's allocated symbols and sc in its
 * band, else hop-1 value if t is in hop 1's allocated symbols, else 0).  Everything else as srsce_estimate. */
int srsce_estimate_compact(const srsce_plan* plan, const void* rx_grid, const int64_t rx_strides[3],
                           const void* pilots, int32_t pilot_share, double beta, void* ch_sc, void* sym_rot,
                           double* scalars, int64_t n_units, void* stream);

/* srsce_estimate with an indirection on the received grid: unit u reads grid rx_index[u] (a device array of
 * n_units int64), located at rx_grids + rx_index[u]*rx_strides[0].  This is how a caller of the reference
 * passes the SAME received grid to several calls (all UEs multiplexed on one rx antenna in one slot) without
 * materialising copies; the grid is then read from HBM once and from L2 afterwards.  (A plan may run a different
 * CTA shape for this entry point and for srsce_estimate_compact than for srsce_estimate: the channel grids are
 * bit-identical, the scalars agree to float32 summation order.) */
int srsce_estimate_indexed(const srsce_plan* plan, const void* rx_grids, const int64_t rx_strides[3],
                           const int64_t* rx_index, const void* pilots, int32_t pilot_share, double beta,
                           void* ch_est, double* scalars, int64_t n_units, void* stream);

/* The fill operator alone -- the reference's plug-in seam fill_ch_est_cdm_fn of process_hop (:513-517, wired at
 * :838-841), i.e. ce_rule_tensorized.fill_ch_est_cdm (:237-354) for a SRSCE_FILL_LINEAR plan and
 * ce_dl_cnn.fill_ch_est_cdm (ce_dl_cnn.py:233-322: the fixed-weight 3-tap "CNN" inpainting) for a SRSCE_FILL_CNN plan,
 * applied to every CDM group of every hop of a batch of units on a zero grid (what process_hop does at :727):
 *   h_pilots  complex64 [n_units][n_hops][n_re][n_layers] contiguous: the (smoothed) pilot estimates h_p of each hop
 *   ch_est    complex64 [n_units][n_sc][n_sym][n_layers], fully overwritten (zero outside band / allocated symbols;
 *             hop 2 wins where the hops overlap, as in srsce_estimate; no CFO rotation -- that is not part of the fill)
 *   impl      SRSCE_FILL_IMPL_DIRECT: operator rows applied with float32 FMAs by the fused kernel's writer warps;
 *             SRSCE_FILL_IMPL_TENSOR: the same operator as a banded GEMM on the tcgen05 tensor cores (bf16 operand
 *             terms, float32 accumulation in tensor memory).  Available for single-hop plans on a 14-symbol grid
 *             whose operator tiles fit in shared memory; SRSCE_ERR_UNSUPPORTED otherwise (srsce_plan_info.fill_tensor).
 * Both implementations are HBM-write bound (0.4 FLOP/B); see DESIGN.md section 3.1. */
enum { SRSCE_FILL_IMPL_DIRECT = 0, SRSCE_FILL_IMPL_TENSOR = 1 };
int srsce_fill(const srsce_plan* plan, const void* h_pilots, void* ch_est, int64_t n_units, int32_t impl, void* stream);

/* srsce_fill (direct implementation) with the output left as per-hop columns, like srsce_estimate_compact:
 *   ch_sc  complex64 [n_units][n_hops][n_sc][n_layers]: the interpolated band of each hop (0 outside it); the replication
 *          over the allocated symbols (:346-347) is left to the consumer -- srsce_replicate_host(..., SRSCE_REPLICATE_FILL)
 *          for a caller that wants the dense grid in host memory.  Not part of the reference contract. */
int srsce_fill_compact(const srsce_plan* plan, const void* h_pilots, void* ch_sc, int64_t n_units, void* stream);

/* EXTENSION (no counterpart in the reference, which takes the pilots as an input tensor,
 * ce_rule_tensorized.py:746-747 -- parity unpinned): low-PAPR pilot sequences of TS 38.211 5.2.2 generated on the device
 * in the estimator's pilot layout.
 *   pilots  complex64 [n_ue][m_zc][n_dmrs][n_layers]: pilot n of UE b on every pilot symbol is rbar_{u,v}(n) for even
 *           layers and e^{j pi n} rbar_{u,v}(n) for odd layers (cyclic shifts 0 and pi: the OCC-2 pair of a comb)
 *   m_zc    sequence length = pilots per CDM group (30, or >= 36: Zadoff-Chu based with N_zc the largest prime < m_zc;
 *           6 / 12 / 18 / 24 are table defined in the specification and not supported)
 *   u_of    device int32 [n_ue]: sequence group u in 0..29;  v_of: device int32 [n_ue] base sequence number 0 / 1, or NULL */
int srsce_gen_pilots(void* pilots, int64_t n_ue, int32_t m_zc, int32_t n_dmrs, int32_t n_layers, const int32_t* u_of,
                     const int32_t* v_of, void* stream);

/* Host half of a host-buffer call (the reference's callers hold host arrays: scripts/validation/validate_all.py:367-382,
 * :549-561).  The dense grid repeats one column over the allocated symbols (ce_rule_tensorized.py:346-347), so 13/14 of
 * the bytes a device->host copy of it moves are copies; this call rebuilds the grid in HOST memory from the per-hop
 * columns srsce_estimate_compact returned, on host threads, with non-temporal stores:
 *   ch_sc      HOST complex64 [n_units][n_hops][n_sc][n_layers]  (srsce_estimate_compact's ch_sc, copied to the host)
 *   ch_est     HOST complex64 [n_units][n_sc][n_sym][n_layers], fully overwritten -- bit-identical to what
 *              srsce_estimate writes on the device for the same inputs
 *   n_threads  host threads to use (<= 0: all the cores); a small persistent pool inside the library
 *   flags      0, or SRSCE_REPLICATE_FILL: the columns come from srsce_fill_compact (the fill never rotates)
 * It performs NO arithmetic, only the copy: a plan whose grid carries the per-symbol CFO rotation (:922-929; cfo_defined
 * and cfo_compensate) fails with SRSCE_ERR_UNSUPPORTED unless SRSCE_REPLICATE_FILL is set -- such plans take
 * srsce_estimate and a dense device->host copy.
 * Needs no CUDA device; blocking; may be called from several threads (calls are serialised on the pool). */
enum { SRSCE_REPLICATE_FILL = 1 };
int srsce_replicate_host(const srsce_plan* plan, const void* ch_sc, void* ch_est, int64_t n_units, int32_t n_threads,
                         int32_t flags);

/* Number of kernel launches the last srsce_estimate on this thread issued (bench bookkeeping). */
int srsce_last_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* SRSCE_H_ */
