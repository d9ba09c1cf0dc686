"""CPU oracle for the srsRAN port channel estimator hot path (numpy restatement).

TEST INFRASTRUCTURE ONLY.  Nothing on the product path may import this module: only
``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline / ``--impl reference`` legs of
``bench.py`` are allowed to.  The product (``srsran_ce_pytorch_b200``) runs hand-written CUDA
through the C ABI in ``include/srsce.h`` and has no CPU fallback.

What it restates (all citations are into /root/reference/src):
  * ``ce_rule_tensorized.py``  srs_channel_estimator :745-937, process_hop :495-739,
    compensate_cfo :357-451, create_virtual_pilots :69-140, _unwrap_1d :35-66,
    get_rc_filter :184-234, _rcosdesign_normal :143-181, _conv_same_1d_complex :459-493,
    fill_ch_est_cdm :237-354
  * ``ce_rule_baseline.py``    same algorithm, loop form (fill :237-360) -> ``fill="baseline"``
  * ``ce_dl_cnn.py``           fill_ch_est_cdm :233-322, _cnn_inpaint_1d_complex :473-508,
    _cnn_lowpass_1d_complex :454-470, _conv1d_same_real_reflect :433-451, blend :712-717

Parity pin: the reference itself is Python and runs in the build container, so this port is
pinned against outputs of the *unmodified reference* on seeded inputs
(``oracle/gen_golden.py`` -> ``tests/golden/*.npz``; checked by ``tests/test_oracle_golden.py``).
The srsRAN/MATLAB ``testvector_outputs`` files are git-ignored upstream and absent here, so
parity against the MATLAB vectors themselves is UNPINNED.

The precision islands of the reference are kept: complex64 data, float32 |.|^2 sums and
``angle``; float64 for the virtual-pilot regression, phase unwrap, raised-cosine taps and
convolution, CFO phase tables and every scalar accumulator.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

NRE = 12            # resource elements per PRB
TA_FFT = 4096       # ce_rule_tensorized.py:677
TA_HALF_CP = 144    # floor((144/2)*4096/2048), :682

c64 = np.complex64
f32 = np.float32
f64 = np.float64


# --------------------------------------------------------------------------------------
# geometry containers (plain numpy mirrors of HopConfig / EstimatorConfig, :13-29)
# --------------------------------------------------------------------------------------
@dataclass
class HopSpec:
    dmrs_symbols: np.ndarray      # bool [n_sym] (empty => hop absent)
    re_mask: np.ndarray           # bool [12, >=n_cdm]
    prb_start: int
    n_prbs: int
    mask_prbs: np.ndarray         # bool [n_prb_grid]
    start_symbol: int
    n_alloc_symbols: int

    @staticmethod
    def absent() -> "HopSpec":
        z = np.zeros((0,), dtype=bool)
        return HopSpec(z, np.zeros((12, 0), dtype=bool), 0, 0, z, 0, 0)

    @property
    def present(self) -> bool:
        return self.dmrs_symbols.size != 0 and bool(self.dmrs_symbols.any())


@dataclass
class EstSpec:
    scs: float                                    # Hz
    cp_ms: np.ndarray                             # float64 [>=14], cyclic prefix durations in ms
    smoothing: str = "filter"                     # filter | mean | none
    cfo_compensate: bool = True
    cnn_alpha: float = 0.0                        # ce_dl_cnn.py:864-867 (only used by fill="cnn")


def _np(x, dtype=None):
    """Accept numpy arrays, lists or CPU torch tensors."""
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    a = np.asarray(x)
    return a.astype(dtype) if dtype is not None else a


def hop_from_any(h) -> HopSpec:
    """Duck-typed conversion from a reference-style HopConfig (MATLAB-cased fields)."""
    if isinstance(h, HopSpec):
        return h
    sym = _np(h.DMRSsymbols)
    if sym.size == 0:
        return HopSpec.absent()
    rm = _np(h.DMRSREmask, bool)
    if rm.ndim == 1:
        rm = rm.reshape(12, -1)
    return HopSpec(sym.astype(bool), rm, int(h.PRBstart), int(h.nPRBs), _np(h.maskPRBs, bool),
                   int(h.startSymbol), int(h.nAllocatedSymbols))


def est_from_any(c) -> EstSpec:
    if isinstance(c, EstSpec):
        return c
    sm = getattr(c, "Smoothing", None)
    return EstSpec(float(c.scs), _np(c.CyclicPrefixDurations, f64),
                   "filter" if sm is None else str(sm),
                   bool(getattr(c, "CFOCompensate", True)),
                   float(getattr(c, "CNNSmoothingAlpha", 0.0) or 0.0))


# --------------------------------------------------------------------------------------
# raised-cosine taps (get_rc_filter :184-234, _rcosdesign_normal :143-181)
# --------------------------------------------------------------------------------------
def rc_taps(stride: int, n_rbs: int) -> np.ndarray:
    """Raised cosine (roll-off 0.2, span n_rbs symbols, 10 samples/symbol), centre tap kept,
    every ``stride``-th sample on both sides, normalised to unit sum.  float64."""
    if stride <= 0:
        raise ValueError("stride must be >= 1")
    if n_rbs <= 0:
        raise ValueError("n_rbs must be >= 1")
    sps, beta = 10, 0.2
    half = (n_rbs * sps) // 2                               # sps is even -> symmetric grid
    n = np.arange(-half, half + 1, dtype=f64)
    t = n / float(sps)
    with np.errstate(divide="ignore", invalid="ignore"):
        sinc = np.where(t == 0, 1.0, np.sin(math.pi * t) / (math.pi * t))
        h = sinc * np.cos(math.pi * beta * t) / (1.0 - (2.0 * beta * t) ** 2)
    t0 = 1.0 / (2.0 * beta)
    sing = ~np.isfinite(h) | (np.abs(np.abs(t) - t0) < (1.0 / sps) * 1e-6)
    h = np.where(sing, (math.pi * beta / 2.0) * math.sin(1.0 / (2.0 * beta)), h)
    length = h.size
    centre = (length - 1) // 2
    reach = ((length // 2) // stride) * stride
    taps = h[centre - reach: centre + reach + 1: stride].copy()
    return taps / taps.sum()


def n_virtual(n_taps: int, n_prbs_in_mask: int, re_per_prb: int) -> int:
    """:644-647"""
    return min(12, n_taps // 2) if n_prbs_in_mask > 1 else re_per_prb


# --------------------------------------------------------------------------------------
# virtual pilots (create_virtual_pilots :69-140, _unwrap_1d :35-66)
# --------------------------------------------------------------------------------------
def _unwrap(ph: np.ndarray) -> np.ndarray:
    if ph.size <= 1:
        return ph
    two_pi = 2.0 * math.pi
    d = ph[1:] - ph[:-1]
    # floored modulo, as torch.remainder
    dm = np.mod(d + math.pi, two_pi) - math.pi
    dm = np.where((dm == -math.pi) & (d > 0), dm + two_pi, dm)
    corr = np.where(np.abs(d) < math.pi, 0.0, dm - d)
    return ph + np.concatenate([[0.0], np.cumsum(corr)])


def virtual_pilots(edge: np.ndarray, n_virt: int) -> np.ndarray:
    """Extrapolate ``n_virt`` pilots in front of ``edge`` (complex64 [n]) from straight-line
    fits of modulus and unwrapped phase.  float64 inside, complex64 out."""
    if n_virt == 0:
        return np.zeros((0,), dtype=c64)
    n = edge.size
    if n == 1:                                             # :95-101 constant extrapolation
        amp = np.abs(edge).astype(f32)
        ph = np.angle(edge).astype(f32)
        return np.repeat((amp * np.exp(1j * ph.astype(c64))).astype(c64), n_virt)
    x = np.arange(n, dtype=f64)
    mx = x.mean()
    sxx = np.linalg.norm(x) ** 2
    den = sxx - n * mx * mx
    k = np.arange(-n_virt, 0, dtype=f64)

    def line(y):
        my = y.mean()
        a = ((x * y).sum() - n * mx * my) / den
        return a * k + (my - a * mx)

    mod = line(np.abs(edge).astype(f32).astype(f64))
    arg = line(_unwrap(np.angle(edge).astype(f32).astype(f64)))
    return (mod * np.exp(1j * arg)).astype(c64)


def conv_same(x: np.ndarray, taps: np.ndarray) -> np.ndarray:
    """conv(x, taps, 'same') on complex64 x with float64 real taps, float64 accumulation,
    zero padding len(taps)//2 each side (:459-493)."""
    k = taps.size
    if k == 0:
        return x.copy()
    pad = k // 2
    xr = np.concatenate([np.zeros(pad), x.real.astype(f64), np.zeros(pad)])
    xi = np.concatenate([np.zeros(pad), x.imag.astype(f64), np.zeros(pad)])
    # full correlation with flipped taps == convolution; keep len(x) (+ (k even) extra sample)
    yr = np.convolve(xr, taps, mode="valid")
    yi = np.convolve(xi, taps, mode="valid")
    return (yr + 1j * yi).astype(c64)


# --------------------------------------------------------------------------------------
# the fixed-stencil "CNN" of ce_dl_cnn.py (:433-508)
# --------------------------------------------------------------------------------------
_H3 = np.array([0.25, 0.5, 0.25], dtype=f64)


def _stencil_reflect(v: np.ndarray) -> np.ndarray:
    """[.25,.5,.25] along axis 0 with reflect (length-1: replicate) padding; float64."""
    if v.shape[0] == 1:
        ext = np.concatenate([v, v, v], axis=0)
    else:
        ext = np.concatenate([v[1:2], v, v[-2:-1]], axis=0)
    return 0.25 * ext[:-2] + 0.5 * ext[1:-1] + 0.25 * ext[2:]


def cnn_lowpass(x: np.ndarray, passes: int = 2) -> np.ndarray:
    """x complex64 [n, ...]; ``passes`` stencil passes, float64 inside (:454-470)."""
    if x.shape[0] <= 2:
        return x
    yr, yi = x.real.astype(f64), x.imag.astype(f64)
    for _ in range(max(1, int(passes))):
        yr, yi = _stencil_reflect(yr), _stencil_reflect(yi)
    return (yr + 1j * yi).astype(c64)


def cnn_inpaint(sparse: np.ndarray, known: np.ndarray, n_iters: int) -> np.ndarray:
    """Partial-convolution inpainting of complex64 ``sparse`` [n, cols] known at ``known`` [n]
    (:473-508).  The complex64 round trip after every iteration is kept."""
    if known.all():
        return cnn_lowpass(sparse, 2)
    cur = sparse.copy()
    m = known.astype(f64)
    kn = known[:, None]
    for _ in range(max(1, int(n_iters))):
        den = _stencil_reflect(m)
        mm = m[:, None]
        num_r = _stencil_reflect(cur.real.astype(f64) * mm)
        num_i = _stencil_reflect(cur.imag.astype(f64) * mm)
        d = (den + 1e-12)[:, None]
        prop = (num_r / d + 1j * (num_i / d)).astype(c64)
        m = np.maximum(m, (den > 1e-12).astype(f64))
        cur = np.where(kn, sparse, prop)
    return np.where(kn, sparse, cnn_lowpass(cur, 2))


# --------------------------------------------------------------------------------------
# grid fill (fill_ch_est_cdm)
# --------------------------------------------------------------------------------------
def _fill_band(h_grp: np.ndarray, hop: HopSpec, grp: int, fill: str) -> np.ndarray:
    """Expand pilots of one CDM group ([n_re, layers_in_group] complex64) to every subcarrier
    of the hop band -> [12*n_prbs, layers_in_group] complex64."""
    n_band = NRE * hop.n_prbs
    mask = np.tile(hop.re_mask[:, grp], hop.n_prbs)
    at = np.flatnonzero(mask)
    if fill == "cnn":                                       # ce_dl_cnn.py:274-295
        sparse = np.zeros((n_band, h_grp.shape[1]), dtype=c64)
        sparse[at] = h_grp
        return cnn_inpaint(sparse, mask, max(6, n_band // 8))
    pos = np.arange(n_band)
    out = np.empty((n_band, h_grp.shape[1]), dtype=c64)
    if fill == "baseline":                                  # ce_rule_baseline.py:302-349
        out[at] = h_grp
        for a, b in zip(at[:-1], at[1:]):
            gap = b - a - 1
            if gap <= 0:
                continue
            w = (np.arange(1, gap + 1, dtype=f64) / float(gap + 1)).astype(f32)
            out[a + 1:b] = out[a][None, :] + w[:, None].astype(c64) * (out[b] - out[a])[None, :]
        out[at[-1]:] = out[at[-1]]
        out[:at[0] + 1] = out[at[0]]
        return out
    # tensorized form (:311-340): every strictly-interior subcarrier (pilot positions included)
    # is left + alpha*(right-left) with float32 alpha; edges are flat.
    lo, hi = at[0], at[-1]
    out[pos <= lo] = h_grp[0]
    out[pos >= hi] = h_grp[-1]
    mid = pos[(pos > lo) & (pos < hi)]
    if mid.size:
        r_ord = np.searchsorted(at, mid, side="left")
        l_ord = r_ord - 1
        lp, rp = at[l_ord], at[r_ord]
        alpha = ((mid.astype(f32) - lp.astype(f32)) / (rp.astype(f32) - lp.astype(f32)))[:, None]
        lv, rv = h_grp[l_ord], h_grp[r_ord]
        out[mid] = lv + alpha.astype(c64) * (rv - lv)
    return out


# --------------------------------------------------------------------------------------
# one hop (process_hop :495-739)
# --------------------------------------------------------------------------------------
@dataclass
class _Acc:
    epre: float = 0.0
    noise: float = 0.0
    rsrp: float = 0.0
    ta: float = 0.0
    cfo: Optional[float] = None
    ta_margin: list = field(default_factory=list)   # diagnostic: top-2 relative gap per hop
    cfo_kappa: float = 1.0                          # diagnostic: worst summation condition number of a CFO correlation


def symbol_start_times(cp_ms: np.ndarray, scs: float) -> np.ndarray:
    """cumsum([CP0, CP1+1, ... CP13+1]) in OFDM-symbol units (:809-820)."""
    cpd = np.asarray(cp_ms, dtype=f64) * scs / 1000.0
    if cpd.size < 14:
        raise ValueError("config.CyclicPrefixDurations must have length >= 14 to match MATLAB code.")
    return np.cumsum(np.concatenate([cpd[:1], cpd[1:14] + 1.0]))


def _hop(rx: np.ndarray, pilots: np.ndarray, beta: float, hop: HopSpec, cfg: EstSpec,
         t_sym: np.ndarray, grid: np.ndarray, acc: _Acc, fill: str) -> None:
    n_layers = pilots.shape[2]
    n_cdm = (n_layers + 1) // 2
    sym_idx = np.flatnonzero(hop.dmrs_symbols)
    n_dmrs = sym_idx.size
    n_prb_mask = int(hop.mask_prbs.sum())
    per_prb = int(hop.re_mask[:, 0].sum())

    # --- gather + EPRE + LS (:571-593)
    y = np.empty((pilots.shape[0], n_dmrs, n_cdm), dtype=c64)
    x = np.empty_like(pilots)
    re_sel = None
    for g in range(n_cdm):
        re_sel = np.tile(hop.re_mask[:, g], hop.mask_prbs.size) & np.repeat(hop.mask_prbs, NRE)
        yg = rx[re_sel][:, sym_idx]
        y[:, :, g] = yg
        acc.epre += float((np.abs(yg) ** 2).sum(dtype=f32))
        lo, hi = 2 * g, min(n_layers, 2 * g + 2)
        x[:, :, lo:hi] = yg[:, :, None] * np.conj(pilots[:, :, lo:hi])

    # --- CFO (:357-451, :596-609)
    cfo_hop = None
    if n_dmrs >= 2:
        cpd = cfg.cp_ms.astype(f64) * (cfg.scs / 1000.0)
        prod = np.conj(x[:, 0, :]) * x[:, 1, :]
        # (per layer over a contiguous vector, i.e. pairwise like torch's sum -- see the mean smoothing below)
        inner = np.array([np.ascontiguousarray(prod[:, l]).sum(dtype=c64) for l in range(n_layers)], dtype=c64)
        n_even = (n_layers // 2) * 2
        tot = 0.0
        if n_even:
            tot += float(np.angle(inner[:n_even].reshape(-1, 2).sum(axis=1, dtype=c64)).astype(f32)
                         .astype(f64).sum())
        if n_layers % 2:
            tot += float(f32(np.angle(inner[-1])))
        # diagnostic only (tests/tolerances.py): sum |terms| / |sum terms| of each correlation whose angle is taken --
        # a layer with no signal leaves a float32 sum of noise products that nearly cancels, and its angle is then
        # determined by rounding (the reference takes the angle all the same)
        terms = (np.conj(x[:, 0, :]) * x[:, 1, :]).astype(np.complex128)
        groups = [terms[:, k:k + 2].sum(axis=1) for k in range(0, n_even, 2)] + ([terms[:, -1]] if n_layers % 2 else [])
        for t in groups:
            tot_abs = abs(t.sum())
            acc.cfo_kappa = max(acc.cfo_kappa, float(np.abs(t).sum() / tot_abs) if tot_abs > 0 else float("inf"))
        n_samp = float(sym_idx[1] - sym_idx[0]) + float(cpd[sym_idx[0] + 1: sym_idx[1] + 1].sum())
        cfo_hop = tot / (2.0 * math.pi * n_samp) / float(n_cdm)
        if cfg.cfo_compensate:
            if cpd.size < 14:
                raise ValueError("cyclic_prefix_durations must have length >= 14 to match MATLAB code.")
            t_loc = np.cumsum(np.concatenate([cpd[:1], cpd[1:14] + 1.0]))
            rot = np.exp(-1j * (2.0 * math.pi * t_loc * cfo_hop)[sym_idx]).astype(c64)
            x = x * rot[None, :, None]
        acc.cfo = cfo_hop if acc.cfo is None else (acc.cfo + cfo_hop) / 2

    # --- symbol average, CDM pair average (:613-628)
    h = (x.sum(axis=1, dtype=c64) / c64(f32(beta))) / c64(f32(n_dmrs))
    if n_layers >= 2:
        m = h.shape[0] // 2
        if m:
            avg = (h[0:2 * m:2] + h[1:2 * m:2]) / c64(2)
            h[0:2 * m:2] = avg
            h[1:2 * m:2] = avg

    # --- smoothing (:633-668)
    if cfg.smoothing == "mean":
        # (per layer over a contiguous vector: numpy then sums pairwise, like torch's cascade sum -- a reduction along
        #  axis 0 of the [N, L] array would accumulate row by row, and over >= 1 000 cancelling pilots that sequential
        #  float32 sum is 30x further from the float64 result than the reference's own float32 mean)
        mean = np.array([np.ascontiguousarray(h[:, l]).sum(dtype=c64) for l in range(h.shape[1])], dtype=c64) / c64(h.shape[0])
        h = np.ones_like(h) * mean[None, :]
    elif cfg.smoothing == "filter":
        taps = rc_taps(12 // per_prb, min(3, n_prb_mask))
        nv = n_virtual(taps.size, n_prb_mask, per_prb)
        for l in range(n_layers):
            col = h[:, l]
            head = virtual_pilots(col[:nv], nv)
            tail = virtual_pilots(col[::-1][:nv] if nv else col[:0], nv)
            ext = np.concatenate([head, col, tail[::-1]])
            sm = conv_same(ext, taps)[nv: nv + col.size]
            if fill == "cnn" and cfg.cnn_alpha > 0.0:       # ce_dl_cnn.py:712-717
                a = float(max(0.0, min(1.0, cfg.cnn_alpha)))
                sm = (sm + c64(a) * (cnn_lowpass(sm, 1) - sm)).astype(c64)
            h[:, l] = sm
    elif cfg.smoothing != "none":
        raise ValueError(f"Unknown smoothing strategy {cfg.smoothing}.")

    # --- time alignment (:674-698); positions of the LAST CDM group, all layers
    stuffed = np.zeros((re_sel.size, n_layers), dtype=c64)
    stuffed[re_sel] = h
    prof = (np.abs(np.fft.ifft(stuffed, n=TA_FFT, axis=0).astype(c64)) ** 2).sum(axis=1, dtype=f32)
    head, tail = prof[:TA_HALF_CP], prof[-TA_HALF_CP:]
    i_d, i_a = int(np.argmax(head)), int(np.argmax(tail))
    i_max = i_d if head[i_d] >= tail[i_a] else -(TA_HALF_CP - i_a)
    acc.ta += float(i_max) / float(TA_FFT) / float(cfg.scs)
    both = np.sort(np.concatenate([head, tail]))[::-1]
    acc.ta_margin.append(float((both[0] - both[1]) / both[0]) if both[0] > 0 else 0.0)

    # --- reconstruction, fill, noise, RSRP (:701-730)
    y_hat = np.zeros_like(y)
    for g in range(n_cdm):
        lo, hi = 2 * g, min(n_layers, 2 * g + 2)
        hs = h[:, None, lo:hi]
        if cfg.cfo_compensate and cfo_hop is not None:
            ph = np.exp(1j * (2.0 * math.pi * t_sym * cfo_hop)[hop.dmrs_symbols]).astype(c64)
            hs = hs * ph[None, :, None]
        y_hat[:, :, g] = c64(f32(beta)) * (pilots[:, :, lo:hi] * hs).sum(axis=2, dtype=c64)
        band = _fill_band(h[:, lo:hi], hop, g, fill)
        s0, s1 = hop.start_symbol, hop.start_symbol + hop.n_alloc_symbols
        b0 = NRE * hop.prb_start
        grid[b0: b0 + band.shape[0], s0:s1, lo:hi] = band[:, None, :]
    acc.noise += float((np.abs(y - y_hat) ** 2).sum(dtype=f32))
    acc.rsrp += (float(beta) ** 2) * float((np.abs(h) ** 2).sum(dtype=f32)) * float(n_dmrs)


# --------------------------------------------------------------------------------------
# entry point (srs_channel_estimator :745-937)
# --------------------------------------------------------------------------------------
def estimate(received_rg, pilots, beta_dmrs, hop1, hop2, config, fill: str = "tensorized",
             diagnostics: Optional[dict] = None):
    """Returns (channel_est_rg c64 [n_sc, n_sym, L], noise, rsrp, epre, time_alignment,
    cfo_hz or None).  ``fill``: "tensorized" | "baseline" | "cnn"."""
    rx = _np(received_rg)
    pil = _np(pilots)
    rx = rx.astype(c64) if not np.iscomplexobj(rx) else rx.astype(c64, copy=False)
    pil = pil.astype(c64) if not np.iscomplexobj(pil) else pil.astype(c64, copy=False)
    h1, h2, cfg = hop_from_any(hop1), hop_from_any(hop2), est_from_any(config)
    beta = float(beta_dmrs)
    n_layers = pil.shape[2]
    grid = np.zeros((rx.shape[0], rx.shape[1], n_layers), dtype=c64)
    acc = _Acc()
    n1 = int(h1.dmrs_symbols.sum())
    t_sym = symbol_start_times(cfg.cp_ms, cfg.scs) if cfg.cfo_compensate else np.zeros((0,))

    _hop(rx, pil[:, :n1, :], beta, h1, cfg, t_sym, grid, acc, fill)
    all_sym = h1.dmrs_symbols.copy()
    if h2.present:
        assert not (all_sym & h2.dmrs_symbols).any(), "Hops should not overlap."
        assert np.array_equal(h1.re_mask, h2.re_mask), "The DM-RS mask should be the same for the two hops."
        all_sym |= h2.dmrs_symbols
        _hop(rx, pil[:, n1:, :], beta, h2, cfg, t_sym, grid, acc, fill)

    n_pil = h1.n_prbs * int(h1.re_mask[:, 0].sum()) * int(all_sym.sum())
    rsrp = acc.rsrp / float(n_pil) / float(n_layers)
    epre = acc.epre / float(n_pil)
    with np.errstate(divide="ignore", invalid="ignore"):      # one pilot in total: x/0 like the reference's tensors
        noise = float(np.float64(acc.noise) / np.float64(((n_layers + 1) // 2) * n_pil - 1))
    ta = acc.ta / 2.0 if h2.present else acc.ta
    if cfg.cfo_compensate and acc.cfo is not None:
        rot = np.exp(1j * (2.0 * math.pi * t_sym * acc.cfo)).astype(c64)
        grid = grid * rot[None, :, None]
    if diagnostics is not None:
        diagnostics["ta_margin"] = acc.ta_margin
        diagnostics["cfo_kappa"] = acc.cfo_kappa
    return grid, noise, rsrp, epre, ta, (None if acc.cfo is None else acc.cfo * cfg.scs)
