"""Host side of the estimator: reference-compatible entry points over the C ABI.

``srs_channel_estimator`` has the signature, argument meaning, return tuple and error behaviour of
the reference function of the same name (/root/reference src/ce_rule_tensorized.py:745-937; the
same in ce_rule_baseline.py:761 and ce_dl_cnn.py:802) but only accepts CUDA tensors: everything is
computed by the fused sm_100a kernel in ``libsrsce.so``.  ``srs_channel_estimator_batched`` is the
same call over many independent units (slot x UE x rx antenna) in one launch.

PyTorch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
import weakref
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import lib

NRE = 12
_SCALAR_FIELDS = ("noise_est", "rsrp", "epre", "time_alignment", "cfo_hz", "cfo_valid")


def _host(x) -> np.ndarray:
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    return np.asarray(x)


def _hop_desc(hop, n_sym: int, n_prb_grid: int, tag: str) -> Tuple[_lib.HopDesc, Optional[np.ndarray]]:
    """HopConfig (reference field names, :13-21) -> ABI hop descriptor + its RE mask [12, n_cdm]."""
    d = _lib.HopDesc()
    sym = _host(hop.DMRSsymbols).astype(bool).reshape(-1)
    if sym.size == 0 or not sym.any():
        return d, None                                    # "hop absent" sentinel (validate_all.py:449-457)
    if sym.size != n_sym:
        raise IndexError(f"{tag}.DMRSsymbols has {sym.size} entries for a grid of {n_sym} symbols")
    prbs = _host(hop.maskPRBs).astype(bool).reshape(-1)
    if prbs.size != n_prb_grid:
        raise IndexError(f"{tag}.maskPRBs has {prbs.size} entries for a grid of {n_prb_grid} PRBs")
    start, n = int(hop.PRBstart), int(hop.nPRBs)
    want = np.zeros_like(prbs)
    want[max(0, start):max(0, start + n)] = True
    if not np.array_equal(prbs, want):
        raise ValueError(f"{tag}.maskPRBs must be exactly the contiguous range [PRBstart, PRBstart+nPRBs)")
    mask = 0
    for t in np.flatnonzero(sym):
        mask |= 1 << int(t)
    d.dmrs_symbol_mask = mask
    d.prb_start, d.n_prbs = start, n
    d.start_symbol, d.n_alloc_symbols = int(hop.startSymbol), int(hop.nAllocatedSymbols)
    rm = _host(hop.DMRSREmask).astype(bool)
    if rm.ndim == 1:
        rm = rm.reshape(NRE, -1)
    if rm.shape[0] != NRE:
        raise ValueError(f"{tag}.DMRSREmask must have 12 rows")
    return d, rm


class Plan:
    """Data-independent part of a call (geometry, filter taps, operator tables), created once and
    cached.  Thin owner of an ``srsce_plan*``."""

    def __init__(self, desc: _lib.PlanDesc):
        self.desc = desc
        handle = C.c_void_p()
        _lib.raise_for(lib.srsce_plan_create(C.byref(desc), C.byref(handle)))
        self._h = handle
        info = _lib.PlanInfo()
        _lib.raise_for(lib.srsce_plan_query(self._h, C.byref(info)))
        self.info = info

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and lib is not None:                      # lib is None during interpreter shutdown
            lib.srsce_plan_destroy(h)

    # ---- host tables (tests of the host logic; no device needed)
    def tables(self):
        i = self.info
        taps = np.zeros(max(1, i.n_taps), dtype=np.float64)
        rows = NRE * self.desc.hop[0].n_prbs
        idx = np.zeros((i.n_cdm, rows), dtype=np.int32)
        coef = np.zeros((i.n_cdm, rows, i.fill_terms), dtype=np.float32)
        _lib.raise_for(lib.srsce_plan_tables(self._h, taps.ctypes.data, idx.ctypes.data, coef.ctypes.data))
        return taps[:i.n_taps], idx, coef

    def fill_tensor_tables(self):
        """Operator tiles of the tensor-core fill as float64 [kinds, n_cdm, 96, K] (bf16 terms summed), the tile -> kind
        map and the first pilot of every tile's window."""
        i = self.info
        a = np.zeros((i.fill_tensor_kinds, i.n_cdm, i.fill_tensor_a_terms, 12, i.fill_tensor_k // 8, 8, 8), dtype=np.uint16)
        kind = np.zeros(i.fill_tensor_tiles, dtype=np.int32)
        k0 = np.zeros(i.fill_tensor_tiles, dtype=np.int32)
        _lib.raise_for(lib.srsce_plan_fill_tensor_tables(self._h, a.ctypes.data, kind.ctypes.data, k0.ctypes.data))
        f = (a.astype(np.uint32) << 16).view(np.float32).astype(np.float64).sum(axis=2)    # [kind, g, row/8, k/8, row%8, k%8]
        return f.transpose(0, 1, 2, 4, 3, 5).reshape(i.fill_tensor_kinds, i.n_cdm, 96, i.fill_tensor_k), kind, k0

    # ---- argument checks shared by the compute calls
    @staticmethod
    def _materialize(t: torch.Tensor) -> torch.Tensor:
        """The kernel reads raw storage: lazily conjugated / negated views (``x.conj()`` keeps the storage of
        ``x`` and only sets a flag) must be resolved before their data_ptr() crosses the ABI."""
        return t.resolve_conj().resolve_neg()

    def _check_inputs(self, rx, pilots, pilot_share, n_units=None):
        """rx [*, n_sc, n_sym] (any strides) and pilots [n_units // share, n_re, S, L]; returns
        (n_units, rx, pilots) with conj/neg views resolved and the pilots contiguous."""
        i = self.info
        if not rx.is_cuda or not pilots.is_cuda:
            raise RuntimeError("srsran_ce_pytorch_b200 computes on CUDA tensors only (no CPU fallback)")
        if rx.dtype != torch.complex64 or pilots.dtype != torch.complex64:
            raise TypeError("rx and pilots must be complex64")
        if rx.dim() != 3 or tuple(rx.shape[1:]) != (i.n_sc, i.n_sym):
            raise IndexError(f"received grid shape {tuple(rx.shape)} != [B, {i.n_sc}, {i.n_sym}]")
        if pilots.device != rx.device:
            raise RuntimeError("rx and pilots must live on the same CUDA device")
        n_units = int(rx.shape[0]) if n_units is None else int(n_units)
        if pilot_share < 1 or n_units % pilot_share:
            raise ValueError("number of units must be a multiple of pilot_share")
        want = (n_units // pilot_share, i.n_re, i.n_dmrs_total, i.n_layers)
        if tuple(pilots.shape) != want:
            raise IndexError(f"pilots shape {tuple(pilots.shape)} != {want}")
        return n_units, self._materialize(rx), self._materialize(pilots).contiguous()

    def _check_outputs(self, n_units, device, out, scalars):
        """Caller-supplied result buffers are written by the kernel without any further check: shape, dtype,
        contiguity and device must be exactly what the ABI documents."""
        i = self.info
        if out is None:
            out = torch.empty((n_units, i.n_sc, i.n_sym, i.n_layers), dtype=torch.complex64, device=device)
        elif tuple(out.shape) != (n_units, i.n_sc, i.n_sym, i.n_layers) or not out.is_contiguous() \
                or out.dtype != torch.complex64 or out.device != device or out.is_conj() or out.is_neg():
            raise ValueError("out must be a contiguous complex64 [B, n_sc, n_sym, L] tensor on rx.device")
        if scalars is None:
            scalars = torch.empty((n_units, 6), dtype=torch.float64, device=device)
        elif tuple(scalars.shape) != (n_units, 6) or not scalars.is_contiguous() or scalars.dtype != torch.float64 \
                or scalars.device != device:
            raise ValueError("scalars must be a contiguous float64 [B, 6] tensor on rx.device")
        return out, scalars

    # ---- the compute calls
    def estimate(self, rx: torch.Tensor, pilots: torch.Tensor, beta: float, pilot_share: int = 1,
                 out: Optional[torch.Tensor] = None, scalars: Optional[torch.Tensor] = None):
        """rx [B, n_sc, n_sym] complex64 CUDA (any strides), pilots [B // pilot_share, n_re, S, L]
        complex64.  Returns (ch_est [B, n_sc, n_sym, L], scalars [B, 6] float64)."""
        n_units, rx, pilots = self._check_inputs(rx, pilots, pilot_share)
        out, scalars = self._check_outputs(n_units, rx.device, out, scalars)
        strides = (C.c_int64 * 3)(*rx.stride())
        with torch.cuda.device(rx.device):
            stream = torch.cuda.current_stream().cuda_stream
            _lib.raise_for(lib.srsce_estimate(self._h, rx.data_ptr(), strides, pilots.data_ptr(), int(pilot_share),
                                              float(beta), out.data_ptr(), scalars.data_ptr(), n_units,
                                              C.c_void_p(stream)))
        return out, scalars

    def estimate_indexed(self, rx_grids: torch.Tensor, rx_index: torch.Tensor, pilots: torch.Tensor, beta: float,
                         pilot_share: int = 1, out: Optional[torch.Tensor] = None,
                         scalars: Optional[torch.Tensor] = None, check_index: bool = True):
        """Like ``estimate`` but unit u reads received grid ``rx_grids[rx_index[u]]``: the way to hand the same
        grid (one rx antenna, one slot) to the units of every UE multiplexed on it, as a caller of the reference
        does by passing the same tensor to several calls.  rx_grids [n_grids, n_sc, n_sym]; rx_index int64 [B].
        ``check_index`` verifies 0 <= rx_index < n_grids, which reads two values back from the device and so
        synchronises the stream; callers that build the index themselves pass False to stay asynchronous."""
        if not rx_index.is_cuda or rx_index.device != rx_grids.device:
            raise RuntimeError("rx_index must live on the CUDA device of rx_grids")
        if rx_index.dtype != torch.int64 or rx_index.dim() != 1:
            raise TypeError("rx_index must be a 1-D int64 tensor")
        n_units, rx_grids, pilots = self._check_inputs(rx_grids, pilots, pilot_share, n_units=rx_index.numel())
        if check_index and n_units and (int(rx_index.min()) < 0 or int(rx_index.max()) >= rx_grids.shape[0]):
            raise IndexError("rx_index out of range")
        rx_index = rx_index.contiguous()
        out, scalars = self._check_outputs(n_units, rx_grids.device, out, scalars)
        strides = (C.c_int64 * 3)(*rx_grids.stride())
        with torch.cuda.device(rx_grids.device):
            stream = torch.cuda.current_stream().cuda_stream
            _lib.raise_for(lib.srsce_estimate_indexed(self._h, rx_grids.data_ptr(), strides, rx_index.data_ptr(),
                                                      pilots.data_ptr(), int(pilot_share), float(beta), out.data_ptr(),
                                                      scalars.data_ptr(), n_units, C.c_void_p(stream)))
        return out, scalars

    def estimate_compact(self, rx: torch.Tensor, pilots: torch.Tensor, beta: float, pilot_share: int = 1):
        """Compact output (not part of the reference contract): returns (ch_sc [B, n_hops, n_sc, L],
        sym_rot [B, n_sym], scalars [B, 6]); ``expand_compact`` rebuilds the dense grid."""
        i = self.info
        n_units, rx, pilots = self._check_inputs(rx, pilots, pilot_share)
        n_hops = 2 if self.desc.hop[1].dmrs_symbol_mask else 1
        ch = torch.empty((n_units, n_hops, i.n_sc, i.n_layers), dtype=torch.complex64, device=rx.device)
        rot = torch.empty((n_units, i.n_sym), dtype=torch.complex64, device=rx.device)
        scalars = torch.empty((n_units, 6), dtype=torch.float64, device=rx.device)
        strides = (C.c_int64 * 3)(*rx.stride())
        with torch.cuda.device(rx.device):
            stream = torch.cuda.current_stream().cuda_stream
            _lib.raise_for(lib.srsce_estimate_compact(self._h, rx.data_ptr(), strides, pilots.data_ptr(),
                                                      int(pilot_share), float(beta), ch.data_ptr(), rot.data_ptr(),
                                                      scalars.data_ptr(), n_units, C.c_void_p(stream)))
        return ch, rot, scalars

    def _check_fill_input(self, h_pilots: torch.Tensor) -> Tuple[int, int, torch.Tensor]:
        i = self.info
        n_hops = 2 if self.desc.hop[1].dmrs_symbol_mask else 1
        if not h_pilots.is_cuda:
            raise RuntimeError("srsran_ce_pytorch_b200 computes on CUDA tensors only (no CPU fallback)")
        if h_pilots.dtype != torch.complex64:
            raise TypeError("h_pilots must be complex64")
        if h_pilots.dim() != 4 or tuple(h_pilots.shape[1:]) != (n_hops, i.n_re, i.n_layers):
            raise IndexError(f"h_pilots shape {tuple(h_pilots.shape)} != [B, {n_hops}, {i.n_re}, {i.n_layers}]")
        return int(h_pilots.shape[0]), n_hops, self._materialize(h_pilots).contiguous()

    def fill(self, h_pilots: torch.Tensor, impl: str = "direct", out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """The fill operator alone (``srsce_fill``): h_pilots [B, n_hops, n_re, L] complex64 (the smoothed pilot
        estimates of each hop) -> dense [B, n_sc, n_sym, L] grid, zero outside the band / allocated symbols.
        ``impl`` = "direct" (float32 FMAs in the fused kernel's writer warps) or "tensor" (tcgen05 banded GEMM;
        raises NotImplementedError when ``info.fill_tensor`` is 0)."""
        i = self.info
        if impl not in _lib.FILL_IMPL:
            raise ValueError(f"unknown fill implementation {impl!r}")
        n_units, _, h_pilots = self._check_fill_input(h_pilots)
        if out is None:
            out = torch.empty((n_units, i.n_sc, i.n_sym, i.n_layers), dtype=torch.complex64, device=h_pilots.device)
        elif tuple(out.shape) != (n_units, i.n_sc, i.n_sym, i.n_layers) or not out.is_contiguous() \
                or out.dtype != torch.complex64 or out.device != h_pilots.device or out.is_conj() or out.is_neg():
            raise ValueError("out must be a contiguous complex64 [B, n_sc, n_sym, L] tensor on h_pilots.device")
        with torch.cuda.device(h_pilots.device):
            stream = torch.cuda.current_stream().cuda_stream
            _lib.raise_for(lib.srsce_fill(self._h, h_pilots.data_ptr(), out.data_ptr(), n_units, _lib.FILL_IMPL[impl],
                                          C.c_void_p(stream)))
        return out

    def fill_compact(self, h_pilots: torch.Tensor) -> torch.Tensor:
        """``srsce_fill_compact``: the fill with the output left as per-hop columns [B, n_hops, n_sc, L] (the band of
        each hop interpolated, zero outside it); ``expand_compact(ch, ones)`` or ``replicate_host(ch, out,
        from_fill=True)`` rebuild the dense grid of ``fill``."""
        i = self.info
        n_units, n_hops, h_pilots = self._check_fill_input(h_pilots)
        ch = torch.empty((n_units, n_hops, i.n_sc, i.n_layers), dtype=torch.complex64, device=h_pilots.device)
        with torch.cuda.device(h_pilots.device):
            stream = torch.cuda.current_stream().cuda_stream
            _lib.raise_for(lib.srsce_fill_compact(self._h, h_pilots.data_ptr(), ch.data_ptr(), n_units, C.c_void_p(stream)))
        return ch

    def replicate_host(self, ch: torch.Tensor, out: torch.Tensor, n_threads: int = 0, from_fill: bool = False) -> torch.Tensor:
        """``srsce_replicate_host``: HOST tensors ch [B, n_hops, n_sc, L] (``estimate_compact``'s first output, copied
        to the host) -> the dense [B, n_sc, n_sym, L] grid in ``out`` (host), written by the library's host threads.
        A copy, no arithmetic: plans with a final CFO rotation raise NotImplementedError (use ``estimate``) unless the
        columns come from ``fill_compact`` (``from_fill=True``: the fill never rotates)."""
        i = self.info
        n_hops = 2 if self.desc.hop[1].dmrs_symbol_mask else 1
        for t in (ch, out):
            if t.is_cuda or t.dtype != torch.complex64 or not t.is_contiguous() or t.is_conj() or t.is_neg():
                raise ValueError("replicate_host takes contiguous complex64 host tensors")
        n_units = int(ch.shape[0])
        if tuple(ch.shape) != (n_units, n_hops, i.n_sc, i.n_layers):
            raise IndexError(f"ch shape {tuple(ch.shape)} != [B, {n_hops}, {i.n_sc}, {i.n_layers}]")
        if tuple(out.shape) != (n_units, i.n_sc, i.n_sym, i.n_layers):
            raise IndexError(f"out shape {tuple(out.shape)} != [{n_units}, {i.n_sc}, {i.n_sym}, {i.n_layers}]")
        _lib.raise_for(lib.srsce_replicate_host(self._h, ch.data_ptr(), out.data_ptr(), n_units, int(n_threads),
                                                _lib.REPLICATE_FILL if from_fill else 0))
        return out

    @property
    def has_final_rotation(self) -> bool:
        """The dense grid carries the per-symbol CFO rotation (:922-929): it cannot be rebuilt by replication alone."""
        return bool(self.desc.cfo_compensate) and bool(self.info.cfo_defined)

    def expand_compact(self, ch: torch.Tensor, rot: torch.Tensor) -> torch.Tensor:
        """[B, n_hops, n_sc, L] x [B, n_sym] -> the dense [B, n_sc, n_sym, L] grid of the reference contract
        (plain torch ops; for consumers and tests -- the fused dense path is ``estimate``)."""
        i, d = self.info, self.desc
        sym = torch.arange(i.n_sym, device=ch.device)
        grid = torch.zeros((ch.shape[0], i.n_sc, i.n_sym, i.n_layers), dtype=torch.complex64, device=ch.device)
        for hh in range(ch.shape[1]):
            h = d.hop[hh]
            in_sym = (sym >= h.start_symbol) & (sym < h.start_symbol + h.n_alloc_symbols)
            sc0, sc1 = 12 * h.prb_start, 12 * (h.prb_start + h.n_prbs)
            if hh == 0:
                grid[:, :, in_sym, :] = ch[:, 0, :, None, :].expand(-1, -1, int(in_sym.sum()), -1)
            else:
                grid[:, sc0:sc1, in_sym, :] = ch[:, hh, sc0:sc1, None, :].expand(-1, -1, int(in_sym.sum()), -1)
        return grid * rot[:, None, :, None]


_PLANS: dict = {}
_PLANS_LOCK = threading.Lock()
_MEMO: dict = {}                 # (ids of hop1 / hop2 / config, shapes, fill, device) -> (fingerprint, weakrefs, plan)
_MEMO_MAX = 256
_HOP_FIELDS = ("DMRSsymbols", "DMRSREmask", "PRBstart", "nPRBs", "maskPRBs", "startSymbol", "nAllocatedSymbols")
_CFG_FIELDS = ("scs", "CyclicPrefixDurations", "Smoothing", "CFOCompensate", "CNNSmoothingAlpha")


def _fingerprint(obj, fields):
    """Identity + in-place-modification counter of every tensor field, value of every plain field: equal
    fingerprints of the same (live) object mean its contents have not changed since the plan was made."""
    out = []
    for f in fields:
        v = getattr(obj, f, None)
        if isinstance(v, torch.Tensor):
            out.append((id(v), v._version))
        elif isinstance(v, np.ndarray):
            return None                                # numpy arrays carry no version counter: never memoised
        else:
            out.append(v)
    return tuple(out)


def _device_index(device) -> int:
    if not torch.cuda.is_available():
        return -1                                      # host-only plan (CPU-side tests of the planner)
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    return dev.index if dev.index is not None else torch.cuda.current_device()


def make_plan(n_prb_grid: int, n_sym: int, n_layers: int, hop1, hop2, config, fill: str = "linear",
              device=None) -> Plan:
    """Build (or fetch from the cache) the plan for HopConfig x2 + EstimatorConfig + tensor shapes.
    A plan's tables live on one CUDA device (``device``, default: the current one); the cache is per device.

    Turning the dataclasses into a descriptor reads their mask tensors on the host (a device->host copy and a
    stream synchronisation if they live on the GPU, as the reference's own ``.item()`` calls do).  That is done
    once per (hop1, hop2, config) object: repeated calls with the same, unmodified objects hit a memo keyed on
    object identity and tensor version counters and touch neither the tensors nor the stream."""
    dev_index = _device_index(device)
    memo_key = (id(hop1), id(hop2), id(config), int(n_prb_grid), int(n_sym), int(n_layers), fill, dev_index)
    fp = (_fingerprint(hop1, _HOP_FIELDS), _fingerprint(hop2, _HOP_FIELDS), _fingerprint(config, _CFG_FIELDS))
    hit = _MEMO.get(memo_key)
    if hit is not None and None not in fp and hit[0] == fp and all(r() is o for r, o in zip(hit[1], (hop1, hop2, config))):
        return hit[2]
    plan = _make_plan_uncached(n_prb_grid, n_sym, n_layers, hop1, hop2, config, fill, dev_index)
    if None not in fp:
        try:
            refs = tuple(weakref.ref(o) for o in (hop1, hop2, config))
        except TypeError:                              # objects that cannot be weakly referenced are not memoised
            return plan
        with _PLANS_LOCK:
            if len(_MEMO) >= _MEMO_MAX:
                _MEMO.clear()
            _MEMO[memo_key] = (fp, refs, plan)
    return plan


def _make_plan_uncached(n_prb_grid, n_sym, n_layers, hop1, hop2, config, fill, dev_index) -> Plan:
    desc = _lib.PlanDesc()
    desc.n_prb_grid, desc.n_sym, desc.n_layers = int(n_prb_grid), int(n_sym), int(n_layers)
    smoothing = getattr(config, "Smoothing", None)
    smoothing = "filter" if smoothing is None else str(smoothing)       # :785-787
    if smoothing not in _lib.SMOOTH:
        raise ValueError(f"Unknown smoothing strategy {smoothing}.")    # :668
    desc.smoothing = _lib.SMOOTH[smoothing]
    desc.cfo_compensate = 1 if bool(getattr(config, "CFOCompensate", True)) else 0   # :803-806
    desc.fill = _lib.FILL_CNN if fill == "cnn" else _lib.FILL_LINEAR
    alpha = getattr(config, "CNNSmoothingAlpha", 0.0) if fill == "cnn" else 0.0      # ce_dl_cnn.py:864-867
    desc.cnn_alpha = float(alpha or 0.0)
    desc.scs_hz = float(config.scs)
    cp = _host(config.CyclicPrefixDurations).astype(np.float64).reshape(-1)
    desc.n_cp = int(min(cp.size, _lib.MAX_SYMBOLS))
    for k in range(desc.n_cp):
        desc.cp_ms[k] = float(cp[k])
    h1, rm1 = _hop_desc(hop1, n_sym, n_prb_grid, "hop1")
    h2, rm2 = _hop_desc(hop2, n_sym, n_prb_grid, "hop2")
    if rm1 is None:
        raise ValueError("hop1 has no pilot symbol")
    if rm2 is not None:
        if h1.dmrs_symbol_mask & h2.dmrs_symbol_mask:
            raise AssertionError("Hops should not overlap.")                                    # :862
        assert rm1.shape == rm2.shape and np.array_equal(rm1, rm2), \
            "The DM-RS mask should be the same for the two hops."                               # :869
    if desc.cfo_compensate and n_sym != 14 and desc.n_cp >= 14 and any(
            bin(h.dmrs_symbol_mask).count("1") >= 2 for h in (h1, h2)):
        # the reference indexes its 14-entry phase table with the grid's symbol mask (:717) and broadcasts it over
        # the grid (:929): both fail with an IndexError-class shape mismatch unless the grid has 14 symbols
        raise IndexError(f"CFO compensation with two pilot symbols in a hop needs a 14-symbol grid (got {n_sym})")
    n_cdm = (n_layers + 1) // 2
    if rm1.shape[1] < n_cdm:
        raise IndexError(f"DMRSREmask has {rm1.shape[1]} columns, {n_cdm} CDM groups needed")
    for g in range(n_cdm):
        desc.re_mask[g] = int(sum(1 << k for k in range(NRE) if rm1[k, g]))
    desc.hop[0], desc.hop[1] = h1, h2
    key = (dev_index, bytes(desc))
    with _PLANS_LOCK:
        plan = _PLANS.get(key)
        if plan is None:
            if dev_index >= 0:
                with torch.cuda.device(dev_index):
                    plan = Plan(desc)
            else:
                plan = Plan(desc)
            _PLANS[key] = plan
    return plan


def _as_c64(x, device=None) -> torch.Tensor:
    t = torch.as_tensor(x, device=device)
    if t.dtype != torch.complex64:
        t = t.to(torch.complex64)                       # real inputs are promoted like the reference (:773-779)
    return t


def srs_channel_estimator_batched(received_rg, pilots, beta_dmrs, hop1, hop2, config, *, fill: str = "linear",
                                  pilot_share: Optional[int] = None, out=None, scalars=None):
    """Many independent reference calls in one launch.

    received_rg [B, n_sc, n_sym], pilots [U, n_re, n_dmrs_total, L] with B = U * pilot_share (the
    pilot_share consecutive units of a UE -- its rx antennas -- share one pilot block).
    Returns (channel_est_rg [B, n_sc, n_sym, L] complex64, scalars [B, 6] float64 with columns
    noise_est, rsrp, epre, time_alignment, cfo_hz, cfo_valid).  Unit b equals the reference called
    on (received_rg[b], pilots[b // pilot_share])."""
    rx = _as_c64(received_rg)
    pil = _as_c64(pilots, device=rx.device)
    if rx.dim() != 3 or pil.dim() != 4:
        raise IndexError("expected received_rg [B, n_sc, n_sym] and pilots [U, n_re, n_dmrs, L]")
    if rx.shape[1] % NRE:
        raise IndexError("n_sc must be a multiple of 12")
    if pilot_share is None:
        if pil.shape[0] == 0 or rx.shape[0] % pil.shape[0]:
            raise ValueError("B must be a multiple of the number of pilot blocks")
        pilot_share = rx.shape[0] // pil.shape[0]
    plan = make_plan(rx.shape[1] // NRE, rx.shape[2], pil.shape[3], hop1, hop2, config, fill,
                     device=rx.device if rx.is_cuda else None)
    return plan.estimate(rx, pil, float(beta_dmrs), pilot_share, out=out, scalars=scalars)


def srs_channel_estimator(received_rg, pilots, beta_dmrs, hop1, hop2, config, *, fill: str = "linear"):
    """Drop-in for the reference ``srs_channel_estimator`` (one received grid, one UE).

    Returns (channel_est_rg [n_sc, n_sym, L] complex64, noise_est, rsrp, epre, time_alignment
    float64 0-dim, cfo_hz float64 0-dim or an empty (0,) tensor when no hop has two pilot symbols),
    all on ``received_rg.device``.  The call itself never synchronises; the first call with a given
    (hop1, hop2, config) reads their mask tensors on the host once (see ``make_plan``)."""
    rx = _as_c64(received_rg)
    pil = _as_c64(pilots, device=rx.device)
    if rx.dim() != 2 or pil.dim() != 3:
        raise IndexError("expected received_rg [n_sc, n_sym] and pilots [n_re, n_dmrs, L]")
    if rx.shape[0] % NRE:
        raise IndexError("n_sc must be a multiple of 12")
    plan = make_plan(rx.shape[0] // NRE, rx.shape[1], pil.shape[2], hop1, hop2, config, fill,
                     device=rx.device if rx.is_cuda else None)
    grid, sc = plan.estimate(rx.unsqueeze(0), pil.unsqueeze(0), float(beta_dmrs), 1)
    cfo = sc[0, 4] if plan.info.cfo_defined else torch.empty((0,), dtype=torch.float64, device=rx.device)
    return grid[0], sc[0, 0], sc[0, 1], sc[0, 2], sc[0, 3], cfo


_ACCEPT_HOST_ARRAYS = os.environ.get("SRSCE_ACCEPT_HOST_ARRAYS", "0") not in ("", "0")


def accept_host_arrays(enable: bool = True) -> None:
    """Opt in (or out) of host arrays in the ``compat`` drop-in modules; also set by SRSCE_ACCEPT_HOST_ARRAYS=1."""
    global _ACCEPT_HOST_ARRAYS
    _ACCEPT_HOST_ARRAYS = bool(enable)


def srs_channel_estimator_any(received_rg, pilots, beta_dmrs, hop1, hop2, config, *, fill: str = "linear"):
    """The entry point of the ``compat`` modules.  CUDA tensors go straight to ``srs_channel_estimator``.  HOST
    arrays are refused by default (no CPU fallback -- raise on CPU tensors).  After ``accept_host_arrays()``
    (or with SRSCE_ACCEPT_HOST_ARRAYS=1) CPU tensors / numpy arrays -- what every caller of the reference holds
    (scripts/validation/validate_all.py:367-382, :543) -- are copied to the current CUDA device, estimated there, and
    the six results come back as CPU tensors, so the reference's scripts run unmodified on top of ``compat/``.  That
    is a transfer, not a fallback: without a CUDA device the call raises, there is no CPU compute path."""
    rx = torch.as_tensor(received_rg)
    if rx.is_cuda or not _ACCEPT_HOST_ARRAYS:
        return srs_channel_estimator(rx, pilots, beta_dmrs, hop1, hop2, config, fill=fill)
    if not torch.cuda.is_available():
        raise RuntimeError("srsran_ce_pytorch_b200 needs a CUDA device (sm_100a): host arrays are copied to it and back, "
                           "there is no CPU compute path")
    dev = torch.device("cuda", torch.cuda.current_device())
    out = srs_channel_estimator(_as_c64(rx).to(dev), _as_c64(torch.as_tensor(pilots)).to(dev), beta_dmrs, hop1, hop2,
                                config, fill=fill)
    return tuple(t.cpu() for t in out)


def unpack_scalars(scalars: torch.Tensor) -> dict:
    return {name: scalars[:, k] for k, name in enumerate(_SCALAR_FIELDS)}


def fill_ch_est_cdm(channel_in: torch.Tensor, estimated: torch.Tensor, hop, i_cdm: int, *, fill: str = "linear",
                    impl: str = "direct") -> torch.Tensor:
    """Drop-in for the reference's ``fill_ch_est_cdm(channel_in, estimated, hop, i_cdm)`` -- the plug-in seam
    ``fill_ch_est_cdm_fn`` of process_hop (ce_rule_tensorized.py:237-354 for ``fill="linear"``,
    ce_dl_cnn.py:233-322 for ``fill="cnn"``): returns a copy of ``channel_in`` [n_sc, n_sym, L] whose band rows /
    allocated symbols hold, for the layers of CDM group ``i_cdm`` (1-based, as in the reference), the pilots
    ``estimated`` [n_re, layers in the group] interpolated over the hop band.  The interpolation and the
    replication over the symbols run in ``srsce_fill``; the copy into ``channel_in`` is a torch slice assignment."""
    ch = torch.as_tensor(channel_in)
    if not ch.is_cuda:
        raise RuntimeError("srsran_ce_pytorch_b200 computes on CUDA tensors only (no CPU fallback)")
    if not torch.is_complex(ch):
        ch = ch.to(torch.complex64)
    est = _as_c64(estimated, device=ch.device)
    if ch.dim() != 3 or est.dim() != 2:
        raise IndexError("expected channel_in [n_sc, n_sym, L] and estimated [n_re, layers in the CDM group]")
    n_sc, n_sym, n_lg = ch.shape[0], ch.shape[1], est.shape[1]
    rm = _host(hop.DMRSREmask).astype(bool)
    if rm.ndim == 1:
        rm = rm.reshape(NRE, -1)
    col = rm[:, int(i_cdm) - 1]
    if not col.any() or int(hop.nPRBs) == 0:
        return ch.clone()                                               # ce_dl_cnn.py:289-291
    sub = type("Hop", (), {})()
    sub.DMRSsymbols, sub.maskPRBs = hop.DMRSsymbols, hop.maskPRBs
    sub.DMRSREmask = torch.as_tensor(col).reshape(NRE, 1)
    sub.PRBstart, sub.nPRBs = hop.PRBstart, hop.nPRBs
    sub.startSymbol, sub.nAllocatedSymbols = hop.startSymbol, hop.nAllocatedSymbols
    none = type("Hop", (), {})()
    none.DMRSsymbols = torch.zeros((0,), dtype=torch.bool)
    cfg = type("Cfg", (), {})()
    cfg.scs, cfg.CyclicPrefixDurations, cfg.Smoothing, cfg.CFOCompensate = 30e3, torch.zeros(14, dtype=torch.float64), "none", False
    plan = make_plan(n_sc // NRE, n_sym, n_lg, sub, none, cfg, fill, device=ch.device)
    grid = plan.fill(est.reshape(1, 1, est.shape[0], n_lg), impl=impl)[0]
    out = ch.clone()
    sc0, sc1 = NRE * int(hop.PRBstart), NRE * (int(hop.PRBstart) + int(hop.nPRBs))
    t0, t1 = int(hop.startSymbol), int(hop.startSymbol) + int(hop.nAllocatedSymbols)
    l0 = (int(i_cdm) - 1) * 2
    out[sc0:sc1, t0:t1, l0:l0 + n_lg] = grid[sc0:sc1, t0:t1, :].to(out.dtype)
    return out
