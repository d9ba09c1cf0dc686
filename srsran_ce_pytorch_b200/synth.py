"""Synthetic uplink pilot workloads (harness side; no estimator arithmetic here).

Builds reference-style ``HopConfig`` / ``EstimatorConfig`` objects and seeded received grids +
pilot tensors for the configurations listed in BASELINE.json / SURVEY.md section 8(d):
multi-tap channel per (unit, layer), frequency-domain OCC on odd layers, a common CFO phase ramp
over OFDM symbols and AWGN.  Works on CPU and CUDA tensors (torch only; used by the tests, the
golden-vector generator and bench.py).
"""
from __future__ import annotations

import math
from typing import Optional, Sequence

import torch

from .types import EstimatorConfig, HopConfig, CnnEstimatorConfig

NRE = 12

# 12-bit pilot RE patterns of one PRB, one entry per CDM group
RE_PATTERNS = {
    "comb2": ["101010101010", "010101010101"],          # DM-RS type 1
    "comb4": ["100010001000", "001000100010"],          # SRS-like comb-4, offsets 0 and 2
    "type2": ["110000110000", "001100001100"],          # DM-RS type 2 (third group 000011000011)
    "comb3": ["100100100100", "010010010010"],
    "comb6": ["100000100000", "000100000100"],
    "comb12": ["100000000000", "000000100000"],
    "dense": ["111111111111", "111111111111"],
}


def re_mask(pattern: str | Sequence[str], n_cdm: int = 2) -> torch.Tensor:
    cols = RE_PATTERNS[pattern] if isinstance(pattern, str) else list(pattern)
    cols = cols[:max(1, n_cdm)]
    return torch.tensor([[c == "1" for c in col] for col in cols], dtype=torch.bool).t().contiguous()


def cp_durations_ms(scs_hz: float, n_syms: int = 14, validator_style: bool = False) -> torch.Tensor:
    """Normal-CP durations in ms.  ``validator_style`` reproduces the reference validator's
    choice (/root/reference scripts/validation/validate_all.py:269-283: CP samples scaled by
    15 kHz/scs at a 2048-point symbol); otherwise the physical 160/144 of 2048."""
    if validator_style:
        scale = 15000.0 / scs_hz
        first, rest = int(round(160 * scale)), int(round(144 * scale))
    else:
        first, rest = 160, 144
    ts = 1.0 / (scs_hz * 2048)
    cp = torch.tensor([first] + [rest] * (n_syms - 1), dtype=torch.float64)
    return cp * ts * 1000.0


def make_hop(n_prb_grid: int, prb_start: int, n_prbs: int, dmrs_symbols: Sequence[int],
             pattern: str | Sequence[str] = "comb2", n_cdm: int = 2, n_sym: int = 14,
             start_symbol: int = 0, n_alloc: Optional[int] = None) -> HopConfig:
    sym = torch.zeros(n_sym, dtype=torch.bool)
    sym[list(dmrs_symbols)] = True
    prbs = torch.zeros(n_prb_grid, dtype=torch.bool)
    prbs[prb_start:prb_start + n_prbs] = True
    return HopConfig(DMRSsymbols=sym, DMRSREmask=re_mask(pattern, n_cdm), PRBstart=prb_start,
                     nPRBs=n_prbs, maskPRBs=prbs, startSymbol=start_symbol,
                     nAllocatedSymbols=n_sym - start_symbol if n_alloc is None else n_alloc)


def no_hop() -> HopConfig:
    """The reference's "hop 2 absent" sentinel (validate_all.py:449-457)."""
    return HopConfig(DMRSsymbols=torch.zeros(0, dtype=torch.bool),
                     DMRSREmask=torch.zeros((12, 0), dtype=torch.bool), PRBstart=0, nPRBs=0,
                     maskPRBs=torch.zeros(0, dtype=torch.bool), startSymbol=0, nAllocatedSymbols=0)


def make_config(scs_hz: float = 30e3, smoothing: str = "filter", cfo_compensate: bool = True,
                cnn_alpha: Optional[float] = None, validator_cp: bool = False) -> EstimatorConfig:
    cp = cp_durations_ms(scs_hz, validator_style=validator_cp)
    if cnn_alpha is None:
        return EstimatorConfig(scs=scs_hz, CyclicPrefixDurations=cp, Smoothing=smoothing,
                               CFOCompensate=cfo_compensate)
    return CnnEstimatorConfig(scs=scs_hz, CyclicPrefixDurations=cp, Smoothing=smoothing,
                              CFOCompensate=cfo_compensate, CNNSmoothingAlpha=cnn_alpha)


def _pilot_subcarriers(hop: HopConfig, grp: int) -> torch.Tensor:
    col = torch.as_tensor(hop.DMRSREmask)[:, grp].bool()
    prbs = torch.as_tensor(hop.maskPRBs).bool()
    return torch.nonzero(col.repeat(prbs.numel()) & prbs.repeat_interleave(NRE)).flatten()


def symbol_times(cfg: EstimatorConfig) -> torch.Tensor:
    cpd = torch.as_tensor(cfg.CyclicPrefixDurations, dtype=torch.float64) * float(cfg.scs) / 1000.0
    return torch.cumsum(torch.cat([cpd[:1], cpd[1:14] + 1.0]), 0)


def make_units(hop1: HopConfig, hop2: HopConfig, cfg: EstimatorConfig, n_layers: int,
               n_units: int, pilot_share: int = 1, seed: int = 0, device="cpu",
               snr_db: float = 20.0, cfo_cycles: float = 0.002, beta: float = 1.0,
               n_taps: int = 4, max_delay: float = 100.0, symbol_major: bool = False):
    """Returns dict(rx [B, n_sc, n_sym] complex64, pilots [B // pilot_share, n_re, S_total, L]
    complex64, beta).  Every unit has its own received grid; ``pilot_share`` consecutive units
    (the rx antennas of one UE) share one pilot tensor.  With ``symbol_major`` the grid memory is
    laid out [B, n_sym, n_sc] and returned as a transposed view (same logical shape)."""
    assert n_units % pilot_share == 0
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(seed))
    n_sc = torch.as_tensor(hop1.maskPRBs).numel() * NRE
    n_sym = torch.as_tensor(hop1.DMRSsymbols).numel()
    n_cdm = (n_layers + 1) // 2
    hops = [hop1] + ([hop2] if torch.as_tensor(hop2.DMRSsymbols).numel() and
                     bool(torch.as_tensor(hop2.DMRSsymbols).any()) else [])
    n_re = int(torch.as_tensor(hop1.DMRSREmask)[:, 0].sum()) * int(hop1.nPRBs)
    s_tot = sum(int(torch.as_tensor(h.DMRSsymbols).sum()) for h in hops)
    n_ue = n_units // pilot_share

    def crandn(*shape):
        re = torch.randn(*shape, generator=gen, device=dev, dtype=torch.float32)
        im = torch.randn(*shape, generator=gen, device=dev, dtype=torch.float32)
        return torch.complex(re, im) * (0.5 ** 0.5)

    # pilots: unit modulus, random phase per (UE, RE, symbol, CDM group); odd layers carry the
    # frequency-domain OCC [+1, -1, +1, ...]
    ph = torch.rand(n_ue, n_re, s_tot, n_cdm, generator=gen, device=dev) * (2 * math.pi)
    base = torch.polar(torch.ones_like(ph), ph)
    occ = (1.0 - 2.0 * (torch.arange(n_re, device=dev) % 2)).to(torch.float32)
    pilots = torch.empty(n_ue, n_re, s_tot, n_layers, dtype=torch.complex64, device=dev)
    for l in range(n_layers):
        pilots[..., l] = base[..., l // 2] * (occ[None, :, None] if l % 2 else 1.0)

    # channel: n_taps paths per (unit, layer)
    gains = crandn(n_units, n_layers, n_taps) * torch.exp(
        -0.5 * torch.arange(n_taps, device=dev, dtype=torch.float32))[None, None, :]
    delays = torch.rand(n_units, n_layers, n_taps, generator=gen, device=dev) * max_delay
    if symbol_major:
        rx = crandn(n_units, n_sym, n_sc).transpose(1, 2)
    else:
        rx = crandn(n_units, n_sc, n_sym)
    noise_amp = 10.0 ** (-snr_db / 20.0)
    t_sym = symbol_times(cfg).to(dev) if n_sym == 14 else torch.zeros(n_sym, dtype=torch.float64, device=dev)
    col0 = 0
    for hop in hops:
        syms = torch.nonzero(torch.as_tensor(hop.DMRSsymbols).bool()).flatten().to(dev)
        rot = torch.exp(1j * (2 * math.pi * cfo_cycles * t_sym[syms])).to(torch.complex64)
        for g in range(n_cdm):
            sc = _pilot_subcarriers(hop, g).to(dev)
            val = torch.zeros(n_units, n_re, syms.numel(), dtype=torch.complex64, device=dev)
            for l in range(2 * g, min(n_layers, 2 * g + 2)):
                arg = (-2 * math.pi / 4096.0) * sc.to(torch.float32)[None, :, None] * delays[:, l, None, :]
                hf = (gains[:, l, None, :] * torch.polar(torch.ones_like(arg), arg)).sum(-1)
                p = pilots[:, :, col0:col0 + syms.numel(), l].repeat_interleave(pilot_share, dim=0)
                val += beta * hf[:, :, None] * p
            val = val * rot[None, None, :] + noise_amp * crandn(n_units, n_re, syms.numel())
            rx[:, sc[:, None], syms[None, :]] = val
        col0 += syms.numel()
    return {"rx": rx, "pilots": pilots, "beta": float(beta)}
