"""Configuration containers of the estimator API.

Field names and meanings are the reference's (MATLAB-cased; /root/reference
src/ce_rule_tensorized.py:13-29), because callers construct them by keyword.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any


@dataclass
class HopConfig:
    """Geometry of one frequency hop.

    DMRSsymbols        bool [n_sym]        pilot-carrying OFDM symbols of this hop
                                           (zero-length => the hop is absent)
    DMRSREmask         bool [12, n_cdm]    pilot REs inside a PRB, one column per CDM group
    PRBstart / nPRBs   int                 contiguous PRB range of the hop (0-based)
    maskPRBs           bool [n_prb_grid]   same range as a mask over the whole grid
    startSymbol / nAllocatedSymbols  int   symbols the estimate is replicated over
    """
    DMRSsymbols: Any
    DMRSREmask: Any
    PRBstart: int
    nPRBs: int
    maskPRBs: Any
    startSymbol: int
    nAllocatedSymbols: int


@dataclass
class EstimatorConfig:
    """scs in Hz; CyclicPrefixDurations in ms (>= 14 entries when CFO compensation is on);
    Smoothing in {"filter", "mean", "none"}."""
    scs: float
    CyclicPrefixDurations: Any
    Smoothing: str = "filter"
    CFOCompensate: bool = True


@dataclass
class CnnEstimatorConfig(EstimatorConfig):
    """EstimatorConfig plus the optional blend factor ce_dl_cnn reads with hasattr()
    (/root/reference src/ce_dl_cnn.py:864-867)."""
    CNNSmoothingAlpha: float = 0.0
