"""Loader for tests/golden/*.npz (made by oracle/gen_golden.py from the unmodified reference)."""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch

from srsran_ce_pytorch_b200.types import CnnEstimatorConfig, EstimatorConfig, HopConfig

GOLDEN_DIR = Path(__file__).resolve().parent / "golden"


def case_names():
    return sorted(p.stem for p in GOLDEN_DIR.glob("*.npz"))


def load_case(path_or_name):
    p = Path(path_or_name)
    if not p.suffix:
        p = GOLDEN_DIR / f"{p.name}.npz"
    z = np.load(p)
    hops = []
    for i in (1, 2):
        ps, n, s0, na = (int(v) for v in z[f"hop{i}_geom"])
        sym = z[f"hop{i}_dmrs_symbols"]
        rm = z["re_mask"] if sym.size else np.zeros((12, 0), dtype=bool)
        hops.append(HopConfig(DMRSsymbols=torch.from_numpy(sym), DMRSREmask=torch.from_numpy(rm),
                              PRBstart=ps, nPRBs=n, maskPRBs=torch.from_numpy(z[f"hop{i}_mask_prbs"]),
                              startSymbol=s0, nAllocatedSymbols=na))
    kw = dict(scs=float(z["scs"]), CyclicPrefixDurations=torch.from_numpy(z["cp_ms"]),
              Smoothing=str(z["smoothing"]), CFOCompensate=bool(z["cfo_compensate"]))
    alpha = float(z["cnn_alpha"])
    cfg = CnnEstimatorConfig(CNNSmoothingAlpha=alpha, **kw) if alpha else EstimatorConfig(**kw)
    ref = {k: z[k] for k in z.files if k.endswith("_grid") or k.endswith("_scalars")}
    return dict(name=p.stem, rx=z["rx"], pilots=z["pilots"], beta=float(z["beta"]),
                n_layers=int(z["n_layers"]), hop1=hops[0], hop2=hops[1], cfg=cfg, ref=ref)
