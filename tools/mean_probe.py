"""Mean smoothing at large pilot counts: kernel vs oracle NMSE on the synthetic 100-sample delay spread (the band mean
cancels to ~1 % of its terms, so float32 summation order shows; see tests/tolerances.py).  Worst NMSE over 24 units per geometry."""
import sys; from pathlib import Path; sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from oracle import ce_oracle
from srsran_ce_pytorch_b200 import estimator, synth
from tests import tolerances as tol
for two in (False, True):
    for grid in (106, 273, 341):
        hop1 = synth.make_hop(grid, 0, grid, [3], "dense", 1, start_symbol=0, n_alloc=7 if two else 14)
        hop2 = synth.make_hop(grid, 0, grid, [10], "dense", 1, start_symbol=7, n_alloc=7) if two else synth.no_hop()
        cfg = synth.make_config(30e3, "mean", True)
        plan = estimator.make_plan(grid, 14, 2, hop1, hop2, cfg)
        data = synth.make_units(hop1, hop2, cfg, 2, 24, seed=5, device="cuda")
        out, sc = plan.estimate(data["rx"], data["pilots"], 1.0, 1)
        rx, pil = data["rx"].cpu().numpy(), data["pilots"].cpu().numpy()
        worst = 0
        for b in range(24):
            want = ce_oracle.estimate(rx[b], pil[b], 1.0, hop1, hop2, cfg)
            mx, nmse = tol.grid_errors(out[b].cpu().numpy(), want[0])
            worst = max(worst, nmse)
        print("two_hops", two, "grid", grid, "h_in_global", plan.info.h_in_global, "worst NMSE over 24 units", worst, flush=True)
