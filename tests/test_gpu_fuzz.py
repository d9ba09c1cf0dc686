"""Randomised differential test: arbitrary 12-bit RE masks (irregular combs, R not dividing 12), odd pilot
counts, 1-PRB hops, every smoothing / CFO / fill combination -- CUDA path vs the oracle."""
import os
import random

import numpy as np
import pytest
import torch

from oracle import ce_oracle
from tests import tolerances as tol

pytestmark = pytest.mark.gpu


def _cases(n=48, seed=777, big=False):
    """``big``: full-size grids (273 / 341 PRB) with hops of 24 ... 136 PRB instead of the small ones (one-off hunts)."""
    rnd = random.Random(seed)
    out = []
    while len(out) < n:
        r = rnd.choice([1, 2, 3, 4, 5, 6, 7, 12])
        layers = rnd.choice([1, 2, 3, 4])
        n_cdm = (layers + 1) // 2
        cols = []
        for _ in range(n_cdm):                                   # every CDM group selects r REs of the PRB
            pos = sorted(rnd.sample(range(12), r))
            cols.append("".join("1" if k in pos else "0" for k in range(12)))
        grid = rnd.choice([273, 341] if big else [25, 52, 106])
        n_prb = rnd.choice([24, 48, 96, 136] if big else [1, 1, 2, 3, 5, 11, 24])
        n_prb = min(n_prb, grid // 2 - 1)
        two = rnd.random() < 0.35
        s_tot = rnd.choice([1, 2, 3, 4])
        syms = sorted(rnd.sample(range(14), s_tot + (1 if two else 0)))
        s1, s2 = (syms[:-1], syms[-1:]) if two else (syms, None)
        if two and not s1:
            continue
        start = rnd.randrange(0, grid - 2 * n_prb)
        start2 = rnd.randrange(0, grid - n_prb) if rnd.random() < 0.5 else start + n_prb   # may overlap hop 1's band
        alloc0 = rnd.choice([0, 0, 2])
        alloc_n = 14 - alloc0 - rnd.choice([0, 0, 3])
        out.append(dict(cols=cols, layers=layers, grid=grid, n_prb=n_prb, s1=s1, s2=s2, start=start, start2=start2,
                        sm=rnd.choice(["filter", "filter", "mean", "none"]), cfo=rnd.random() < 0.6,
                        alloc=(alloc0, alloc_n), beta=rnd.choice([1.0, 1.4125, 0.5]), snr=rnd.choice([5.0, 20.0, 35.0])))
    return out


# FUZZ_N / FUZZ_SEED widen the campaign for one-off hunts (the default set is what CI runs)
CASES = _cases(int(os.environ.get("FUZZ_N", "48")), int(os.environ.get("FUZZ_SEED", "777")),
               os.environ.get("FUZZ_BIG", "0") not in ("", "0"))


@pytest.mark.parametrize("k", range(len(CASES)))
def test_random_geometry_matches_oracle(k):
    from srsran_ce_pytorch_b200 import estimator, synth
    c = CASES[k]
    n_cdm = (c["layers"] + 1) // 2
    a0, an = c["alloc"]
    hop1 = synth.make_hop(c["grid"], c["start"], c["n_prb"], c["s1"], c["cols"], n_cdm, start_symbol=a0, n_alloc=an)
    hop2 = synth.make_hop(c["grid"], c["start2"], c["n_prb"], c["s2"], c["cols"], n_cdm, start_symbol=a0, n_alloc=an) \
        if c["s2"] else synth.no_hop()
    cfg = synth.make_config(30e3, c["sm"], c["cfo"])
    data = synth.make_units(hop1, hop2, cfg, c["layers"], 3, pilot_share=1, seed=500 + k, device="cuda", beta=c["beta"],
                            snr_db=c["snr"])
    rx, pil = data["rx"].cpu().numpy(), data["pilots"].cpu().numpy()
    for fill in ("linear", "cnn"):
        out, sc = estimator.srs_channel_estimator_batched(data["rx"], data["pilots"], c["beta"], hop1, hop2, cfg,
                                                          fill=fill, pilot_share=1)
        out, sc = out.cpu().numpy(), sc.cpu().numpy()
        for b in range(3):
            diag = {}
            want = ce_oracle.estimate(rx[b], pil[b], c["beta"], hop1, hop2, cfg,
                                      fill="cnn" if fill == "cnn" else "tensorized", diagnostics=diag)
            if not np.isfinite(want[1]):
                continue                                   # a single pilot in total: the reference divides by zero
            mx, nmse = tol.grid_errors(out[b], want[0])
            gate = tol.grid_nmse_gate((c["s1"], c["s2"]), c["cfo"], c["sm"], c["n_prb"] * c["cols"][0].count("1"),
                                      cfo_kappa=diag.get("cfo_kappa"), n_cdm=n_cdm)
            mx_gate = tol.grid_max_rel_gate((c["s1"], c["s2"]), c["cfo"], diag.get("cfo_kappa"), n_cdm)
            assert mx <= mx_gate and nmse <= gate, (c, fill, b, mx, nmse, mx_gate, gate, diag.get("cfo_kappa"))
            got = [sc[b, 0], sc[b, 1], sc[b, 2], sc[b, 3], sc[b, 4] if sc[b, 5] else None]
            tol.check_scalars(got, want[1:], 30e3, diag["ta_margin"], tag=f"fuzz{k}/{fill}/{b}")
