"""The oracle's band fill against outputs of the unmodified reference ``fill_ch_est_cdm`` functions
(tests/golden_fill, made by oracle/gen_golden_fill.py): pins the checker of the srsce_fill GPU tests."""
from pathlib import Path

import numpy as np
import pytest

from oracle import ce_oracle

FILL_DIR = Path(__file__).resolve().parent / "golden_fill"
NAMES = sorted(p.stem for p in FILL_DIR.glob("*.npz"))


def load_fill_case(name):
    z = np.load(FILL_DIR / f"{name}.npz")
    ps, n, s0, na = (int(v) for v in z["geom"])
    hop = ce_oracle.HopSpec(z["dmrs_symbols"], z["re_mask"], ps, n, z["mask_prbs"], s0, na)
    return z, hop


def oracle_fill(z, hop, variant):
    """fill_ch_est_cdm restated with the oracle's band fill: copy of channel_in with the band rows / allocated symbols
    of the group's layers overwritten."""
    g = int(z["i_cdm"]) - 1
    band = ce_oracle._fill_band(z["estimated"], hop, g, variant)
    out = z["channel_in"].copy()
    sc0 = 12 * hop.prb_start
    for l in range(band.shape[1]):
        out[sc0:sc0 + band.shape[0], hop.start_symbol:hop.start_symbol + hop.n_alloc_symbols, 2 * g + l] = band[:, l:l + 1]
    return out


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("variant", ["tensorized", "baseline", "cnn"])
def test_oracle_band_fill_matches_reference(name, variant):
    z, hop = load_fill_case(name)
    want = z[f"out_{variant}"]
    got = oracle_fill(z, hop, variant)
    scale = np.abs(want).max()
    assert np.abs(got - want).max() <= 2e-7 * scale, (name, variant, np.abs(got - want).max() / scale)
