"""Golden vectors of the fill seam: the UNMODIFIED reference functions ``fill_ch_est_cdm`` of
/root/reference/src/{ce_rule_tensorized,ce_rule_baseline,ce_dl_cnn}.py run in this container on seeded inputs.

    python oracle/gen_golden_fill.py            # writes tests/golden_fill/*.npz

/root/reference does not exist on the GPU box, so the fixtures are committed; this script is the record of how they
were made.  TEST INFRASTRUCTURE ONLY (see oracle/ce_oracle.py header)."""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
REF_SRC = "/root/reference/src"

from srsran_ce_pytorch_b200 import synth  # noqa: E402

# name: grid PRBs, PRB start, PRBs, pattern, CDM group (1-based), layers in the group, total layers, start symbol,
#       allocated symbols, grid symbols, random channel_in?
CASES = {
    "comb4_g1_L2of4": (25, 3, 9, "comb4", 1, 2, 4, 0, 14, 14, False),
    "comb4_g2_L2of4_alloc": (25, 0, 25, "comb4", 2, 2, 4, 2, 9, 14, True),
    "comb2_g1_L1of1": (25, 5, 17, "comb2", 1, 1, 1, 0, 14, 14, False),
    "comb2_g2_L1of3": (25, 0, 8, "comb2", 2, 1, 3, 0, 14, 14, True),
    "type2_g1_L2of2": (25, 2, 20, "type2", 1, 2, 2, 0, 14, 14, False),
    "comb3_g2_L2of4": (25, 16, 9, "comb3", 2, 2, 4, 0, 14, 14, False),
    "comb6_g1_L2of2": (25, 1, 24, "comb6", 1, 2, 2, 5, 9, 14, False),
    "comb12_g1_L1of1": (25, 4, 10, "comb12", 1, 1, 1, 0, 14, 14, False),
    "dense_g1_L2of2": (25, 6, 11, "dense", 1, 2, 2, 0, 14, 14, False),
    "comb4_1prb": (25, 12, 1, "comb4", 1, 2, 2, 0, 14, 14, False),
    "comb2_2prb": (25, 12, 2, "comb2", 2, 2, 4, 0, 14, 14, False),
    "irr_R5": (25, 3, 10, ["110010010100", "001101001010"], 2, 2, 4, 0, 14, 14, False),
    "comb4_96prb": (106, 5, 96, "comb4", 1, 2, 2, 0, 14, 14, False),
}


def main():
    sys.path.insert(0, REF_SRC)
    import ce_dl_cnn          # noqa
    import ce_rule_baseline   # noqa
    import ce_rule_tensorized # noqa
    out_dir = ROOT / "tests" / "golden_fill"
    out_dir.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(1)
    for k, (name, c) in enumerate(CASES.items()):
        grid, start, n, pat, i_cdm, n_lg, n_l, s0, n_alloc, n_sym, rnd = c
        g = torch.Generator().manual_seed(4000 + k)
        hop = synth.make_hop(grid, start, n, [min(2, n_sym - 1)], pat, 2, n_sym, s0, n_alloc)
        n_re = n * int(hop.DMRSREmask[:, i_cdm - 1].sum())
        est = torch.complex(torch.randn(n_re, n_lg, generator=g), torch.randn(n_re, n_lg, generator=g)).to(torch.complex64)
        if rnd:
            ch = torch.complex(torch.randn(12 * grid, n_sym, n_l, generator=g), torch.randn(12 * grid, n_sym, n_l, generator=g)).to(torch.complex64)
        else:
            ch = torch.zeros((12 * grid, n_sym, n_l), dtype=torch.complex64)
        store = {"channel_in": ch.numpy(), "estimated": est.numpy(), "i_cdm": np.int64(i_cdm),
                 "re_mask": hop.DMRSREmask.numpy(), "dmrs_symbols": hop.DMRSsymbols.numpy(), "mask_prbs": hop.maskPRBs.numpy(),
                 "geom": np.array([hop.PRBstart, hop.nPRBs, hop.startSymbol, hop.nAllocatedSymbols], dtype=np.int64)}
        with torch.no_grad():
            for tag, mod in (("tensorized", ce_rule_tensorized), ("baseline", ce_rule_baseline), ("cnn", ce_dl_cnn)):
                out = mod.fill_ch_est_cdm(ch, est, mod.HopConfig(**vars(hop)), i_cdm)
                assert out.dtype == torch.complex64 and out.shape == ch.shape
                store[f"out_{tag}"] = out.numpy()
        np.savez_compressed(out_dir / f"{name}.npz", **store)
        print(f"{name:24s} {(out_dir / f'{name}.npz').stat().st_size / 1024:8.1f} KiB")


if __name__ == "__main__":
    main()
