"""Generate golden vectors by running the UNMODIFIED reference in this container.

    python oracle/gen_golden.py            # writes tests/golden/*.npz

Imports /root/reference/src/{ce_rule_baseline,ce_rule_tensorized,ce_dl_cnn}.py (read-only, never
copied), feeds them seeded synthetic inputs from ``srsran_ce_pytorch_b200.synth`` and stores
inputs, geometry and all six outputs of every variant.  /root/reference does not exist on the GPU
box, so the fixtures are committed; this script is only the record of how they were made.
TEST INFRASTRUCTURE ONLY (see oracle/ce_oracle.py header).
"""
from __future__ import annotations

import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
REF_SRC = "/root/reference/src"

from srsran_ce_pytorch_b200 import synth  # noqa: E402

# name -> dict(grid PRBs, hop-1 (start, n, symbols), optional hop-2 (start, symbols), pattern,
#              layers, scs, smoothing, cfo flag, beta, start symbol / allocation, n_sym)
D = dict(grid=25, h1=(4, 6, [2, 11]), h2=None, pat="comb2", L=1, scs=30e3, sm="filter", cfo=True,
         beta=1.0, start=0, alloc=14, n_sym=14, vcp=False, snr=20.0, cfo_cyc=0.002, cnn=True, base=True)


def case(**kw):
    c = dict(D)
    c.update(kw)
    return c


CASES = {
    # the three geometries the reference's own single-case scripts hard-code (validate_case{0,4,8}.py)
    "case0_shape": case(grid=52, h1=(40, 3, [0, 4, 8, 12]), pat="comb2", L=1, scs=15e3, beta=1.4125, vcp=True),
    "case4_shape": case(grid=52, h1=(3, 3, [0, 4]), h2=(28, [8, 12]), pat="comb2", L=1, scs=15e3, beta=1.4125, vcp=True),
    "case8_shape": case(grid=52, h1=(1, 3, [2, 11]), pat="comb2", L=2, scs=15e3, beta=1.4125, vcp=True),
    # BASELINE.json config 2 and its comb-2 / multi-symbol sibling (full 273-PRB band)
    "cfg2_273prb_comb4_L4_S1": case(grid=273, h1=(0, 273, [13]), pat="comb4", L=4),
    "cfg2_273prb_comb2_L4_S2": case(grid=273, h1=(0, 273, [3, 10]), pat="comb2", L=4, cnn=False, base=False),
    # layers
    "comb2_L2": case(L=2), "comb2_L3": case(L=3), "comb2_L4": case(L=4),
    "comb4_L1": case(pat="comb4", L=1), "comb4_L3": case(pat="comb4", L=3), "comb4_L4_S4": case(pat="comb4", L=4, h1=(4, 6, [2, 5, 8, 11])),
    # masks
    "type2_L4": case(pat="type2", L=4), "comb3_L2": case(pat="comb3", L=2),
    "comb6_L2": case(pat="comb6", L=2), "comb12_L1": case(pat="comb12", L=1, h1=(4, 6, [2, 11])),
    "comb12_L2_4prb": case(pat="comb12", L=2, h1=(3, 4, [7])),
    "dense_L2": case(pat="dense", L=2, h1=(4, 4, [2, 11])),
    # hop widths (filter span / virtual pilot count change with 1, 2, >=3 PRBs)
    "comb2_1prb": case(h1=(7, 1, [2, 11]), L=2), "comb2_2prb": case(h1=(7, 2, [2, 11]), L=2),
    "comb4_1prb": case(pat="comb4", h1=(0, 1, [5]), L=2), "comb4_2prb": case(pat="comb4", h1=(23, 2, [5]), L=4),
    "comb6_1prb_L1": case(pat="comb6", h1=(5, 1, [3]), L=1),
    "comb12_2prb_L1": case(pat="comb12", h1=(5, 2, [3, 9]), L=1),
    "comb2_24prb_L4": case(h1=(0, 24, [2, 7, 11]), L=4),
    # smoothing modes and CFO switch
    "mean_L2": case(sm="mean", L=2), "none_L4": case(sm="none", L=4, pat="comb4"),
    "cfo_off_L2": case(cfo=False, L=2), "cfo_off_S1_L1": case(cfo=False, h1=(4, 6, [6])),
    "mean_cfo_off_type2_L3": case(sm="mean", cfo=False, pat="type2", L=3),
    # one pilot symbol (CFO output empty), partial allocations
    "S1_L1": case(h1=(4, 6, [2])), "S1_alloc_10_4": case(h1=(4, 6, [11]), start=10, alloc=4, L=2),
    "S2_alloc_2_9": case(h1=(2, 9, [3, 8]), start=2, alloc=9, L=2),
    # pilot-symbol-only grids (n_sym == 1) and CFO-off short grid
    "nsym1_L4": case(n_sym=1, h1=(0, 25, [0]), pat="comb4", L=4, start=0, alloc=1),
    "nsym2_cfo_off": case(n_sym=2, h1=(3, 5, [0, 1]), L=2, cfo=False, start=0, alloc=2),
    # two hops
    "2hop_L2_comb2": case(h1=(1, 5, [2]), h2=(17, [11]), L=2),
    "2hop_L4_comb4_S2": case(h1=(0, 8, [1, 3]), h2=(12, [8, 10]), pat="comb4", L=4),
    "2hop_overlap_prbs": case(h1=(4, 6, [2]), h2=(7, [11]), L=1),
    "2hop_mean_L3": case(h1=(2, 4, [2, 4]), h2=(16, [9, 11]), L=3, sm="mean"),
    "2hop_cfo_off": case(h1=(1, 5, [2, 5]), h2=(17, [9, 12]), L=2, cfo=False),
    # SNR extremes (noise estimate conditioning) and larger CFO
    "snr0_L2": case(L=2, snr=0.0), "snr40_L2": case(L=2, snr=40.0, cfo_cyc=0.0005),
    "big_cfo_L2": case(L=2, cfo_cyc=0.02),
    # irregular RE masks (arbitrary 12-bit columns, R not dividing 12, odd pilot counts): the masks the fuzz test draws
    "irr_R5_L2": case(pat=["110010010100", "001101001010"], L=2, h1=(3, 5, [2, 9])),
    "irr_R7_L3_2hop": case(pat=["111010100110", "010111010011"], L=3, h1=(1, 3, [4]), h2=(14, [10]), sm="mean"),
    "irr_R1_off5_L1": case(pat=["000001000000"], L=1, h1=(6, 11, [3, 4])),
    "irr_R4_L4_none": case(pat=["100100011000", "011000100100"], L=4, h1=(0, 2, [1, 6, 12]), sm="none"),
    "irr_R2_1prb_L2": case(pat=["010000000010", "000100100000"], L=2, h1=(9, 1, [7])),
    "irr_R3_L4_alloc": case(pat=["100001000100", "010000100010"], L=4, h1=(2, 7, [5, 8]), start=2, alloc=9, cfo=False),
    # CNN blend switch (ce_dl_cnn CNNSmoothingAlpha)
    "cnn_alpha_0p5_L2": case(L=2, cnn_alpha=0.5), "cnn_alpha_1_comb4_L4": case(L=4, pat="comb4", cnn_alpha=1.0),
}


def build(name: str, c: dict, seed: int):
    start, n, syms1 = c["h1"]
    hop1 = synth.make_hop(c["grid"], start, n, syms1, c["pat"], (c["L"] + 1) // 2, c["n_sym"],
                          c["start"], c["alloc"])
    if c["h2"] is None:
        hop2 = synth.no_hop()
    else:
        hop2 = synth.make_hop(c["grid"], c["h2"][0], n, c["h2"][1], c["pat"], (c["L"] + 1) // 2,
                              c["n_sym"], c["start"], c["alloc"])
    cfg = synth.make_config(c["scs"], c["sm"], c["cfo"], cnn_alpha=c.get("cnn_alpha"), validator_cp=c["vcp"])
    data = synth.make_units(hop1, hop2, cfg, c["L"], 1, seed=seed, snr_db=c["snr"],
                            cfo_cycles=c["cfo_cyc"], beta=c["beta"])
    return hop1, hop2, cfg, data["rx"][0], data["pilots"][0], c["beta"]


def pack_outputs(tag, out, store):
    grid, noise, rsrp, epre, ta, cfo = out
    store[f"{tag}_grid"] = grid.numpy()
    store[f"{tag}_scalars"] = np.array([noise.item(), rsrp.item(), epre.item(), ta.item(),
                                        cfo.item() if cfo.numel() else np.nan], dtype=np.float64)
    assert grid.dtype == torch.complex64 and noise.dtype == torch.float64


def main():
    only_new = "--only-new" in sys.argv          # keep the committed fixtures byte for byte, add the missing ones
    sys.path.insert(0, REF_SRC)
    import ce_dl_cnn          # noqa
    import ce_rule_baseline   # noqa
    import ce_rule_tensorized # noqa
    mods = {"tensorized": ce_rule_tensorized, "baseline": ce_rule_baseline, "cnn": ce_dl_cnn}
    out_dir = ROOT / "tests" / "golden"
    out_dir.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(1)
    total = 0
    for k, (name, c) in enumerate(CASES.items()):
        if only_new and (out_dir / f"{name}.npz").exists():
            continue
        hop1, hop2, cfg, rx, pilots, beta = build(name, c, seed=1000 + sum(map(ord, name)) if name.startswith("irr_") else 1000 + k)
        store = {
            "rx": rx.numpy(), "pilots": pilots.numpy(), "beta": np.float64(beta),
            "n_layers": np.int64(c["L"]), "scs": np.float64(c["scs"]),
            "cp_ms": cfg.CyclicPrefixDurations.numpy(), "smoothing": np.str_(c["sm"]),
            "cfo_compensate": np.bool_(c["cfo"]),
            "cnn_alpha": np.float64(c.get("cnn_alpha") or 0.0),
            "re_mask": hop1.DMRSREmask.numpy(),
        }
        for i, h in enumerate((hop1, hop2), start=1):
            store[f"hop{i}_dmrs_symbols"] = torch.as_tensor(h.DMRSsymbols).numpy()
            store[f"hop{i}_mask_prbs"] = torch.as_tensor(h.maskPRBs).numpy()
            store[f"hop{i}_geom"] = np.array([h.PRBstart, h.nPRBs, h.startSymbol, h.nAllocatedSymbols], dtype=np.int64)
        with torch.no_grad():
            for tag, mod in mods.items():
                if (tag == "cnn" and not c["cnn"]) or (tag == "baseline" and not c["base"]):
                    continue
                mh = [mod.HopConfig(**vars(h)) for h in (hop1, hop2)]
                if c.get("cnn_alpha") is not None:
                    mcfg = cfg                                   # carries CNNSmoothingAlpha
                else:
                    mcfg = mod.EstimatorConfig(scs=cfg.scs, CyclicPrefixDurations=cfg.CyclicPrefixDurations,
                                               Smoothing=cfg.Smoothing, CFOCompensate=cfg.CFOCompensate)
                pack_outputs(tag, mod.srs_channel_estimator(rx, pilots, beta, mh[0], mh[1], mcfg), store)
        path = out_dir / f"{name}.npz"
        np.savez_compressed(path, **store)
        total += os.path.getsize(path)
        print(f"{name:28s} {os.path.getsize(path) / 1024:8.1f} KiB")
    print(f"{len(CASES)} cases, {total / 2**20:.2f} MiB")


if __name__ == "__main__":
    main()
