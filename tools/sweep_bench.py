"""Throughput sweep for BASELINE configs 4 and 5 (device-timed, inputs resident): narrow-band / hopping /
comb / layer / smoothing / pilot-symbol points plus the ce_dl_cnn fill at 273 PRB.  Writes one JSON line per
point to stdout and a table to profiles/ when --out is given.  Not the bench contract (bench.py is)."""
import argparse
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

from srsran_ce_pytorch_b200 import estimator, synth

PEAK = 6416.7
if (Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json").exists():
    PEAK = float(json.loads((Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])

POINTS = [
    # name, grid PRBs, hop PRBs, hop-1 symbols, hop-2 symbols, pattern, layers, smoothing, fill, rx antennas
    ("cfg2 S=1", 273, 273, [13], None, "comb4", 4, "filter", "linear", 4),
    ("cfg2 S=2 (CFO)", 273, 273, [3, 10], None, "comb4", 4, "filter", "linear", 4),
    ("cfg2 S=4 (CFO)", 273, 273, [2, 5, 8, 11], None, "comb4", 4, "filter", "linear", 4),
    ("cfg4 ce_dl_cnn fill", 273, 273, [13], None, "comb4", 4, "filter", "cnn", 4),
    ("cfg4 ce_dl_cnn comb-2", 273, 273, [13], None, "comb2", 4, "filter", "cnn", 4),
    ("comb-2 S=1", 273, 273, [13], None, "comb2", 4, "filter", "linear", 4),
    ("comb-2 S=2", 273, 273, [3, 10], None, "comb2", 4, "filter", "linear", 4),
    ("type-2 mask", 273, 273, [13], None, "type2", 4, "filter", "linear", 4),
    ("mean smoothing", 273, 273, [13], None, "comb4", 4, "mean", "linear", 4),
    ("no smoothing", 273, 273, [13], None, "comb4", 4, "none", "linear", 4),
    ("2 ports", 273, 273, [13], None, "comb4", 2, "filter", "linear", 4),
    ("1 port", 273, 273, [13], None, "comb4", 1, "filter", "linear", 4),
    ("3 ports", 273, 273, [13], None, "comb4", 3, "filter", "linear", 4),
] + [(f"hop {n} PRB, 1 hop, S=1", 273, n, [6], None, "comb4", 4, "filter", "linear", 4) for n in (4, 8, 16, 24, 48, 96, 136, 272)] + [
    ("hop 24 PRB, 2 hops, S=1+1", 273, 24, [3], [10], "comb4", 4, "filter", "linear", 4),
    ("hop 48 PRB, 2 hops, S=2+2 (CFO)", 273, 48, [2, 4], [9, 11], "comb2", 2, "filter", "linear", 4),
    ("hop 96 PRB, 2 hops, S=1+1, comb-2", 273, 96, [3], [10], "comb2", 4, "filter", "linear", 4),
    ("hop 136 PRB, 2 hops, S=2+2", 273, 136, [2, 4], [9, 11], "comb4", 4, "filter", "linear", 4),
]


def run(point, n_units, reps):
    name, grid, nprb, s1, s2, pat, layers, sm, fill, share = point
    n_cdm = (layers + 1) // 2
    hop1 = synth.make_hop(grid, 0, nprb, s1, pat, n_cdm)
    hop2 = synth.make_hop(grid, grid - nprb, nprb, s2, pat, n_cdm) if s2 else synth.no_hop()
    cfg = synth.make_config(30e3, sm, True)
    gen = 256
    d = synth.make_units(hop1, hop2, cfg, layers, gen, pilot_share=share, seed=7, device="cuda")
    rx = d["rx"].repeat(n_units // gen, 1, 1)
    pil = d["pilots"].repeat(n_units // gen, 1, 1, 1)
    plan = estimator.make_plan(grid, 14, layers, hop1, hop2, cfg, fill)
    i = plan.info
    out = torch.empty((n_units, i.n_sc, 14, layers), dtype=torch.complex64, device="cuda")
    sc = torch.empty((n_units, 6), dtype=torch.float64, device="cuda")
    for _ in range(3):
        plan.estimate(rx, pil, 1.0, share, out=out, scalars=sc)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        plan.estimate(rx, pil, 1.0, share, out=out, scalars=sc)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    st = i.n_dmrs_total
    bpu = 8 * (i.n_cdm * i.n_re * st + i.n_re * st * layers // share + i.n_sc * 14 * layers) + 48
    return {"point": name, "units": n_units, "ms": round(ms, 4), "units_per_s": round(n_units / ms * 1e3),
            "estimates_per_s": round(layers * n_units / ms * 1e3), "algorithmic_GBps": round(bpu * n_units / ms / 1e6, 1),
            "frac_of_measured_hbm_peak": round(bpu * n_units / ms / 1e6 / PEAK, 3), "smem_bytes": i.smem_bytes,
            "cta_threads": i.threads, "fft_size": i.fft_size}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--units", type=int, default=8192)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--out", default="")
    ap.add_argument("--only", default="", help="comma-separated substring filters on the point name")
    a = ap.parse_args()
    rows = []
    for p in POINTS:
        if a.only and not any(o in p[0] for o in a.only.split(",")):
            continue
        r = run(p, a.units, a.reps)
        rows.append(r)
        print(json.dumps(r), flush=True)
    if a.out:
        Path(a.out).write_text(json.dumps({"gpu": torch.cuda.get_device_name(0), "hbm_peak_GBps": PEAK, "rows": rows}, indent=1))


if __name__ == "__main__":
    main()
