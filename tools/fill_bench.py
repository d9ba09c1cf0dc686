"""Device-timed throughput of srsce_fill (BASELINE config 4: ce_dl_cnn fill of 4096 UE-antenna pilot vectors at 273 PRB),
direct vs tcgen05 implementation.  One JSON line per point."""
import argparse
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

from srsran_ce_pytorch_b200 import estimator, synth

ROOT = Path(__file__).resolve().parents[1]
PEAK = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]) if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--units", type=int, default=4096)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--points", default="comb4:4,comb2:4,comb4:2,comb4:1,comb4:3,dense:2")
    a = ap.parse_args()
    for pt in a.points.split(","):
        pat, L = pt.split(":")
        L = int(L)
        hop1 = synth.make_hop(273, 0, 273, [13], pat, (L + 1) // 2)
        for fill in ("cnn", "linear"):
            plan = estimator.make_plan(273, 14, L, hop1, synth.no_hop(), synth.make_config(30e3, "filter", True), fill)
            i = plan.info
            h = torch.randn(a.units, 1, i.n_re, L, 2, device="cuda").view(torch.float32)
            h = torch.view_as_complex(h.contiguous())
            out = torch.empty((a.units, i.n_sc, 14, L), dtype=torch.complex64, device="cuda")
            bytes_unit = 8 * (i.n_re * L + i.n_sc * 14 * L)
            for impl in ("direct", "tensor"):
                if impl == "tensor" and not i.fill_tensor:
                    continue
                for _ in range(3):
                    plan.fill(h, impl=impl, out=out)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(a.reps):
                    plan.fill(h, impl=impl, out=out)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / a.reps
                print(json.dumps({"pattern": pat, "layers": L, "fill": fill, "impl": impl, "units": a.units, "ms": round(ms, 4),
                                  "estimates_per_s": round(L * a.units / ms * 1e3), "GBps": round(bytes_unit * a.units / ms / 1e6, 1),
                                  "frac_of_measured_hbm_peak": round(bytes_unit * a.units / ms / 1e6 / PEAK, 3),
                                  "k": i.fill_tensor_k, "products": i.fill_tensor_terms, "units_per_mma": i.fill_tensor_units}), flush=True)
            del h, out
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
