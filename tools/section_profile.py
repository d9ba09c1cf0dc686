"""Per-section cycle counters of the fused kernel (profiling build only: `make -C srsran_ce_pytorch_b200/csrc prof`).

    python tools/section_profile.py [--only substring] [--units N]

Thread 0 of the estimator warps and lane 0 of the first writer warp of every CTA accumulate clock64() deltas per
section (PROF_MARK in srsce_kernels.cu); this prints, per sweep point, the cycles per unit of each section
averaged over the CTAs.  The instrumented build is slower than the product; read SHARES, not throughput."""
import argparse
import ctypes as C
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
os.environ["SRSCE_LIBRARY"] = str(ROOT / "srsran_ce_pytorch_b200" / "libsrsce_prof.so")
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from srsran_ce_pytorch_b200 import _lib, estimator, synth  # noqa: E402
from tools.sweep_bench import POINTS  # noqa: E402

SECTIONS = ["wait stage", "cfo", "ls", "virt pilots", "wait writers", "fir(+blend)", "mean", "table+release",
            "noise/rsrp", "ta: argmax", "stage issue", "W wait full", "W stream", "ta: pass 1", "ta: pass 2-3", "ta: profile"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--units", type=int, default=4096)
    ap.add_argument("--only", default="", help="comma-separated substring filters on the point name")
    ap.add_argument("--ablate", type=int, default=0, help="SRSCE_DEBUG_ABLATE bits (2: no dense write = phase A alone)")
    a = ap.parse_args()
    if a.ablate:
        os.environ["SRSCE_DEBUG_ABLATE"] = str(a.ablate)
    lib = _lib.lib
    lib.srsce_prof_set_buffer.argtypes = [C.c_void_p]
    lib.srsce_prof_set_buffer.restype = C.c_int
    for point in POINTS:
        name, grid, nprb, s1, s2, pat, layers, sm, fill, share = point
        if a.only and not any(o in name for o in a.only.split(",")):
            continue
        n_cdm = (layers + 1) // 2
        hop1 = synth.make_hop(grid, 0, nprb, s1, pat, n_cdm)
        hop2 = synth.make_hop(grid, grid - nprb, nprb, s2, pat, n_cdm) if s2 else synth.no_hop()
        cfg = synth.make_config(30e3, sm, True)
        d = synth.make_units(hop1, hop2, cfg, layers, 256, pilot_share=share, seed=7, device="cuda")
        rx = d["rx"].repeat(a.units // 256, 1, 1)
        pil = d["pilots"].repeat(a.units // 256, 1, 1, 1)
        plan = estimator.make_plan(grid, 14, layers, hop1, hop2, cfg, fill)
        out = torch.empty((a.units, plan.info.n_sc, 14, layers), dtype=torch.complex64, device="cuda")
        sc = torch.empty((a.units, 6), dtype=torch.float64, device="cuda")
        for _ in range(2):
            plan.estimate(rx, pil, 1.0, share, out=out, scalars=sc)
        buf = torch.zeros((1024, 16), dtype=torch.int64, device="cuda")
        assert lib.srsce_prof_set_buffer(buf.data_ptr()) == 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        plan.estimate(rx, pil, 1.0, share, out=out, scalars=sc)
        e1.record()
        torch.cuda.synchronize()
        lib.srsce_prof_set_buffer(None)
        ms = e0.elapsed_time(e1)
        t = buf.cpu().double()
        ctas = int((t.sum(1) > 0).sum())
        per_unit = t[:ctas].sum(0) / a.units          # cycles per unit (every unit is counted by exactly one CTA)
        est_total = float(per_unit[:11].sum() + per_unit[13:16].sum())
        print(f"== {name}: {ms:.3f} ms for {a.units} units, {ctas} CTAs, threads {plan.info.threads}, smem {plan.info.smem_bytes}")
        print("   estimator cycles/unit %.0f | writer cycles/unit %.0f" % (est_total, float(per_unit[11:13].sum())))
        for k, sname in enumerate(SECTIONS):
            if per_unit[k] > 0:
                ref = float(per_unit[11:13].sum()) if k in (11, 12) else est_total
                print(f"   {sname:14s} {float(per_unit[k]):9.0f} cyc  {100 * float(per_unit[k]) / ref:5.1f} %")
        del rx, pil, out, sc
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
