"""Scratch timing of the BASELINE config-1 shape (validate_case0 geometry) batched: 52-PRB grid, 3-PRB hop,
1 layer, 4 pilot symbols, comb-2, CFO on."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from srsran_ce_pytorch_b200 import estimator, synth
n_units = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
hop1 = synth.make_hop(52, 40, 3, [0, 4, 8, 12], "comb2", 1)
cfg = synth.make_config(15e3, "filter", True, validator_cp=True)
gen = 1024
d = synth.make_units(hop1, synth.no_hop(), cfg, 1, gen, pilot_share=1, seed=1, device="cuda", beta=1.4125)
rx = d["rx"].repeat(n_units // gen, 1, 1); pil = d["pilots"].repeat(n_units // gen, 1, 1, 1)
plan = estimator.make_plan(52, 14, 1, hop1, synth.no_hop(), cfg)
i = plan.info
out = torch.empty((n_units, i.n_sc, 14, 1), dtype=torch.complex64, device="cuda"); sc = torch.empty((n_units, 6), dtype=torch.float64, device="cuda")
for _ in range(3): plan.estimate(rx, pil, 1.4125, 1, out=out, scalars=sc)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(5): plan.estimate(rx, pil, 1.4125, 1, out=out, scalars=sc)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
bpu = 8 * (i.n_re * 4 + i.n_re * 4 + i.n_sc * 14) + 48
print(f"case0 shape: units={n_units} ms={ms:.3f} units/s={n_units/ms*1e3:.0f} algGB/s={bpu*n_units/ms/1e6:.1f} frac={bpu*n_units/ms/1e6/6416.7:.3f} smem={i.smem_bytes} threads={i.threads} bytes/unit={bpu}")
