"""Scratch timing of the fused kernel (not the bench contract; see bench.py).

    python tools/quick_time.py UNITS [SYMS] [PATTERN] [FILL] [SMOOTHING] [LAYERS] [NPRB_HOP] [HOP2SYMS]
"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from srsran_ce_pytorch_b200 import estimator, synth

a = sys.argv[1:] + [""] * 9
n_units = int(a[0] or 4096)
syms = [int(s) for s in (a[1] or "13").split(",")]
pattern, fill, smoothing = a[2] or "comb4", a[3] or "linear", a[4] or "filter"
layers = int(a[5] or 4)
nprb = int(a[6] or 273)
hop2_syms = [int(s) for s in a[7].split(",")] if a[7] else None
share = 4
hop1 = synth.make_hop(273, 0, nprb, syms, pattern, (layers + 1) // 2)
hop2 = synth.make_hop(273, 273 - nprb, nprb, hop2_syms, pattern, (layers + 1) // 2) if hop2_syms else synth.no_hop()
cfg = synth.make_config(30e3, smoothing, True)
gen = 256
data = synth.make_units(hop1, hop2, cfg, layers, gen, pilot_share=share, seed=1, device="cuda")
rx = data["rx"].repeat(n_units // gen, 1, 1)
pil = data["pilots"].repeat(n_units // gen, 1, 1, 1)
plan = estimator.make_plan(273, 14, layers, hop1, hop2, cfg, fill)
i = plan.info
out = torch.empty((n_units, i.n_sc, 14, layers), dtype=torch.complex64, device="cuda")
sc = torch.empty((n_units, 6), dtype=torch.float64, device="cuda")
for _ in range(3):
    plan.estimate(rx, pil, 1.0, share, out=out, scalars=sc)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 5
e0.record()
for _ in range(reps):
    plan.estimate(rx, pil, 1.0, share, out=out, scalars=sc)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
st = i.n_dmrs_total
bpu = 8 * (i.n_cdm * i.n_re * st + i.n_re * st * layers / share + i.n_sc * 14 * layers) + 48
print(f"units={n_units} syms={syms} hop2={hop2_syms} {pattern} L={layers} nprb={nprb} {fill} {smoothing}: ms={ms:.3f} "
      f"units/s={n_units/ms*1e3:.0f} est/s={layers*n_units/ms*1e3:.0f} algGB/s={bpu*n_units/ms/1e6:.1f} "
      f"frac={bpu*n_units/ms/1e6/6416.7:.3f} smem={i.smem_bytes}")
