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

# single calls, as the reference's own scripts make them (one unit per call): drop-in entry point (wall clock per call,
# synchronised), Plan.estimate into preallocated buffers, and the same launch replayed from a CUDA graph
import time
rx1, pil1 = d["rx"][0].contiguous(), d["pilots"][0].contiguous()
for _ in range(20): estimator.srs_channel_estimator(rx1, pil1, 1.4125, hop1, synth.no_hop(), cfg)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(200): estimator.srs_channel_estimator(rx1, pil1, 1.4125, hop1, synth.no_hop(), cfg)
torch.cuda.synchronize(); us_dropin = (time.perf_counter() - t0) / 200 * 1e6
o1, s1 = out[:1], sc[:1]
r1, p1 = rx[:1], pil[:1]
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(200): plan.estimate(r1, p1, 1.4125, 1, out=o1, scalars=s1)
torch.cuda.synchronize(); us_plan = (time.perf_counter() - t0) / 200 * 1e6
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    plan.estimate(r1, p1, 1.4125, 1, out=o1, scalars=s1)
for _ in range(10): g.replay()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(200): g.replay()
torch.cuda.synchronize(); us_graph = (time.perf_counter() - t0) / 200 * 1e6
e0.record()
for _ in range(200): g.replay()
e1.record(); torch.cuda.synchronize()
print(f"case0 single unit per call: drop-in {us_dropin:.1f} us, Plan.estimate {us_plan:.1f} us, CUDA-graph replay {us_graph:.1f} us wall "
      f"({e0.elapsed_time(e1) / 200 * 1e3:.1f} us on the device)")
