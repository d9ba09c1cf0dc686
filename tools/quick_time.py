"""Scratch timing of the fused kernel at BASELINE config 2 (not the bench contract; see bench.py)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from srsran_ce_pytorch_b200 import estimator, synth

n_units = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
syms = [int(s) for s in sys.argv[2].split(",")] if len(sys.argv) > 2 else [13]
hop1 = synth.make_hop(273, 0, 273, syms, "comb4", 2)
cfg = synth.make_config(30e3, "filter", True)
gen = 256
data = synth.make_units(hop1, synth.no_hop(), cfg, 4, gen, pilot_share=4, seed=1, device="cuda")
rx = data["rx"].repeat(n_units // gen, 1, 1)
pil = data["pilots"].repeat(n_units // gen, 1, 1, 1)
import os, ctypes, glob
if os.environ.get("L2GRAN"):
    import nvidia.cuda_runtime as cr
    rt = ctypes.CDLL(glob.glob(os.path.dirname(cr.__file__) + "/lib/libcudart.so*")[0])
    rc = rt.cudaDeviceSetLimit(5, ctypes.c_size_t(int(os.environ["L2GRAN"])))
    val = ctypes.c_size_t()
    rt.cudaDeviceGetLimit(ctypes.byref(val), 5)
    print("cudaLimitMaxL2FetchGranularity rc", rc, "now", val.value)
plan = estimator.make_plan(273, 14, 4, hop1, synth.no_hop(), cfg)
out = torch.empty((n_units, 3276, 14, 4), dtype=torch.complex64, device="cuda")
sc = torch.empty((n_units, 6), dtype=torch.float64, device="cuda")
for _ in range(3):
    plan.estimate(rx, pil, 1.0, 4, out=out, scalars=sc)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 5
e0.record()
for _ in range(reps):
    plan.estimate(rx, pil, 1.0, 4, out=out, scalars=sc)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
bpu = 8 * (2 * 819 * len(syms) + 819 * len(syms) * 4 / 4 + 3276 * 14 * 4) + 48
print(f"units={n_units} S={len(syms)} ms={ms:.3f} units/s={n_units/ms*1e3:.0f} est/s={4*n_units/ms*1e3:.0f} "
      f"GB/s={bpu*n_units/ms/1e6:.1f} frac_of_6416.7={bpu*n_units/ms/1e6/6416.7:.3f} smem={plan.info.smem_bytes}")
