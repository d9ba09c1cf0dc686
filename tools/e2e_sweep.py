"""Scratch: end-to-end (host buffers in/out) throughput vs chunk size / stream count."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import bench
from srsran_ce_pytorch_b200 import estimator
from srsran_ce_pytorch_b200.hostio import HostPipeline

hop1, hop2, cfg = bench.geometry()
plan = estimator.make_plan(273, 14, 4, hop1, hop2, cfg)
rx, pil = bench.make_step_inputs(8, 1, torch.device("cuda"))
n = rx.shape[0]
h_rx, h_pil = rx.cpu().pin_memory(), pil.cpu().pin_memory()
h_out = torch.empty((n, 3276, 14, 4), dtype=torch.complex64).pin_memory()
h_sc = torch.empty((n, 6), dtype=torch.float64).pin_memory()
for chunk, streams in ((64, 2), (128, 2), (256, 2), (128, 3), (256, 3), (512, 2), (128, 4)):
    pipe = HostPipeline(plan, chunk, 4, n_streams=streams)
    pipe.run(h_rx, h_pil, 1.0, h_out, h_sc)
    t0 = time.perf_counter()
    for _ in range(4):
        pipe.run(h_rx, h_pil, 1.0, h_out, h_sc)
    dt = (time.perf_counter() - t0) / 4
    print(f"chunk={chunk} streams={streams}: {4*n/dt:.0f} est/s, D2H {pipe.d2h_bytes/dt/1e9:.1f} GB/s, H2D {pipe.h2d_bytes/dt/1e9:.1f} GB/s")
    del pipe
