"""End-to-end (host buffers in / dense grid out) throughput vs replication side, chunk size, stream count, host threads."""
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
ref_out, ref_sc = plan.estimate(rx[:64], pil[:16], 1.0, 4)
ref_out, ref_sc = ref_out.cpu(), ref_sc.cpu()
for mode, chunk, streams, threads in (("device", 64, 2, 0), ("host", 64, 2, 0), ("host", 128, 2, 0), ("host", 256, 2, 0),
                                      ("host", 128, 3, 0), ("host", 256, 3, 0), ("host", 512, 2, 0), ("host", 256, 2, 8),
                                      ("host", 256, 2, 12), ("host", 128, 4, 0)):
    pipe = HostPipeline(plan, chunk, 4, n_streams=streams, replicate=mode, host_threads=threads)
    h_out.zero_()
    pipe.run(h_rx, h_pil, 1.0, h_out, h_sc)
    ok = torch.equal(h_out[:64], ref_out) and torch.equal(h_sc[:64], ref_sc) and torch.equal(h_out[-1], h_out[-1 - 0]) \
        and bool((h_out[-1] != 0).any())
    t0 = time.perf_counter()
    for _ in range(4):
        pipe.run(h_rx, h_pil, 1.0, h_out, h_sc)
    dt = (time.perf_counter() - t0) / 4
    print(f"replicate={mode} chunk={chunk} streams={streams} threads={threads or 'all'}: {4*n/dt:.0f} est/s, "
          f"D2H {pipe.d2h_bytes/dt/1e9:.1f} GB/s, H2D {pipe.h2d_bytes/dt/1e9:.1f} GB/s, grid written {n*3276*14*4*8/dt/1e9:.1f} GB/s, "
          f"bit-identical to the device path: {ok}", flush=True)
    del pipe
