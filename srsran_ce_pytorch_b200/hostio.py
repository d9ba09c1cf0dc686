"""Host-buffer entry point: the reference call as its callers make it -- received grids and pilots
in host memory in, channel grids and scalars in host memory out -- pipelined over two CUDA streams
so that the host->device copy of chunk k+1, the fused kernel on chunk k and the device->host copy
of chunk k-1 overlap (PCIe is full duplex).  PyTorch provides pinned memory, streams and copies;
the arithmetic is ``Plan.estimate`` (libsrsce.so).

The dense [n_sc, n_sym, L] grid the reference returns is 99 % of the bytes that cross PCIe, so this
path is bounded by the device->host link, not by the kernel.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from .estimator import Plan


class HostPipeline:
    """Reusable staging buffers for ``chunk_units`` units per stream (allocated once)."""

    def __init__(self, plan: Plan, chunk_units: int, pilot_share: int, device: Optional[torch.device] = None,
                 n_streams: int = 2, compact: bool = False):
        if chunk_units % pilot_share:
            raise ValueError("chunk_units must be a multiple of pilot_share")
        i = plan.info
        self.plan, self.chunk, self.share = plan, int(chunk_units), int(pilot_share)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(n_streams)]
        c64, dev = torch.complex64, self.device
        self.d_rx = [torch.empty((self.chunk, i.n_sc, i.n_sym), dtype=c64, device=dev) for _ in self.streams]
        self.d_pil = [torch.empty((self.chunk // self.share, i.n_re, i.n_dmrs_total, i.n_layers), dtype=c64, device=dev)
                      for _ in self.streams]
        self.compact = bool(compact)
        if not self.compact:
            self.d_out = [torch.empty((self.chunk, i.n_sc, i.n_sym, i.n_layers), dtype=c64, device=dev) for _ in self.streams]
        self.d_sc = [torch.empty((self.chunk, 6), dtype=torch.float64, device=dev) for _ in self.streams]
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self.launches = 0

    def run(self, rx: torch.Tensor, pilots: torch.Tensor, beta: float, out: torch.Tensor, scalars: torch.Tensor
            ) -> Tuple[torch.Tensor, torch.Tensor]:
        """rx [B, n_sc, n_sym], pilots [B // share, n_re, S, L], out [B, n_sc, n_sym, L] (compact mode: the pair
        (ch_sc [B, hops, n_sc, L], sym_rot [B, n_sym])), scalars [B, 6]: host tensors (pinned for real overlap).
        Returns (out, scalars) once every copy has landed."""
        for t in (rx, pilots, scalars) + (tuple(out) if self.compact else (out,)):
            if t.is_cuda:
                raise ValueError("HostPipeline takes host tensors; call Plan.estimate for device tensors")
        n = int(rx.shape[0])
        if n % self.share:
            raise ValueError("number of units must be a multiple of pilot_share")
        self.h2d_bytes = self.d2h_bytes = self.launches = 0
        for k, b0 in enumerate(range(0, n, self.chunk)):
            b1 = min(n, b0 + self.chunk)
            m = b1 - b0
            s = k % len(self.streams)
            with torch.cuda.stream(self.streams[s]):
                d_rx, d_pil = self.d_rx[s][:m], self.d_pil[s][:m // self.share]
                d_rx.copy_(rx[b0:b1], non_blocking=True)
                d_pil.copy_(pilots[b0 // self.share:b1 // self.share], non_blocking=True)
                if self.compact:
                    ch, rot, sc = self.plan.estimate_compact(d_rx, d_pil, beta, self.share)
                    out[0][b0:b1].copy_(ch, non_blocking=True)
                    out[1][b0:b1].copy_(rot, non_blocking=True)
                    scalars[b0:b1].copy_(sc, non_blocking=True)
                    for t in (ch, rot, sc):                       # freed on the host before the copies ran
                        t.record_stream(self.streams[s])
                    d2h = (ch.numel() + rot.numel()) * 8 + m * 48
                else:
                    self.plan.estimate(d_rx, d_pil, beta, self.share, out=self.d_out[s][:m], scalars=self.d_sc[s][:m])
                    out[b0:b1].copy_(self.d_out[s][:m], non_blocking=True)
                    scalars[b0:b1].copy_(self.d_sc[s][:m], non_blocking=True)
                    d2h = m * self.d_out[s][0].numel() * 8 + m * 48
            self.h2d_bytes += d_rx.numel() * 8 + d_pil.numel() * 8
            self.d2h_bytes += d2h
            self.launches += 1
        for st in self.streams:
            st.synchronize()
        return out, scalars


def srs_channel_estimator_host(plan: Plan, rx, pilots, beta_dmrs: float, pilot_share: int = 1,
                               chunk_units: int = 256, pipeline: Optional[HostPipeline] = None):
    """Convenience wrapper: allocate pinned outputs and run one pipelined pass."""
    i = plan.info
    n = int(rx.shape[0])
    pipe = pipeline or HostPipeline(plan, min(chunk_units, max(pilot_share, n)), pilot_share)
    out = torch.empty((n, i.n_sc, i.n_sym, i.n_layers), dtype=torch.complex64).pin_memory()
    sc = torch.empty((n, 6), dtype=torch.float64).pin_memory()
    return pipe.run(rx, pilots, float(beta_dmrs), out, sc)
