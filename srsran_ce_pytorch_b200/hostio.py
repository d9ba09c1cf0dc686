"""Host-buffer entry point: the reference call as its callers make it -- received grids and pilots
in host memory in, channel grids and scalars in host memory out -- pipelined over two CUDA streams
so that the host->device copy of chunk k+1, the fused kernel on chunk k and the device->host copy
of chunk k-1 overlap (PCIe is full duplex).  PyTorch provides pinned memory, streams and copies;
the arithmetic is ``Plan.estimate`` (libsrsce.so).

The dense [n_sc, n_sym, L] grid the reference returns is 99 % of the bytes that cross PCIe, so a
dense device->host copy bounds this path by the link (54 GB/s), not by the kernel.  13/14 of those
bytes are the replication of one column over the allocated symbols (ce_rule_tensorized.py:346-347):
``replicate="host"`` moves the per-hop columns instead (``srsce_estimate_compact``) and lets the
library's host threads rebuild the grid in the caller's buffer (``srsce_replicate_host`` -- a copy
with non-temporal stores, no arithmetic, bit-identical to the device-written grid).  Plans whose grid
carries the final CFO rotation keep the dense copy (the host side never multiplies).
"""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor
from typing import Optional, Tuple

import torch

from .estimator import Plan


class HostPipeline:
    """Reusable staging buffers for ``chunk_units`` units per stream (allocated once)."""

    def __init__(self, plan: Plan, chunk_units: int, pilot_share: int, device: Optional[torch.device] = None,
                 n_streams: int = 2, compact: bool = False, replicate: str = "device", host_threads: int = 0):
        """``replicate``: "device" -- the fused kernel writes the dense grid, which then crosses PCIe; "host" -- the
        per-hop columns cross PCIe and host threads replicate them into ``out`` ("auto": "host" unless the plan has a
        final CFO rotation).  ``compact`` returns the factors themselves (not the reference contract)."""
        if replicate not in ("device", "host", "auto"):
            raise ValueError(f"unknown replicate mode {replicate!r}")
        if replicate == "auto":
            replicate = "device" if plan.has_final_rotation else "host"
        if replicate == "host" and plan.has_final_rotation:
            raise NotImplementedError("this plan's grid carries the final CFO rotation; the host side only replicates")
        if compact and replicate == "host":
            raise ValueError("compact output and host replication are alternatives")
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
        self.replicate, self.host_threads = replicate, int(host_threads)
        if replicate == "host":
            n_hops = 2 if plan.desc.hop[1].dmrs_symbol_mask else 1
            self.h_ch = [torch.empty((self.chunk, n_hops, i.n_sc, i.n_layers), dtype=c64).pin_memory() for _ in self.streams]
            self._worker = ThreadPoolExecutor(max_workers=1, thread_name_prefix="srsce-replicate")
        elif not self.compact:
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
        pending = [None] * len(self.streams)          # host replication still reading each stream's staging buffer
        for k, b0 in enumerate(range(0, n, self.chunk)):
            b1 = min(n, b0 + self.chunk)
            m = b1 - b0
            s = k % len(self.streams)
            if pending[s] is not None:
                pending[s].result()
                pending[s] = None
            with torch.cuda.stream(self.streams[s]):
                d_rx, d_pil = self.d_rx[s][:m], self.d_pil[s][:m // self.share]
                d_rx.copy_(rx[b0:b1], non_blocking=True)
                d_pil.copy_(pilots[b0 // self.share:b1 // self.share], non_blocking=True)
                if self.replicate == "host":
                    ch, _rot, sc = self.plan.estimate_compact(d_rx, d_pil, beta, self.share)
                    self.h_ch[s][:m].copy_(ch, non_blocking=True)
                    scalars[b0:b1].copy_(sc, non_blocking=True)
                    for t in (ch, _rot, sc):
                        t.record_stream(self.streams[s])
                    landed = torch.cuda.Event()
                    landed.record(self.streams[s])
                    pending[s] = self._worker.submit(self._replicate, landed, self.h_ch[s][:m], out[b0:b1])
                    d2h = ch.numel() * 8 + m * 48
                elif self.compact:
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
        for f in pending:
            if f is not None:
                f.result()
        for st in self.streams:
            st.synchronize()
        return out, scalars

    def _replicate(self, landed: torch.cuda.Event, ch_host: torch.Tensor, out: torch.Tensor) -> None:
        landed.synchronize()
        self.plan.replicate_host(ch_host, out, self.host_threads)


def srs_channel_estimator_host(plan: Plan, rx, pilots, beta_dmrs: float, pilot_share: int = 1,
                               chunk_units: int = 256, pipeline: Optional[HostPipeline] = None):
    """Convenience wrapper: allocate pinned outputs and run one pipelined pass."""
    i = plan.info
    n = int(rx.shape[0])
    pipe = pipeline or HostPipeline(plan, min(chunk_units, max(pilot_share, n)), pilot_share, n_streams=3, replicate="auto")
    out = torch.empty((n, i.n_sc, i.n_sym, i.n_layers), dtype=torch.complex64).pin_memory()
    sc = torch.empty((n, 6), dtype=torch.float64).pin_memory()
    return pipe.run(rx, pilots, float(beta_dmrs), out, sc)


class FillHostPipeline:
    """The fill seam for a caller that holds host arrays: smoothed pilots [B, n_hops, n_re, L] in host memory in, dense
    grids [B, n_sc, n_sym, L] in host memory out (``srsce_fill`` semantics).  ``replicate="host"`` (default): the
    per-hop columns of ``srsce_fill_compact`` cross PCIe and ``srsce_replicate_host`` rebuilds the grid (the fill has no
    CFO rotation, so this is always a pure copy); ``"device"``: ``srsce_fill`` writes the dense grid on the device
    (``impl`` direct or tensor) and the whole grid crosses the link."""

    def __init__(self, plan: Plan, chunk_units: int = 256, device: Optional[torch.device] = None, n_streams: int = 3,
                 replicate: str = "host", impl: str = "direct", host_threads: int = 0):
        if replicate not in ("device", "host"):
            raise ValueError(f"unknown replicate mode {replicate!r}")
        i = plan.info
        n_hops = 2 if plan.desc.hop[1].dmrs_symbol_mask else 1
        self.plan, self.chunk, self.replicate, self.impl, self.host_threads = plan, int(chunk_units), replicate, impl, int(host_threads)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(n_streams)]
        c64, dev = torch.complex64, self.device
        self.d_h = [torch.empty((self.chunk, n_hops, i.n_re, i.n_layers), dtype=c64, device=dev) for _ in self.streams]
        if replicate == "host":
            self.h_ch = [torch.empty((self.chunk, n_hops, i.n_sc, i.n_layers), dtype=c64).pin_memory() for _ in self.streams]
            self._worker = ThreadPoolExecutor(max_workers=1, thread_name_prefix="srsce-replicate")
        else:
            self.d_out = [torch.empty((self.chunk, i.n_sc, i.n_sym, i.n_layers), dtype=c64, device=dev) for _ in self.streams]
        self.h2d_bytes = self.d2h_bytes = self.launches = 0

    def run(self, h_pilots: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        if h_pilots.is_cuda or out.is_cuda:
            raise ValueError("FillHostPipeline takes host tensors; call Plan.fill for device tensors")
        n = int(h_pilots.shape[0])
        self.h2d_bytes = self.d2h_bytes = self.launches = 0
        pending = [None] * len(self.streams)
        for k, b0 in enumerate(range(0, n, self.chunk)):
            b1 = min(n, b0 + self.chunk)
            m = b1 - b0
            s = k % len(self.streams)
            if pending[s] is not None:
                pending[s].result()
                pending[s] = None
            with torch.cuda.stream(self.streams[s]):
                d_h = self.d_h[s][:m]
                d_h.copy_(h_pilots[b0:b1], non_blocking=True)
                if self.replicate == "host":
                    ch = self.plan.fill_compact(d_h)
                    self.h_ch[s][:m].copy_(ch, non_blocking=True)
                    ch.record_stream(self.streams[s])
                    landed = torch.cuda.Event()
                    landed.record(self.streams[s])
                    pending[s] = self._worker.submit(self._replicate, landed, self.h_ch[s][:m], out[b0:b1])
                    self.d2h_bytes += ch.numel() * 8
                else:
                    self.plan.fill(d_h, impl=self.impl, out=self.d_out[s][:m])
                    out[b0:b1].copy_(self.d_out[s][:m], non_blocking=True)
                    self.d2h_bytes += m * self.d_out[s][0].numel() * 8
            self.h2d_bytes += d_h.numel() * 8
            self.launches += 1
        for f in pending:
            if f is not None:
                f.result()
        for st in self.streams:
            st.synchronize()
        return out

    def _replicate(self, landed: torch.cuda.Event, ch_host: torch.Tensor, out: torch.Tensor) -> None:
        landed.synchronize()
        self.plan.replicate_host(ch_host, out, self.host_threads, from_fill=True)
