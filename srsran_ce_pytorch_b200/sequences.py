"""Pilot sequence generation -- an EXTENSION with no counterpart in the reference, which takes the pilots
as an input tensor (/root/reference src/ce_rule_tensorized.py:746-747): parity is unpinned; the checks are properties
(constant modulus, ideal periodic autocorrelation of the Zadoff-Chu core, the OCC relation between the two ports of a
comb) and an independent float64 restatement of the TS 38.211 5.2.2 formulas in tests/.

``low_papr_pilots`` runs on the device through ``srsce_gen_pilots`` and returns the estimator's pilot layout
[n_ue, n_re, n_dmrs, n_layers]; there is no host fallback."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import lib


def low_papr_pilots(u: torch.Tensor, n_re: int, n_dmrs: int, n_layers: int, v: Optional[torch.Tensor] = None) -> torch.Tensor:
    """u: int32 CUDA tensor [n_ue] of sequence groups (0..29); v: optional int32 [n_ue] base sequence numbers (0 / 1).
    n_re = pilots per CDM group = sequence length M (30 or >= 36)."""
    if not u.is_cuda:
        raise RuntimeError("srsran_ce_pytorch_b200 computes on CUDA tensors only (no CPU fallback)")
    u = u.to(torch.int32).contiguous()
    if v is not None:
        v = v.to(device=u.device, dtype=torch.int32).contiguous()
        if v.shape != u.shape:
            raise IndexError("u and v must have the same shape")
    out = torch.empty((u.numel(), n_re, n_dmrs, n_layers), dtype=torch.complex64, device=u.device)
    with torch.cuda.device(u.device):
        stream = torch.cuda.current_stream().cuda_stream
        _lib.raise_for(lib.srsce_gen_pilots(out.data_ptr(), u.numel(), int(n_re), int(n_dmrs), int(n_layers), u.data_ptr(),
                                            v.data_ptr() if v is not None else None, C.c_void_p(stream)))
    return out
