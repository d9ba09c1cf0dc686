"""ctypes binding of libsrsce.so (C ABI declared in include/srsce.h).

The library is the product: if it is missing the import fails loudly -- there is no CPU or
PyTorch fallback for the compute entry points.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("SRSCE_LIBRARY", _HERE / "libsrsce.so"))   # override: A/B runs of two builds

MAX_SYMBOLS = 14
ABI_VERSION = 3

SMOOTH = {"none": 0, "mean": 1, "filter": 2}
FILL_LINEAR, FILL_CNN = 0, 1
FILL_IMPL = {"direct": 0, "tensor": 1}
REPLICATE_FILL = 1

ERR_INVALID, ERR_ASSERT, ERR_UNSUPPORTED, ERR_NO_DEVICE, ERR_CUDA = -1, -2, -3, -4, -5


class HopDesc(C.Structure):
    _fields_ = [("dmrs_symbol_mask", C.c_uint32), ("prb_start", C.c_int32), ("n_prbs", C.c_int32),
                ("start_symbol", C.c_int32), ("n_alloc_symbols", C.c_int32)]


class PlanDesc(C.Structure):
    _fields_ = [("n_prb_grid", C.c_int32), ("n_sym", C.c_int32), ("n_layers", C.c_int32), ("n_cp", C.c_int32),
                ("re_mask", C.c_uint16 * 2), ("smoothing", C.c_int32), ("cfo_compensate", C.c_int32),
                ("fill", C.c_int32), ("cnn_alpha", C.c_double), ("scs_hz", C.c_double),
                ("cp_ms", C.c_double * MAX_SYMBOLS), ("hop", HopDesc * 2)]


class PlanInfo(C.Structure):
    _fields_ = [("n_sc", C.c_int32), ("n_sym", C.c_int32), ("n_layers", C.c_int32), ("n_cdm", C.c_int32),
                ("n_re", C.c_int32), ("n_dmrs_total", C.c_int32), ("n_taps", C.c_int32),
                ("n_virtual", C.c_int32), ("fft_size", C.c_int32), ("cfo_defined", C.c_int32),
                ("smem_bytes", C.c_int32), ("threads", C.c_int32), ("fill_terms", C.c_int32),
                ("has_device", C.c_int32), ("bytes_per_unit", C.c_int64), ("fill_tensor", C.c_int32),
                ("fill_tensor_k", C.c_int32), ("fill_tensor_terms", C.c_int32), ("fill_tensor_units", C.c_int32),
                ("fill_tensor_tiles", C.c_int32), ("fill_tensor_kinds", C.c_int32), ("fill_tensor_a_terms", C.c_int32),
                ("h_in_global", C.c_int32)]


def _load() -> C.CDLL:
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            f"or `make -C {_HERE / 'csrc'}`.  srsran_ce_pytorch_b200 has no CPU / PyTorch fallback.")
    lib = C.CDLL(os.fspath(LIB_PATH))
    lib.srsce_abi_version.restype = C.c_int
    lib.srsce_last_error.restype = C.c_char_p
    lib.srsce_plan_create.argtypes = [C.POINTER(PlanDesc), C.POINTER(C.c_void_p)]
    lib.srsce_plan_create.restype = C.c_int
    lib.srsce_plan_destroy.argtypes = [C.c_void_p]
    lib.srsce_plan_destroy.restype = None
    lib.srsce_plan_query.argtypes = [C.c_void_p, C.POINTER(PlanInfo)]
    lib.srsce_plan_query.restype = C.c_int
    lib.srsce_plan_tables.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.srsce_plan_tables.restype = C.c_int
    lib.srsce_estimate.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.c_void_p, C.c_int32, C.c_double,
                                   C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    lib.srsce_estimate.restype = C.c_int
    lib.srsce_estimate_compact.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.c_void_p, C.c_int32,
                                           C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    lib.srsce_estimate_compact.restype = C.c_int
    lib.srsce_estimate_indexed.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.c_void_p, C.c_void_p,
                                           C.c_int32, C.c_double, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    lib.srsce_estimate_indexed.restype = C.c_int
    lib.srsce_plan_fill_tensor_tables.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.srsce_plan_fill_tensor_tables.restype = C.c_int
    lib.srsce_fill.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]
    lib.srsce_fill.restype = C.c_int
    lib.srsce_gen_pilots.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.srsce_gen_pilots.restype = C.c_int
    lib.srsce_fill_compact.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    lib.srsce_fill_compact.restype = C.c_int
    lib.srsce_replicate_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32]
    lib.srsce_replicate_host.restype = C.c_int
    lib.srsce_last_launch_count.restype = C.c_int
    if lib.srsce_abi_version() != ABI_VERSION:
        raise ImportError(f"libsrsce.so ABI {lib.srsce_abi_version()} != binding ABI {ABI_VERSION}; rebuild")
    return lib


lib = _load()

EXPORTS = ("srsce_abi_version", "srsce_last_error", "srsce_plan_create", "srsce_plan_destroy", "srsce_plan_query",
           "srsce_plan_tables", "srsce_plan_fill_tensor_tables", "srsce_estimate", "srsce_estimate_compact", "srsce_estimate_indexed", "srsce_fill", "srsce_fill_compact", "srsce_gen_pilots", "srsce_replicate_host",
           "srsce_last_launch_count")


def last_error() -> str:
    return (lib.srsce_last_error() or b"").decode()


def raise_for(rc: int) -> None:
    """Map ABI error codes to the exception types the reference raises for the same condition."""
    if rc == 0:
        return
    msg = last_error()
    if rc == ERR_INVALID:
        raise ValueError(msg)
    if rc == ERR_ASSERT:
        raise AssertionError(msg)
    if rc == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError(f"srsce error {rc}: {msg}")
