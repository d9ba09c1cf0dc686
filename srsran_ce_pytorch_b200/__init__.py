"""B200-native drop-in for the channel estimator of pjookim/srsran-ce-pytorch.

Python host side (dataclasses, plan cache, ctypes binding, batching/sharding) over a C-ABI
library of hand-written sm_100a CUDA kernels (``csrc/``, ``include/srsce.h``).  There is no CPU
fallback: every compute entry point raises if ``libsrsce.so`` is missing or the tensors are not
on a CUDA device.
"""
from .types import CnnEstimatorConfig, EstimatorConfig, HopConfig

__all__ = ["HopConfig", "EstimatorConfig", "CnnEstimatorConfig"]
