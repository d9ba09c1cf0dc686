"""B200-native drop-in for the channel estimator of pjookim/srsran-ce-pytorch.

Python host side (dataclasses, plan cache, ctypes binding, batching/sharding) over a C-ABI
library of hand-written sm_100a CUDA kernels (``csrc/``, ``include/srsce.h``).  There is no CPU
fallback: importing the estimator raises if ``libsrsce.so`` is missing, and every compute entry
point raises if the tensors are not on a CUDA device.
"""
from .types import CnnEstimatorConfig, EstimatorConfig, HopConfig

__all__ = ["HopConfig", "EstimatorConfig", "CnnEstimatorConfig", "srs_channel_estimator",
           "srs_channel_estimator_batched", "make_plan"]


def __getattr__(name):
    if name in ("srs_channel_estimator", "srs_channel_estimator_batched", "make_plan", "Plan", "unpack_scalars"):
        from . import estimator
        return getattr(estimator, name)
    raise AttributeError(name)
