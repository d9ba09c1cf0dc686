"""Import-compatible stand-ins for the reference's three modules.

Putting this directory on ``sys.path`` keeps ``from ce_rule_tensorized import EstimatorConfig,
HopConfig, srs_channel_estimator`` (and ce_rule_baseline / ce_dl_cnn) working unchanged, with the
computation done by the sm_100a kernels.  CUDA tensors in -> CUDA tensors out; CPU tensors raise (there is
no CPU fallback).  Opt-in for callers that hold host arrays, as the reference's scripts do:
``srsran_ce_pytorch_b200.estimator.accept_host_arrays()`` or SRSCE_ACCEPT_HOST_ARRAYS=1 -- host arrays are then
copied to the current CUDA device and the results copied back (a transfer; without a CUDA device it raises).
"""
