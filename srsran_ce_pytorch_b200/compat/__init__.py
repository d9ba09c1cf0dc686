"""Import-compatible stand-ins for the reference's three modules.

Putting this directory on ``sys.path`` keeps ``from ce_rule_tensorized import EstimatorConfig,
HopConfig, srs_channel_estimator`` (and ce_rule_baseline / ce_dl_cnn) working unchanged, with the
computation done by the sm_100a kernels.  CUDA tensors in -> CUDA tensors out; host tensors (what the
reference's scripts pass) are copied to the current CUDA device and the results copied back
(``estimator.srs_channel_estimator_any``).  Without a CUDA device the call raises: no CPU compute path.
"""
