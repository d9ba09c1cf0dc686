"""Import-compatible stand-ins for the reference's three modules.

Putting this directory on ``sys.path`` keeps ``from ce_rule_tensorized import EstimatorConfig,
HopConfig, srs_channel_estimator`` (and ce_rule_baseline / ce_dl_cnn) working unchanged, with the
computation done by the sm_100a kernels.  Inputs must be CUDA tensors.
"""
