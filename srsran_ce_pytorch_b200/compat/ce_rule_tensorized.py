"""Drop-in for /root/reference src/ce_rule_tensorized.py (entry point :745, dataclasses :13-29)."""
from srsran_ce_pytorch_b200.estimator import srs_channel_estimator_any as _estimate
from srsran_ce_pytorch_b200.types import EstimatorConfig, HopConfig

__all__ = ["HopConfig", "EstimatorConfig", "srs_channel_estimator"]


def srs_channel_estimator(received_rg, pilots, beta_dmrs, hop1, hop2, config):
    return _estimate(received_rg, pilots, beta_dmrs, hop1, hop2, config, fill="linear")
