"""Drop-in for /root/reference src/ce_rule_baseline.py (entry point :761).  The baseline and the
tensorized reference compute the same estimator (they differ by <= 3.3e-7 relative),
so both names are served by the same fused kernel."""
from srsran_ce_pytorch_b200.estimator import srs_channel_estimator_any as _estimate
from srsran_ce_pytorch_b200.types import EstimatorConfig, HopConfig

__all__ = ["HopConfig", "EstimatorConfig", "srs_channel_estimator"]


def srs_channel_estimator(received_rg, pilots, beta_dmrs, hop1, hop2, config):
    return _estimate(received_rg, pilots, beta_dmrs, hop1, hop2, config, fill="linear")
