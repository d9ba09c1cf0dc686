"""Drop-in for /root/reference src/ce_dl_cnn.py (entry point :802): the estimator whose
frequency interpolation is the fixed-stencil partial-convolution inpainting (:433-508) and whose
smoothed pilots may be blended with a low-pass copy (``CNNSmoothingAlpha``, :712-717, :864-867)."""
from srsran_ce_pytorch_b200.estimator import srs_channel_estimator_any as _estimate
from srsran_ce_pytorch_b200.types import CnnEstimatorConfig, EstimatorConfig, HopConfig

__all__ = ["HopConfig", "EstimatorConfig", "CnnEstimatorConfig", "srs_channel_estimator"]


def srs_channel_estimator(received_rg, pilots, beta_dmrs, hop1, hop2, config):
    return _estimate(received_rg, pilots, beta_dmrs, hop1, hop2, config, fill="cnn")
