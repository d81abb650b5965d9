"""Default configuration: the fields of the reference's ``config.py`` that the hot path and its
caller (``Agent``) read, with ``cfg.filters`` naming this package's drop-in classes in the
reference's order (config.py:19-22).  A reference ``cfg`` object works just as well: only attribute
access is used."""
from __future__ import annotations


class Cfg(dict):
    """dict with attribute access (same behaviour as the reference's util.Dict, util.py:67-99)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def make_cfg(**overrides) -> Cfg:
    from . import filters as F

    cfg = Cfg()
    cfg.filters = [
        F.ExposureFilter, F.GammaFilter, F.CCMFilter, F.SharpenFilter, F.DenoiseFilter,
        F.ToneFilter, F.ContrastFilter, F.SaturationPlusFilter, F.WNBFilter, F.ImprovedWhiteBalanceFilter,
    ]
    cfg.filter_runtime_penalty = False
    cfg.filters_runtime = [1.7, 2.0, 1.9, 6.3, 10, 2.7, 2.1, 2.0, 1.9, 1.7]   # config.py:24
    cfg.filter_runtime_penalty_lambda = 0.01
    cfg.curve_steps = 8
    cfg.gamma_range = 3
    cfg.exposure_range = 3.5
    cfg.wb_range = 1.1
    cfg.color_curve_range = (0.90, 1.10)
    cfg.lab_curve_range = (0.90, 1.10)
    cfg.tone_curve_range = (0.5, 2)
    cfg.usm_sharpen_range = (0.0, 2.0)
    cfg.sharpen_range = (0.0, 10.0)
    cfg.ccm_range = (-2.0, 2.0)
    cfg.denoise_range = (0.0, 1.0)
    cfg.masking = False
    cfg.minimum_strength = 0.3
    cfg.maximum_sharpness = 1
    cfg.clamp = False
    cfg.filter_usage_penalty = 1.0
    cfg.img_include_states = True
    cfg.exploration = 0.05
    cfg.exploration_penalty = 0.05
    cfg.early_stop_penalty = 1.0
    cfg.base_channels = 32
    cfg.dropout_keep_prob = 0.5
    cfg.shared_feature_extractor = True
    cfg.fc1_size = 128
    cfg.feature_extractor_dims = 4096
    cfg.z_type = "uniform"
    cfg.z_dim_per_filter = 16
    cfg.test_steps = 5
    cfg.update(overrides)
    cfg.num_state_dim = 3 + len(cfg.filters)
    cfg.z_dim = 3 + len(cfg.filters) * cfg.z_dim_per_filter
    return cfg
