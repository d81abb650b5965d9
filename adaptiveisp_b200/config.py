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


# name -> value, grouped as in the reference (config.py:24-86); values are the reference's defaults
_FILTER_FIELDS = dict(
    filter_runtime_penalty=False,
    filters_runtime=[1.7, 2.0, 1.9, 6.3, 10, 2.7, 2.1, 2.0, 1.9, 1.7],   # relative cost table, cfg.filters order
    filter_runtime_penalty_lambda=0.01,
    curve_steps=8, gamma_range=3, exposure_range=3.5, wb_range=1.1,
    color_curve_range=(0.90, 1.10), lab_curve_range=(0.90, 1.10), tone_curve_range=(0.5, 2),
    usm_sharpen_range=(0.0, 2.0), sharpen_range=(0.0, 10.0), ccm_range=(-2.0, 2.0), denoise_range=(0.0, 1.0),
    masking=False, minimum_strength=0.3, maximum_sharpness=1, clamp=False,
)
_RL_FIELDS = dict(
    filter_usage_penalty=1.0, img_include_states=True, exploration=0.05, exploration_penalty=0.05,
    early_stop_penalty=1.0, test_steps=5,
    replay_memory_size=128, maximum_trajectory_length=7, over_length_keep_prob=0.5,
)
_NET_FIELDS = dict(
    base_channels=32, dropout_keep_prob=0.5, shared_feature_extractor=True, fc1_size=128,
    feature_extractor_dims=4096, z_type="uniform", z_dim_per_filter=16,
)


def make_cfg(**overrides) -> Cfg:
    from . import filters as F

    cfg = Cfg(filters=[F.ExposureFilter, F.GammaFilter, F.CCMFilter, F.SharpenFilter, F.DenoiseFilter,
                       F.ToneFilter, F.ContrastFilter, F.SaturationPlusFilter, F.WNBFilter,
                       F.ImprovedWhiteBalanceFilter])
    for group in (_FILTER_FIELDS, _RL_FIELDS, _NET_FIELDS):
        cfg.update({k: (list(v) if isinstance(v, list) else v) for k, v in group.items()})
    cfg.update(overrides)
    # derived sizes (config.py:85-86)
    cfg.num_state_dim = 3 + len(cfg.filters)
    cfg.z_dim = 3 + len(cfg.filters) * cfg.z_dim_per_filter
    return cfg
