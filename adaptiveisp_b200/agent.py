"""Caller of the hot path: the actor's filter bank with select-first dispatch.

The reference's ``Agent.forward`` (agent.py:88-285) runs ALL filters on the whole batch, stacks
``[B,10,3,H,W]``, and keeps one of ten with a one-hot multiply-sum (agent.py:103-116,154).  Here the
selection is computed first and ONE heterogeneous apply runs only the selected filter of each
sample (``aisp_select_apply_*``): identical values for finite inputs (the one-hot sum is bitwise a
pick), 1/10 of the arithmetic and none of the 10x stack traffic.

The small conv/FC nets are the reference's architecture re-instantiated in PyTorch with the SAME
parameter names, so ``load_state_dict(ckpt['agent_model'])`` of a reference checkpoint works
strictly.  They are not the product; the filter application is.

Known, documented deviation: in the reference a NaN/inf produced by an UNselected filter (e.g. a CCM
row summing to 0, isp/filters.py:707) poisons the selected result through ``0 * NaN``
(agent.py:154); the select-first form never evaluates unselected filters.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn

from . import functional as AF

STATE_REWARD_DIM, STATE_STOPPED_DIM, STATE_STEP_DIM, STATE_DROPOUT_BEGIN = 0, 1, 2, 3  # util.py:15-18


def enrich_image_input(cfg, net, states):
    """util.py:58-63: broadcast the state vector into extra image planes."""
    if cfg.img_include_states:
        planes = states[:, :, None, None] + (net[:, 0:1, :, :] * 0)
        net = torch.cat([net, planes], dim=1)
    return net


def pdf_sample(pdf, uniform_noise):
    """agent.py:12-16.  The 10-entry cumsum must stay sequential fp32 (torch.cumsum) to keep the
    knife-edge compare ``cdf < u`` bit-identical to the reference."""
    pdf = pdf / (torch.sum(pdf, dim=1, keepdim=True) + 1e-36)
    cdf = torch.cumsum(pdf, dim=1) - pdf
    return torch.sum(torch.less(cdf, uniform_noise).to(torch.int32), dim=1) - 1


def one_hot(num_class, index):
    """agent.py:18-23 without its ten boolean-index host syncs; index -1 gives an all-zero row, as
    in the reference."""
    classes = torch.arange(num_class, device=index.device, dtype=index.dtype)
    return (index[:, None] == classes[None, :]).to(torch.int64)


class FeatureExtractor(nn.Module):
    """agent.py:26-60: strided 4x4 conv + BN + LeakyReLU pyramid down to 4x4, then dropout."""

    def __init__(self, shape=(14, 64, 64), mid_channels=32, output_dim=4096, dropout_prob=0.5):
        super().__init__()
        self.output_dim = output_dim
        floor = 4
        assert output_dim % (floor * floor) == 0, "output dim=%d" % output_dim
        size = int(shape[2]) // 2
        cin, cout = shape[0], mid_channels
        blocks = [nn.Conv2d(cin, cout, kernel_size=4, stride=2, padding=1), nn.BatchNorm2d(cout),
                  nn.LeakyReLU(negative_slope=0.2)]
        while size > floor:
            cin = cout
            cout = output_dim // (floor * floor) if size == 2 * floor else cout * 2
            assert size % 2 == 0
            size //= 2
            blocks += [nn.Conv2d(cin, cout, kernel_size=4, stride=2, padding=1), nn.BatchNorm2d(cout),
                       nn.LeakyReLU(negative_slope=0.2)]
        self.layers = nn.Sequential(*blocks)
        self.droupout = nn.Dropout(p=dropout_prob)  # (sic) attribute name as in the reference

    def forward(self, x):
        return self.droupout(torch.reshape(self.layers(x), [-1, self.output_dim]))


class Agent(nn.Module):
    """Drop-in for the reference ``Agent`` (same constructor, inputs, outputs and state_dict keys)."""

    def __init__(self, cfg, shape=(16, 64, 64), device="cuda"):
        super().__init__()
        self.cfg = cfg
        self.feature_extractor = FeatureExtractor(shape=shape, mid_channels=cfg.base_channels,
                                                  output_dim=cfg.feature_extractor_dims,
                                                  dropout_prob=1.0 - cfg.dropout_keep_prob)
        self.filters = []
        for cls in cfg.filters:
            flt = cls(cfg, predict=True).to(device)
            self.__setattr__(flt.get_short_name(), flt)
            self.filters.append(flt)
        self.action_selection = FeatureExtractor(shape=shape, mid_channels=cfg.base_channels,
                                                 output_dim=cfg.feature_extractor_dims,
                                                 dropout_prob=1.0 - cfg.dropout_keep_prob)
        self.fc1 = nn.Linear(cfg.feature_extractor_dims, cfg.fc1_size)
        self.lrelu = nn.LeakyReLU(negative_slope=0.2)
        self.fc2 = nn.Linear(cfg.fc1_size, len(self.filters))
        self.down_sample = nn.AdaptiveAvgPool2d((shape[1], shape[2]))
        self.runtime = torch.tensor(cfg.filters_runtime, requires_grad=False).to(device)
        # filter index -> aisp_op code, as a device-side lookup table (no host sync per step)
        self.register_buffer("_op_table", torch.tensor([f.OP for f in self.filters], dtype=torch.int32),
                             persistent=False)
        self._predictor = None

    # ------------------------------------------------------------------------------------------
    def downsample(self, x):
        """agent.py:97.  An image that left this Agent carries the block means its kernels emitted from
        their store path (``_aisp_down``): the next step / the critic reuse them instead of re-reading
        the full-resolution image.  Otherwise evenly dividing CUDA inputs that carry no gradient (the
        training case, train.py:255) take the one-pass block-mean kernel; everything else the PyTorch
        module."""
        oh, ow = self.down_sample.output_size
        cached = getattr(x, "_aisp_down", None)
        if cached is not None and cached.shape == (x.shape[0], 3, oh, ow) and cached.device == x.device:
            return cached
        if x.is_cuda and not x.requires_grad and x.dtype == torch.float32 and x.is_contiguous() \
                and x.shape[2] % oh == 0 and x.shape[3] % ow == 0 and x.shape[0] * 3 <= 65535:
            return AF.block_mean(x, (oh, ow))
        return self.down_sample(x)

    def predict_all_params(self, filter_features, batched=True):
        """Every filter's regressed parameters (tiny ``[B,n]`` tensors) packed as ``[B,F,PSTRIDE]``, plus the
        per-filter tensors in the reference's layouts.  ``batched`` (default, CUDA): one GEMM for the ten
        ``fc1`` layers + one regressor kernel (``filters.BankPredictor``); otherwise the per-module statement."""
        if batched and filter_features.is_cuda:
            if self._predictor is None:
                from .filters import BankPredictor
                self._predictor = BankPredictor(self.filters)
            if self._predictor.usable(filter_features):
                packed = self._predictor(filter_features)
                return packed, self._predictor.split(packed)
        rows, per_filter = [], []
        for flt in self.filters:
            feats, _ = flt.extract_parameters(filter_features)
            p = flt.filter_param_regressor(feats)
            per_filter.append(p)
            rows.append(AF.pack_params(p, flt.get_num_filter_parameters()))
        return torch.stack(rows, dim=1), per_filter

    def policy(self, x_down, states):
        """agent.py:119-133 -> (pdf [B,F] with the exploration mix, entropy [B,1]); stays PyTorch
        (the policy gradient flows through both)."""
        n = len(self.filters)
        feats = self.action_selection(enrich_image_input(self.cfg, x_down, states))
        feats = self.lrelu(self.fc1(feats))
        pdf = torch.softmax(self.fc2(feats), dim=1) + 1e-37
        pdf = pdf * (1 - self.cfg.exploration) + self.cfg.exploration * 1.0 / n
        pdf = pdf / (torch.sum(pdf, dim=1, keepdim=True) + 1e-30)
        entropy = torch.sum(-pdf * torch.log(pdf), dim=1)[:, None]
        return pdf, entropy

    def select(self, x_down, states, selection_noise, selected_filter_id=None):
        """agent.py:119-149 -> (pdf, entropy[B,1], selected int64 [B], one_hot int64 [B,F]) as PyTorch
        ops (kept as the readable statement of the selection; ``forward`` runs ``aisp_select``)."""
        n = len(self.filters)
        pdf, entropy = self.policy(x_down, states)
        if selected_filter_id is not None:
            sel = torch.full((pdf.shape[0],), int(selected_filter_id), dtype=torch.int64, device=pdf.device)
        elif self.training:
            sel = pdf_sample(pdf, selection_noise).to(torch.int64)
        else:
            sel = torch.argmax(pdf, dim=1).to(torch.int32).to(torch.int64)
        return pdf, entropy, sel, one_hot(n, sel)

    def apply_selected(self, x, packed_all, sel):
        """Gather the selected filter's parameter row per sample and run ONE heterogeneous apply."""
        n = len(self.filters)
        valid = sel >= 0
        safe = torch.clamp(sel, 0, n - 1)
        rows = torch.gather(packed_all, 1, safe[:, None, None].expand(-1, 1, packed_all.shape[2]))[:, 0, :]
        ops = torch.where(valid, self._op_table.to(x.device)[safe], torch.full_like(safe, -1, dtype=torch.int32))
        rows = rows * valid[:, None].to(rows.dtype)
        return AF.apply_ops(x, rows, ops, clip=True, family=None), rows, ops

    def forward(self, inp, progress, high_res=None, selected_filter_id=None):
        x, z, states = inp
        selection_noise = z[:, 0:1]
        x_down = self.downsample(x)
        if not self.cfg.shared_feature_extractor:
            raise ValueError("current just support shared_feature_extractor")
        filter_features = self.feature_extractor(enrich_image_input(self.cfg, x_down, states))
        packed_all, per_filter = self.predict_all_params(filter_features)
        pdf, entropy = self.policy(x_down, states)
        # selection, one-hot, parameter-row gather and the state update: ONE launch, no host sync
        if selected_filter_id is not None:
            mode, forced = AF.SELECT_FORCED, int(selected_filter_id)
        else:
            mode, forced = (AF.SELECT_SAMPLE if self.training else AF.SELECT_ARGMAX), 0
        rows, sel, hot, ops, new_states, pens = AF.select_rows(
            pdf, selection_noise, states.to(torch.float32), packed_all, self._op_table.to(x.device), mode, forced,
            float(self.cfg.test_steps), float(self.cfg.early_stop_penalty))
        surrogate = torch.sum(hot * torch.log(pdf + 1e-10), dim=1, keepdim=True)

        # ONE launch set applies the selected filter of every sample to the batch AND to its full-size
        # twin (agent.py:155-157) and emits the 64x64 block means of the result from the store path (what
        # the next step and the critic pool first, agent.py:97 / value.py:63)
        x_in = x
        oh, ow = self.down_sample.output_size
        even = x_in.shape[2] % oh == 0 and x_in.shape[3] % ow == 0 and x_in.shape[0] * 3 <= 65535
        if high_res is not None or even:
            x, high_res_output, x_out_down = AF.apply_ops(x_in, rows, ops, clip=True, family=None, high_res=high_res,
                                                          down_hw=(oh, ow) if even else None)
        else:
            x, high_res_output, x_out_down = AF.apply_ops(x_in, rows, ops, clip=True, family=None), None, None

        ones = self.filters[0].get_mask(x_in)
        filter_debug_info = [{"filter_parameters": f._debug(p), "mask": ones[0]}
                             for f, p in zip(self.filters, per_filter)]
        debug_info = {
            "state": states,
            "selected_filter_id": sel[0],
            "filter_debug_info": filter_debug_info,
            "pdf": pdf[0],
            "selected_filter": sel,
        }

        def debugger(info, combined=True):
            canvas = np.full((64, 64, 3), 0.8, dtype=np.float32)
            i = int(info["selected_filter_id"])
            if 0 <= i < len(self.filters):
                self.filters[i].visualize_filter(info["filter_debug_info"][i], canvas)
            return canvas if combined else [canvas.copy(), canvas.copy(), canvas.copy()]

        debugger.width = int(x.shape[2])

        # new states and the two state-dependent penalties (agent.py:234-259) came out of aisp_select
        usage_penalty, early_stop_penalty = pens[:, 0:1], pens[:, 1:2]

        if self.cfg.clamp:
            x = torch.clip(x, min=0.0, max=5.0)       # a no-op on values: every kernel already clipped to [0,1]
        if x_out_down is not None:
            x._aisp_down = x_out_down                  # travels with the tensor object (see downsample)
        entropy_penalty = (1.0 - progress) * self.cfg.exploration_penalty * (-entropy + math.log(len(self.filters)))
        runtime_penalty = 0.0
        if self.cfg.filter_runtime_penalty:
            runtime_penalty = torch.sum(hot * self.runtime.to(hot.device), dim=1, keepdim=True)
            runtime_penalty = self.cfg.filter_runtime_penalty_lambda * runtime_penalty
        # agent.py:276: mean(relu(x-1)^2) is identically 0 here -- every selected output was clipped to
        # [0,1] inside the kernel -- so the extra full-image pass of the reference is skipped.
        over_range = torch.zeros((x.shape[0], 1), dtype=x.dtype, device=x.device)
        penalty = over_range + entropy_penalty + usage_penalty * self.cfg.filter_usage_penalty + \
            early_stop_penalty + runtime_penalty

        if high_res is None:
            return (x, new_states, surrogate, penalty), debug_info, debugger
        return (x, new_states, high_res_output), debug_info, debugger
