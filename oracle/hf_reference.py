"""The reference's CPU path, rebuilt from the installed ``transformers`` -- TEST / BASELINE
INFRASTRUCTURE, NOT PRODUCT CODE (same import rules as ``oracle/dit_oracle.py``).

``/root/reference`` does not exist on the GPU box, but the library that does all of the
reference's arithmetic (``transformers.BeitModel``) is part of the image.  This module
wraps it exactly the way R:src/layoutdit/modeling/dit_backbone.py:23-62 does, with the one
network call (``AutoConfig.from_pretrained("microsoft/dit-base")``, :25-26) replaced by an
explicit ``BeitConfig``.  It is used

* to time the reference's PyTorch-eager CPU forward on the box's host cores
  (``bench.py`` ``cpu_baseline`` and ``--impl reference``), and
* as a second pin for ``oracle/dit_oracle.py`` in the CPU test-suite.

``oracle/make_golden.py`` checks (in the build container, where /root/reference exists)
that this wrapper and the reference's own class produce identical tensors.
"""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F


def hf_config(cfg_dict: dict):
    from transformers import BeitConfig
    keys = ["hidden_size", "num_hidden_layers", "num_attention_heads", "intermediate_size",
            "layer_norm_eps", "image_size", "patch_size", "num_channels", "use_mask_token",
            "use_absolute_position_embeddings", "use_relative_position_bias",
            "use_shared_relative_position_bias", "layer_scale_init_value", "initializer_range",
            "hidden_act"]
    kw = {k: cfg_dict[k] for k in keys if k in cfg_dict}
    return BeitConfig(output_hidden_states=True, **kw)


class HFDiTBackbone(nn.Module):
    """Line-for-line behavioural twin of the reference ``DiTBackbone`` built offline."""

    def __init__(self, cfg_dict: dict):
        super().__init__()
        from transformers import AutoModel
        config = hf_config(cfg_dict)
        self.dit = AutoModel.from_config(config)          # R:dit_backbone.py:27-31 (pretrained=False arm)
        d = config.num_hidden_layers
        self.layer_idxs = [d // 3, d // 2, 2 * d // 3, d]  # R:dit_backbone.py:33-34
        self.scales = [4.0, 2.0, 1.0, 0.5]                 # R:dit_backbone.py:35
        self.hidden_size = config.hidden_size              # R:dit_backbone.py:36

    def forward(self, x):
        B, _, H, W = x.shape
        Gh, Gw = H // 16, W // 16
        hs = self.dit(x).hidden_states                     # R:dit_backbone.py:47
        feats = OrderedDict()
        for i, (idx, scale) in enumerate(zip(self.layer_idxs, self.scales), start=2):
            t = hs[idx][:, 1:, :].permute(0, 2, 1).view(B, self.hidden_size, Gh, Gw)
            if scale != 1.0:
                t = F.interpolate(t, scale_factor=scale, mode="bilinear", align_corners=False)
            feats[f"p{i}"] = t
        return feats


def build(cfg_dict: dict, state_dict: dict | None = None) -> HFDiTBackbone:
    m = HFDiTBackbone(cfg_dict).eval()
    if state_dict is not None:
        m.dit.load_state_dict(state_dict, strict=True)
    return m
