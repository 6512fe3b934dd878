"""Drop-in replacement for the reference's ``DiTWithFPN``.

Boundary being mirrored: R:src/layoutdit/modeling/dit_backbone.py:65-95 --
``DiTWithFPN(pretrained=True)`` owns ``.backbone`` (a ``DiTBackbone``), ``.fpn`` (torchvision
``FeaturePyramidNetwork([D]*4, 256, extra_blocks=LastLevelMaxPool())``) and ``.out_channels = 256``;
``forward(x[B,3,H,W])`` returns an ``OrderedDict`` with keys ``p2, p3, p4, p5, pool`` -- the names
``MultiScaleRoIAlign`` is configured with at R:model.py:34-38,63.  ``state_dict()`` keys are the
reference's: ``backbone.dit.<HF BeitModel names>`` and ``fpn.inner_blocks.{i}.0.{weight,bias}``,
``fpn.layer_blocks.{i}.0.{weight,bias}`` (TV = torchvision/ops/feature_pyramid_network.py:104-124).

The forward is one launch sequence in libldit_b200 (``Engine._plan`` with ``head="fpn"``): the backbone
kernels, then per tapped layer a 1x1-lateral GEMM on the token grid (before resampling -- the D-channel
taps of ``DiTBackbone`` are never written), a top-down merge kernel per level and the 3x3 output
convolutions as implicit tcgen05 GEMMs.  Same platform rules as ``DiTBackbone``: CUDA only, bf16
activations, inference only, no CPU or eager fallback.
"""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn as nn

from .config import DiTConfig
from .dit_backbone import DiTBackbone


class FPNParameters(nn.Module):
    """Parameter holder with torchvision ``FeaturePyramidNetwork`` names; ``forward`` is never called."""

    def __init__(self, in_channels_list, out_channels: int = 256):
        super().__init__()
        if out_channels % 128:
            raise ValueError("the sm_100a convolution kernel needs out_channels to be a multiple of 128")
        self.out_channels = out_channels
        # Conv2dNormActivation(norm_layer=None, activation_layer=None) == Sequential(Conv2d)  (TV:111-124)
        self.inner_blocks = nn.ModuleList(nn.Sequential(nn.Conv2d(c, out_channels, 1)) for c in in_channels_list)
        self.layer_blocks = nn.ModuleList(nn.Sequential(nn.Conv2d(out_channels, out_channels, 3, padding=1))
                                          for _ in in_channels_list)
        self.reset_parameters()

    @torch.no_grad()
    def reset_parameters(self):
        for m in self.modules():   # TV:126-131
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_uniform_(m.weight, a=1)
                nn.init.constant_(m.bias, 0)

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter holder: the FPN runs in the sm_100a kernels (DiTWithFPN.forward)")


class DiTWithFPN(nn.Module):
    """DiT backbone + FPN (p2..p5 + pool), the module ``FasterRCNN`` is given at R:model.py:44-56."""

    def __init__(self, pretrained: bool = True, config: DiTConfig | None = None, state_dict: dict | None = None,
                 fpn_state_dict: dict | None = None, use_cuda_graph: bool = False, out_dtype: torch.dtype = torch.bfloat16):
        """``out_dtype``: dtype of the returned maps.  bf16 (default) is what the kernels compute in; ``torch.float32``
        makes the 3x3 output convolutions store fp32 (same arithmetic, the cast is the epilogue's store format), which is
        what torchvision's fp32 RPN / RoI heads need when this module is the ``backbone`` of ``FasterRCNN``
        (R:model.py:44-56) outside autocast -- no user-side cast."""
        super().__init__()
        if out_dtype not in (torch.bfloat16, torch.float32):
            raise ValueError("out_dtype must be torch.bfloat16 or torch.float32")
        self.out_dtype = out_dtype
        self.backbone = DiTBackbone(pretrained=pretrained, config=config, state_dict=state_dict)
        in_channels = [self.backbone.hidden_size] * 4          # R:dit_backbone.py:79
        self.fpn = FPNParameters(in_channels, out_channels=256)  # R:dit_backbone.py:80-84
        if fpn_state_dict is not None:
            self.fpn.load_state_dict(fpn_state_dict, strict=True)
        self.out_channels = 256
        self.use_cuda_graph = use_cuda_graph

    def _head(self) -> str:
        return "fpn32" if self.out_dtype == torch.float32 else "fpn"

    def _engine(self):
        eng = self.backbone._get_engine()
        if eng.fpn_params is not self.fpn:
            eng.fpn_params = self.fpn
            eng._pack_key = None
        return eng

    def forward_pages(self, pages, size=(224, 224), mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5), fixed_size=None):
        """Raw pages -> FPN maps: what ``FasterRCNN.forward`` computes up to its RPN (transform + backbone,
        torchvision generalized_rcnn.py), with the transform fused into the patch gather."""
        with torch.no_grad():
            return self._engine().forward_pages(pages, size, mean, std, self._head(), fixed_size)

    def forward(self, x: torch.Tensor) -> "OrderedDict[str, torch.Tensor]":
        if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError(
                "layoutdit_b200.DiTWithFPN implements the inference forward only; call .eval() or wrap "
                "the call in torch.no_grad()")
        eng = self._engine()
        with torch.no_grad():
            return eng.forward_graphed(x, 0, self._head()) if self.use_cuda_graph else eng.forward(x, self._head())
