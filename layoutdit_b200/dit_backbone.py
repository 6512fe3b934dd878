"""Drop-in replacement for the reference's ``DiTBackbone`` module.

Boundary being mirrored:

* class / ctor: ``DiTBackbone(pretrained: bool = True)``, R:src/layoutdit/modeling/dit_backbone.py:23,
  constructed only by ``DiTWithFPN.__init__`` (R:dit_backbone.py:72);
* attributes read by callers: ``.hidden_size`` (R:dit_backbone.py:79), ``.dit`` with
  ``load_state_dict(sd, strict=False)`` over HF ``BeitModel`` keys (R:model.py:70),
  ``.layer_idxs`` / ``.scales`` (R:dit_backbone.py:34-35);
* ``forward(x[B,3,H,W]) -> OrderedDict{"p2","p3","p4","p5"}`` of ``[B, D, h_i, w_i]``
  (R:dit_backbone.py:38-62), consumed in key order by torchvision's FPN.

Differences, all forced by the platform and stated in DESIGN.md: the architecture comes from
an explicit ``DiTConfig`` (the hub is unreachable); the module is CUDA-only with no CPU or
eager fallback; activations are bf16 with an fp32 residual stream and the taps are returned
in bf16; the forward is inference-only (no autograd graph is recorded).
"""
from __future__ import annotations

import os
import warnings
from collections import OrderedDict

import torch
import torch.nn as nn

from .config import DiTConfig, dit_base
from .dit_params import DiTParameters
from .engine import Engine, TAP_SCALES, tap_layer_indices


class DiTBackbone(nn.Module):
    """4-scale DiT feature extractor: hidden states after layers d/3, d/2, 2d/3, d,
    resampled by [4x, 2x, 1x, 0.5x]."""

    def __init__(self, pretrained: bool = True, config: DiTConfig | None = None,
                 state_dict: dict | None = None, use_cuda_graph: bool = False):
        super().__init__()
        config = config if config is not None else dit_base()
        if not isinstance(config, DiTConfig):
            config = DiTConfig.from_hf(config)
        self.config = config
        self.dit = DiTParameters(config)
        # pretrained=True: the reference downloads microsoft/dit-base here (R:dit_backbone.py:27-29).  There is
        # no network on the target boxes, so the weights must come from the caller: `state_dict=` (HF BeitModel
        # keys, bare or prefixed), `pretrained="/path/to/checkpoint"`, the environment variable
        # LDIT_PRETRAINED_PATH, or `.dit.load_state_dict(...)` afterwards exactly as R:model.py:65-70 does.
        # Asking for pretrained weights and getting random ones silently would be the worst outcome, hence the warning.
        from . import checkpoint as _ckpt
        path = pretrained if isinstance(pretrained, (str, os.PathLike)) else None
        if state_dict is not None:
            _ckpt.load_dit_state_dict(self.dit, _ckpt.split_checkpoint(state_dict).dit, strict=True)
            pretrained = False
        elif pretrained:
            path = path or os.environ.get("LDIT_PRETRAINED_PATH")
            if path:
                _ckpt.load_dit_state_dict(self.dit, _ckpt.split_checkpoint(_ckpt.read_state_dict(path)).dit, strict=True)
                pretrained = False
            else:
                warnings.warn("DiTBackbone(pretrained=True): no weights were given (state_dict=, pretrained=<path> or "
                              "LDIT_PRETRAINED_PATH) and the hub is unreachable -- the module holds RANDOM-INIT weights "
                              "until .dit.load_state_dict(...) is called", RuntimeWarning, stacklevel=2)
        self.pretrained = bool(pretrained)   # True = still waiting for real weights
        d = config.num_hidden_layers
        self.layer_idxs = tap_layer_indices(d)
        self.scales = list(TAP_SCALES)
        self.hidden_size = config.hidden_size
        self.use_cuda_graph = use_cuda_graph
        self._engine: Engine | None = None

    def _get_engine(self) -> Engine:
        if self._engine is None:
            self._engine = Engine(self.dit, self.config)
        return self._engine

    def _apply(self, fn, *a, **k):  # .to() / .cuda() / .half(): packed weights are stale afterwards
        out = super()._apply(fn, *a, **k)
        if self._engine is not None:
            self._engine._pack_key = None
        return out

    def forward_host(self, pages: torch.Tensor, result_host: torch.Tensor | None = None, result_key: str = "p5"):
        """Pipelined forward of a HOST batch (see ``Engine.forward_host``): PCIe copies overlap the
        kernels of the neighbouring calls.  Returns ``(feats, done_event)``."""
        with torch.no_grad():
            return self._get_engine().forward_host(pages, result_host, result_key)

    def forward_pages(self, pages, size=(224, 224), mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5), fixed_size=None):
        """Raw pages (list of ``[3, H_i, W_i]`` float CUDA tensors in [0, 1]) -> taps, with the detector's input
        transform (R:model.py:44-56) fused into the patch gather; see ``Engine.forward_pages``."""
        with torch.no_grad():
            return self._get_engine().forward_pages(pages, size, mean, std, "taps", fixed_size)

    def forward(self, x: torch.Tensor) -> "OrderedDict[str, torch.Tensor]":
        if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.dit.parameters()):
            raise NotImplementedError(
                "layoutdit_b200.DiTBackbone implements the inference forward only; call .eval() or wrap "
                "the call in torch.no_grad() (the training backward is listed under 'next' in DESIGN.md)")
        eng = self._get_engine()
        with torch.no_grad():
            return eng.forward_graphed(x) if self.use_cuda_graph else eng.forward(x)
