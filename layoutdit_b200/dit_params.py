"""Parameter tree of the backbone with HuggingFace ``BeitModel`` names.

The reference resumes by calling ``backbone.backbone.dit.load_state_dict(sd, strict=False)``
(R:src/layoutdit/modeling/model.py:65-70) and checkpoints the whole model's
``state_dict()`` (R:model.py:90-121), so ``DiTBackbone.dit`` must be an ``nn.Module`` whose
keys are exactly those of ``transformers`` ``BeitModel``.  The
sub-modules below exist only to own parameters under the right names; their ``forward`` is
never called -- the arithmetic runs in libldit_b200.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .config import DiTConfig


class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter holder: the forward pass runs in the sm_100a kernels")


class _RelPosBias(_Holder):
    def __init__(self, cfg: DiTConfig):
        super().__init__()
        g = cfg.grid
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * g - 1) * (2 * g - 1) + 3, cfg.num_attention_heads))


class _SelfAttention(_Holder):
    def __init__(self, cfg: DiTConfig):
        super().__init__()
        D = cfg.hidden_size
        self.query = nn.Linear(D, D)
        self.key = nn.Linear(D, D, bias=False)       # HF:240
        self.value = nn.Linear(D, D)
        if cfg.use_relative_position_bias:
            self.relative_position_bias = _RelPosBias(cfg)


class _SelfOutput(_Holder):
    def __init__(self, cfg: DiTConfig):
        super().__init__()
        self.dense = nn.Linear(cfg.hidden_size, cfg.hidden_size)


class _Attention(_Holder):
    def __init__(self, cfg: DiTConfig):
        super().__init__()
        self.attention = _SelfAttention(cfg)
        self.output = _SelfOutput(cfg)


class _Dense(_Holder):
    def __init__(self, n_in: int, n_out: int):
        super().__init__()
        self.dense = nn.Linear(n_in, n_out)


class _Layer(_Holder):
    def __init__(self, cfg: DiTConfig):
        super().__init__()
        D, I = cfg.hidden_size, cfg.intermediate_size
        self.attention = _Attention(cfg)
        self.intermediate = _Dense(D, I)
        self.output = _Dense(I, D)
        self.layernorm_before = nn.LayerNorm(D, eps=cfg.layer_norm_eps)
        self.layernorm_after = nn.LayerNorm(D, eps=cfg.layer_norm_eps)
        if cfg.layer_scale_init_value > 0:           # HF:462-467
            self.lambda_1 = nn.Parameter(cfg.layer_scale_init_value * torch.ones(D))
            self.lambda_2 = nn.Parameter(cfg.layer_scale_init_value * torch.ones(D))
        else:
            self.lambda_1, self.lambda_2 = None, None


class _PatchEmbeddings(_Holder):
    def __init__(self, cfg: DiTConfig):
        super().__init__()
        self.projection = nn.Conv2d(cfg.num_channels, cfg.hidden_size, kernel_size=cfg.patch_size, stride=cfg.patch_size)


class _Embeddings(_Holder):
    def __init__(self, cfg: DiTConfig):
        super().__init__()
        D = cfg.hidden_size
        self.cls_token = nn.Parameter(torch.zeros(1, 1, D))
        if cfg.use_mask_token:
            self.mask_token = nn.Parameter(torch.zeros(1, 1, D))
        else:
            self.mask_token = None
        self.patch_embeddings = _PatchEmbeddings(cfg)
        if cfg.use_absolute_position_embeddings:
            self.position_embeddings = nn.Parameter(torch.zeros(1, cfg.grid * cfg.grid + 1, D))
        else:
            self.position_embeddings = None


class _Encoder(_Holder):
    def __init__(self, cfg: DiTConfig):
        super().__init__()
        if cfg.use_shared_relative_position_bias:
            self.relative_position_bias = _RelPosBias(cfg)
        self.layer = nn.ModuleList([_Layer(cfg) for _ in range(cfg.num_hidden_layers)])


class _Pooler(_Holder):
    """Never evaluated (the reference reads only ``.hidden_states``, R:dit_backbone.py:47);
    kept so checkpoints load with ``strict=True``."""

    def __init__(self, cfg: DiTConfig):
        super().__init__()
        self.layernorm = nn.LayerNorm(cfg.hidden_size, eps=cfg.layer_norm_eps)


class DiTParameters(_Holder):
    """``state_dict()``-compatible with ``transformers.BeitModel`` (use_mean_pooling=True)."""

    # old BEiT checkpoints carry these buffers; HF ignores them on load (HF:674)
    _ignored_suffix = "relative_position_index"

    def __init__(self, cfg: DiTConfig):
        super().__init__()
        self.config = cfg
        self.embeddings = _Embeddings(cfg)
        self.encoder = _Encoder(cfg)
        self.pooler = _Pooler(cfg)
        self.reset_parameters()

    @torch.no_grad()
    def reset_parameters(self):
        """HF ``_init_weights`` (HF:677-692): N(0, initializer_range) matrices, zero biases,
        LayerNorm (1, 0), zero cls / mask / position / tables, layer-scale = init value."""
        std = self.config.initializer_range
        for m in self.modules():
            if isinstance(m, (nn.Linear, nn.Conv2d)):
                m.weight.normal_(0.0, std)
                if m.bias is not None:
                    m.bias.zero_()
            elif isinstance(m, nn.LayerNorm):
                m.weight.fill_(1.0)
                m.bias.zero_()
            elif isinstance(m, _RelPosBias):
                m.relative_position_bias_table.zero_()
            elif isinstance(m, _Layer) and m.lambda_1 is not None:
                m.lambda_1.fill_(self.config.layer_scale_init_value)
                m.lambda_2.fill_(self.config.layer_scale_init_value)
        e = self.embeddings
        e.cls_token.zero_()
        if e.mask_token is not None:
            e.mask_token.zero_()
        if e.position_embeddings is not None:
            e.position_embeddings.zero_()

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        state_dict = {k: v for k, v in state_dict.items() if not k.endswith(self._ignored_suffix)}
        return super().load_state_dict(state_dict, strict=strict, assign=assign)
