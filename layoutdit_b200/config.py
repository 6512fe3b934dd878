"""Offline description of the DiT (BEiT-style ViT) backbone architecture.

The reference hard-codes the hub id ``microsoft/dit-base``
(R:src/layoutdit/modeling/dit_backbone.py:25-26) and fetches a HuggingFace
``BeitConfig`` over the network.  There is no network on a B200 box, so the
drop-in takes this explicit struct instead.  Field names and defaults follow
``transformers`` ``BeitConfig`` (configuration_beit.py:72-102, transformers
5.5.0) so that ``DiTConfig(**hf_config.to_dict())``-style construction works.
"""
from __future__ import annotations

from dataclasses import dataclass, asdict


@dataclass(frozen=True)
class DiTConfig:
    hidden_size: int = 768
    num_hidden_layers: int = 12
    num_attention_heads: int = 12
    intermediate_size: int = 3072
    layer_norm_eps: float = 1e-12
    image_size: int = 224            # resolution the position table was trained at
    patch_size: int = 16
    num_channels: int = 3
    use_mask_token: bool = True      # DiT checkpoints carry a mask token
    use_absolute_position_embeddings: bool = True
    use_relative_position_bias: bool = False
    use_shared_relative_position_bias: bool = False
    layer_scale_init_value: float = 0.1
    initializer_range: float = 0.02
    hidden_act: str = "gelu"         # exact erf GELU (HF ACT2FN["gelu"])

    def __post_init__(self):
        if self.hidden_size % self.num_attention_heads:
            raise ValueError(
                f"The hidden size {self.hidden_size} is not a multiple of the number of "
                f"attention heads {self.num_attention_heads}.")
        if self.hidden_size // self.num_attention_heads != 64:
            raise ValueError("the sm_100a attention kernel is built for head_dim == 64 "
                             "(true for DiT-base and DiT-large)")
        if self.patch_size != 16 or self.num_channels != 3:
            raise ValueError("the patch-embed kernel is built for 16x16 patches over 3 channels")
        if self.hidden_act != "gelu":
            raise ValueError("only the exact erf GELU of the reference is implemented")
        if self.hidden_size % 64 or self.intermediate_size % 64:
            raise ValueError("hidden_size and intermediate_size must be multiples of 64")
        if self.use_relative_position_bias and self.use_shared_relative_position_bias:
            # HF allows both flags; the biases would simply add.  Not a DiT configuration.
            raise ValueError("choose per-layer or shared relative position bias, not both")

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_attention_heads

    @property
    def grid(self) -> int:
        return self.image_size // self.patch_size

    def to_dict(self) -> dict:
        return asdict(self)

    @classmethod
    def from_hf(cls, hf_config) -> "DiTConfig":
        """Build from a transformers ``BeitConfig`` (or anything with the same attributes)."""
        d = hf_config.to_dict() if hasattr(hf_config, "to_dict") else dict(hf_config)
        names = cls.__dataclass_fields__.keys()
        kw = {k: d[k] for k in names if k in d}
        for k in ("image_size", "patch_size"):
            if k in kw and not isinstance(kw[k], int):
                kw[k] = int(kw[k][0])
        return cls(**kw)


def dit_base(**kw) -> DiTConfig:
    """microsoft/dit-base: 12 layers, D=768, 12 heads, I=3072, absolute positions."""
    return DiTConfig(**kw)


def dit_large(**kw) -> DiTConfig:
    """microsoft/dit-large: 24 layers, D=1024, 16 heads, I=4096, absolute positions."""
    base = dict(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096)
    base.update(kw)
    return DiTConfig(**base)


def flops_per_image(cfg: DiTConfig, height: int, width: int) -> float:
    """Algorithmic FLOPs of one backbone forward: 2 flops per MAC,
    patch-embed + QKV + out-proj + MLP + QK^T/PV; no credit for padding, pooler excluded."""
    P = (height // cfg.patch_size) * (width // cfg.patch_size)
    N = P + 1
    D, I, L = cfg.hidden_size, cfg.intermediate_size, cfg.num_hidden_layers
    K0 = cfg.num_channels * cfg.patch_size * cfg.patch_size
    return 2.0 * P * K0 * D + L * (8.0 * N * D * D + 4.0 * N * D * I + 4.0 * N * N * D)
