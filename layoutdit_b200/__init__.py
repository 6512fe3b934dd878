"""B200-native DiT backbone forward (drop-in for LayoutDiT's ``DiTBackbone``)."""
from .config import DiTConfig, dit_base, dit_large, flops_per_image  # noqa: F401

__all__ = ["DiTConfig", "dit_base", "dit_large", "flops_per_image", "DiTBackbone", "DiTWithFPN"]


def __getattr__(name):
    if name == "DiTBackbone":
        from .dit_backbone import DiTBackbone
        return DiTBackbone
    if name == "DiTWithFPN":
        from .dit_fpn import DiTWithFPN
        return DiTWithFPN
    raise AttributeError(name)
