"""Checkpoint I/O next to the hot path.

On-disk formats this reads and writes, and where they come from:

* HuggingFace DiT / BEiT checkpoints (``microsoft/dit-base|large``): ``pytorch_model.bin`` or
  ``model.safetensors`` whose keys are ``BeitModel`` names, bare or under the ``beit.`` prefix of the
  ``BeitForMaskedImageModeling`` / ``BeitForImageClassification`` heads they were exported from (the
  reference loads them through ``AutoModel.from_pretrained``, R:src/layoutdit/modeling/dit_backbone.py:27-29,
  which strips that prefix and drops the head);
* LayoutDiT ``.pth`` files written by ``LayoutDetectionModel.save_checkpoint_to_gcs``
  (R:src/layoutdit/modeling/model.py:90-121): ``torch.save(self.state_dict())`` of the whole detector, so the
  backbone sits under ``model.backbone.backbone.dit.`` and the FPN under ``model.backbone.fpn.``
  (R:model.py:45-46, R:dit_backbone.py:72,80).

The reference resumes with ``backbone.backbone.dit.load_state_dict(state_dict, strict=False)``
(R:model.py:65-70); ``DiTBackbone.dit`` keeps exactly that ``nn.Module`` behaviour.  The helpers here do the
prefix bookkeeping that call leaves to the user, infer the architecture from the tensors (there is no hub to
ask for ``config.json``) and round-trip ``state_dict()`` to disk.  Pure host code: nothing here touches the GPU.
"""
from __future__ import annotations

import os
from collections import OrderedDict
from dataclasses import dataclass, field

import torch

from .config import DiTConfig

# most specific first; "" last
DIT_PREFIXES = ("model.backbone.backbone.dit.", "backbone.backbone.dit.", "backbone.dit.", "dit.", "beit.", "")
FPN_PREFIXES = ("model.backbone.fpn.", "backbone.fpn.", "fpn.")
_DIT_ROOTS = ("embeddings.", "encoder.", "pooler.")
_IGNORED_SUFFIX = "relative_position_index"   # buffers of old BEiT checkpoints, ignored by HF on load (HF:674)


@dataclass
class SplitCheckpoint:
    dit: "OrderedDict[str, torch.Tensor]"            # HF BeitModel names
    fpn: "OrderedDict[str, torch.Tensor]"            # torchvision FeaturePyramidNetwork names
    other: list = field(default_factory=list)         # keys that belong to neither (heads, RPN, optimizer, ...)
    dit_prefix: str = ""
    fpn_prefix: str | None = None


def read_state_dict(path: str, map_location="cpu") -> dict:
    """``.safetensors`` through safetensors, anything else through ``torch.load(weights_only=True)``.
    A ``{"state_dict": ...}`` / ``{"model": ...}`` wrapper is unwrapped."""
    if str(path).endswith(".safetensors"):
        from safetensors.torch import load_file
        return load_file(path, device=str(map_location))
    obj = torch.load(path, map_location=map_location, weights_only=True)
    for k in ("state_dict", "model", "model_state_dict"):
        if isinstance(obj, dict) and k in obj and isinstance(obj[k], dict):
            obj = obj[k]
    if not isinstance(obj, dict):
        raise ValueError(f"{path}: expected a state dict, found {type(obj).__name__}")
    return obj


def split_checkpoint(sd: dict) -> SplitCheckpoint:
    """Find the backbone (and FPN, if any) inside a checkpoint of any of the formats above."""
    dit_prefix = None
    for p in DIT_PREFIXES:
        if any(k.startswith(p) and k[len(p):].startswith(_DIT_ROOTS) for k in sd):
            dit_prefix = p
            break
    if dit_prefix is None:
        raise ValueError("no BeitModel tensors found (looked for embeddings./encoder./pooler. under the prefixes "
                         + ", ".join(repr(p) for p in DIT_PREFIXES) + ")")
    fpn_prefix = next((p for p in FPN_PREFIXES if any(k.startswith(p + "inner_blocks.") for k in sd)), None)
    dit, fpn, other = OrderedDict(), OrderedDict(), []
    for k, v in sd.items():
        if k.startswith(dit_prefix) and k[len(dit_prefix):].startswith(_DIT_ROOTS):
            name = k[len(dit_prefix):]
            if not name.endswith(_IGNORED_SUFFIX):
                dit[name] = v
        elif fpn_prefix is not None and k.startswith(fpn_prefix):
            fpn[k[len(fpn_prefix):]] = v
        else:
            other.append(k)
    return SplitCheckpoint(dit=dit, fpn=fpn, other=other, dit_prefix=dit_prefix, fpn_prefix=fpn_prefix)


def infer_config(dit_sd: dict, patch_size: int = 16, **overrides) -> DiTConfig:
    """The architecture, read off the tensors (HF names): what ``config.json`` would have said."""
    cls = dit_sd["embeddings.cls_token"]
    D = int(cls.shape[-1])
    layers = {int(k.split(".")[2]) for k in dit_sd if k.startswith("encoder.layer.")}
    L = max(layers) + 1
    I = int(dit_sd["encoder.layer.0.intermediate.dense.weight"].shape[0])
    per_layer = "encoder.layer.0.attention.attention.relative_position_bias.relative_position_bias_table"
    shared = "encoder.relative_position_bias.relative_position_bias_table"
    table = dit_sd.get(per_layer, dit_sd.get(shared))
    heads = int(table.shape[1]) if table is not None else D // 64
    if "embeddings.position_embeddings" in dit_sd:
        g = int(round((dit_sd["embeddings.position_embeddings"].shape[1] - 1) ** 0.5))
    elif table is not None:
        g = (int(round((table.shape[0] - 3) ** 0.5)) + 1) // 2
    else:
        g = 224 // patch_size
    lam = dit_sd.get("encoder.layer.0.lambda_1")
    kw = dict(hidden_size=D, num_hidden_layers=L, num_attention_heads=heads, intermediate_size=I,
              image_size=g * patch_size, patch_size=patch_size,
              num_channels=int(dit_sd["embeddings.patch_embeddings.projection.weight"].shape[1]),
              use_mask_token="embeddings.mask_token" in dit_sd,
              use_absolute_position_embeddings="embeddings.position_embeddings" in dit_sd,
              use_relative_position_bias=per_layer in dit_sd,
              use_shared_relative_position_bias=shared in dit_sd,
              layer_scale_init_value=0.1 if lam is not None else 0.0)
    kw.update(overrides)
    return DiTConfig(**kw)


# never evaluated by the backbone and absent from the real microsoft/dit-* files: those are
# BeitForMaskedImageModeling exports, built with BeitModel(add_pooling_layer=False) (HF:799)
OPTIONAL_DIT_KEYS = ("pooler.layernorm.weight", "pooler.layernorm.bias")


def load_dit_state_dict(dit, sd: dict, strict: bool = True):
    """``dit.load_state_dict`` with HF ``from_pretrained`` semantics for the pooler: ``pooler.layernorm.*`` may be
    missing (it keeps its initial value, HF warns "newly initialized"); every other mismatch raises when ``strict``."""
    res = dit.load_state_dict(sd, strict=False)
    missing = [k for k in res.missing_keys if k not in OPTIONAL_DIT_KEYS]
    unexpected = [k for k in res.unexpected_keys if not k.endswith(_IGNORED_SUFFIX)]
    if strict and (missing or unexpected):
        raise RuntimeError(f"Error(s) in loading state_dict for {type(dit).__name__}: "
                           f"Missing key(s): {missing}. Unexpected key(s): {unexpected}.")
    return res


def load_checkpoint(module, src, strict: bool = True):
    """Load ``src`` (a path or a state dict in any of the formats above) into a ``DiTBackbone`` or
    ``DiTWithFPN``.  Returns the :class:`SplitCheckpoint` (``.other`` lists what was not consumed)."""
    sd = read_state_dict(src) if isinstance(src, (str, os.PathLike)) else src
    parts = split_checkpoint(sd)
    backbone = module.backbone if hasattr(module, "backbone") else module
    load_dit_state_dict(backbone.dit, parts.dit, strict=strict)
    if hasattr(backbone, "pretrained"):
        backbone.pretrained = False   # real weights are in: nothing left to fetch
    if hasattr(module, "fpn"):
        if parts.fpn:
            module.fpn.load_state_dict(parts.fpn, strict=strict)
        elif strict and parts.fpn_prefix is not None:
            raise KeyError("checkpoint has an FPN prefix but no FPN tensors")
    return parts


def build_from_checkpoint(src, with_fpn: bool | None = None, **kw):
    """Construct the drop-in module a checkpoint describes: architecture inferred from the tensors,
    ``DiTWithFPN`` when the file carries FPN weights (or ``with_fpn=True``), else ``DiTBackbone``."""
    from .dit_backbone import DiTBackbone
    from .dit_fpn import DiTWithFPN
    sd = read_state_dict(src) if isinstance(src, (str, os.PathLike)) else src
    parts = split_checkpoint(sd)
    cfg = infer_config(parts.dit)
    if with_fpn is None:
        with_fpn = bool(parts.fpn)
    m = DiTWithFPN(pretrained=False, config=cfg, **kw) if with_fpn else DiTBackbone(pretrained=False, config=cfg, **kw)
    load_checkpoint(m, sd, strict=True)
    return m


def export_state_dict(module, layout: str = "hf", cpu: bool = True) -> "OrderedDict[str, torch.Tensor]":
    """``layout="hf"``: bare ``BeitModel`` names (what ``AutoModel.from_pretrained`` reads);
    ``layout="layoutdit"``: the keys of the reference's whole-model checkpoint for the parts this module
    owns (``model.backbone.backbone.dit.*`` and, for ``DiTWithFPN``, ``model.backbone.fpn.*``)."""
    backbone = module.backbone if hasattr(module, "backbone") else module
    conv = (lambda t: t.detach().cpu()) if cpu else (lambda t: t.detach())
    out = OrderedDict()
    if layout == "hf":
        for k, v in backbone.dit.state_dict().items():
            out[k] = conv(v)
    elif layout == "layoutdit":
        for k, v in backbone.dit.state_dict().items():
            out["model.backbone.backbone.dit." + k] = conv(v)
        if hasattr(module, "fpn"):
            for k, v in module.fpn.state_dict().items():
                out["model.backbone.fpn." + k] = conv(v)
    else:
        raise ValueError("layout must be 'hf' or 'layoutdit'")
    return out


def save_checkpoint(module, path: str, layout: str = "hf", cpu: bool = True) -> str:
    """R:model.py:90-121 writes ``torch.save(state_dict)`` twice (device tensors and a CPU copy); this writes
    one file -- ``.safetensors`` if the name says so, else a ``torch.save`` pickle -- CPU tensors by default."""
    sd = export_state_dict(module, layout, cpu)
    if str(path).endswith(".safetensors"):
        from safetensors.torch import save_file
        save_file({k: v.contiguous() for k, v in sd.items()}, path)
    else:
        torch.save(sd, path)
    return path
