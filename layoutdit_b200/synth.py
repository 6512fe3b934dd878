"""Deterministic synthetic document pages and HF-named weight sets.

There is no network on the B200 boxes (no PubLayNet, no ``microsoft/dit-base``
checkpoint), so benchmarks and parity tests use:

* pages "already in backbone space": the detector normalises with mean/std 0.5
  (R:src/layoutdit/modeling/model.py:53-54) so values lie in [-1, 1]; a white page with
  dark axis-aligned blocks plus a little noise;
* two weight sets keyed exactly like ``transformers`` ``BeitModel.state_dict()``:
  ``W0`` follows HF's random init (HF:677-692: N(0, .02) matrices, zero biases, zero
  cls/pos/tables, LayerNorm (1, 0), layer-scale 0.1) and ``W1`` is a stress init with every
  fused term non-trivial (biases, position rows, tables, LayerNorm affine, layer-scale).

numpy's PCG64 stream is used instead of torch's CPU generator because it is bit-stable
across machines, so the golden fixtures under tests/golden/ can be regenerated anywhere.
"""
from __future__ import annotations

import numpy as np
import torch

from .config import DiTConfig


def synthetic_pages(batch: int, height: int, width: int, seed: int = 1234) -> torch.Tensor:
    """``[batch, 3, height, width]`` fp32 pages in [-1, 1]."""
    rng = np.random.default_rng(seed)
    x = np.ones((batch, 3, height, width), dtype=np.float32)
    for b in range(batch):
        for _ in range(int(rng.integers(8, 21))):
            h = int(rng.integers(max(2, height // 40), max(3, height // 4)))
            w = int(rng.integers(max(2, width // 10), max(3, (3 * width) // 4)))
            y0 = int(rng.integers(0, max(1, height - h)))
            x0 = int(rng.integers(0, max(1, width - w)))
            x[b, :, y0:y0 + h, x0:x0 + w] = np.float32(rng.uniform(-1.0, -0.2))
    x += rng.standard_normal(x.shape, dtype=np.float32) * np.float32(0.02)
    np.clip(x, -1.0, 1.0, out=x)
    return torch.from_numpy(x)


def state_dict_keys(cfg: DiTConfig):
    """(name, shape) pairs of ``BeitModel.state_dict()`` for ``cfg``."""
    D, I, h = cfg.hidden_size, cfg.intermediate_size, cfg.num_attention_heads
    g = cfg.grid
    nrel = (2 * g - 1) * (2 * g - 1) + 3
    out = [("embeddings.cls_token", (1, 1, D))]
    if cfg.use_mask_token:
        out.append(("embeddings.mask_token", (1, 1, D)))
    if cfg.use_absolute_position_embeddings:
        out.append(("embeddings.position_embeddings", (1, g * g + 1, D)))
    out += [("embeddings.patch_embeddings.projection.weight", (D, cfg.num_channels, cfg.patch_size, cfg.patch_size)),
            ("embeddings.patch_embeddings.projection.bias", (D,))]
    if cfg.use_shared_relative_position_bias:
        out.append(("encoder.relative_position_bias.relative_position_bias_table", (nrel, h)))
    for i in range(cfg.num_hidden_layers):
        p = f"encoder.layer.{i}."
        if cfg.layer_scale_init_value > 0:
            out += [(p + "lambda_1", (D,)), (p + "lambda_2", (D,))]
        out += [(p + "attention.attention.query.weight", (D, D)),
                (p + "attention.attention.query.bias", (D,)),
                (p + "attention.attention.key.weight", (D, D)),
                (p + "attention.attention.value.weight", (D, D)),
                (p + "attention.attention.value.bias", (D,))]
        if cfg.use_relative_position_bias:
            out.append((p + "attention.attention.relative_position_bias.relative_position_bias_table", (nrel, h)))
        out += [(p + "attention.output.dense.weight", (D, D)),
                (p + "attention.output.dense.bias", (D,)),
                (p + "intermediate.dense.weight", (I, D)),
                (p + "intermediate.dense.bias", (I,)),
                (p + "output.dense.weight", (D, I)),
                (p + "output.dense.bias", (D,)),
                (p + "layernorm_before.weight", (D,)),
                (p + "layernorm_before.bias", (D,)),
                (p + "layernorm_after.weight", (D,)),
                (p + "layernorm_after.bias", (D,))]
    out += [("pooler.layernorm.weight", (D,)), ("pooler.layernorm.bias", (D,))]
    return out


def make_state_dict(cfg: DiTConfig, seed: int = 0, stress: bool = False) -> dict:
    """HF-named fp32 CPU state dict.  ``stress=False`` -> W0, ``stress=True`` -> W1."""
    rng = np.random.default_rng(seed)

    def normal(shape, std):
        return (rng.standard_normal(shape, dtype=np.float32) * np.float32(std)).astype(np.float32)

    sd = {}
    for name, shape in state_dict_keys(cfg):
        leaf = name.rsplit(".", 1)[-1]
        is_ln = "layernorm" in name
        if leaf == "weight" and not is_ln:
            v = normal(shape, 0.04 if stress else cfg.initializer_range)
        elif is_ln and leaf == "weight":
            v = 1.0 + normal(shape, 0.1) if stress else np.ones(shape, np.float32)
        elif is_ln and leaf == "bias":
            v = normal(shape, 0.1) if stress else np.zeros(shape, np.float32)
        elif leaf == "bias":
            v = normal(shape, 0.1) if stress else np.zeros(shape, np.float32)
        elif leaf in ("lambda_1", "lambda_2"):
            v = (rng.uniform(0.05, 1.0, shape).astype(np.float32) if stress
                 else np.full(shape, cfg.layer_scale_init_value, np.float32))
        elif leaf == "relative_position_bias_table":
            v = normal(shape, 0.5) if stress else np.zeros(shape, np.float32)
        else:  # cls_token, mask_token, position_embeddings
            v = normal(shape, 0.2) if stress else np.zeros(shape, np.float32)
        sd[name] = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32))
    return sd


def make_fpn_state_dict(in_channels: int, out_channels: int = 256, seed: int = 0, stress: bool = False) -> dict:
    """torchvision ``FeaturePyramidNetwork([in_channels]*4, out_channels)`` state dict (TV:104-131).
    ``stress=False`` follows its init (kaiming_uniform(a=1): U(-sqrt(3/fan_in), +), zero bias);
    ``stress=True`` adds non-zero biases and larger weights so every fused term is exercised."""
    rng = np.random.default_rng(seed)
    sd = {}
    for block, k, cin in (("inner_blocks", 1, in_channels), ("layer_blocks", 3, out_channels)):
        bound = float(np.sqrt(3.0 / (cin * k * k))) * (1.5 if stress else 1.0)
        for i in range(4):
            w = rng.uniform(-bound, bound, (out_channels, cin, k, k)).astype(np.float32)
            b = (rng.standard_normal(out_channels).astype(np.float32) * np.float32(0.1) if stress
                 else np.zeros(out_channels, np.float32))
            sd[f"{block}.{i}.0.weight"] = torch.from_numpy(w)
            sd[f"{block}.{i}.0.bias"] = torch.from_numpy(b)
    return sd


def raw_pages(sizes, seed: int = 0):
    """Raw document pages as the dataset hands them to the detector: a list of ``[3, H, W]`` fp32 tensors in
    [0, 1] (white paper, dark blocks, a little noise), one per ``(H, W)`` in ``sizes``."""
    rng = np.random.default_rng(seed)
    out = []
    for (h, w) in sizes:
        x = np.ones((3, h, w), dtype=np.float32)
        for _ in range(int(rng.integers(4, 12))):
            bh, bw = int(rng.integers(1, max(2, h // 3))), int(rng.integers(1, max(2, w // 2)))
            y0, x0 = int(rng.integers(0, max(1, h - bh))), int(rng.integers(0, max(1, w - bw)))
            x[:, y0:y0 + bh, x0:x0 + bw] = np.float32(rng.uniform(0.0, 0.4))
        x += rng.standard_normal(x.shape, dtype=np.float32) * np.float32(0.01)
        np.clip(x, 0.0, 1.0, out=x)
        out.append(torch.from_numpy(x))
    return out
