"""CPU ORACLE for the FPN on the DiT taps -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Same import rules as ``oracle/dit_oracle.py``: only ``tests/``, ``__graft_entry__.smoke()`` and the CPU
legs of ``bench.py`` may import this file; ``layoutdit_b200`` never does.

What this restates: the reference's ``DiTWithFPN.forward`` (R:src/layoutdit/modeling/dit_backbone.py:92-95),
i.e. ``DiTBackbone.forward`` followed by torchvision's ``FeaturePyramidNetwork.forward`` with
``LastLevelMaxPool``.  The arithmetic lives in a third-party dependency, ``torchvision`` (reference pins
0.19.0, R:uv.lock:1709-1710; 0.26.0 is installed here and is what ``TV:`` line numbers refer to:
``torchvision/ops/feature_pyramid_network.py``).  The convolutions are written out as the matrix products
they are (no ``nn.Module``, no torchvision import) on a plain state dict with torchvision's key names.

Parity pinning: the reference has no test or golden vector for this path; ``oracle/make_golden_fpn.py``
runs the reference's own ``DiTWithFPN`` class (imported from /root/reference/src, hub fetch replaced) on
seeded inputs and commits its outputs under ``tests/golden/fpn_*.npz``; ``tests/test_oracle.py`` checks this
restatement against them.
"""
from __future__ import annotations

from collections import OrderedDict

import torch

from . import dit_oracle


def conv1x1(t, weight, bias):
    """``Conv2d(Cin, Cout, 1)`` (TV:111-116): a matrix product over the channel axis."""
    return torch.einsum("bchw,oc->bohw", t, weight[:, :, 0, 0]) + bias[None, :, None, None]


def conv3x3(t, weight, bias):
    """``Conv2d(C, Cout, 3, padding=1)`` (TV:118-124): nine shifted matrix products over a zero-padded image."""
    B, C, H, W = t.shape
    p = torch.zeros(B, C, H + 2, W + 2, dtype=t.dtype)
    p[:, :, 1:H + 1, 1:W + 1] = t
    out = bias[None, :, None, None].expand(B, weight.shape[0], H, W).clone()
    for ky in range(3):
        for kx in range(3):
            out += torch.einsum("bchw,oc->bohw", p[:, :, ky:ky + H, kx:kx + W], weight[:, :, ky, kx])
    return out


def nearest_resize(t, size):
    """``F.interpolate(t, size=size, mode="nearest")`` (TV:189): src = min(floor(dst * in / out), in - 1),
    the scale evaluated in fp32 as ATen does."""
    H, W = t.shape[-2:]
    oh, ow = size
    sy = torch.tensor(H / oh, dtype=torch.float32)
    sx = torch.tensor(W / ow, dtype=torch.float32)
    iy = torch.clamp(torch.floor(torch.arange(oh, dtype=torch.float32) * sy).long(), max=H - 1)
    ix = torch.clamp(torch.floor(torch.arange(ow, dtype=torch.float32) * sx).long(), max=W - 1)
    return t[:, :, iy][:, :, :, ix]


def fpn_forward(fsd: dict, feats: "OrderedDict[str, torch.Tensor]") -> "OrderedDict[str, torch.Tensor]":
    """``FeaturePyramidNetwork.forward`` (TV:172-204) + ``LastLevelMaxPool`` (TV:231-249)."""
    names = list(feats.keys())
    xs = list(feats.values())
    inner = lambda i, t: conv1x1(t, fsd[f"inner_blocks.{i}.0.weight"], fsd[f"inner_blocks.{i}.0.bias"])
    layer = lambda i, t: conv3x3(t, fsd[f"layer_blocks.{i}.0.weight"], fsd[f"layer_blocks.{i}.0.bias"])
    n = len(xs)
    last_inner = inner(n - 1, xs[-1])                         # TV:181
    results = [layer(n - 1, last_inner)]                      # TV:183
    for idx in range(n - 2, -1, -1):                          # TV:185-193
        lateral = inner(idx, xs[idx])
        last_inner = lateral + nearest_resize(last_inner, lateral.shape[-2:])
        results.insert(0, layer(idx, last_inner))
    results.append(results[-1][:, :, ::2, ::2])               # max_pool2d(kernel 1, stride 2), TV:247
    names.append("pool")
    return OrderedDict(zip(names, results))


def dit_with_fpn_forward(sd: dict, fsd: dict, cfg: dict, x, dtype=torch.float32):
    """R:dit_backbone.py:92-95: ``fpn(backbone(x))``."""
    feats = dit_oracle.dit_backbone_forward(sd, cfg, x, dtype=dtype)
    return fpn_forward({k: v.to(dtype) for k, v in fsd.items()}, feats)
