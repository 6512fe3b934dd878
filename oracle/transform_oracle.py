"""CPU ORACLE for the detector's input transform -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (same import rules as
``oracle/dit_oracle.py``).

Restates what the reference applies to raw pages before its backbone: ``FasterRCNN(..., min_size=224,
max_size=224, fixed_size=(224, 224), image_mean=(.5, .5, .5), image_std=(.5, .5, .5))`` at
R:src/layoutdit/modeling/model.py:44-56 builds torchvision's ``GeneralizedRCNNTransform`` (third-party,
un-vendored: torchvision 0.19.0 pinned at R:uv.lock:1709-1710, 0.26.0 installed; TV =
``torchvision/models/detection/transform.py``): per page ``normalize`` ((image - mean) / std), ``resize``
(``_resize_image_and_masks`` with ``fixed_size``: ``F.interpolate(image[None], size=(224, 224),
mode="bilinear", align_corners=False)``), then ``batch_images`` (a 224x224 batch needs no padding to the
32-pixel stride).  Pinned by ``oracle/make_golden_transform.py`` against the transform object inside the
reference's own ``LayoutDetectionModel``.
"""
from __future__ import annotations

import numpy as np
import torch


def _bilinear_axis_f32(src_len, dst_len, dtype):
    """Dense [dst_len, src_len] matrix of ATen's 1-D bilinear resampling for ``size=`` given and
    align_corners=False, with the coordinate arithmetic in fp32 as ATen does it for fp32 images
    (area_pixel_compute_scale / area_pixel_compute_source_index: ratio = in / out, src = ratio * (dst + .5) - .5,
    clamped at 0; i0 = floor(src); i1 = i0 + (i0 < in - 1); lambda1 = src - i0).  On a 1024 -> 224 resize the
    fp32 rounding of ``src`` moves the weights by up to ~1e-4, so a float64 restatement would not reproduce
    the reference's numbers."""
    f = np.float32
    ratio = f(src_len) / f(dst_len)
    m = torch.zeros(dst_len, src_len, dtype=dtype)
    for o in range(dst_len):
        src = max(f(ratio * f(f(o) + f(0.5))) - f(0.5), f(0.0))
        i0 = min(int(src), src_len - 1)
        i1 = i0 + (1 if i0 < src_len - 1 else 0)
        l1 = f(src - f(i0))
        m[o, i0] += float(f(1.0) - l1)
        m[o, i1] += float(l1)
    return m


def page_transform(pages, size=(224, 224), mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5), dtype=torch.float32):
    """List of ``[3, H_i, W_i]`` pages in [0, 1] -> ``[B, 3, size[0], size[1]]`` backbone input."""
    m = torch.tensor(mean, dtype=dtype)[:, None, None]
    s = torch.tensor(std, dtype=dtype)[:, None, None]
    out = []
    for p in pages:
        p = (p.to(dtype) - m) / s                                    # normalize()
        my = _bilinear_axis_f32(p.shape[1], size[0], dtype)          # size= given: ratio = in / out
        mx = _bilinear_axis_f32(p.shape[2], size[1], dtype)
        out.append(torch.einsum("oy,cyx,px->cop", my, p, mx))        # _resize_image_and_masks(fixed_size)
    return torch.stack(out)                                          # batch_images: equal sizes, 224 % 32 == 0
