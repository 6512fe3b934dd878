"""CPU ORACLE for the DiT backbone forward -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this file.  The product path (``layoutdit_b200``) never
imports it and has no CPU fallback.

What this restates
------------------
The reference hot path is ``DiTBackbone.forward``
(R:src/layoutdit/modeling/dit_backbone.py:38-62).  Its arithmetic lives in a third-party,
un-vendored dependency: HuggingFace ``transformers`` ``BeitModel`` (reference pins
transformers==4.49.0, R:uv.lock:1771-1772; 5.5.0 is what is installed in this image and is
what the line numbers ``HF:`` below refer to:
``transformers/models/beit/modeling_beit.py``).  This file restates that published
algorithm with primitive fp32 (or fp64) tensor ops on the CPU -- no ``nn.Module``, no
``transformers`` import -- operating on a plain ``state_dict`` with HF ``BeitModel`` key
names.

Parity pinning
--------------
The reference ships no golden vectors for this path (its only test file covers the dataset
and errors at collection, R:tests/test_dataset.py:23).  The oracle is therefore pinned
against OUTPUTS OF THE REFERENCE ITSELF: ``oracle/make_golden.py`` imports the reference's
own ``DiTBackbone`` class from /root/reference/src (only the hub fetch at
dit_backbone.py:26-31 replaced by a local ``BeitConfig``), runs it on seeded inputs and
commits the results under ``tests/golden/``; ``tests/test_oracle.py`` checks this
restatement against those fixtures and against the installed ``transformers.BeitModel``.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch


# --------------------------------------------------------------------------- embeddings
def patch_embed(x, weight, bias):
    """``Conv2d(3, D, k=16, s=16)`` then ``flatten(2).transpose(1, 2)``  (HF:209, 218-220).

    Written as the im2col GEMM it is: rows are patches in (gy, gx) raster order, columns
    are (c, py, px) in the order of the flattened conv weight ``[D, 3*16*16]``."""
    B, C, H, W = x.shape
    D, _, ps, _ = weight.shape
    Gh, Gw = H // ps, W // ps
    cols = x[:, :, :Gh * ps, :Gw * ps].reshape(B, C, Gh, ps, Gw, ps)
    cols = cols.permute(0, 2, 4, 1, 3, 5).reshape(B, Gh * Gw, C * ps * ps)
    return cols @ weight.reshape(D, C * ps * ps).t() + bias


def _cubic_weights(t, a=-0.75):
    """Keys cubic convolution coefficients, A=-0.75 (what ATen's upsample_bicubic2d uses)."""
    def c1(x):  # |x| <= 1
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1.0
    def c2(x):  # 1 < |x| < 2
        return ((a * x - 5.0 * a) * x + 8.0 * a) * x - 4.0 * a
    return [c2(t + 1.0), c1(t), c1(1.0 - t), c2(2.0 - t)]


def _bicubic_axis(src_len, dst_len, dtype):
    """Dense [dst_len, src_len] matrix of 1-D bicubic resampling, align_corners=False,
    border indices clamped (ATen ``upsample_bicubic2d`` semantics, no antialias)."""
    m = torch.zeros(dst_len, src_len, dtype=dtype)
    scale = src_len / dst_len
    for o in range(dst_len):
        real = scale * (o + 0.5) - 0.5
        i0 = math.floor(real)
        w = _cubic_weights(real - i0)
        for k in range(4):
            idx = min(max(i0 - 1 + k, 0), src_len - 1)
            m[o, idx] += w[k]
    return m


def interpolate_pos_encoding(pos, native_grid, Gh, Gw, height, width):
    """HF:121-159.  ``pos`` is ``[1, 1+g*g, D]``.  Returned unchanged when the patch count
    matches and the image is square (HF:135-136); otherwise the CLS row is kept and the
    patch rows are resampled bicubically from g x g to Gh x Gw (HF:138-159)."""
    num_positions = pos.shape[1] - 1
    if Gh * Gw == num_positions and height == width:
        return pos
    g = native_grid
    D = pos.shape[-1]
    grid = pos[0, 1:].reshape(g, g, D)
    my = _bicubic_axis(g, Gh, pos.dtype)
    mx = _bicubic_axis(g, Gw, pos.dtype)
    out = torch.einsum("oy,yxd->oxd", my, grid)
    out = torch.einsum("px,oxd->opd", mx, out).reshape(1, Gh * Gw, D)
    return torch.cat([pos[:, :1], out], dim=1)


def embeddings(sd, cfg, x):
    """``BeitEmbeddings.forward`` (HF:161-184) with ``bool_masked_pos=None``, dropout p=0."""
    B, _, H, W = x.shape
    ps = cfg["patch_size"]
    tok = patch_embed(x, sd["embeddings.patch_embeddings.projection.weight"],
                      sd["embeddings.patch_embeddings.projection.bias"])
    cls = sd["embeddings.cls_token"].expand(B, -1, -1)
    emb = torch.cat([cls, tok], dim=1)
    if "embeddings.position_embeddings" in sd:
        emb = emb + interpolate_pos_encoding(sd["embeddings.position_embeddings"],
                                             cfg["image_size"] // ps, H // ps, W // ps, H, W)
    return emb


# ------------------------------------------------------------------ relative position bias
def relative_position_index(Gh, Gw):
    """``generate_relative_position_index`` (HF:522-544): int64 ``[N, N]`` with values in
    ``[0, (2Gh-1)(2Gw-1)+3)``; the three extra rows are cls->token, token->cls, cls->cls."""
    num_rel = (2 * Gh - 1) * (2 * Gw - 1) + 3
    ys, xs = torch.meshgrid(torch.arange(Gh), torch.arange(Gw), indexing="ij")
    ys, xs = ys.flatten(), xs.flatten()
    dy = ys[:, None] - ys[None, :] + (Gh - 1)
    dx = xs[:, None] - xs[None, :] + (Gw - 1)
    idx = torch.zeros(Gh * Gw + 1, Gh * Gw + 1, dtype=torch.int64)
    idx[1:, 1:] = dy * (2 * Gw - 1) + dx
    idx[0, :] = num_rel - 3
    idx[:, 0] = num_rel - 2
    idx[0, 0] = num_rel - 1
    return idx


def _bilinear_axis(src_len, dst_len, dtype, scale=None):
    """Dense [dst_len, src_len] matrix of 1-D bilinear resampling, align_corners=False.
    ``scale`` = the user scale_factor when one was given (ATen then uses 1/scale_factor as
    the coordinate ratio instead of src/dst)."""
    m = torch.zeros(dst_len, src_len, dtype=dtype)
    ratio = (1.0 / scale) if scale is not None else src_len / dst_len
    for o in range(dst_len):
        real = max(ratio * (o + 0.5) - 0.5, 0.0)
        i0 = min(int(math.floor(real)), src_len - 1)
        i1 = min(i0 + 1, src_len - 1)
        l1 = real - i0
        m[o, i0] += 1.0 - l1
        m[o, i1] += l1
    return m


def resized_bias_table(table, native_grid, Gh, Gw):
    """First half of ``BeitRelativePositionBias.forward`` (HF:550-571): bilinearly resize the
    ``(2g-1) x (2g-1)`` part of the table to ``(2Gh-1) x (2Gw-1)`` and re-append the three
    CLS rows.  Returns ``[(2Gh-1)(2Gw-1)+3, heads]``.  (HF reshapes the old table as
    ``(old_width, old_height)``; the native window is square so the two coincide.)"""
    g = native_grid
    old = 2 * g - 1
    nh, nw = 2 * Gh - 1, 2 * Gw - 1
    sub = table[: old * old].reshape(old, old, -1)
    my = _bilinear_axis(old, nh, table.dtype)
    mx = _bilinear_axis(old, nw, table.dtype)
    new = torch.einsum("oy,yxh->oxh", my, sub)
    new = torch.einsum("px,oxh->oph", mx, new).reshape(nh * nw, -1)
    return torch.cat([new, table[old * old:]], dim=0)


def relative_position_bias(table, native_grid, Gh, Gw):
    """``BeitRelativePositionBias.forward`` (HF:546-591, ``interpolate_pos_encoding=False``
    as the reference never sets it): ``[1, heads, N, N]``."""
    new_table = resized_bias_table(table, native_grid, Gh, Gw)
    idx = relative_position_index(Gh, Gw)
    N = Gh * Gw + 1
    bias = new_table[idx.reshape(-1)].reshape(N, N, -1).permute(2, 0, 1)
    return bias.unsqueeze(0)


# ----------------------------------------------------------------------------- one layer
def layer_norm(x, weight, bias, eps):
    """``nn.LayerNorm(D, eps)`` (HF:458, 460): biased variance over the last dim."""
    mean = x.mean(dim=-1, keepdim=True)
    xc = x - mean
    var = (xc * xc).mean(dim=-1, keepdim=True)
    return xc / torch.sqrt(var + eps) * weight + bias


def gelu_erf(x):
    """HF ``ACT2FN["gelu"]`` == exact erf GELU (HFC:77; HF:430)."""
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def self_attention(sd, pfx, x, heads, bias):
    """``BeitSelfAttention.forward`` (HF:249-306; the sdpa twin HF:310-368 computes the same
    thing).  ``key`` has no bias (HF:240).  ``bias`` is ``[1, heads, N, N]`` or None."""
    B, N, D = x.shape
    dh = D // heads
    a = pfx + "attention.attention."
    q = x @ sd[a + "query.weight"].t() + sd[a + "query.bias"]
    k = x @ sd[a + "key.weight"].t()
    v = x @ sd[a + "value.weight"].t() + sd[a + "value.bias"]
    q = q.reshape(B, N, heads, dh).transpose(1, 2)
    k = k.reshape(B, N, heads, dh).transpose(1, 2)
    v = v.reshape(B, N, heads, dh).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(dh)
    if bias is not None:
        s = s + bias
    p = torch.softmax(s, dim=-1)
    ctx = (p @ v).transpose(1, 2).reshape(B, N, D)
    return ctx


def beit_layer(sd, cfg, i, x, shared_bias, Gh, Gw, drop=None):
    """``BeitLayer.forward`` (HF:469-508).  ``drop=None``: eval mode, drop-path is the identity (HF:66-67); else a
    ``[2, B]`` tensor of the factors ``floor(keep + U) / keep`` that ``drop_path`` (HF:61-73) multiplies the layer-scaled
    attention / MLP branch of every image with in training mode (HF:492, 504)."""
    pfx = f"encoder.layer.{i}."
    eps = cfg["layer_norm_eps"]
    heads = cfg["num_attention_heads"]
    bias = None
    tkey = pfx + "attention.attention.relative_position_bias.relative_position_bias_table"
    if tkey in sd:  # per-layer table (HF:245-247, 285-290)
        bias = relative_position_bias(sd[tkey], cfg["image_size"] // cfg["patch_size"], Gh, Gw)
    if shared_bias is not None:  # HF:292-294
        bias = shared_bias if bias is None else bias + shared_bias
    h = layer_norm(x, sd[pfx + "layernorm_before.weight"], sd[pfx + "layernorm_before.bias"], eps)
    ctx = self_attention(sd, pfx, h, heads, bias)
    attn = ctx @ sd[pfx + "attention.output.dense.weight"].t() + sd[pfx + "attention.output.dense.bias"]
    if pfx + "lambda_1" in sd:
        attn = sd[pfx + "lambda_1"] * attn
    if drop is not None:
        attn = attn * drop[0].to(attn.dtype).reshape(-1, 1, 1)
    x = attn + x
    h = layer_norm(x, sd[pfx + "layernorm_after.weight"], sd[pfx + "layernorm_after.bias"], eps)
    h = gelu_erf(h @ sd[pfx + "intermediate.dense.weight"].t() + sd[pfx + "intermediate.dense.bias"])
    h = h @ sd[pfx + "output.dense.weight"].t() + sd[pfx + "output.dense.bias"]
    if pfx + "lambda_2" in sd:
        h = sd[pfx + "lambda_2"] * h
    if drop is not None:
        h = h * drop[1].to(h.dtype).reshape(-1, 1, 1)
    return h + x


def hidden_states(sd, cfg, x):
    """``BeitModel.forward(...).hidden_states`` (HF:720-764 -> HF:616-663): a list of L+1
    tensors ``[B, N, D]``; entry 0 is the embedding output, entry i the output of layer i."""
    B, _, H, W = x.shape
    ps = cfg["patch_size"]
    Gh, Gw = H // ps, W // ps
    h = embeddings(sd, cfg, x)
    shared = None
    skey = "encoder.relative_position_bias.relative_position_bias_table"
    if skey in sd:  # HF:598-600, 632-637
        shared = relative_position_bias(sd[skey], cfg["image_size"] // ps, Gh, Gw)
    out = [h]
    for i in range(cfg["num_hidden_layers"]):
        h = beit_layer(sd, cfg, i, h, shared, Gh, Gw)
        out.append(h)
    return out


# ---------------------------------------------------------------------------------- taps
def resample_bilinear(t, scale):
    """``F.interpolate(t, scale_factor=scale, mode="bilinear", align_corners=False)`` for
    ``t`` of shape ``[B, D, h, w]`` (R:dit_backbone.py:56-59).  Output size is
    ``floor(h*scale)``; the coordinate ratio is ``1/scale`` (scale_factor given and
    recompute_scale_factor unset)."""
    B, D, h, w = t.shape
    oh, ow = int(math.floor(h * scale)), int(math.floor(w * scale))
    my = _bilinear_axis(h, oh, t.dtype, scale)
    mx = _bilinear_axis(w, ow, t.dtype, scale)
    out = torch.einsum("oy,bdyx->bdox", my, t)
    return torch.einsum("px,bdox->bdop", mx, out)


def tap_layer_indices(num_layers):
    """R:dit_backbone.py:33-34."""
    d = num_layers
    return [d // 3, d // 2, 2 * d // 3, d]


TAP_SCALES = [4.0, 2.0, 1.0, 0.5]  # R:dit_backbone.py:35


def dit_backbone_forward(sd, cfg, x, dtype=torch.float32):
    """``DiTBackbone.forward`` (R:src/layoutdit/modeling/dit_backbone.py:38-62):
    ``x [B,3,H,W]`` -> ``OrderedDict{p2,p3,p4,p5}`` of ``[B, D, h_i, w_i]``."""
    sd = {k: v.detach().to("cpu", dtype) for k, v in sd.items() if v.is_floating_point()}
    x = x.detach().to("cpu", dtype)
    B, _, H, W = x.shape
    ps = cfg["patch_size"]
    Gh, Gw = H // ps, W // ps
    D = cfg["hidden_size"]
    hs = hidden_states(sd, cfg, x)
    feats = OrderedDict()
    for i, (idx, scale) in enumerate(zip(tap_layer_indices(cfg["num_hidden_layers"]), TAP_SCALES), start=2):
        t = hs[idx][:, 1:, :].permute(0, 2, 1).reshape(B, D, Gh, Gw)
        if scale != 1.0:
            t = resample_bilinear(t, scale)
        feats[f"p{i}"] = t
    return feats


def as_cfg_dict(cfg) -> dict:
    """Accept a ``DiTConfig``, a transformers ``BeitConfig`` or a dict."""
    if isinstance(cfg, dict):
        return cfg
    d = cfg.to_dict()
    for k in ("image_size", "patch_size"):
        if not isinstance(d[k], int):
            d[k] = int(d[k][0])
    return d
