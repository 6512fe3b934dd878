"""Training path: ``DiTBackbone.forward`` differentiable end to end -- embeddings, every
``BeitLayer`` and the four taps as ``torch.autograd.Function``s over the C ABI (``TrainableBackbone``; ``TrainableEncoder``
is the layer stack alone) -- and the bucketed gradient all-reduce of a data-parallel step.

What the reference does here: ``LayoutDetectionModel`` is trained with ``torch.autograd`` through HF ``BeitLayer``
(HF:469-508) under autocast, ``scaler.scale(loss).backward()`` (R:src/layoutdit/training/trainer.py:164-183); it has no
distributed code (R:README.md:59) -- BASELINE config 5 asks for DDP with an NCCL gradient all-reduce, which is new.

Scope (the rest is listed in DESIGN.md "next"): every configuration the forward covers; relative-position tables
(per layer or shared) train at their native window, through the flash-style attention backward.  Drop-path (HF's training-mode stochastic depth, HF:61-73: a per-image
Bernoulli scaling of the two branches) is a constructor argument, active in ``train()`` mode.  The attention backward is a tcgen05 kernel: up to 256 tokens (224 x 224
pages have 197) one self-contained CTA per (image, head); beyond that a flash-style kernel over key tiles that recomputes P
from the row statistics the forward writes (``ldit_attention_lse``).  The eight GEMMs of a layer's backward run on the
forward's tcgen05 kernel; see ``csrc/backward.cuh``.

Dtypes: bf16 activations and activation gradients, fp32 residual stream / residual-stream gradients / parameter
gradients, exactly the forward's contract.  PyTorch is used for memory, autograd bookkeeping, weight casts / transposes
(data movement) and ``torch.distributed``; every arithmetic step of the layer is a library kernel.
"""
from __future__ import annotations

import torch
import torch.distributed as dist
import torch.nn as nn

from . import _lib
from .config import DiTConfig
from .dit_params import DiTParameters

_BF = torch.bfloat16
# sequences longer than this take the flash-style attention backward (the self-contained kernel stops at 256 tokens);
# LDIT_BWD_FLASH_ABOVE=0 sends every length through it (A/B knob)
_FLASH_ABOVE = min(256, int(__import__("os").environ.get("LDIT_BWD_FLASH_ABOVE", "256")))


def _st(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


class _K:
    """Thin typed wrappers over the C ABI (every call checks its status code)."""

    def __init__(self, dev):
        self.lib, self.dev = _lib.load(), dev

    def layernorm(self, x, w, b, eps):
        M, D = x.shape
        y = torch.empty(M, D, device=self.dev, dtype=_BF)
        _lib.check(self.lib.ldit_layernorm(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), M, D, eps, _st(self.dev)), "ldit_layernorm")
        return y

    def gemm(self, a, w, bias=None):
        """bf16 [M, N] = a [M, K] x w [N, K]^T (+ bias f32 [N])."""
        M, K = a.shape
        N = w.shape[0]
        out = torch.empty(M, N, device=self.dev, dtype=_BF)
        _lib.check(self.lib.ldit_gemm_bias(a.data_ptr(), w.data_ptr(), _p(bias), out.data_ptr(), M, N, K, _st(self.dev)), "ldit_gemm_bias")
        return out

    def gemm_acc(self, a, w, acc):
        """acc f32 [M, N] += a [M, K] x w [N, K]^T  (the wgrad form: fp32 reduce-add epilogue)."""
        M, K = a.shape
        N = w.shape[0]
        _lib.check(self.lib.ldit_gemm_accumulate(a.data_ptr(), w.data_ptr(), acc.data_ptr(), M, N, K, _st(self.dev)), "ldit_gemm_accumulate")

    def dgrad(self, dy, w):
        """bf16 [M, Kin] = dy [M, Nout] x w [Nout, Kin] (w as the forward holds it: no transposed copy)."""
        M, Nout = dy.shape
        Kin = w.shape[1]
        out = torch.empty(M, Kin, device=self.dev, dtype=_BF)
        _lib.check(self.lib.ldit_gemm_dgrad(dy.data_ptr(), w.data_ptr(), out.data_ptr(), M, Nout, Kin, _st(self.dev)), "ldit_gemm_dgrad")
        return out

    def wgrad(self, dy, a, acc):
        """acc f32 [Nw, Kw] += dy^T a for dy bf16 [T, Nw], a bf16 [T, Kw] (no transposed copies)."""
        T, Nw = dy.shape
        Kw = a.shape[1]
        _lib.check(self.lib.ldit_gemm_wgrad(dy.data_ptr(), a.data_ptr(), acc.data_ptr(), T, Nw, Kw, _st(self.dev)), "ldit_gemm_wgrad")

    def attention(self, qkv, B, N, heads, Gh, Gw):
        D = heads * 64
        ctx = torch.empty(B * N, D, device=self.dev, dtype=_BF)
        _lib.check(self.lib.ldit_attention(qkv.data_ptr(), ctx.data_ptr(), None, B, N, heads, Gh, Gw, _st(self.dev)), "ldit_attention")
        return ctx

    def attention_bwd(self, qkv, dctx, B, N, heads):
        dqkv = torch.empty_like(qkv)
        _lib.check(self.lib.ldit_attention_bwd(qkv.data_ptr(), dctx.data_ptr(), dqkv.data_ptr(), B, N, heads, _st(self.dev)), "ldit_attention_bwd")
        return dqkv

    def attention_lse(self, qkv, B, N, heads, Gh, Gw, table=None):
        """Forward attention that also returns the rows' log2-sum-exp; ``table``: f32 [heads, T] relative-position table or None."""
        D = heads * 64
        ctx = torch.empty(B * N, D, device=self.dev, dtype=_BF)
        lse = torch.empty(B, heads, N, device=self.dev, dtype=torch.float32)
        _lib.check(self.lib.ldit_attention_lse(qkv.data_ptr(), ctx.data_ptr(), _p(table), lse.data_ptr(), B, N, heads, Gh, Gw, _st(self.dev)),
                   "ldit_attention_lse")
        return ctx, lse

    def attention_bwd_flash(self, qkv, ctx, lse, dctx, B, N, heads, Gh, Gw, table=None, dtable=None):
        """Any sequence length: key-tile CTAs, P recomputed from the forward's row statistics; with ``table`` also
        ``dtable`` f32 [heads, T] += the table's gradient."""
        D = heads * 64
        dqkv = torch.empty_like(qkv)
        dq_acc = torch.empty(B * N, D, device=self.dev, dtype=torch.float32)
        delta = torch.empty(B * heads * N, device=self.dev, dtype=torch.float32)
        _lib.check(self.lib.ldit_attention_bwd_flash(qkv.data_ptr(), ctx.data_ptr(), lse.data_ptr(), dctx.data_ptr(), dqkv.data_ptr(),
                                                     dq_acc.data_ptr(), delta.data_ptr(), _p(table), _p(dtable), B, N, heads, Gh, Gw,
                                                     _st(self.dev)), "ldit_attention_bwd_flash")
        return dqkv

    def transpose(self, t):
        """[R, C] -> [C, R padded to a multiple of 8] (zero padding: the token dimension becomes the wgrad GEMM's K)."""
        R, C = t.shape
        ld = (R + 7) // 8 * 8
        out = torch.zeros(C, ld, device=self.dev, dtype=_BF) if ld != R else torch.empty(C, ld, device=self.dev, dtype=_BF)
        _lib.check(self.lib.ldit_transpose_bf16(t.data_ptr(), out.data_ptr(), R, C, ld, _st(self.dev)), "ldit_transpose_bf16")
        return out

    def colsum(self, t, acc, cols=None, col0=0):
        R, ld = t.shape
        C = ld if cols is None else cols
        _lib.check(self.lib.ldit_colsum_bf16(t.data_ptr() + 2 * col0, acc.data_ptr(), R, C, ld, _st(self.dev)), "ldit_colsum_bf16")

    def gelu(self, pre):
        h = torch.empty_like(pre)
        _lib.check(self.lib.ldit_gelu(pre.data_ptr(), h.data_ptr(), pre.numel(), _st(self.dev)), "ldit_gelu")
        return h

    def gelu_bwd(self, dh, pre):
        out = torch.empty_like(pre)
        _lib.check(self.lib.ldit_gelu_bwd(dh.data_ptr(), pre.data_ptr(), out.data_ptr(), pre.numel(), _st(self.dev)), "ldit_gelu_bwd")
        return out

    def scale_residual(self, x, branch, lam, row_scale=None, rows_per_image=1):
        """y = x + row_scale[image] * lam (.) branch  (row_scale: the drop-path factor of each image, or None)."""
        y = torch.empty_like(x)
        _lib.check(self.lib.ldit_scale_residual_rows(x.data_ptr(), branch.data_ptr(), _p(lam), _p(row_scale), rows_per_image, y.data_ptr(),
                                                     x.shape[0], x.shape[1], _st(self.dev)), "ldit_scale_residual_rows")
        return y

    def scale_residual_bwd(self, dy, branch, lam, dlam, row_scale=None, rows_per_image=1):
        out = torch.empty_like(branch)
        _lib.check(self.lib.ldit_scale_residual_rows_bwd(dy.data_ptr(), branch.data_ptr(), _p(lam), _p(row_scale), rows_per_image, out.data_ptr(),
                                                         _p(dlam), dy.shape[0], dy.shape[1], _st(self.dev)), "ldit_scale_residual_rows_bwd")
        return out

    def layernorm_bwd(self, x, w, dy, dx_in, dw, db, eps):
        out = torch.empty_like(x)
        _lib.check(self.lib.ldit_layernorm_bwd(x.data_ptr(), w.data_ptr(), dy.data_ptr(), _p(dx_in), out.data_ptr(), dw.data_ptr(), db.data_ptr(),
                                               x.shape[0], x.shape[1], eps, _st(self.dev)), "ldit_layernorm_bwd")
        return out


# order of the parameter tensors handed to BeitLayerFunction (lambda_1 / lambda_2 may be None)
PARAM_NAMES = ("ln1_w", "ln1_b", "wq", "bq", "wk", "wv", "bv", "wo", "bo", "lam1", "ln2_w", "ln2_b", "w1", "b1", "w2", "b2", "lam2",
               "rel_table")     # relative_position_bias_table [T, heads] (the layer's own or the shared one) or None


def layer_params(layer, shared_table=None) -> tuple:
    """The tensors of a ``dit_params._Layer`` (HF names) in ``PARAM_NAMES`` order; ``shared_table``: the encoder's shared
    relative-position table when the configuration has one (HF:598-600)."""
    at = layer.attention.attention
    own = getattr(at, "relative_position_bias", None)
    table = own.relative_position_bias_table if own is not None else shared_table
    return _layer_params17(layer) + (table,)


def _layer_params17(layer) -> tuple:
    at = layer.attention.attention
    return (layer.layernorm_before.weight, layer.layernorm_before.bias, at.query.weight, at.query.bias, at.key.weight, at.value.weight,
            at.value.bias, layer.attention.output.dense.weight, layer.attention.output.dense.bias, layer.lambda_1,
            layer.layernorm_after.weight, layer.layernorm_after.bias, layer.intermediate.dense.weight, layer.intermediate.dense.bias,
            layer.output.dense.weight, layer.output.dense.bias, layer.lambda_2)


class BeitLayerFunction(torch.autograd.Function):
    """``y = BeitLayer(x)`` (HF:469-508, eval-mode arithmetic) with a hand-written backward.

    ``x``: fp32 ``[B*N, D]`` residual stream; ``geom = (B, N, heads, Gh, Gw, eps)``; ``drop``: ``None`` or f32 ``[2, B]``,
    the drop-path factors (0 or 1 / keep_prob, HF:61-73) of the attention and the MLP branch of every image; then the 17
    parameter tensors."""

    @staticmethod
    def forward(ctx, x, geom, drop, *params):
        B, N, heads, Gh, Gw, eps = geom
        d1 = d2 = None
        if drop is not None:
            drop = drop.detach().to(x.device, torch.float32).contiguous()
            d1, d2 = drop[0], drop[1]
        p = dict(zip(PARAM_NAMES, params))
        dev = x.device
        k = _K(dev)
        f32 = lambda t: None if t is None else t.detach().to(dev, torch.float32).contiguous()
        bf = lambda t: t.detach().to(dev, _BF).contiguous()
        x = x.detach().contiguous()
        wqkv = bf(torch.cat([p["wq"], p["wk"], p["wv"]], dim=0))
        bqkv = f32(torch.cat([p["bq"], torch.zeros_like(p["bq"]), p["bv"]], dim=0))      # key has no bias (HF:240)
        wo, w1, w2 = bf(p["wo"]), bf(p["w1"]), bf(p["w2"])
        lam1, lam2 = f32(p["lam1"]), f32(p["lam2"])
        a1 = k.layernorm(x, f32(p["ln1_w"]), f32(p["ln1_b"]), eps)
        qkv = k.gemm(a1, wqkv, bqkv)
        table = None
        if p["rel_table"] is not None:
            T = (2 * Gh - 1) * (2 * Gw - 1) + 3
            if p["rel_table"].shape[0] != T:
                raise NotImplementedError("training with a relative-position table runs at the table's native window "
                                          "(the bilinear window resize, HF:556-571, has no backward yet)")
            table = p["rel_table"].detach().to(dev, torch.float32).t().contiguous()     # [heads, T] as the kernels take it
        if N > _FLASH_ABOVE or table is not None:     # the self-contained backward kernel covers two 128-row tiles without a table;
            att, lse = k.attention_lse(qkv, B, N, heads, Gh, Gw, table)                 # otherwise keep the row statistics
        else:
            att, lse = k.attention(qkv, B, N, heads, Gh, Gw), x.new_empty(0)
        br1 = k.gemm(att, wo, f32(p["bo"]))
        xm = k.scale_residual(x, br1, lam1, d1, N)
        a2 = k.layernorm(xm, f32(p["ln2_w"]), f32(p["ln2_b"]), eps)
        pre = k.gemm(a2, w1, f32(p["b1"]))
        h = k.gelu(pre)
        br2 = k.gemm(h, w2, f32(p["b2"]))
        y = k.scale_residual(xm, br2, lam2, d2, N)
        ctx.geom = geom
        ctx.has_lam = (p["lam1"] is not None, p["lam2"] is not None)
        ctx.drop = drop
        ctx.table = table
        ctx.save_for_backward(lse, x, xm, a1, qkv, att, br1, a2, pre, h, br2, wqkv, wo, w1, w2,
                              f32(p["ln1_w"]), f32(p["ln2_w"]), lam1 if lam1 is not None else x.new_empty(0),
                              lam2 if lam2 is not None else x.new_empty(0))
        return y

    @staticmethod
    def backward(ctx, dy):
        B, N, heads, Gh, Gw, eps = ctx.geom
        (lse, x, xm, a1, qkv, att, br1, a2, pre, h, br2, wqkv, wo, w1, w2, g1w, g2w, lam1, lam2) = ctx.saved_tensors
        lam1 = lam1 if ctx.has_lam[0] else None
        lam2 = lam2 if ctx.has_lam[1] else None
        d1, d2 = (ctx.drop[0], ctx.drop[1]) if ctx.drop is not None else (None, None)
        dev = dy.device
        k = _K(dev)
        M, D = x.shape
        I = w1.shape[0]
        dy = dy.detach().to(torch.float32).contiguous()
        z = lambda *s: torch.zeros(*s, device=dev, dtype=torch.float32)

        # ---- MLP half: y = xm + lam2 (.) (h W2^T + b2)
        dlam2 = z(D) if lam2 is not None else None
        g2 = k.scale_residual_bwd(dy, br2, lam2, dlam2, d2, N)             # d(branch 2), bf16
        db2 = z(D); k.colsum(g2, db2)
        dw2 = z(D, I); k.wgrad(g2, h, dw2)                                 # [M, D]^T [M, I]
        dh = k.dgrad(g2, w2)                                               # [M, D] x W2 [D, I]
        dpre = k.gelu_bwd(dh, pre)
        db1 = z(I); k.colsum(dpre, db1)
        dw1 = z(I, D); k.wgrad(dpre, a2, dw1)
        da2 = k.dgrad(dpre, w1)                                            # [M, I] x W1 [I, D]
        dg2, dbt2 = z(D), z(D)
        dxm = k.layernorm_bwd(xm, g2w, da2, dy, dg2, dbt2, eps)            # dy (residual path) + LayerNorm-2 path

        # ---- attention half: xm = x + lam1 (.) (ctx Wo^T + bo)
        dlam1 = z(D) if lam1 is not None else None
        g1 = k.scale_residual_bwd(dxm, br1, lam1, dlam1, d1, N)
        dbo = z(D); k.colsum(g1, dbo)
        dwo = z(D, D); k.wgrad(g1, att, dwo)
        datt = k.dgrad(g1, wo)
        dtable = None
        if ctx.table is not None:
            dtab = torch.zeros_like(ctx.table)
            dqkv = k.attention_bwd_flash(qkv, att, lse, datt, B, N, heads, Gh, Gw, ctx.table, dtab)
            dtable = dtab.t()                                              # back to HF's [T, heads]
        elif N <= _FLASH_ABOVE:
            dqkv = k.attention_bwd(qkv, datt, B, N, heads)
        else:
            dqkv = k.attention_bwd_flash(qkv, att, lse, datt, B, N, heads, Gh, Gw)
        dbqkv = z(3 * D); k.colsum(dqkv, dbqkv)
        dwqkv = z(3 * D, D); k.wgrad(dqkv, a1, dwqkv)
        da1 = k.dgrad(dqkv, wqkv)
        dg1, dbt1 = z(D), z(D)
        dx = k.layernorm_bwd(x, g1w, da1, dxm, dg1, dbt1, eps)

        grads = dict(ln1_w=dg1, ln1_b=dbt1, wq=dwqkv[:D], bq=dbqkv[:D], wk=dwqkv[D:2 * D], wv=dwqkv[2 * D:], bv=dbqkv[2 * D:],
                     wo=dwo, bo=dbo, lam1=dlam1, ln2_w=dg2, ln2_b=dbt2, w1=dw1, b1=db1, w2=dw2, b2=db2, lam2=dlam2, rel_table=dtable)
        return (dx, None, None) + tuple(grads[n] for n in PARAM_NAMES)


class TrainableEncoder(nn.Module):
    """``BeitEncoder`` layers (HF:594-663) over a ``DiTParameters`` tree, differentiable through the kernels above.
    Input / output: the fp32 residual stream ``[B, N, D]`` (``hidden_states[0]`` -> ``hidden_states[L]``)."""

    def __init__(self, params: DiTParameters, cfg: DiTConfig, drop_path_rate: float = 0.0):
        super().__init__()
        if not 0.0 <= drop_path_rate < 1.0:
            raise ValueError("drop_path_rate must be in [0, 1)")
        self.params_tree, self.cfg = params, cfg
        # stochastic depth as HF builds it: rate i of L grows linearly from 0 to drop_path_rate (HF:602-604), active in train()
        L = cfg.num_hidden_layers
        self.drop_rates = [drop_path_rate * i / max(L - 1, 1) for i in range(L)]

    def shared_table(self):
        """The encoder-level relative-position table (HF:598-600) or None."""
        rp = getattr(self.params_tree.encoder, "relative_position_bias", None)
        return None if rp is None else rp.relative_position_bias_table

    def drop_factors(self, i: int, B: int, device):
        """f32 [2, B] drop-path factors of layer i for this step (HF:61-73: floor(keep + U[0,1)) / keep), or None."""
        p = self.drop_rates[i]
        if not self.training or p == 0.0:
            return None
        keep = 1.0 - p
        return torch.floor(keep + torch.rand(2, B, device=device)) / keep

    def forward(self, hidden: torch.Tensor, Gh: int, Gw: int) -> torch.Tensor:
        B, N, D = hidden.shape
        if N != Gh * Gw + 1:
            raise ValueError("hidden must be [B, Gh*Gw + 1, D]")
        if not hidden.is_cuda:
            raise _lib.LditError("TrainableEncoder needs CUDA tensors (there is no CPU path)")
        geom = (B, N, self.cfg.num_attention_heads, Gh, Gw, float(self.cfg.layer_norm_eps))
        x = hidden.reshape(B * N, D).float()
        for i, layer in enumerate(self.params_tree.encoder.layer):
            x = BeitLayerFunction.apply(x, geom, self.drop_factors(i, B, x.device), *layer_params(layer, self.shared_table()))
        return x.reshape(B, N, D)


# ----------------------------------------------------------------------------------- embeddings and taps
class PatchEmbedFunction(torch.autograd.Function):
    """``BeitEmbeddings.forward`` (HF:161-184, ``bool_masked_pos=None``, dropout 0) at the native grid:
    ``x [B*N, D] f32 = cat(cls, conv16(pixels)) + position_embeddings``, differentiable in the projection weight / bias,
    the CLS token and the position table (not in the pixels).  Forward = the inference entry point ``ldit_patch_embed``
    (fp32 pages: gather pass + CLS rows + GEMM); backward: the projection's wgrad is the same transposed-copy GEMM as the
    layers', its A operand the im2col matrix the forward left in its scratch."""

    @staticmethod
    def forward(ctx, pixels, geom, w, b, cls, pos):
        B, H, W, D = geom
        Gh, Gw = H // 16, W // 16
        P, N = Gh * Gw, Gh * Gw + 1
        dev = pixels.device
        k = _K(dev)
        st = _st(dev)
        px = pixels.detach().to(torch.float32).contiguous()
        wb = w.detach().reshape(D, -1).to(dev, _BF).contiguous()
        bias = b.detach().to(dev, torch.float32).contiguous()
        clsv = cls.detach().reshape(D).to(dev, torch.float32).contiguous()
        resize = None
        if pos is not None:
            g = int(round((pos.shape[1] - 1) ** 0.5))
            if g * g + 1 != pos.shape[1]:
                raise ValueError("position table must hold a square grid + the CLS row")
            native = (Gh * Gw == g * g and H == W)             # HF:138-141
            gh, gw = (Gh, Gw) if native else (g, g)
            pt = pos.detach().to(dev, torch.float32).contiguous()
            pos_bias = torch.empty(P, D, device=dev, dtype=torch.float32)
            cls_pos = torch.empty(D, device=dev, dtype=torch.float32)
            _lib.check(k.lib.ldit_resize_rows(pt[0, 1:].data_ptr(), pos_bias.data_ptr(), bias.data_ptr(), gh, gw, Gh, Gw, D, 1, st), "ldit_resize_rows")
            _lib.check(k.lib.ldit_resize_rows(pt[0, :1].data_ptr(), cls_pos.data_ptr(), clsv.data_ptr(), 1, 1, 1, 1, D, 1, st), "ldit_resize_rows")
            if not native:
                # the bicubic resize (HF:143-159) is linear in the table: its matrix R [P, g*g] is the resize of the identity
                # (one unit image per native cell); the backward applies R^T
                eye = torch.eye(g * g, device=dev, dtype=torch.float32)
                resize = torch.empty(P, g * g, device=dev, dtype=torch.float32)
                _lib.check(k.lib.ldit_resize_rows(eye.data_ptr(), resize.data_ptr(), None, g, g, Gh, Gw, g * g, 1, st), "ldit_resize_rows")
        else:
            pos_bias, cls_pos = bias.expand(P, D).contiguous(), clsv
        scratch = torch.empty(k.lib.ldit_patch_embed_scratch_bytes(B, H, W) // 2, device=dev, dtype=_BF)
        x = torch.empty(B * N, D, device=dev, dtype=torch.float32)
        _lib.check(k.lib.ldit_patch_embed(px.data_ptr(), _lib.DTYPE_F32, wb.data_ptr(), pos_bias.data_ptr(), cls_pos.data_ptr(),
                                          scratch.data_ptr(), x.data_ptr(), B, H, W, D, st), "ldit_patch_embed")
        ctx.geom, ctx.has_pos, ctx.wshape = (B, Gh, Gw, D), pos is not None, tuple(w.shape)
        ctx.save_for_backward(scratch, resize if resize is not None else x.new_empty(0))
        return x

    @staticmethod
    def backward(ctx, dx):
        B, Gh, Gw, D = ctx.geom
        P, N = Gh * Gw, Gh * Gw + 1
        scratch, resize = ctx.saved_tensors
        dev = dx.device
        k = _K(dev)
        st = _st(dev)
        dx = dx.detach().to(torch.float32).contiguous()
        a = scratch[: B * P * 768].view(B * P, 768)                        # im2col(pixels), bf16
        dtok = torch.empty(B * P, D, device=dev, dtype=_BF)                # patch rows of dx, bf16 (scale 1 = CLS-less cast)
        _lib.check(k.lib.ldit_resample_taps(dx.data_ptr(), dtok.data_ptr(), B, Gh, Gw, D, 1.0, st), "ldit_resample_taps")
        db = torch.zeros(D, device=dev, dtype=torch.float32)
        k.colsum(dtok, db)
        dw = torch.zeros(D, 768, device=dev, dtype=torch.float32)
        k.wgrad(dtok, a, dw)                                               # [B P, D]^T [B P, 768]
        dsum = torch.zeros(N * D, device=dev, dtype=torch.float32)         # every image adds the same cls / position rows
        _lib.check(k.lib.ldit_batch_sum(dx.data_ptr(), dsum.data_ptr(), B, N * D, st), "ldit_batch_sum")
        dcls = dsum[:D].clone().view(1, 1, D)
        dpos = None
        if ctx.has_pos:
            dpos = dsum.view(1, N, D)
            if resize.numel():   # non-native grid: rows 1.. go back through the resize (a [g*g, P] x [P, D] product, fp32)
                dpos = torch.cat([dpos[:, :1], (resize.t() @ dpos[0, 1:]).unsqueeze(0)], dim=1)
        return None, None, dw.view(ctx.wshape), db, dcls, dpos


class TapFunction(torch.autograd.Function):
    """One feature tap (R:dit_backbone.py:50-61): patch rows of a hidden state -> ``[B, D, floor(Gh s), floor(Gw s)]``
    bf16 (channels-last memory), bilinear with ``scale_factor = s``; backward = ``ldit_resample_taps_bwd``."""

    @staticmethod
    def forward(ctx, x, geom):
        B, Gh, Gw, D, scale = geom
        dev = x.device
        oh, ow = int(Gh * scale), int(Gw * scale)
        out = torch.empty(B, oh, ow, D, device=dev, dtype=_BF)
        xc = x.detach().contiguous()
        _lib.check(_lib.load().ldit_resample_taps(xc.data_ptr(), out.data_ptr(), B, Gh, Gw, D, float(scale), _st(dev)), "ldit_resample_taps")
        ctx.geom = geom
        return out.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, dout):
        B, Gh, Gw, D, scale = ctx.geom
        dev = dout.device
        dn = dout.detach().permute(0, 2, 3, 1).to(_BF).contiguous()
        dx = torch.zeros(B * (Gh * Gw + 1), D, device=dev, dtype=torch.float32)
        _lib.check(_lib.load().ldit_resample_taps_bwd(dn.data_ptr(), dx.data_ptr(), B, Gh, Gw, D, float(scale), _st(dev)), "ldit_resample_taps_bwd")
        return dx, None


class TrainableBackbone(nn.Module):
    """``DiTBackbone.forward`` (R:dit_backbone.py:38-62) differentiable end to end through the library's kernels:
    embeddings -> L x ``BeitLayerFunction`` -> four taps.  Same outputs as the inference module
    (``OrderedDict{p2..p5}``, bf16, channels-last); same restrictions as ``TrainableEncoder`` plus the native grid."""

    SCALES = (4.0, 2.0, 1.0, 0.5)

    def __init__(self, params: DiTParameters, cfg: DiTConfig, drop_path_rate: float = 0.0):
        super().__init__()
        self.encoder = TrainableEncoder(params, cfg, drop_path_rate)     # validates the configuration, owns the drop-path schedule
        self.params_tree, self.cfg = params, cfg
        d = cfg.num_hidden_layers
        self.layer_idxs = [d // 3, d // 2, 2 * d // 3, d]

    def forward(self, pixels: torch.Tensor):
        from collections import OrderedDict
        if pixels.dim() != 4 or pixels.shape[1] != 3 or pixels.shape[2] % 16 or pixels.shape[3] % 16:
            raise ValueError("pixels must be [B, 3, H, W] with H and W multiples of 16")
        if not pixels.is_cuda:
            raise _lib.LditError("TrainableBackbone needs CUDA tensors (there is no CPU path)")
        B, _, H, W = pixels.shape
        cfg, e = self.cfg, self.params_tree.embeddings
        D, Gh, Gw = cfg.hidden_size, H // 16, W // 16
        pos = getattr(e, "position_embeddings", None)
        x = PatchEmbedFunction.apply(pixels, (B, H, W, D), e.patch_embeddings.projection.weight, e.patch_embeddings.projection.bias,
                                     e.cls_token, pos)
        geom = (B, Gh * Gw + 1, cfg.num_attention_heads, Gh, Gw, float(cfg.layer_norm_eps))
        taps = {}
        for i, layer in enumerate(self.params_tree.encoder.layer, start=1):
            x = BeitLayerFunction.apply(x, geom, self.encoder.drop_factors(i - 1, B, x.device), *layer_params(layer, self.encoder.shared_table()))
            for j, idx in enumerate(self.layer_idxs):          # shallow models tap one layer more than once
                if idx == i:
                    taps[j] = TapFunction.apply(x, (B, Gh, Gw, D, self.SCALES[j]))
        return OrderedDict((f"p{j + 2}", taps[j]) for j in range(4))


class TrainableDiTWithFPN(nn.Module):
    """The training-time counterpart of ``DiTWithFPN`` (R:dit_backbone.py:65-95): ``.backbone`` is the differentiable
    ``TrainableBackbone`` over a ``DiTParameters`` tree exposed as ``.backbone.dit`` (the attribute the reference loads
    checkpoints into, R:model.py:70), ``.fpn`` is torchvision's own ``FeaturePyramidNetwork`` + ``LastLevelMaxPool`` under
    torch autograd (the library's FPN kernels are inference-only so far), ``out_channels = 256``.  It is what
    ``FasterRCNN(backbone=...)`` takes in ``R:model.py:33-56`` when the detector is trained."""

    class _Backbone(nn.Module):
        def __init__(self, cfg, drop_path_rate):
            super().__init__()
            self.dit = DiTParameters(cfg)
            self.net = TrainableBackbone(self.dit, cfg, drop_path_rate)

        def forward(self, x):
            return self.net(x)

    def __init__(self, cfg: DiTConfig, out_channels: int = 256, drop_path_rate: float = 0.0):
        super().__init__()
        from torchvision.ops import FeaturePyramidNetwork
        from torchvision.ops.feature_pyramid_network import LastLevelMaxPool
        self.backbone = self._Backbone(cfg, drop_path_rate)
        self.fpn = FeaturePyramidNetwork([cfg.hidden_size] * 4, out_channels, extra_blocks=LastLevelMaxPool())
        self.out_channels = out_channels

    def forward(self, x):
        return self.fpn(self.backbone(x))


# ----------------------------------------------------------------------------------- data-parallel gradients
class GradientBuckets:
    """Bucketed gradient all-reduce of a data-parallel step (BASELINE config 5): parameters are packed, in reverse
    registration order (the order their gradients become ready in a backward), into flat fp32 buckets of at most
    ``bucket_bytes``; ``all_reduce()`` launches one asynchronous ``dist.all_reduce`` per bucket (NCCL over NVLink on the
    GPUs, gloo in the CPU tests), divides by the world size and scatters the averages back into ``.grad``.
    Parameters without a gradient contribute zeros (every rank must reduce the same buckets)."""

    def __init__(self, params, bucket_bytes: int = 25 << 20, group=None, overlap: bool = False):
        """``overlap=True``: every parameter gets a post-accumulate-grad hook and a bucket's all-reduce is launched the
        moment its last gradient of the step has been written, i.e. under the rest of the backward (what DDP does);
        ``all_reduce()`` then launches whatever is left, waits and scatters the averages back."""
        self.params = [p for p in params if p.requires_grad][::-1]
        self.group = group
        self.buckets, cur, size = [], [], 0
        for p in self.params:
            nbytes = p.numel() * 4
            if cur and size + nbytes > bucket_bytes:
                self.buckets.append(cur); cur, size = [], 0
            cur.append(p); size += nbytes
        if cur:
            self.buckets.append(cur)
        self._flat = [None] * len(self.buckets)
        self._works = [None] * len(self.buckets)
        self._pending = [len(b) for b in self.buckets]
        self._bucket_of = {id(p): i for i, b in enumerate(self.buckets) for p in b}
        self._hooks = []
        if overlap:
            for p in self.params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))

    def _active(self):
        return dist.is_initialized() and dist.get_world_size(self.group) > 1

    def _on_grad(self, p):
        i = self._bucket_of[id(p)]
        self._pending[i] -= 1
        if self._pending[i] == 0 and self._active():
            self._launch(i)

    def _launch(self, i):
        bucket = self.buckets[i]
        dev = bucket[0].device
        n = sum(p.numel() for p in bucket)
        if self._flat[i] is None or self._flat[i].device != dev:
            self._flat[i] = torch.empty(n, device=dev, dtype=torch.float32)
        flat, off = self._flat[i], 0
        for p in bucket:
            seg = flat[off: off + p.numel()]
            if p.grad is None:
                seg.zero_()
            else:
                seg.copy_(p.grad.detach().reshape(-1))
            off += p.numel()
        self._works[i] = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def all_reduce(self):
        if not self._active():
            self._pending = [len(b) for b in self.buckets]
            return
        world = dist.get_world_size(self.group)
        for i in range(len(self.buckets)):
            if self._works[i] is None:          # not launched by a hook (no overlap, or a parameter without a gradient this step)
                self._launch(i)
        for i, bucket in enumerate(self.buckets):
            self._works[i].wait()
            flat, off = self._flat[i], 0
            flat.div_(world)
            for p in bucket:
                seg = flat[off: off + p.numel()].view_as(p)
                if p.grad is None:
                    p.grad = seg.clone()
                else:
                    p.grad.copy_(seg)
                off += p.numel()
            self._works[i] = None
        self._pending = [len(b) for b in self.buckets]
