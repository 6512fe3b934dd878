"""Training path, first vertical slice: one ``BeitLayer`` forward + backward as a
``torch.autograd.Function`` over the C ABI, an encoder that chains such layers, and the bucketed gradient all-reduce of a
data-parallel step.

What the reference does here: ``LayoutDetectionModel`` is trained with ``torch.autograd`` through HF ``BeitLayer``
(HF:469-508) under autocast, ``scaler.scale(loss).backward()`` (R:src/layoutdit/training/trainer.py:164-183); it has no
distributed code (R:README.md:59) -- BASELINE config 5 asks for DDP with an NCCL gradient all-reduce, which is new.

Scope of this slice (the rest is listed in DESIGN.md "next"): absolute-position configurations (no relative-position
bias: its table gradient is not written yet), drop-path rate 0 (HF's training-mode stochastic depth, HF:61-73, is a
per-sample Bernoulli scaling of the two branches), sequences of at most 256 tokens (the attention backward is a
correctness-first CUDA-core kernel; 224 x 224 pages have 197).  The eight GEMMs of a layer's backward run on the
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

    def attention(self, qkv, B, N, heads, Gh, Gw):
        D = heads * 64
        ctx = torch.empty(B * N, D, device=self.dev, dtype=_BF)
        _lib.check(self.lib.ldit_attention(qkv.data_ptr(), ctx.data_ptr(), None, B, N, heads, Gh, Gw, _st(self.dev)), "ldit_attention")
        return ctx

    def attention_bwd(self, qkv, dctx, B, N, heads):
        dqkv = torch.empty_like(qkv)
        _lib.check(self.lib.ldit_attention_bwd(qkv.data_ptr(), dctx.data_ptr(), dqkv.data_ptr(), B, N, heads, _st(self.dev)), "ldit_attention_bwd")
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

    def scale_residual(self, x, branch, lam):
        y = torch.empty_like(x)
        _lib.check(self.lib.ldit_scale_residual(x.data_ptr(), branch.data_ptr(), _p(lam), y.data_ptr(), x.shape[0], x.shape[1], _st(self.dev)),
                   "ldit_scale_residual")
        return y

    def scale_residual_bwd(self, dy, branch, lam, dlam):
        out = torch.empty_like(branch)
        _lib.check(self.lib.ldit_scale_residual_bwd(dy.data_ptr(), branch.data_ptr(), _p(lam), out.data_ptr(), _p(dlam), dy.shape[0], dy.shape[1],
                                                    _st(self.dev)), "ldit_scale_residual_bwd")
        return out

    def layernorm_bwd(self, x, w, dy, dx_in, dw, db, eps):
        out = torch.empty_like(x)
        _lib.check(self.lib.ldit_layernorm_bwd(x.data_ptr(), w.data_ptr(), dy.data_ptr(), _p(dx_in), out.data_ptr(), dw.data_ptr(), db.data_ptr(),
                                               x.shape[0], x.shape[1], eps, _st(self.dev)), "ldit_layernorm_bwd")
        return out


# order of the parameter tensors handed to BeitLayerFunction (lambda_1 / lambda_2 may be None)
PARAM_NAMES = ("ln1_w", "ln1_b", "wq", "bq", "wk", "wv", "bv", "wo", "bo", "lam1", "ln2_w", "ln2_b", "w1", "b1", "w2", "b2", "lam2")


def layer_params(layer) -> tuple:
    """The 17 tensors of a ``dit_params._Layer`` (HF names) in ``PARAM_NAMES`` order."""
    at = layer.attention.attention
    return (layer.layernorm_before.weight, layer.layernorm_before.bias, at.query.weight, at.query.bias, at.key.weight, at.value.weight,
            at.value.bias, layer.attention.output.dense.weight, layer.attention.output.dense.bias, layer.lambda_1,
            layer.layernorm_after.weight, layer.layernorm_after.bias, layer.intermediate.dense.weight, layer.intermediate.dense.bias,
            layer.output.dense.weight, layer.output.dense.bias, layer.lambda_2)


class BeitLayerFunction(torch.autograd.Function):
    """``y = BeitLayer(x)`` (HF:469-508, eval-mode arithmetic) with a hand-written backward.

    ``x``: fp32 ``[B*N, D]`` residual stream; ``geom = (B, N, heads, Gh, Gw, eps)``; then the 17 parameter tensors."""

    @staticmethod
    def forward(ctx, x, geom, *params):
        B, N, heads, Gh, Gw, eps = geom
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
        att = k.attention(qkv, B, N, heads, Gh, Gw)
        br1 = k.gemm(att, wo, f32(p["bo"]))
        xm = k.scale_residual(x, br1, lam1)
        a2 = k.layernorm(xm, f32(p["ln2_w"]), f32(p["ln2_b"]), eps)
        pre = k.gemm(a2, w1, f32(p["b1"]))
        h = k.gelu(pre)
        br2 = k.gemm(h, w2, f32(p["b2"]))
        y = k.scale_residual(xm, br2, lam2)
        ctx.geom = geom
        ctx.has_lam = (p["lam1"] is not None, p["lam2"] is not None)
        ctx.save_for_backward(x, xm, a1, qkv, att, br1, a2, pre, h, br2, wqkv, wo, w1, w2,
                              f32(p["ln1_w"]), f32(p["ln2_w"]), lam1 if lam1 is not None else x.new_empty(0),
                              lam2 if lam2 is not None else x.new_empty(0))
        return y

    @staticmethod
    def backward(ctx, dy):
        B, N, heads, Gh, Gw, eps = ctx.geom
        (x, xm, a1, qkv, att, br1, a2, pre, h, br2, wqkv, wo, w1, w2, g1w, g2w, lam1, lam2) = ctx.saved_tensors
        lam1 = lam1 if ctx.has_lam[0] else None
        lam2 = lam2 if ctx.has_lam[1] else None
        dev = dy.device
        k = _K(dev)
        M, D = x.shape
        I = w1.shape[0]
        dy = dy.detach().to(torch.float32).contiguous()
        z = lambda *s: torch.zeros(*s, device=dev, dtype=torch.float32)
        tr = lambda w: w.t().contiguous()                                  # weight transposes: data movement

        # ---- MLP half: y = xm + lam2 (.) (h W2^T + b2)
        dlam2 = z(D) if lam2 is not None else None
        g2 = k.scale_residual_bwd(dy, br2, lam2, dlam2)                    # d(branch 2), bf16
        db2 = z(D); k.colsum(g2, db2)
        dw2 = z(D, I); k.gemm_acc(k.transpose(g2), k.transpose(h), dw2)    # [D, M] x [I, M]^T
        dh = k.gemm(g2, tr(w2))                                            # [M, D] x (W2^T [I, D])^T
        dpre = k.gelu_bwd(dh, pre)
        db1 = z(I); k.colsum(dpre, db1)
        dw1 = z(I, D); k.gemm_acc(k.transpose(dpre), k.transpose(a2), dw1)
        da2 = k.gemm(dpre, tr(w1))                                         # [M, I] x (W1^T [D, I])^T
        dg2, dbt2 = z(D), z(D)
        dxm = k.layernorm_bwd(xm, g2w, da2, dy, dg2, dbt2, eps)            # dy (residual path) + LayerNorm-2 path

        # ---- attention half: xm = x + lam1 (.) (ctx Wo^T + bo)
        dlam1 = z(D) if lam1 is not None else None
        g1 = k.scale_residual_bwd(dxm, br1, lam1, dlam1)
        dbo = z(D); k.colsum(g1, dbo)
        dwo = z(D, D); k.gemm_acc(k.transpose(g1), k.transpose(att), dwo)
        datt = k.gemm(g1, tr(wo))
        dqkv = k.attention_bwd(qkv, datt, B, N, heads)
        dbqkv = z(3 * D); k.colsum(dqkv, dbqkv)
        dwqkv = z(3 * D, D); k.gemm_acc(k.transpose(dqkv), k.transpose(a1), dwqkv)
        da1 = k.gemm(dqkv, tr(wqkv))
        dg1, dbt1 = z(D), z(D)
        dx = k.layernorm_bwd(x, g1w, da1, dxm, dg1, dbt1, eps)

        grads = dict(ln1_w=dg1, ln1_b=dbt1, wq=dwqkv[:D], bq=dbqkv[:D], wk=dwqkv[D:2 * D], wv=dwqkv[2 * D:], bv=dbqkv[2 * D:],
                     wo=dwo, bo=dbo, lam1=dlam1, ln2_w=dg2, ln2_b=dbt2, w1=dw1, b1=db1, w2=dw2, b2=db2, lam2=dlam2)
        return (dx, None) + tuple(grads[n] for n in PARAM_NAMES)


class TrainableEncoder(nn.Module):
    """``BeitEncoder`` layers (HF:594-663) over a ``DiTParameters`` tree, differentiable through the kernels above.
    Input / output: the fp32 residual stream ``[B, N, D]`` (``hidden_states[0]`` -> ``hidden_states[L]``)."""

    def __init__(self, params: DiTParameters, cfg: DiTConfig):
        super().__init__()
        if cfg.use_relative_position_bias or cfg.use_shared_relative_position_bias:
            raise NotImplementedError("the backward slice covers absolute-position configurations (no relative-position bias yet)")
        self.params_tree, self.cfg = params, cfg

    def forward(self, hidden: torch.Tensor, Gh: int, Gw: int) -> torch.Tensor:
        B, N, D = hidden.shape
        if N != Gh * Gw + 1:
            raise ValueError("hidden must be [B, Gh*Gw + 1, D]")
        if not hidden.is_cuda:
            raise _lib.LditError("TrainableEncoder needs CUDA tensors (there is no CPU path)")
        geom = (B, N, self.cfg.num_attention_heads, Gh, Gw, float(self.cfg.layer_norm_eps))
        x = hidden.reshape(B * N, D).float()
        for layer in self.params_tree.encoder.layer:
            x = BeitLayerFunction.apply(x, geom, *layer_params(layer))
        return x.reshape(B, N, D)


# ----------------------------------------------------------------------------------- data-parallel gradients
class GradientBuckets:
    """Bucketed gradient all-reduce of a data-parallel step (BASELINE config 5): parameters are packed, in reverse
    registration order (the order their gradients become ready in a backward), into flat fp32 buckets of at most
    ``bucket_bytes``; ``all_reduce()`` launches one asynchronous ``dist.all_reduce`` per bucket (NCCL over NVLink on the
    GPUs, gloo in the CPU tests), divides by the world size and scatters the averages back into ``.grad``.
    Parameters without a gradient contribute zeros (every rank must reduce the same buckets)."""

    def __init__(self, params, bucket_bytes: int = 25 << 20, group=None):
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

    def all_reduce(self):
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return
        world = dist.get_world_size(self.group)
        works = []
        for i, bucket in enumerate(self.buckets):
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
            works.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        for i, (bucket, work) in enumerate(zip(self.buckets, works)):
            work.wait()
            flat, off = self._flat[i], 0
            flat.div_(world)
            for p in bucket:
                seg = flat[off: off + p.numel()].view_as(p)
                if p.grad is None:
                    p.grad = seg.clone()
                else:
                    p.grad.copy_(seg)
                off += p.numel()
