"""Host-side orchestration of the backbone forward: weight packing, per-geometry
workspaces, and the launch sequence over the C ABI (include/ldit.h).

PyTorch is used here for device memory, streams and CUDA-graph capture only; every
arithmetic step is a kernel in libldit_b200.so -- including the weight-preparation steps
executed once per (weights, H, W): resizing the position table (HF:138-159) and the
relative-position bias tables (HF:556-571) to the current patch grid (``ldit_resize_rows``).
torch only casts, concatenates and transposes parameters into the kernels' layouts.
"""
from __future__ import annotations

import contextlib
import dataclasses
import math
import os
from collections import OrderedDict
from dataclasses import dataclass, field

import torch

from . import _lib
from .config import DiTConfig
from .dit_params import DiTParameters

_DTYPE_CODE = {torch.float32: _lib.DTYPE_F32, torch.float16: _lib.DTYPE_F16, torch.bfloat16: _lib.DTYPE_BF16}
TAP_SCALES = (4.0, 2.0, 1.0, 0.5)  # R:dit_backbone.py:35


def tap_layer_indices(num_layers: int):
    d = num_layers
    return [d // 3, d // 2, 2 * d // 3, d]  # R:dit_backbone.py:33-34


def _ptr(t):
    return None if t is None else t.data_ptr()


@dataclass
class _LayerPack:
    ln1_w: torch.Tensor
    ln1_b: torch.Tensor
    wqkv: torch.Tensor      # bf16 [3D, D]
    bqkv: torch.Tensor      # f32 [3D]  (key third is zero, HF:240)
    wo: torch.Tensor        # bf16 [D, D]
    bo: torch.Tensor
    lam1: torch.Tensor | None
    ln2_w: torch.Tensor
    ln2_b: torch.Tensor
    w1: torch.Tensor        # bf16 [I, D]
    b1: torch.Tensor
    w2: torch.Tensor        # bf16 [D, I]
    b2: torch.Tensor
    lam2: torch.Tensor | None
    rel_table: torch.Tensor | None   # f32 [T0, heads] native-window table (per-layer), or None


@dataclass
class _Geometry:
    """Everything that depends on (B, H, W): workspaces and resized tables."""
    B: int
    H: int
    W: int
    Gh: int
    Gw: int
    N: int
    M: int
    x: torch.Tensor          # f32 [M, D] residual stream
    a: torch.Tensor          # bf16 [M, D] LayerNorm output / attention context
    big: torch.Tensor        # bf16 [M, max(3D, I, 768)] QKV, MLP hidden and im2col scratch (disjoint lifetimes)
    pos_bias: torch.Tensor   # f32 [P, D]
    cls_pos: torch.Tensor    # f32 [D]
    bias_tables: list        # per layer: f32 [heads, T] or None
    graph: object = None
    graph_in: torch.Tensor | None = None
    graph_out: OrderedDict | None = None
    launches: int = 0
    head: str = "taps"       # "taps": the four D-channel taps of DiTBackbone; "fpn": DiTWithFPN's p2..p5 + pool
    extra: dict = field(default_factory=dict)


@dataclass
class _PageList:
    """Raw pages for ``ldit_patch_embed_pages``: device arrays of pointers and (H, W), plus the tensors they
    point into (kept alive until the launch sequence has been enqueued)."""
    ptrs: torch.Tensor       # int64 [B] device
    hw: torch.Tensor         # int32 [B, 2] device
    max_w: int               # widest page (sizes the row staging of the gather kernel)
    dtype: torch.dtype
    mean: tuple
    std: tuple
    keep: list


@dataclass
class _FpnPack:
    """torchvision FeaturePyramidNetwork weights (TV:ops/feature_pyramid_network.py:104-124) in kernel layout."""
    w_lat: list      # 4 x bf16 [C, D]       inner_blocks[i][0].weight[:, :, 0, 0]
    b_lat: list      # 4 x f32 [C]
    w_out: list      # 4 x bf16 [C, 9*C]     layer_blocks[i][0].weight.permute(0, 2, 3, 1): (ky, kx, cin) columns
    b_out: list      # 4 x f32 [C]
    C: int


class Engine:
    def __init__(self, params: DiTParameters, cfg: DiTConfig):
        self.params = params
        self.fpn_params = None   # set by DiTWithFPN: holder of the torchvision-named FPN parameters
        self._fpn: _FpnPack | None = None
        self.cfg = cfg
        self.lib = _lib.load()
        self._pack_key = None
        self._layers: list[_LayerPack] = []
        self._geoms: "OrderedDict" = OrderedDict()   # LRU over (B, H, W[, slot, head]); see _remember()
        self._max_geoms = int(os.environ.get("LDIT_MAX_GEOMETRIES", "8"))
        self.device = None
        self.tap_idx = tap_layer_indices(cfg.num_hidden_layers)
        # fc1 + fc2 of a layer as one persistent kernel with a balanced tile schedule (ldit_mlp_fused): experimental,
        # measured slower; needs a library built with -DLDIT_EXPERIMENTAL
        self.mlp_fused = os.environ.get("LDIT_MLP_FUSED", "0") != "0" and bool(self.lib.ldit_has_experimental())
        # L2 access-policy window (persisting) over the fp32 residual stream on every launch (ldit_set_l2_window)
        # On by default (LDIT_L2_PERSIST=0 switches it off): measured -3 % step time at base224.  Process-level side
        # effect: grows the device's persisting-L2 set-aside to the window size (at most the device maximum, 79 MB).
        self.l2_persist = os.environ.get("LDIT_L2_PERSIST", "1") != "0"
        self._persist_cap = int(os.environ.get("LDIT_L2_PERSIST_CAP_MB", "64")) << 20
        self._persist_partial = os.environ.get("LDIT_L2_PERSIST_PARTIAL", "0") != "0"   # experiment: partial window for x > cap
        # Residual adds deferred into the LayerNorm that follows them: out-projection / fc2 end in a bf16 store of the
        # layer-scaled branch (ldit_gemm_bias_scale) and ldit_add_layernorm does x += branch; a = LN(x).  Only the last
        # layer's fc2 (no LayerNorm behind it) keeps the fp32 reduce-add epilogue.
        # Measured (same-box A/B, bench.py): base224 -1.6 % step time (out-projection 24.7 -> 17.5 us, fc2 52.9 -> 45.7,
        # each LayerNorm +4.6 us); base512 +2.4 % and large224 +3.3 %, where x and the branch no longer sit in the L2 window
        # and the add streams the residual through HBM twice.  Hence "auto": only when x and a are both inside the window.
        # LDIT_DEFER_RESID = 0 never, 1 auto (default), 2 always.
        self.defer_residual = int(os.environ.get("LDIT_DEFER_RESID", "1"))
        self.patch_tma = os.environ.get("LDIT_PATCH_TMA", "1") != "0"   # fp16 pages: TMA-fed patch GEMM (bf16 pages: inside the library)

    # ------------------------------------------------------------------ weight packing
    def _weights_key(self):
        """Changes whenever a parameter is written in place (``_version``), replaced (``data_ptr``) or moved.  The
        parameter list itself is cached: walking the module tree costs more than the 211 attribute reads."""
        ps = self.__dict__.get("_param_list")
        if ps is None or self.__dict__.get("_param_list_fpn") is not self.fpn_params:
            ps = list(self.params.parameters())
            if self.fpn_params is not None:
                ps += list(self.fpn_params.parameters())
            self._param_list, self._param_list_fpn = ps, self.fpn_params
        return tuple(p._version for p in ps), tuple(p.data_ptr() for p in ps)

    def refresh_weights(self, force: bool = False):
        key = self._weights_key()
        if not force and key == self._pack_key:
            return
        cfg, P = self.cfg, self.params
        dev = P.embeddings.cls_token.device
        if dev.type != "cuda":
            raise _lib.LditError("DiTBackbone parameters must live on a CUDA device (there is no CPU path)")
        f32 = lambda t: t.detach().to(dev, torch.float32).contiguous()
        bf16 = lambda t: t.detach().to(dev, torch.bfloat16).contiguous()
        with torch.no_grad():
            e = P.embeddings
            D = cfg.hidden_size
            self.w_patch = bf16(e.patch_embeddings.projection.weight.reshape(D, -1))
            # fp16 copy for fp16 pages (what the reference feeds on CUDA): the TMA-fed patch GEMM needs both operands in one type
            self.w_patch_f16 = e.patch_embeddings.projection.weight.detach().reshape(D, -1).to(dev, torch.float16).contiguous()
            self.b_patch = f32(e.patch_embeddings.projection.bias)
            self.cls = f32(e.cls_token.reshape(D))
            self.pos = None if e.position_embeddings is None else f32(e.position_embeddings)
            enc = P.encoder
            self.shared_table = (f32(enc.relative_position_bias.relative_position_bias_table)
                                 if cfg.use_shared_relative_position_bias else None)
            layers = []
            for L in enc.layer:
                at = L.attention.attention
                wqkv = torch.cat([at.query.weight, at.key.weight, at.value.weight], dim=0)
                bqkv = torch.cat([at.query.bias, torch.zeros_like(at.query.bias), at.value.bias], dim=0)
                layers.append(_LayerPack(
                    ln1_w=f32(L.layernorm_before.weight), ln1_b=f32(L.layernorm_before.bias),
                    wqkv=bf16(wqkv), bqkv=f32(bqkv),
                    wo=bf16(L.attention.output.dense.weight), bo=f32(L.attention.output.dense.bias),
                    lam1=None if L.lambda_1 is None else f32(L.lambda_1),
                    ln2_w=f32(L.layernorm_after.weight), ln2_b=f32(L.layernorm_after.bias),
                    w1=bf16(L.intermediate.dense.weight), b1=f32(L.intermediate.dense.bias),
                    w2=bf16(L.output.dense.weight), b2=f32(L.output.dense.bias),
                    lam2=None if L.lambda_2 is None else f32(L.lambda_2),
                    rel_table=(f32(at.relative_position_bias.relative_position_bias_table)
                               if cfg.use_relative_position_bias else None)))
            self._layers = layers
            if self.fpn_params is not None:
                F_ = self.fpn_params
                C = F_.out_channels
                self._fpn = _FpnPack(
                    w_lat=[bf16(m[0].weight.reshape(C, -1)) for m in F_.inner_blocks],
                    b_lat=[f32(m[0].bias) for m in F_.inner_blocks],
                    w_out=[bf16(m[0].weight.permute(0, 2, 3, 1).reshape(C, -1)) for m in F_.layer_blocks],
                    b_out=[f32(m[0].bias) for m in F_.layer_blocks], C=C)
        self._pack_key = key
        self._geoms.clear()  # resized tables and captured graphs hold the old weights
        self.device = dev

    # --------------------------------------------------------------- per-geometry state
    def _resize_rows(self, src, h, w, oh, ow, add=None, bicubic=False):
        """``ldit_resize_rows``: src f32 [h*w, C] -> f32 [oh*ow, C] (+ add[C]).  Weight preparation."""
        src = src.contiguous()
        C = src.shape[-1]
        dst = torch.empty(oh * ow, C, device=self.device, dtype=torch.float32)
        with self._on_device():
            _lib.check(self.lib.ldit_resize_rows(src.data_ptr(), dst.data_ptr(), _ptr(add), h, w, oh, ow, C, int(bicubic),
                                                 torch.cuda.current_stream(self.device).cuda_stream), "ldit_resize_rows")
        return dst

    def _pos_rows(self, Gh, Gw, H, W):
        """Position rows 1..P for this grid + the conv bias, and cls_token + position row 0.
        HF ``interpolate_pos_encoding`` (HF:121-159): native table when the patch count matches and the image is
        square, else the patch rows resized bicubically g x g -> Gh x Gw, the CLS row kept.  Once per (H, W)."""
        g, D = self.cfg.grid, self.cfg.hidden_size
        native = Gh * Gw == g * g and H == W
        patch_rows = self.pos[0, 1:]
        pos_bias = self._resize_rows(patch_rows, *((Gh, Gw, Gh, Gw) if native else (g, g, Gh, Gw)), add=self.b_patch, bicubic=True)
        cls_pos = self._resize_rows(self.pos[0, :1], 1, 1, 1, 1, add=self.cls).reshape(D)
        return pos_bias, cls_pos

    def _resized_table(self, table, Gh, Gw):
        """First half of ``BeitRelativePositionBias.forward`` (HF:550-571), once per (H, W):
        -> f32 [heads, (2Gh-1)(2Gw-1)+3]; the gather by index (HF:573-581) happens in-tile."""
        g = self.cfg.grid
        old = 2 * g - 1
        nh, nw = 2 * Gh - 1, 2 * Gw - 1
        new = self._resize_rows(table[: old * old], old, old, nh, nw)          # bilinear; exact copy at the native window
        return torch.cat([new, table[old * old:]], dim=0).t().contiguous()    # data movement only

    def _geometry(self, B, H, W, slot: int = 0, head: str = "taps") -> _Geometry:
        """Workspaces / tables / graph for one (B, H, W).  ``slot`` > 0 gives pipelined callers their own captured
        graph, graph input and static outputs; the workspaces (x, a, big, tables) are those of slot 0 -- every
        forward of this engine runs on the caller's one compute stream, so they are never live in two forwards at
        once, and sharing them keeps the L2 working set (and the persistence window) that of a single forward."""
        key = (B, H, W) if (slot == 0 and head == "taps") else (B, H, W, slot, head)
        geo = self._geoms.get(key)
        if geo is not None:
            self._geoms.move_to_end(key)
            return geo
        if slot != 0:
            base = self._geometry(B, H, W, 0, head)
            geo = dataclasses.replace(base, graph=None, graph_in=None, graph_out=None, launches=0, extra=dict(base.extra))
            return self._remember(key, geo)
        cfg, dev = self.cfg, self.device
        D, I = cfg.hidden_size, cfg.intermediate_size
        Gh, Gw = H // 16, W // 16
        P = Gh * Gw
        N = P + 1
        M = B * N
        with torch.no_grad():
            if self.pos is not None:
                pos_bias, cls_pos = self._pos_rows(Gh, Gw, H, W)
            else:
                pos_bias = self.b_patch.expand(P, D).contiguous()
                cls_pos = self.cls.clone()
            tables = []
            shared = None if self.shared_table is None else self._resized_table(self.shared_table, Gh, Gw)
            for L in self._layers:
                t = None if L.rel_table is None else self._resized_table(L.rel_table, Gh, Gw)
                tables.append(t if t is not None else shared)
        wide = max(3 * D, I, 768)  # 768 = im2col row (3*16*16): the patch-embed scratch lives here too
        # residual stream and LayerNorm-output / context buffer share one allocation, x first: the L2 persistence window
        # (ldit_set_l2_persist) is a single address range
        ws = torch.empty(int(self.lib.ldit_workspace_bytes(B, H, W, D, I)), device=dev, dtype=torch.uint8)
        up = lambda v: (v + 1023) // 1024 * 1024
        o_a = M * D * 4                      # a starts right behind x (not rounded: the window must be one range)
        o_big = up(M * D * 4) + up(M * D * 2)
        geo = _Geometry(B=B, H=H, W=W, Gh=Gh, Gw=Gw, N=N, M=M,
                        x=ws[: M * D * 4].view(torch.float32).view(M, D),
                        a=ws[o_a: o_a + M * D * 2].view(torch.bfloat16).view(M, D),
                        big=ws[o_big: o_big + M * wide * 2].view(torch.bfloat16),
                        pos_bias=pos_bias, cls_pos=cls_pos, bias_tables=tables, head=head)
        if self.mlp_fused:
            stride = int(self.lib.ldit_mlp_schedule(M, D, I, None, 0))
            if stride > 0:   # shapes the fused kernel is built for; otherwise the two-call form is used
                host = torch.empty(int(self.lib.ldit_mlp_clusters()) * stride, dtype=torch.int32)
                _lib.check(min(0, int(self.lib.ldit_mlp_schedule(M, D, I, host.data_ptr(), host.numel()))), "ldit_mlp_schedule")
                geo.extra["mlp_sched"] = (host.to(dev), stride)
                geo.extra["mlp_ready"] = torch.zeros(2 * ((M + 255) // 256), device=dev, dtype=torch.int32)
        if head in ("fpn", "fpn32"):
            C = self._fpn.C
            bf = dict(device=dev, dtype=torch.bfloat16)
            geo.extra["tok"] = torch.empty(B * P, D, **bf)                        # tap tokens without CLS, bf16
            geo.extra["lat"] = [torch.empty(B * P, C, **bf) for _ in TAP_SCALES]   # laterals on the token grid
            geo.extra["inner"] = [torch.empty(B, *self._tap_hw(geo, s), C, **bf) for s in TAP_SCALES]
        return self._remember(key, geo)

    def _remember(self, key, geo):
        """Bounded cache: every entry owns workspaces (M*D*6 + M*max(3D, I)*2 bytes), resized tables and possibly a
        captured graph with static inputs / outputs, so a stream of distinct page or batch sizes must not grow GPU
        memory without limit.  Least-recently-used geometries beyond LDIT_MAX_GEOMETRIES (default 8) are dropped."""
        self._geoms[key] = geo
        while len(self._geoms) > max(1, self._max_geoms):
            self._geoms.popitem(last=False)
        return geo

    # -------------------------------------------------------------------- launch sequence
    @staticmethod
    def _tap_hw(geo, s):
        return int(math.floor(geo.Gh * s)), int(math.floor(geo.Gw * s))

    def _alloc_outputs(self, geo: _Geometry):
        bf = dict(device=self.device, dtype=torch.bfloat16)
        if geo.head in ("fpn", "fpn32"):   # p2..p5 of the FPN + "pool" (TV:231-249), C channels each; "fpn32": fp32 maps
            C = self._fpn.C
            od = dict(device=self.device, dtype=torch.float32 if geo.head == "fpn32" else torch.bfloat16)
            outs = [torch.empty(geo.B, *self._tap_hw(geo, s), C, **od) for s in TAP_SCALES]
            h5, w5 = self._tap_hw(geo, TAP_SCALES[-1])
            return outs + [torch.empty(geo.B, (h5 + 1) // 2, (w5 + 1) // 2, C, **od)]
        return [torch.empty(geo.B, *self._tap_hw(geo, s), self.cfg.hidden_size, **bf) for s in TAP_SCALES]

    def _plan(self, geo: _Geometry, x, outs, stream: int):
        """The forward as an ordered list of (name, C-ABI function, args): one entry per
        library call, in stream order.  ``x`` is the page batch tensor, or a ``_PageList`` (raw pages of
        any size: the input transform is fused into the patch gather)."""
        lib, cfg = self.lib, self.cfg
        D, I, heads = cfg.hidden_size, cfg.intermediate_size, cfg.num_attention_heads
        M, N, B = geo.M, geo.N, geo.B
        xr, a, big = geo.x.data_ptr(), geo.a.data_ptr(), geo.big.data_ptr()
        eps = float(cfg.layer_norm_eps)
        if isinstance(x, _PageList):
            plan = [("ldit_patch_embed_pages", lib.ldit_patch_embed_pages,
                     (x.ptrs.data_ptr(), x.hw.data_ptr(), x.max_w, _DTYPE_CODE[x.dtype], *x.mean, *x.std, self.w_patch.data_ptr(),
                      geo.pos_bias.data_ptr(), geo.cls_pos.data_ptr(), big, xr, B, geo.H, geo.W, D, stream))]
        elif x.dtype == torch.float16 and self.patch_tma and lib.ldit_patch_embed_tma_preferred(B, geo.H, geo.W, D):
            plan = [("ldit_patch_embed_tma", lib.ldit_patch_embed_tma,
                     (x.data_ptr(), _DTYPE_CODE[x.dtype], self.w_patch_f16.data_ptr(), geo.pos_bias.data_ptr(),
                      geo.cls_pos.data_ptr(), xr, B, geo.H, geo.W, D, stream))]
        else:
            plan = [("ldit_patch_embed", lib.ldit_patch_embed,
                     (x.data_ptr(), _DTYPE_CODE[x.dtype], self.w_patch.data_ptr(), geo.pos_bias.data_ptr(),
                      geo.cls_pos.data_ptr(), big, xr, B, geo.H, geo.W, D, stream))]

        fpn = self._fpn if geo.head in ("fpn", "fpn32") else None
        conv_fn, conv_name = ((lib.ldit_conv3x3_bias_f32, "ldit_conv3x3_bias_f32") if geo.head == "fpn32"
                              else (lib.ldit_conv3x3_bias, "ldit_conv3x3_bias"))
        sub_fn, sub_name = ((lib.ldit_subsample2_f32, "ldit_subsample2_f32") if geo.head == "fpn32"
                            else (lib.ldit_subsample2, "ldit_subsample2"))

        def emit_tap(layer_no):
            # hidden_states[layer_no] is the residual stream right now (HF:628-630, 654-655)
            for slot, idx in enumerate(self.tap_idx):
                if idx != layer_no:
                    continue
                if fpn is None:
                    plan.append(("ldit_resample_taps", lib.ldit_resample_taps,
                                 (xr, outs[slot].data_ptr(), B, geo.Gh, geo.Gw, D, TAP_SCALES[slot], stream)))
                else:
                    # FPN head: the 1x1 lateral (TV:187) runs on the token grid, before R:57-59's resampling
                    # (both linear, they commute): CLS-less bf16 copy of the tokens, then a GEMM to C channels
                    tok, lat = geo.extra["tok"].data_ptr(), geo.extra["lat"][slot].data_ptr()
                    plan.append(("ldit_resample_taps", lib.ldit_resample_taps, (xr, tok, B, geo.Gh, geo.Gw, D, 1.0, stream)))
                    plan.append(("ldit_gemm_bias", lib.ldit_gemm_bias,
                                 (tok, fpn.w_lat[slot].data_ptr(), fpn.b_lat[slot].data_ptr(), lat, B * geo.Gh * geo.Gw, fpn.C, D, stream)))

        emit_tap(0)
        defer = "mlp_sched" not in geo.extra and (
            self.defer_residual == 2 or (self.defer_residual == 1 and self._persist_bytes(geo) >= geo.x.numel() * 6))
        nl = len(self._layers)
        for i, L in enumerate(self._layers):
            if defer:
                # x <- x + branch of the previous layer's fc2 (waiting in `a`) fused into LayerNorm 1, which overwrites `a`
                # in place; hidden_states[i] is complete only now, so its tap is emitted here
                if i == 0:
                    plan.append(("ldit_layernorm", lib.ldit_layernorm, (xr, L.ln1_w.data_ptr(), L.ln1_b.data_ptr(), a, M, D, eps, stream)))
                else:
                    plan.append(("ldit_add_layernorm", lib.ldit_add_layernorm, (xr, a, L.ln1_w.data_ptr(), L.ln1_b.data_ptr(), a, M, D, eps, stream)))
                    emit_tap(i)
                plan += [
                    ("ldit_gemm_bias", lib.ldit_gemm_bias, (a, L.wqkv.data_ptr(), L.bqkv.data_ptr(), big, M, 3 * D, D, stream)),
                    ("ldit_attention", lib.ldit_attention, (big, a, _ptr(geo.bias_tables[i]), B, N, heads, geo.Gh, geo.Gw, stream)),
                    # out-projection branch -> the (now dead) QKV buffer
                    ("ldit_gemm_bias_scale", lib.ldit_gemm_bias_scale, (a, L.wo.data_ptr(), L.bo.data_ptr(), _ptr(L.lam1), big, M, D, D, stream)),
                    ("ldit_add_layernorm", lib.ldit_add_layernorm, (xr, big, L.ln2_w.data_ptr(), L.ln2_b.data_ptr(), a, M, D, eps, stream)),
                ]
                # fc1 writes the MLP hidden over the branch it no longer needs; big is [M, max(3D, I)]: hidden at offset 0 would
                # overlap nothing live (the branch was consumed by the add above)
                plan.append(("ldit_gemm_bias_gelu", lib.ldit_gemm_bias_gelu, (a, L.w1.data_ptr(), L.b1.data_ptr(), big, M, I, D, stream)))
                if i + 1 < nl:   # fc2 branch -> `a` (LayerNorm 2's output is dead once fc1 has run)
                    plan.append(("ldit_gemm_bias_scale", lib.ldit_gemm_bias_scale, (big, L.w2.data_ptr(), L.b2.data_ptr(), _ptr(L.lam2), a, M, D, I, stream)))
                else:
                    plan.append(("ldit_gemm_bias_scale_residual", lib.ldit_gemm_bias_scale_residual,
                                 (big, L.w2.data_ptr(), L.b2.data_ptr(), _ptr(L.lam2), xr, M, D, I, stream)))
                    emit_tap(i + 1)
                continue
            plan += [
                ("ldit_layernorm", lib.ldit_layernorm, (xr, L.ln1_w.data_ptr(), L.ln1_b.data_ptr(), a, M, D, eps, stream)),
                ("ldit_gemm_bias", lib.ldit_gemm_bias, (a, L.wqkv.data_ptr(), L.bqkv.data_ptr(), big, M, 3 * D, D, stream)),
                ("ldit_attention", lib.ldit_attention,
                 (big, a, _ptr(geo.bias_tables[i]), B, N, heads, geo.Gh, geo.Gw, stream)),
                ("ldit_gemm_bias_scale_residual", lib.ldit_gemm_bias_scale_residual,
                 (a, L.wo.data_ptr(), L.bo.data_ptr(), _ptr(L.lam1), xr, M, D, D, stream)),
                ("ldit_layernorm", lib.ldit_layernorm, (xr, L.ln2_w.data_ptr(), L.ln2_b.data_ptr(), a, M, D, eps, stream)),
            ]
            if "mlp_sched" in geo.extra:
                sched, stride = geo.extra["mlp_sched"]
                plan.append(("ldit_mlp_fused", lib.ldit_mlp_fused,
                             (a, L.w1.data_ptr(), L.b1.data_ptr(), big, L.w2.data_ptr(), L.b2.data_ptr(), _ptr(L.lam2), xr, M, D, I,
                              sched.data_ptr(), stride, geo.extra["mlp_ready"].data_ptr(), stream)))
            else:
                plan += [
                    ("ldit_gemm_bias_gelu", lib.ldit_gemm_bias_gelu, (a, L.w1.data_ptr(), L.b1.data_ptr(), big, M, I, D, stream)),
                    ("ldit_gemm_bias_scale_residual", lib.ldit_gemm_bias_scale_residual,
                     (big, L.w2.data_ptr(), L.b2.data_ptr(), _ptr(L.lam2), xr, M, D, I, stream)),
                ]
            emit_tap(i + 1)
        if fpn is not None:
            # top-down pathway, coarsest level first (TV:181-193), then the 3x3 output convolutions and "pool"
            C, inner = fpn.C, geo.extra["inner"]
            for slot in range(len(TAP_SCALES) - 1, -1, -1):
                top = inner[slot + 1] if slot + 1 < len(TAP_SCALES) else None
                th, tw = (top.shape[1], top.shape[2]) if top is not None else (0, 0)
                plan.append(("ldit_fpn_merge", lib.ldit_fpn_merge,
                             (geo.extra["lat"][slot].data_ptr(), _ptr(top), inner[slot].data_ptr(), B, geo.Gh, geo.Gw, C,
                              TAP_SCALES[slot], th, tw, stream)))
            for slot in range(len(TAP_SCALES)):
                h, w = inner[slot].shape[1], inner[slot].shape[2]
                plan.append((conv_name, conv_fn,
                             (inner[slot].data_ptr(), fpn.w_out[slot].data_ptr(), fpn.b_out[slot].data_ptr(),
                              outs[slot].data_ptr(), B, h, w, C, C, stream)))
            h5, w5 = outs[3].shape[1], outs[3].shape[2]
            plan.append((sub_name, sub_fn, (outs[3].data_ptr(), outs[4].data_ptr(), B, h5, w5, C, stream)))
        return plan

    def _persist_bytes(self, geo: _Geometry) -> int:
        """Size of the L2 persistence window for this geometry, 0 = none.  The set-aside comes out of the 126 MB L2
        every other buffer lives in, so it only pays while it stays a modest part of it (measured: 58 MB over x and a
        at base224 -3.4 % step time; 77 MB at large224 +3.5 %; 79 of x's 100 MB at base512 +37 %): x and a when they
        fit 64 MB, else x alone when it does."""
        if not self.l2_persist:
            return 0
        xb = geo.x.numel() * 4
        cap = self._persist_cap
        if xb * 3 // 2 <= cap:
            return xb * 3 // 2
        if xb <= cap or self._persist_partial:
            return xb          # a window larger than the cap persists the fraction cap / window of its lines
        return 0

    def _enqueue(self, geo: _Geometry, x: torch.Tensor, outs, stream: int, limit: int | None = None):
        """Enqueue the whole forward on ``stream``.  Returns the number of kernels launched."""
        with self._on_device():   # function attributes, SM count and the L2 limit are those of the CURRENT device
            n0 = self.lib.ldit_launch_count()
            persist = self._persist_bytes(geo)
            if persist:   # keep the residual stream (and the LayerNorm / context buffer behind it) resident in L2
                if self.lib.ldit_set_l2_window(stream, geo.x.data_ptr(), persist, self._persist_cap) != 0:
                    self.l2_persist, persist = False, 0     # an optimisation the device refused (MPS, MIG slice ...): run without it
            try:
                for name, fn, args in self._plan(geo, x, outs, stream)[:limit]:
                    _lib.check(fn(*args), name)
            finally:
                if persist:
                    self.lib.ldit_set_l2_window(stream, None, 0, 0)
            return int(self.lib.ldit_launch_count() - n0)

    def _on_device(self):
        """Make the engine's device current for the duration of a launch sequence (a model on cuda:1 driven from a
        process whose current device is cuda:0 would otherwise launch into the wrong context)."""
        if self.device is None or torch.cuda.current_device() == self.device.index:
            return contextlib.nullcontext()
        return torch.cuda.device(self.device)

    @staticmethod
    def _as_feats(outs):
        feats = OrderedDict()
        for name, o in zip(("p2", "p3", "p4", "p5", "pool"), outs):   # "pool" only with the FPN head (R:model.py:63)
            feats[name] = o.permute(0, 3, 1, 2)  # [B, C, oh, ow] view of channels-last memory
        return feats

    def prepare_input(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() != 4:
            raise ValueError(f"expected pixel_values of shape [B, 3, H, W], got {tuple(x.shape)}")
        if x.shape[1] != self.cfg.num_channels:
            raise ValueError("Make sure that the channel dimension of the pixel values match with the one set in the configuration.")
        if not x.is_cuda:
            raise _lib.LditError("DiTBackbone.forward needs a CUDA tensor (there is no CPU path)")
        if x.dtype not in _DTYPE_CODE:
            x = x.float()
        H, W = x.shape[2], x.shape[3]
        if H < 16 or W < 16:
            raise ValueError("image smaller than one 16x16 patch")
        if H % 16 or W % 16:  # Conv2d(k16, s16) ignores the ragged border (HF:218)
            x = x[:, :, : H // 16 * 16, : W // 16 * 16]
        return x.contiguous()

    def forward(self, x: torch.Tensor, head: str = "taps"):
        self.refresh_weights()
        H0, W0 = x.shape[2], x.shape[3]
        x = self.prepare_input(x)
        geo = (self._geometry(x.shape[0], H0, W0, 0, head) if (H0 % 16 == 0 and W0 % 16 == 0)
               else self._geometry_ragged(x, H0, W0, head))
        outs = self._alloc_outputs(geo)
        geo.launches = self._enqueue(geo, x, outs, torch.cuda.current_stream(self.device).cuda_stream)
        return self._as_feats(outs)

    def forward_pages(self, pages, size=(224, 224), mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5), head: str = "taps",
                      fixed_size=None):
        """Raw pages -> features: ``GeneralizedRCNNTransform`` as the detector configures it
        (R:model.py:44-56: normalize, bilinear resize to ``fixed_size``, batch) fused into the patch gather,
        then the backbone (and FPN for ``head="fpn"``).  ``pages``: a list of ``[3, H_i, W_i]`` CUDA tensors of
        any sizes (one floating dtype), or one ``[B, 3, Hs, Ws]`` tensor.

        ``size`` is **(height, width)**, the order ``F.interpolate(size=...)`` takes.  torchvision's ``fixed_size`` is
        **(width, height)** (``_resize_image_and_masks`` resizes to ``[fixed_size[1], fixed_size[0]]``): pass that
        tuple unchanged as ``fixed_size=`` instead and it is swapped here.  Both sides must be multiples of 32:
        torchvision's ``batch_images`` zero-pads the batch to ``size_divisible=32``, which this fused path does not
        emulate (the reference's 224 x 224 needs no padding)."""
        if fixed_size is not None:
            size = (int(fixed_size[1]), int(fixed_size[0]))
        if isinstance(pages, torch.Tensor):
            if pages.dim() != 4:
                raise ValueError(f"expected [B, 3, H, W] pages or a list of [3, H, W] pages, got {tuple(pages.shape)}")
            pages = list(pages.contiguous().unbind(0))
        if not pages:
            raise ValueError("empty page list")
        self.refresh_weights()   # also tells which device the model is on
        keep = []
        for p in pages:
            if p.dim() != 3 or p.shape[0] != self.cfg.num_channels:
                raise ValueError(f"images is expected to be a list of 3d tensors of shape [C, H, W], got {tuple(p.shape)}")
            if not p.is_floating_point():
                raise TypeError(f"Expected input images to be of floating type (in range [0, 1]), but found type {p.dtype} instead")
            if not p.is_cuda:
                raise _lib.LditError("DiTBackbone.forward_pages needs CUDA tensors (there is no CPU path)")
            if self.device is not None and p.device != self.device:
                raise _lib.LditError(f"page on {p.device}, model on {self.device}: the gather kernel reads the pages through "
                                     "raw device pointers and cannot cross devices")
            keep.append(p if p.dtype in _DTYPE_CODE else p.float())
        dt = keep[0].dtype
        keep = [p.to(dt).contiguous() for p in keep]
        H, W = int(size[0]), int(size[1])
        if H % 32 or W % 32 or H < 32 or W < 32:
            raise ValueError("the fixed size must be a multiple of 32 in both directions (torchvision pads batches to "
                             "size_divisible=32; that padding is not emulated)")
        self.refresh_weights()
        B = len(keep)
        host = torch.tensor([p.data_ptr() for p in keep], dtype=torch.int64)
        hw = torch.tensor([[p.shape[1], p.shape[2]] for p in keep], dtype=torch.int32)
        pl = _PageList(ptrs=host.to(self.device), hw=hw.to(self.device), max_w=max(p.shape[2] for p in keep), dtype=dt,
                       mean=tuple(float(v) for v in mean), std=tuple(float(v) for v in std), keep=keep)
        geo = self._geometry(B, H, W, 0, head)
        outs = self._alloc_outputs(geo)
        geo.launches = self._enqueue(geo, pl, outs, torch.cuda.current_stream(self.device).cuda_stream)
        return self._as_feats(outs)

    def _geometry_ragged(self, x, H0, W0, head="taps"):
        # the position-table rule looks at the ORIGINAL height/width (HF:135), the conv at the cropped ones
        return dataclasses.replace(self._geometry(x.shape[0], H0, W0, 0, head), H=x.shape[2], W=x.shape[3])

    # ----------------------------------------------------------------------- CUDA graphs
    def forward_graphed(self, x: torch.Tensor, slot: int = 0, head: str = "taps"):
        """Replay a captured CUDA graph of the whole forward for this (B, H, W, dtype).
        Outputs are STATIC buffers overwritten by the next call with the same geometry (and slot)."""
        self.refresh_weights()
        if x.shape[2] % 16 or x.shape[3] % 16:
            return self.forward(x, head)
        x = self.prepare_input(x)
        geo = self._geometry(x.shape[0], x.shape[2], x.shape[3], slot, head)
        if geo.graph is None or geo.graph_in.dtype != x.dtype:
            if geo.graph_in is None or geo.graph_in.dtype != x.dtype:
                geo.graph_in = torch.empty_like(x)
            if x.data_ptr() != geo.graph_in.data_ptr():
                geo.graph_in.copy_(x)
            outs = self._alloc_outputs(geo)
            side = torch.cuda.Stream(self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):  # warm-up outside capture: sets func attributes, loads modules
                self._enqueue(geo, geo.graph_in, outs, side.cuda_stream)
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                geo.launches = self._enqueue(geo, geo.graph_in, outs, torch.cuda.current_stream(self.device).cuda_stream)
            geo.graph, geo.graph_out = graph, self._as_feats(outs)
        if x.data_ptr() != geo.graph_in.data_ptr():
            geo.graph_in.copy_(x, non_blocking=True)
        geo.graph.replay()
        return geo.graph_out

    def graph_input_buffer(self, B, H, W, dtype=torch.float32, slot: int = 0):
        """The static input tensor of the captured graph for this geometry: writing pixels
        straight into it (e.g. the H2D copy) skips the staging copy in ``forward_graphed``."""
        self.refresh_weights()
        geo = self._geometry(B, H, W, slot)
        if geo.graph_in is None or geo.graph_in.dtype != dtype:
            geo.graph = None
            geo.graph_in = torch.zeros(B, 3, H, W, device=self.device, dtype=dtype)
            self.forward_graphed(geo.graph_in, slot)
        return geo.graph_in

    # ------------------------------------------------------------- host-fed, pipelined forward
    def forward_host(self, pages: torch.Tensor, result_host: torch.Tensor | None = None, result_key: str = "p5",
                     depth: int = 2):
        """Forward of a HOST page batch (pinned memory recommended) with the PCIe copies taken off the
        compute stream: the H2D copy of call i+1 and the D2H copy of call i's ``result_key`` tap run on
        their own streams under the kernels of call i (``depth`` independent slots, each with its own
        input buffer, workspaces, captured graph and static outputs).  Returns (feats, done_event):
        ``feats`` are the slot's static device taps (overwritten ``depth`` calls later), ``done_event``
        fires when ``result_host`` (if given) holds the tap in channels-last order."""
        if pages.is_cuda:
            raise ValueError("forward_host takes a host tensor; call forward() for device tensors")
        if pages.dim() != 4 or pages.shape[2] % 16 or pages.shape[3] % 16:
            raise ValueError("forward_host needs [B, 3, H, W] pages with H, W multiples of 16")
        if pages.dtype not in _DTYPE_CODE:
            pages = pages.float()
        self.refresh_weights()
        B, H, W = pages.shape[0], pages.shape[2], pages.shape[3]
        st = self.__dict__.setdefault("_host_pipe", {"n": 0, "h2d": None, "d2h": None, "ev": {}})
        if st["h2d"] is None:
            st["h2d"], st["d2h"] = torch.cuda.Stream(self.device), torch.cuda.Stream(self.device)
        slot = 1 + st["n"] % depth          # slot 0 stays the plain forward_graphed() state
        st["n"] += 1
        cur = torch.cuda.current_stream(self.device)
        buf = self.graph_input_buffer(B, H, W, pages.dtype, slot)
        ev = st["ev"].setdefault((B, H, W, slot), {})
        with torch.cuda.stream(st["h2d"]):
            if "compute" in ev:
                st["h2d"].wait_event(ev["compute"])        # the slot's previous forward has consumed its input
            buf.copy_(pages, non_blocking=True)
            ev["h2d"] = torch.cuda.Event(); ev["h2d"].record(st["h2d"])
        cur.wait_event(ev["h2d"])
        if "d2h" in ev:
            cur.wait_event(ev["d2h"])                      # the slot's previous result has left its static tap
        feats = self.forward_graphed(buf, slot)
        ev["compute"] = torch.cuda.Event(); ev["compute"].record(cur)
        done = ev["compute"]
        if result_host is not None:
            tap = feats[result_key].permute(0, 2, 3, 1)    # channels-last memory of the static output buffer
            with torch.cuda.stream(st["d2h"]):
                st["d2h"].wait_event(ev["compute"])
                result_host.view(tap.shape).copy_(tap, non_blocking=True)
                ev["d2h"] = torch.cuda.Event(); ev["d2h"].record(st["d2h"])
            done = ev["d2h"]
        return feats, done

    def last_launches(self, B, H, W) -> int:
        geo = self._geoms.get((B, H, W))
        return 0 if geo is None else geo.launches
