"""GPU: every C-ABI entry point against a plain fp32 PyTorch statement of the same op on the
same (bf16-rounded) inputs.  All calls go through ctypes -> libldit_b200.so."""
import math

import pytest
import torch
import torch.nn.functional as F

from layoutdit_b200 import _lib

pytestmark = pytest.mark.gpu


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _rel_fro(got, ref):
    return float((got.double() - ref.double()).norm() / ref.double().norm().clamp_min(1e-30))


@pytest.fixture(scope="module")
def lib(cuda_device):
    lib = _lib.load()
    yield lib
    lib.ldit_set_gemm_tile_n(0)
    lib.ldit_set_gemm_cta_pair(2)
    lib.ldit_set_attention_impl(0)


# ----------------------------------------------------------------------------- LayerNorm
@pytest.mark.parametrize("rows,D", [(1, 128), (1001, 768), (1576, 1024), (12608, 768)])
def test_layernorm(lib, rows, D):
    g = torch.Generator(device="cuda").manual_seed(rows + D)
    x = torch.randn(rows, D, device="cuda", generator=g) * 3 + 0.7
    x[0, :] = 5.0                                   # constant row: variance 0, eps=1e-12 must not blow up to NaN
    w = torch.randn(D, device="cuda", generator=g)
    b = torch.randn(D, device="cuda", generator=g)
    y = torch.empty(rows, D, device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.ldit_layernorm(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), rows, D, 1e-12, _stream()), "ln")
    ref = F.layer_norm(x, (D,), w, b, 1e-12)
    assert torch.isfinite(y.float()).all()
    torch.testing.assert_close(y.float()[1:], ref[1:], rtol=2 ** -7, atol=2e-2)
    assert _rel_fro(y.float()[1:], ref[1:]) < 3e-3


# ---------------------------------------------------------------------------------- GEMMs
GEMM_SHAPES = [(128, 128, 64), (200, 256, 128), (333, 768, 768), (1000, 2304, 768), (520, 3072, 768),
               (257, 768, 3072), (130, 1024, 1024), (4100, 768, 768)]


def _gemm_inputs(M, N, K, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = (torch.randn(M, K, device="cuda", generator=g)).to(torch.bfloat16)
    W = (torch.randn(N, K, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", generator=g)
    return A, W, bias


@pytest.mark.parametrize("ctas", [2, 1])
@pytest.mark.parametrize("bn", [0, 128, 192, 256])
@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_bias(lib, M, N, K, bn, ctas):
    lib.ldit_set_gemm_cta_pair(ctas)
    lib.ldit_set_gemm_tile_n(bn)
    A, W, bias = _gemm_inputs(M, N, K, M + N + K)
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.ldit_gemm_bias(A.data_ptr(), W.data_ptr(), bias.data_ptr(), out.data_ptr(), M, N, K, _stream()), "gemm")
    ref = A.float() @ W.float().t() + bias
    assert torch.isfinite(out.float()).all()
    assert _rel_fro(out.float(), ref) < 4e-3
    torch.testing.assert_close(out.float(), ref, rtol=2 ** -7, atol=2e-2)


@pytest.mark.parametrize("ctas", [2, 1])
@pytest.mark.parametrize("M,N,K", [(333, 3072, 768), (12608, 3072, 768)])
def test_gemm_bias_gelu(lib, M, N, K, ctas):
    lib.ldit_set_gemm_cta_pair(ctas)
    lib.ldit_set_gemm_tile_n(0)
    A, W, bias = _gemm_inputs(M, N, K, 7)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.ldit_gemm_bias_gelu(A.data_ptr(), W.data_ptr(), bias.data_ptr(), out.data_ptr(), M, N, K, _stream()), "gemm")
    ref = F.gelu(A.float() @ W.float().t() + bias)          # exact erf GELU
    assert _rel_fro(out.float(), ref) < 4e-3
    torch.testing.assert_close(out.float(), ref, rtol=2 ** -7, atol=2e-2)


@pytest.mark.parametrize("ctas", [2, 1])
@pytest.mark.parametrize("with_scale", [True, False])
@pytest.mark.parametrize("M,N,K", [(333, 768, 768), (12608, 768, 3072), (197, 1024, 4096)])
def test_gemm_bias_scale_residual(lib, M, N, K, with_scale, ctas):
    lib.ldit_set_gemm_cta_pair(ctas)
    lib.ldit_set_gemm_tile_n(0)
    A, W, bias = _gemm_inputs(M, N, K, 11)
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(M, N, device="cuda", generator=g)
    x0 = x.clone()
    scale = torch.rand(N, device="cuda", generator=g) + 0.05 if with_scale else None
    ref = x + (scale if with_scale else 1.0) * (A.float() @ W.float().t() + bias)
    _lib.check(lib.ldit_gemm_bias_scale_residual(A.data_ptr(), W.data_ptr(), bias.data_ptr(),
                                                 scale.data_ptr() if with_scale else None, x.data_ptr(), M, N, K, _stream()),
               "gemm")
    # the branch is rounded to bf16 before it is added (so that this call and ldit_gemm_bias_scale + ldit_add_layernorm agree
    # bit for bit): half a bf16 ulp of the branch on top of the accumulation-order noise over K bf16 products
    branch = ref - x0
    torch.testing.assert_close(x, x0 + branch.to(torch.bfloat16).float(), rtol=1e-4, atol=2e-3 + 2 ** -7 * float(branch.abs().max()))   # a rounding boundary may flip: one bf16 ulp
    assert _rel_fro(x - x0, branch) < 4e-3
    # ... and it is the same contribution, bit for bit, as the two-step form
    x2 = x0.clone()
    br = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.ldit_gemm_bias_scale(A.data_ptr(), W.data_ptr(), bias.data_ptr(), scale.data_ptr() if with_scale else None,
                                        br.data_ptr(), M, N, K, _stream()), "gemm")
    if N % 128 == 0:
        ones, zeros = torch.ones(N, device="cuda"), torch.zeros(N, device="cuda")
        y = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        _lib.check(lib.ldit_add_layernorm(x2.data_ptr(), br.data_ptr(), ones.data_ptr(), zeros.data_ptr(), y.data_ptr(), M, N, 1e-12, _stream()), "add_ln")
        assert torch.equal(x2, x)
    # the unrounded accumulate used by the wgrad GEMMs
    acc = x0.clone()
    _lib.check(lib.ldit_gemm_accumulate(A.data_ptr(), W.data_ptr(), acc.data_ptr(), M, N, K, _stream()), "gemm")
    torch.testing.assert_close(acc, x0 + A.float() @ W.float().t(), rtol=1e-4, atol=2e-3)


def test_gemm_rejects_bad_arguments(lib):
    A, W, bias = _gemm_inputs(128, 128, 64, 1)
    out = torch.empty(128, 128, device="cuda", dtype=torch.bfloat16)
    assert lib.ldit_gemm_bias(A.data_ptr(), W.data_ptr(), bias.data_ptr(), out.data_ptr(), 128, 100, 64, _stream()) == -2
    assert lib.ldit_gemm_bias(A.data_ptr(), None, bias.data_ptr(), out.data_ptr(), 128, 128, 64, _stream()) == -1


# ---------------------------------------------------------------------------- patch embed
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("B,H,W,D", [(2, 64, 96, 128), (3, 224, 224, 768), (1, 80, 112, 128), (2, 512, 512, 768), (1, 16, 16, 128),
                                     (5, 320, 224, 1024)])
def test_patch_embed(lib, B, H, W, D, dtype):
    """fp32 pixels: cast / gather pass + GEMM; fp16 / bf16 pixels: ONE kernel whose A operand is gathered by 5-D TMA boxes
    straight out of the NCHW batch (fp16 pixels stay fp16 in the MMA, the weights bf16), CLS rows included."""
    lib.ldit_set_gemm_cta_pair(2)
    lib.ldit_set_gemm_tile_n(0)
    g = torch.Generator(device="cuda").manual_seed(B * H + W)
    px = (torch.rand(B, 3, H, W, device="cuda", generator=g) * 2 - 1).to(dtype)
    w = (torch.randn(D, 3, 16, 16, device="cuda", generator=g) * 0.04).to(torch.bfloat16)
    cb = torch.randn(D, device="cuda", generator=g) * 0.1
    P = (H // 16) * (W // 16)
    pos = torch.randn(P + 1, D, device="cuda", generator=g) * 0.2
    cls = torch.randn(D, device="cuda", generator=g) * 0.2
    pos_bias = (pos[1:] + cb).contiguous()
    cls_pos = (cls + pos[0]).contiguous()
    scratch = torch.empty(lib.ldit_patch_embed_scratch_bytes(B, H, W) // 2, device="cuda", dtype=torch.bfloat16)
    x = torch.full((B, P + 1, D), float("nan"), device="cuda")
    code = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}[dtype]
    wf = w.float()
    if dtype == torch.float16:     # the TMA-fed GEMM wants both operands in one 16-bit type: an fp16 copy of the weights
        w16 = w.float().to(torch.float16)
        wf = w16.float()
        _lib.check(lib.ldit_patch_embed_tma(px.data_ptr(), code, w16.reshape(D, -1).data_ptr(), pos_bias.data_ptr(), cls_pos.data_ptr(),
                                            x.data_ptr(), B, H, W, D, _stream()), "patch_embed_tma")
    else:
        _lib.check(lib.ldit_patch_embed(px.data_ptr(), code, w.reshape(D, -1).data_ptr(), pos_bias.data_ptr(), cls_pos.data_ptr(),
                                        scratch.data_ptr(), x.data_ptr(), B, H, W, D, _stream()), "patch_embed")
    pxr = px.float() if dtype == torch.float16 else px.to(torch.bfloat16).float()   # A operand: fp16 as given, else bf16
    tok = F.conv2d(pxr, wf, cb, stride=16).flatten(2).transpose(1, 2)
    ref = torch.cat([cls.expand(B, 1, D), tok], dim=1) + pos
    torch.testing.assert_close(x, ref, rtol=1e-4, atol=2e-3)


# ------------------------------------------------------------------------------ attention
def _attn_ref(qkv, B, N, heads, bias):
    D = heads * 64
    q, k, v = qkv.float().reshape(B, N, 3, heads, 64).permute(2, 0, 3, 1, 4)
    s = q @ k.transpose(-1, -2) / 8.0
    if bias is not None:
        s = s + bias
    return (torch.softmax(s, dim=-1) @ v).transpose(1, 2).reshape(B * N, D)


def _dense_bias(table_t, Gh, Gw):
    """[heads, T] -> [1, heads, N, N] by the index rule of HF:522-544 (torch statement)."""
    T = table_t.shape[1]
    ys, xs = torch.meshgrid(torch.arange(Gh), torch.arange(Gw), indexing="ij")
    ys, xs = ys.flatten(), xs.flatten()
    idx = torch.zeros(Gh * Gw + 1, Gh * Gw + 1, dtype=torch.long)
    idx[1:, 1:] = (ys[:, None] - ys[None, :] + Gh - 1) * (2 * Gw - 1) + (xs[:, None] - xs[None, :] + Gw - 1)
    idx[0, :] = T - 3
    idx[:, 0] = T - 2
    idx[0, 0] = T - 1
    return table_t[:, idx.to(table_t.device)].unsqueeze(0)


@pytest.mark.parametrize("impl", [0, 1, 2, 4])
@pytest.mark.parametrize("qk_scale", [1.5, 6.0])
@pytest.mark.parametrize("with_bias", [False, True])
@pytest.mark.parametrize("B,heads,Gh,Gw", [(2, 2, 4, 4), (3, 12, 14, 14), (2, 4, 14, 20), (1, 3, 32, 32), (2, 2, 5, 7),
                                           (1, 2, 13, 16), (2, 1, 1, 1),
                                           # N % 128 in 1..4: the rows past the last full tile take the CUDA-core tail path
                                           (2, 2, 16, 16), (1, 2, 16, 8), (2, 1, 43, 3), (1, 2, 131, 1), (3, 5, 32, 16)])
def test_attention(lib, B, heads, Gh, Gw, with_bias, qk_scale, impl):
    """impl 0 = the product kernel (attention_v3); 1 / 2 / 4 = the superseded variants of -DLDIT_EXPERIMENTAL builds.
    qk_scale 6 gives logits of +-100 and rows whose maximum jumps by far more than 2^8 between key tiles: the lazy
    rescale of the running maximum has to fire (and the polynomial exp2 path sees arguments below -126)."""
    if impl != 0 and not lib.ldit_has_experimental():
        pytest.skip("superseded attention variants are only in -DLDIT_EXPERIMENTAL builds")
    if qk_scale != 1.5 and impl != 0:
        pytest.skip("stress scale only for the product kernel")
    lib.ldit_set_attention_impl(impl)
    N, D = Gh * Gw + 1, heads * 64
    g = torch.Generator(device="cuda").manual_seed(N + heads)
    qkv = (torch.randn(B * N, 3 * D, device="cuda", generator=g) * 1.5).to(torch.bfloat16)
    if qk_scale != 1.5:
        qkv[:, : 2 * D] = (qkv[:, : 2 * D].float() * (qk_scale / 1.5)).to(torch.bfloat16)
    table = None
    bias = None
    if with_bias:
        T = (2 * Gh - 1) * (2 * Gw - 1) + 3
        table = torch.randn(heads, T, device="cuda", generator=g)
        bias = _dense_bias(table, Gh, Gw)
    ctx = torch.full((B * N, D), float("nan"), device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.ldit_attention(qkv.data_ptr(), ctx.data_ptr(), None if table is None else table.data_ptr(),
                                  B, N, heads, Gh, Gw, _stream()), "attention")
    ref = _attn_ref(qkv, B, N, heads, bias)
    assert torch.isfinite(ctx.float()).all()
    assert _rel_fro(ctx.float(), ref) < (8e-3 if qk_scale == 1.5 else 1.5e-2)
    torch.testing.assert_close(ctx.float(), ref, rtol=2e-2, atol=2e-2)


# ----------------------------------------------------------------------------------- taps
@pytest.mark.parametrize("scale", [4.0, 2.0, 1.0, 0.5])
@pytest.mark.parametrize("B,Gh,Gw,D", [(2, 4, 4, 128), (2, 5, 7, 128), (3, 14, 14, 768), (1, 20, 14, 1024)])
def test_resample_taps(lib, B, Gh, Gw, D, scale):
    g = torch.Generator(device="cuda").manual_seed(Gh * Gw)
    x = torch.randn(B, Gh * Gw + 1, D, device="cuda", generator=g)
    oh, ow = int(math.floor(Gh * scale)), int(math.floor(Gw * scale))
    out = torch.full((B, oh, ow, D), float("nan"), device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.ldit_resample_taps(x.data_ptr(), out.data_ptr(), B, Gh, Gw, D, scale, _stream()), "taps")
    t = x[:, 1:, :].permute(0, 2, 1).reshape(B, D, Gh, Gw)
    ref = t if scale == 1.0 else F.interpolate(t, scale_factor=scale, mode="bilinear", align_corners=False)
    got = out.permute(0, 3, 1, 2).float()
    assert got.shape == ref.shape
    torch.testing.assert_close(got, ref, rtol=2 ** -8, atol=1e-6)
    # bit-exact against the bf16 rounding of the fp32 reference for the exact-weight scales
    assert (got == ref.to(torch.bfloat16).float()).float().mean() > 0.999


# ------------------------------------------------------------------ weight-preparation resizes
@pytest.mark.parametrize("h,w,oh,ow,C,cubic", [(14, 14, 32, 32, 768, True), (14, 14, 20, 14, 768, True), (4, 4, 6, 4, 128, True),
                                                (14, 14, 7, 9, 64, True), (27, 27, 63, 63, 12, False), (7, 7, 11, 7, 2, False),
                                                (27, 27, 13, 19, 12, False), (5, 5, 5, 5, 16, True), (1, 1, 1, 1, 768, False)])
def test_resize_rows(lib, h, w, oh, ow, C, cubic):
    """ldit_resize_rows vs F.interpolate(size=..., align_corners=False): HF:138-159 (bicubic) / HF:556-571 (bilinear)."""
    g = torch.Generator(device="cuda").manual_seed(h * 100 + oh)
    src = torch.randn(h * w, C, device="cuda", generator=g)
    add = torch.randn(C, device="cuda", generator=g)
    dst = torch.empty(oh * ow, C, device="cuda")
    _lib.check(lib.ldit_resize_rows(src.data_ptr(), dst.data_ptr(), add.data_ptr(), h, w, oh, ow, C, int(cubic), _stream()), "resize")
    img = src.reshape(1, h, w, C).permute(0, 3, 1, 2)
    ref = img if (h, w) == (oh, ow) else F.interpolate(img, size=(oh, ow), mode="bicubic" if cubic else "bilinear", align_corners=False)
    ref = ref.permute(0, 2, 3, 1).reshape(oh * ow, C) + add
    torch.testing.assert_close(dst, ref, rtol=1e-5, atol=1e-5)


# ------------------------------------------------------------------------- fused MLP (fc1 + fc2)
@pytest.mark.parametrize("M,D,I", [(12608, 768, 3072), (12608, 1024, 4096), (197, 768, 3072), (3000, 768, 3072), (32800, 768, 3072)])
def test_mlp_fused_matches_the_two_gemms_bit_for_bit(lib, M, D, I):
    """ldit_mlp_fused (one persistent kernel, balanced tile lists, fc2 tiles gated on per-row-block readiness counters)
    against ldit_gemm_bias_gelu + ldit_gemm_bias_scale_residual: identical h and x, counters re-armed, repeatedly."""
    import ctypes
    if not lib.ldit_has_experimental():
        pytest.skip("ldit_mlp_fused is only in -DLDIT_EXPERIMENTAL builds")
    g = torch.Generator(device="cuda").manual_seed(M + D)
    a = torch.randn(M, D, device="cuda", generator=g).to(torch.bfloat16)
    W1 = (torch.randn(I, D, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    W2 = (torch.randn(D, I, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    b1, b2, lam = torch.randn(I, device="cuda", generator=g), torch.randn(D, device="cuda", generator=g), torch.rand(D, device="cuda", generator=g)
    x0 = torch.randn(M, D, device="cuda", generator=g)
    st = _stream()
    h_ref, x_ref = torch.empty(M, I, device="cuda", dtype=torch.bfloat16), x0.clone()
    _lib.check(lib.ldit_gemm_bias_gelu(a.data_ptr(), W1.data_ptr(), b1.data_ptr(), h_ref.data_ptr(), M, I, D, st), "fc1")
    _lib.check(lib.ldit_gemm_bias_scale_residual(h_ref.data_ptr(), W2.data_ptr(), b2.data_ptr(), lam.data_ptr(), x_ref.data_ptr(), M, D, I, st), "fc2")
    stride = lib.ldit_mlp_schedule(M, D, I, None, 0)
    assert stride > 0
    host = torch.empty(lib.ldit_mlp_clusters() * stride, dtype=torch.int32)
    assert lib.ldit_mlp_schedule(M, D, I, host.data_ptr(), host.numel()) == stride
    sched = host.cuda()
    ready = torch.zeros(2 * ((M + 255) // 256), device="cuda", dtype=torch.int32)
    for rep in range(3):
        h, x = torch.full((M, I), float("nan"), device="cuda", dtype=torch.bfloat16), x0.clone()
        _lib.check(lib.ldit_mlp_fused(a.data_ptr(), W1.data_ptr(), b1.data_ptr(), h.data_ptr(), W2.data_ptr(), b2.data_ptr(), lam.data_ptr(),
                                      x.data_ptr(), M, D, I, sched.data_ptr(), stride, ready.data_ptr(), st), "mlp")
        torch.cuda.synchronize()
        assert torch.equal(h, h_ref), f"h differs (rep {rep})"
        assert torch.equal(x, x_ref), f"x differs (rep {rep})"
        assert int(ready.abs().sum()) == 0, "counters not re-armed"


# ------------------------------------------------------- residual add deferred into the LayerNorm
@pytest.mark.parametrize("rows,D", [(5, 128), (1001, 768), (1576, 1024), (12608, 768)])
def test_add_layernorm(lib, rows, D):
    g = torch.Generator(device="cuda").manual_seed(rows * 3 + D)
    x = torch.randn(rows, D, device="cuda", generator=g) * 2 + 0.3
    br = (torch.randn(rows, D, device="cuda", generator=g) * 0.7).to(torch.bfloat16)
    w, b = torch.randn(D, device="cuda", generator=g), torch.randn(D, device="cuda", generator=g)
    x_ref = x + br.float()
    y_ref = F.layer_norm(x_ref, (D,), w, b, 1e-12)
    xs, ys = x.clone(), br.clone()                       # y aliases the branch buffer, as in the forward
    _lib.check(lib.ldit_add_layernorm(xs.data_ptr(), ys.data_ptr(), w.data_ptr(), b.data_ptr(), ys.data_ptr(), rows, D, 1e-12, _stream()), "add_ln")
    torch.cuda.synchronize()
    assert torch.equal(xs, x_ref)                        # one fp32 add per element: exact
    torch.testing.assert_close(ys.float(), y_ref, rtol=2 ** -7, atol=2e-2)
    ya = torch.empty(rows, D, device="cuda", dtype=torch.bfloat16)
    xs2 = x.clone()
    _lib.check(lib.ldit_add_layernorm(xs2.data_ptr(), br.data_ptr(), w.data_ptr(), b.data_ptr(), ya.data_ptr(), rows, D, 1e-12, _stream()), "add_ln")
    assert torch.equal(ya, ys)


@pytest.mark.parametrize("M,N,K", [(12608, 768, 768), (12608, 768, 3072), (300, 1024, 4096), (64, 128, 256)])
def test_gemm_bias_scale(lib, M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    bias, lam = torch.randn(N, device="cuda", generator=g), torch.rand(N, device="cuda", generator=g) + 0.1
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.ldit_gemm_bias_scale(a.data_ptr(), w.data_ptr(), bias.data_ptr(), lam.data_ptr(), out.data_ptr(), M, N, K, _stream()), "gemm")
    ref = lam * (a.float() @ w.float().t() + bias)
    assert torch.isfinite(out.float()).all()
    assert _rel_fro(out.float(), ref) < 4e-3
    out2 = torch.empty_like(out)
    _lib.check(lib.ldit_gemm_bias_scale(a.data_ptr(), w.data_ptr(), bias.data_ptr(), None, out2.data_ptr(), M, N, K, _stream()), "gemm")
    assert _rel_fro(out2.float(), a.float() @ w.float().t() + bias) < 4e-3
