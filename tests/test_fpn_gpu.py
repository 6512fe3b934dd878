"""GPU parity of the FPN head: ``DiTWithFPN`` through the C ABI against

* fixtures produced by the reference's own ``DiTWithFPN`` (tests/golden/fpn_*.npz, oracle/make_golden_fpn.py),
* the CPU oracle (oracle/fpn_oracle.py) at BASELINE.json's full batch,
* and each new C-ABI entry point against a plain fp32 PyTorch statement of the same op.

Stated tolerance: as for the backbone (bf16 tensor-core math, fp32 accumulation, against the fp32
reference): per map relative-Frobenius error <= 1e-2 and max-abs error <= 5e-2 of the map's abs-max.
"""
import pytest
import torch
import torch.nn.functional as F

from conftest import build_fpn_case, compare_to_golden, fpn_golden_index
from layoutdit_b200 import DiTWithFPN, _lib
from layoutdit_b200.config import dit_base
from layoutdit_b200.synth import make_fpn_state_dict, make_state_dict, synthetic_pages
from oracle import fpn_oracle

pytestmark = pytest.mark.gpu

KEYS = ("p2", "p3", "p4", "p5", "pool")
REL_FRO = 1e-2
MAX_ABS_REL = 5e-2


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _rel_fro(got, ref):
    return float((got.double().cpu() - ref.double().cpu()).norm() / ref.double().cpu().norm().clamp_min(1e-30))


@pytest.fixture(scope="module")
def lib(cuda_device):
    return _lib.load()


# ------------------------------------------------------------------------- entry points
@pytest.mark.parametrize("B,H,W,Cin,Cout", [(1, 7, 7, 256, 256), (2, 20, 28, 256, 256), (3, 56, 56, 256, 256),
                                            (1, 16, 32, 64, 128), (2, 5, 3, 128, 256), (1, 128, 128, 256, 256)])
def test_conv3x3_bias(lib, B, H, W, Cin, Cout):
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + H * 10 + W)
    x = torch.randn(B, H, W, Cin, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) * 0.03).to(torch.bfloat16)
    bias = torch.randn(Cout, device="cuda", generator=g)
    out = torch.full((B, H, W, Cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    wp = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous()
    _lib.check(lib.ldit_conv3x3_bias(x.data_ptr(), wp.data_ptr(), bias.data_ptr(), out.data_ptr(), B, H, W, Cin, Cout, _stream()), "conv")
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, padding=1).permute(0, 2, 3, 1)
    assert torch.isfinite(out.float()).all()        # every output pixel written, none twice with garbage
    assert _rel_fro(out.float(), ref) < 4e-3
    torch.testing.assert_close(out.float(), ref, rtol=2 ** -7, atol=2e-2)


@pytest.mark.parametrize("Gh,Gw,scale,top", [(14, 14, 4.0, True), (14, 14, 2.0, True), (14, 14, 1.0, True), (14, 14, 0.5, False),
                                             (5, 7, 4.0, True), (5, 7, 0.5, False), (5, 7, 1.0, True), (32, 32, 2.0, True)])
def test_fpn_merge(lib, Gh, Gw, scale, top):
    B, C = 2, 256
    g = torch.Generator(device="cuda").manual_seed(int(Gh * 100 + Gw + scale * 7))
    lat = torch.randn(B, Gh, Gw, C, device="cuda", generator=g).to(torch.bfloat16)
    oh, ow = int(Gh * scale), int(Gw * scale)
    th, tw = (int(Gh * scale * 0.5), int(Gw * scale * 0.5)) if top else (0, 0)
    t = torch.randn(B, th, tw, C, device="cuda", generator=g).to(torch.bfloat16) if top else None
    out = torch.empty(B, oh, ow, C, device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.ldit_fpn_merge(lat.data_ptr(), None if t is None else t.data_ptr(), out.data_ptr(), B, Gh, Gw, C, scale, th, tw,
                                  _stream()), "merge")
    ref = lat.float().permute(0, 3, 1, 2)
    if scale != 1.0:
        ref = F.interpolate(ref, scale_factor=scale, mode="bilinear", align_corners=False)
    if t is not None:
        ref = ref + F.interpolate(t.float().permute(0, 3, 1, 2), size=(oh, ow), mode="nearest")
    torch.testing.assert_close(out.float(), ref.permute(0, 2, 3, 1), rtol=2 ** -7, atol=2e-2)


@pytest.mark.parametrize("H,W", [(7, 7), (2, 3), (16, 16)])
def test_subsample2(lib, H, W):
    x = torch.randn(3, H, W, 256, device="cuda").to(torch.bfloat16)
    out = torch.empty(3, (H + 1) // 2, (W + 1) // 2, 256, device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.ldit_subsample2(x.data_ptr(), out.data_ptr(), 3, H, W, 256, _stream()), "pool")
    ref = F.max_pool2d(x.float().permute(0, 3, 1, 2), kernel_size=1, stride=2).permute(0, 2, 3, 1)
    assert torch.equal(out.float(), ref)


# --------------------------------------------------------------------------- the module
def _model(cfg, sd, fsd, **kw):
    return DiTWithFPN(pretrained=False, config=cfg, state_dict=sd, fpn_state_dict=fsd, **kw).cuda().eval()


@pytest.mark.parametrize("name", sorted(fpn_golden_index().keys()))
def test_matches_reference_fixture(cuda_device, name):
    cfg, sd, fsd, x, gold, meta = build_fpn_case(name)
    feats = _model(cfg, sd, fsd)(x.cuda())
    assert list(feats.keys()) == list(KEYS)
    errs = compare_to_golden(feats, gold, meta, rel_fro=REL_FRO, max_abs_rel=MAX_ABS_REL, keys=KEYS)
    print(name, {k: (f"{e:.2e}", f"{m:.2e}") for k, (e, m) in errs.items()})


def test_graph_replay_equals_eager_and_backbone_still_works(cuda_device):
    cfg, sd, fsd, x, _, _ = build_fpn_case("fpn_tiny_native")
    m = _model(cfg, sd, fsd)
    a = {k: v.clone() for k, v in m(x.cuda()).items()}
    g = _model(cfg, sd, fsd, use_cuda_graph=True)
    b = {k: v.clone() for k, v in g(x.cuda()).items()}
    c = g(x.cuda())
    for k in a:
        assert torch.equal(a[k], b[k]) and torch.equal(a[k], c[k])
    taps = m.backbone(x.cuda())                       # the D-channel taps remain available from the same engine
    assert list(taps.keys()) == ["p2", "p3", "p4", "p5"] and taps["p2"].shape[1] == cfg.hidden_size


def test_full_size_c2_fpn(cuda_device):
    """BASELINE config 2 batch (DiT-base, 64 x 224x224) through the FPN head: shapes, strides, finiteness,
    batch independence, and the oracle on two images of the batch."""
    cfg = dit_base()
    sd, fsd = make_state_dict(cfg, 1, True), make_fpn_state_dict(768, 256, 2, True)
    x = synthetic_pages(64, 224, 224, 1234)
    m = _model(cfg, sd, fsd)
    full = m(x.cuda())
    shapes = {"p2": (64, 256, 56, 56), "p3": (64, 256, 28, 28), "p4": (64, 256, 14, 14), "p5": (64, 256, 7, 7), "pool": (64, 256, 4, 4)}
    for k, v in full.items():
        assert tuple(v.shape) == shapes[k] and v.dtype == torch.bfloat16 and v.stride(1) == 1
        assert torch.isfinite(v.float()).all()
    solo = m(x[21:22].cuda())
    for k in full:
        assert torch.equal(full[k][21:22], solo[k])
    ref = fpn_oracle.dit_with_fpn_forward(sd, fsd, cfg.to_dict(), x[[0, 63]])
    for k in ref:
        got = full[k][[0, 63]].float()
        assert _rel_fro(got, ref[k]) < REL_FRO
        assert float((got.cpu() - ref[k]).abs().max() / ref[k].abs().max()) < MAX_ABS_REL
