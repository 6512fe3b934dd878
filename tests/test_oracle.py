"""CPU: the oracle restatement against (a) fixtures produced by the reference itself and
(b) the installed transformers BeitModel wrapped like the reference."""
import numpy as np
import pytest
import torch

from conftest import (build_case, build_fpn_case, build_transform_case, compare_to_golden, fpn_golden_index, golden_index,
                      transform_golden_index)
from oracle import dit_oracle, fpn_oracle, hf_reference, transform_oracle

ALL = sorted(golden_index().keys())
FAST = [n for n in ALL if n.startswith("tiny")] + ["base_224_w1", "base_224_relpos_w1", "base_224_shared_relpos_w1"]


@pytest.mark.parametrize("name", FAST)
def test_oracle_matches_reference_fixture(name):
    cfg, sd, x, gold, meta = build_case(name)
    feats = dit_oracle.dit_backbone_forward(sd, cfg.to_dict(), x)
    # fp32 oracle vs fp32 reference: only summation-order noise is allowed
    compare_to_golden(feats, gold, meta, rel_fro=2e-5, max_abs_rel=2e-4)


def test_oracle_fp64_is_closer_than_tolerance():
    cfg, sd, x, gold, meta = build_case("tiny_abs_native")
    feats = dit_oracle.dit_backbone_forward(sd, cfg.to_dict(), x, dtype=torch.float64)
    compare_to_golden(feats, gold, meta, rel_fro=1e-5, max_abs_rel=1e-4)


@pytest.mark.parametrize("name", ["tiny_abs_interp", "tiny_relpos_interp", "tiny_shared_relpos"])
def test_oracle_matches_installed_transformers(name):
    cfg, sd, x, _, _ = build_case(name)
    with torch.no_grad():
        ref = hf_reference.build(cfg.to_dict(), sd)(x)
    got = dit_oracle.dit_backbone_forward(sd, cfg.to_dict(), x)
    for k in ref:
        assert ref[k].shape == got[k].shape
        torch.testing.assert_close(got[k], ref[k], rtol=1e-4, atol=2e-5)


def test_relative_position_index_conventions():
    idx = dit_oracle.relative_position_index(14, 14)
    assert idx.shape == (197, 197) and idx.dtype == torch.int64
    assert int(idx.min()) == 0 and int(idx.max()) == 731
    assert int(idx[0, 0]) == 731 and int(idx[0, 5]) == 729 and int(idx[5, 0]) == 730


def test_tap_closed_forms():
    """x0.5 == avg_pool2d(2) (floor on odd dims); x2 / x4 impulse responses."""
    t = torch.randn(1, 3, 5, 7)
    half = dit_oracle.resample_bilinear(t, 0.5)
    torch.testing.assert_close(half, torch.nn.functional.avg_pool2d(t, 2), rtol=1e-6, atol=1e-6)
    imp = torch.zeros(1, 1, 1, 9); imp[..., 4] = 1
    r2 = dit_oracle.resample_bilinear(imp.expand(1, 1, 2, 9).contiguous(), 2.0)[0, 0, 0]
    assert np.allclose(r2[7:11].numpy(), [0.25, 0.75, 0.75, 0.25])
    r4 = dit_oracle.resample_bilinear(imp.expand(1, 1, 2, 9).contiguous(), 4.0)[0, 0, 0]
    assert np.allclose(r4[14:22].numpy(), [.125, .375, .625, .875, .875, .625, .375, .125])


def test_tap_layer_indices():
    assert dit_oracle.tap_layer_indices(12) == [4, 6, 8, 12]
    assert dit_oracle.tap_layer_indices(24) == [8, 12, 16, 24]


FPN_KEYS = ("p2", "p3", "p4", "p5", "pool")


@pytest.mark.parametrize("name", sorted(fpn_golden_index().keys()))
def test_fpn_oracle_matches_reference_fixture(name):
    """oracle/fpn_oracle.py vs the outputs of the reference's own DiTWithFPN (oracle/make_golden_fpn.py)."""
    cfg, sd, fsd, x, gold, meta = build_fpn_case(name)
    feats = fpn_oracle.dit_with_fpn_forward(sd, fsd, cfg.to_dict(), x)
    assert list(feats.keys()) == list(FPN_KEYS)
    compare_to_golden(feats, gold, meta, rel_fro=3e-5, max_abs_rel=3e-4, keys=FPN_KEYS)


def test_fpn_oracle_matches_installed_torchvision():
    from torchvision.ops import FeaturePyramidNetwork
    from torchvision.ops.feature_pyramid_network import LastLevelMaxPool
    from collections import OrderedDict
    from layoutdit_b200.synth import make_fpn_state_dict
    fsd = make_fpn_state_dict(64, 256, 5, True)
    fpn = FeaturePyramidNetwork([64] * 4, 256, extra_blocks=LastLevelMaxPool()).eval()
    fpn.load_state_dict(fsd, strict=True)
    g = torch.Generator().manual_seed(3)
    feats = OrderedDict((f"p{i + 2}", torch.randn(2, 64, h, w, generator=g))
                        for i, (h, w) in enumerate([(20, 28), (10, 14), (5, 7), (2, 3)]))
    with torch.no_grad():
        ref = fpn(feats)
    got = fpn_oracle.fpn_forward(fsd, feats)
    assert list(ref.keys()) == list(got.keys()) == list(FPN_KEYS)
    for k in ref:
        torch.testing.assert_close(got[k], ref[k], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("name", sorted(transform_golden_index().keys()))
def test_transform_oracle_matches_reference_fixture(name):
    """oracle/transform_oracle.py vs the transform inside the reference's own LayoutDetectionModel."""
    pages, samples, shape, meta = build_transform_case(name)
    got = transform_oracle.page_transform(pages)
    assert tuple(got.shape) == shape
    np.testing.assert_allclose(got.reshape(-1)[:: meta["stride"]].numpy(), samples, rtol=0, atol=2e-6)
