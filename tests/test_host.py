"""CPU: host-side logic and the C-ABI surface (no kernel is launched)."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT
from layoutdit_b200 import _lib
from layoutdit_b200.config import DiTConfig, dit_base, dit_large, flops_per_image
from layoutdit_b200.dit_params import DiTParameters
from layoutdit_b200.synth import make_state_dict, state_dict_keys, synthetic_pages
from oracle import hf_reference


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "ldit.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ldit_[a-z_0-9]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    from layoutdit_b200 import build
    path = build.build()
    lib = ctypes.CDLL(path)
    declared = _declared_symbols()
    assert len(declared) >= 10
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ldit.h but not exported"
    assert sorted(_lib.SIGNATURES) == declared, "ctypes table and header drifted apart"
    assert lib.ldit_version() >= 100


def test_argument_validation_without_gpu():
    lib = _lib.load()
    assert lib.ldit_layernorm(None, None, None, None, 4, 768, 1e-12, None) == -1
    buf = ctypes.create_string_buffer(64)
    p = ctypes.addressof(buf)
    p16 = (p + 15) & ~15
    assert lib.ldit_layernorm(p16, p16, p16, p16, 4, 100, 1e-12, None) == -2      # D not a multiple of 128
    assert lib.ldit_gemm_bias(p16, p16, None, p16 + 2, 128, 768, 768, None) == -3   # misaligned output
    assert lib.ldit_attention(p16, p16, None, 1, 100, 12, 14, 14, None) == -2       # N != Gh*Gw+1
    assert lib.ldit_patch_embed(p16, 7, p16, p16, p16, p16, p16, 1, 32, 32, 768, None) == -4
    assert lib.ldit_error_string(-3).decode().startswith("pointer")
    assert lib.ldit_patch_embed_scratch_bytes(64, 224, 224) == 64 * 196 * 768 * 2


@pytest.mark.parametrize("cfg", [dit_base(), dit_large(),
                                 DiTConfig(use_relative_position_bias=True, use_absolute_position_embeddings=False),
                                 DiTConfig(use_shared_relative_position_bias=True, use_mask_token=False,
                                           layer_scale_init_value=0.0)])
def test_parameter_tree_matches_hf_state_dict(cfg):
    ours = DiTParameters(cfg).state_dict()
    hf = hf_reference.build(cfg.to_dict()).dit.state_dict()
    assert list(ours) == list(hf)
    assert all(ours[k].shape == hf[k].shape for k in hf)
    assert dict(state_dict_keys(cfg)) == {k: tuple(v.shape) for k, v in hf.items()}


def test_state_dict_roundtrip_and_legacy_buffers_ignored():
    cfg = DiTConfig(hidden_size=128, num_hidden_layers=3, num_attention_heads=2, intermediate_size=256, image_size=64,
                    use_relative_position_bias=True)
    sd = make_state_dict(cfg, 3, True)
    sd["encoder.layer.0.attention.attention.relative_position_bias.relative_position_index"] = torch.zeros(17, 17)
    p = DiTParameters(cfg)
    res = p.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    for k, v in p.state_dict().items():
        assert torch.equal(v, sd[k])


def test_hf_init_statistics():
    p = DiTParameters(dit_base())
    sd = p.state_dict()
    assert float(sd["embeddings.cls_token"].abs().max()) == 0.0
    assert float(sd["embeddings.position_embeddings"].abs().max()) == 0.0
    assert abs(float(sd["encoder.layer.3.intermediate.dense.weight"].std()) - 0.02) < 1e-3
    assert float(sd["encoder.layer.3.lambda_1"].mean()) == pytest.approx(0.1)
    assert float(sd["encoder.layer.0.layernorm_before.weight"].mean()) == 1.0


def test_config_rules_and_flops():
    with pytest.raises(ValueError):
        DiTConfig(hidden_size=100)
    with pytest.raises(ValueError):
        DiTConfig(hidden_size=768, num_attention_heads=8)
    assert flops_per_image(dit_base(), 224, 224) / 1e9 == pytest.approx(35.126, abs=2e-3)
    assert flops_per_image(dit_base(), 512, 512) / 1e9 == pytest.approx(214.054, abs=2e-3)
    assert flops_per_image(dit_large(), 224, 224) / 1e9 == pytest.approx(123.107, abs=2e-3)


def test_backbone_boundary_attributes_and_cpu_refusal():
    from layoutdit_b200 import DiTBackbone
    m = DiTBackbone(pretrained=False)
    assert m.hidden_size == 768 and m.layer_idxs == [4, 6, 8, 12] and m.scales == [4.0, 2.0, 1.0, 0.5]
    assert isinstance(m.dit, torch.nn.Module)
    res = m.dit.load_state_dict({"embeddings.cls_token": torch.ones(1, 1, 768)}, strict=False)
    assert "embeddings.cls_token" not in res.missing_keys
    with pytest.raises(_lib.LditError):   # no CPU fallback: fails loudly
        m.eval()(torch.zeros(1, 3, 224, 224))


def test_synthetic_pages_are_deterministic_and_in_range():
    a, b = synthetic_pages(2, 64, 96, 7), synthetic_pages(2, 64, 96, 7)
    assert torch.equal(a, b) and a.shape == (2, 3, 64, 96)
    assert float(a.min()) >= -1.0 and float(a.max()) <= 1.0


def test_fpn_module_boundary_and_state_dict_names():
    """R:dit_backbone.py:65-95: .backbone / .fpn / .out_channels and the reference's state_dict keys."""
    from torchvision.ops import FeaturePyramidNetwork
    from torchvision.ops.feature_pyramid_network import LastLevelMaxPool
    from layoutdit_b200 import DiTWithFPN
    from layoutdit_b200.synth import make_fpn_state_dict
    cfg = DiTConfig(hidden_size=128, num_hidden_layers=3, num_attention_heads=2, intermediate_size=256, image_size=64)
    m = DiTWithFPN(pretrained=False, config=cfg)
    assert m.out_channels == 256 and m.backbone.hidden_size == 128
    tv = FeaturePyramidNetwork([128] * 4, 256, extra_blocks=LastLevelMaxPool())
    assert {k: tuple(v.shape) for k, v in m.fpn.state_dict().items()} == {k: tuple(v.shape) for k, v in tv.state_dict().items()}
    keys = list(m.state_dict().keys())
    assert "backbone.dit.embeddings.cls_token" in keys and "fpn.inner_blocks.0.0.weight" in keys
    fsd = make_fpn_state_dict(128, 256, 1, True)
    assert not m.fpn.load_state_dict(fsd, strict=True).missing_keys
    with pytest.raises(_lib.LditError):   # CUDA only, fails loudly
        m.eval()(torch.zeros(1, 3, 64, 64))


def test_fpn_argument_validation_without_gpu():
    lib = _lib.load()
    buf = ctypes.create_string_buffer(64)
    p16 = (ctypes.addressof(buf) + 15) & ~15
    assert lib.ldit_conv3x3_bias(None, p16, None, p16, 1, 8, 8, 256, 256, None) == -1
    assert lib.ldit_conv3x3_bias(p16, p16, None, p16, 1, 8, 8, 100, 256, None) == -2   # Cin not a multiple of 64
    assert lib.ldit_fpn_merge(p16, None, p16, 1, 4, 4, 100, 2.0, 0, 0, None) == -2     # C not a multiple of 8
    assert lib.ldit_subsample2(p16, p16 + 2, 1, 4, 4, 256, None) == -3


def test_fused_mlp_schedule_is_a_balanced_partition():
    """ldit_mlp_schedule (host function): every fc1 / fc2 tile exactly once, a pair's fc1 tiles before its fc2 tiles,
    both in increasing order, lists no longer than two separate launches would make them."""
    import numpy as np
    lib = _lib.load()
    C = lib.ldit_mlp_clusters()
    assert C >= 1
    if not lib.ldit_has_experimental():
        assert lib.ldit_mlp_schedule(12608, 768, 3072, None, 0) == -6      # LDIT_E_UNSUPPORTED: not in the product build
        pytest.skip("ldit_mlp_* are only in -DLDIT_EXPERIMENTAL builds")
    for (M, D, I) in [(12608, 768, 3072), (32800, 768, 3072), (12608, 1024, 4096), (197, 768, 3072), (1, 768, 3072)]:
        stride = lib.ldit_mlp_schedule(M, D, I, None, 0)
        assert stride > 0
        buf = np.full(C * stride, -7, dtype=np.int32)
        assert lib.ldit_mlp_schedule(M, D, I, buf.ctypes.data, buf.size) == stride
        assert lib.ldit_mlp_schedule(M, D, I, buf.ctypes.data, buf.size - 1) == -2      # capacity too small
        s = buf.reshape(C, stride)
        bn = 192 if D % 192 == 0 and I % 192 == 0 else 256
        mbs = (M + 255) // 256
        T1, T2 = mbs * (I // bn), mbs * (D // bn)
        assert sorted(s[s >= 0].tolist()) == list(range(T1 + T2))
        w = 1.15 * I / D
        longest = 0.0
        for c in range(C):
            row = s[c]
            n = int((row >= 0).sum())
            assert (row[:n] >= 0).all() and (row[n:] == -1).all()
            l = row[:n]
            second = l >= T1
            assert (np.diff(second.astype(int)) >= 0).all()                  # fc1 tiles first
            assert (np.diff(l[~second]) > 0).all() and (np.diff(l[second]) > 0).all()
            longest = max(longest, float((~second).sum() + w * second.sum()))
        separate = -(-T1 // C) + w * -(-T2 // C)
        assert longest <= separate + 1e-6
    assert lib.ldit_mlp_schedule(34, 128, 256, None, 0) == -2                            # widths the kernel is not built for


def test_l2_persistence_window_policy():
    """Engine._persist_bytes: x + a up to the cap, else x alone, else nothing."""
    import types
    from layoutdit_b200.engine import Engine
    eng = Engine.__new__(Engine)
    eng.l2_persist, eng._persist_cap, eng._persist_partial = True, 64 << 20, False
    geo = lambda rows, D: types.SimpleNamespace(x=torch.empty(rows, D, device="meta"), extra={})
    assert eng._persist_bytes(geo(12608, 768)) == 12608 * 768 * 6          # base224: x and a (58 MB)
    assert eng._persist_bytes(geo(12608, 1024)) == 12608 * 1024 * 4        # large224: x alone (52 MB)
    assert eng._persist_bytes(geo(32800, 768)) == 0                        # base512: x is 100 MB
    eng.l2_persist = False
    assert eng._persist_bytes(geo(12608, 768)) == 0


def test_argument_validation_of_the_widened_entry_points_without_gpu():
    lib = _lib.load()
    buf = ctypes.create_string_buffer(256)
    p16 = (ctypes.addressof(buf) + 15) & ~15
    # fused transform: page table / sizes are required, H and W multiples of 16, non-zero std
    assert lib.ldit_patch_embed_pages(None, p16, 0, 0, .5, .5, .5, .5, .5, .5, p16, p16, p16, p16, p16, 1, 224, 224, 768, None) == -1
    assert lib.ldit_patch_embed_pages(p16, p16, 0, 0, .5, .5, .5, .5, .5, .5, p16, p16, p16, p16, p16, 1, 220, 224, 768, None) == -2
    assert lib.ldit_patch_embed_pages(p16, p16, 0, 0, .5, .5, .5, 0., .5, .5, p16, p16, p16, p16, p16, 1, 224, 224, 768, None) == -2
    assert lib.ldit_patch_embed_pages(p16, p16, 0, 9, .5, .5, .5, .5, .5, .5, p16, p16, p16, p16, p16, 1, 224, 224, 768, None) == -4
    # weight-preparation resize
    assert lib.ldit_resize_rows(None, p16, None, 14, 14, 20, 20, 768, 1, None) == -1
    assert lib.ldit_resize_rows(p16, p16, None, 14, 0, 20, 20, 768, 1, None) == -2
    # fused MLP: schedule and counters are required; widths must be multiples of 192 or of 256
    if lib.ldit_has_experimental():
        assert lib.ldit_mlp_fused(p16, p16, p16, p16, p16, p16, None, p16, 128, 768, 3072, None, 4, p16, None) == -1
        assert lib.ldit_mlp_fused(p16, p16, p16, p16, p16, p16, None, p16, 128, 128, 320, p16, 4, p16, None) == -2
        assert lib.ldit_mlp_fused(p16, p16, p16, p16 + 4, p16, p16, None, p16, 128, 768, 3072, p16, 4, p16, None) == -3
    else:
        assert lib.ldit_mlp_fused(p16, p16, p16, p16, p16, p16, None, p16, 128, 768, 3072, p16, 4, p16, None) == -6
    # L2 window: removing a stream's window never fails, also without a device
    assert lib.ldit_set_l2_window(None, None, 0, 0) == 0
    # workspace size: x f32 + a bf16 + wide bf16 buffer, each rounded up to 1 KB
    M = 64 * 197
    up = lambda v: (v + 1023) // 1024 * 1024
    assert lib.ldit_workspace_bytes(64, 224, 224, 768, 3072) == up(M * 768 * 4) + up(M * 768 * 2) + up(M * 3072 * 2)
    assert lib.ldit_workspace_bytes(0, 224, 224, 768, 3072) == 0


def test_argument_validation_of_the_backward_entry_points_without_gpu():
    """Every training entry point rejects null / misaligned pointers and impossible shapes before touching the device."""
    lib = _lib.load()
    buf = ctypes.create_string_buffer(256)
    p16 = (ctypes.addressof(buf) + 15) & ~15
    assert lib.ldit_gemm_dgrad(None, p16, p16, 128, 768, 768, None) == -1
    assert lib.ldit_gemm_dgrad(p16, p16, p16, 128, 770, 768, None) == -2           # N_out not a multiple of 8
    assert lib.ldit_gemm_dgrad(p16, p16 + 2, p16, 128, 768, 768, None) == -3
    assert lib.ldit_gemm_wgrad(p16, None, p16, 128, 768, 768, None) == -1
    assert lib.ldit_gemm_wgrad(p16, p16, p16, 0, 768, 768, None) == -2
    assert lib.ldit_gemm_wgrad(p16, p16, p16 + 4, 128, 768, 768, None) == -3
    assert lib.ldit_attention_lse(p16, p16, None, None, 1, 197, 12, 14, 14, None) == -1      # the statistics buffer is required
    assert lib.ldit_attention_lse(p16, p16, None, p16, 1, 100, 12, 14, 14, None) == -2       # N != Gh*Gw + 1
    assert lib.ldit_attention_bwd(p16, p16, p16, 1, 257, 12, None) == -6                     # beyond two query tiles
    assert lib.ldit_attention_bwd_flash(p16, p16, None, p16, p16, p16, p16, None, None, 1, 197, 12, 14, 14, None) == -1
    assert lib.ldit_attention_bwd_flash(p16, p16, p16, p16, p16, p16, p16, p16, None, 1, 197, 12, 14, 14, None) == -1   # table without gradient
    assert lib.ldit_attention_bwd_flash(p16, p16, p16, p16, p16, p16, p16, p16, p16, 1, 100, 12, 14, 14, None) == -2    # table and N != Gh*Gw + 1
    assert lib.ldit_scale_residual_rows(p16, p16, None, p16, 0, p16, 8, 128, None) == -2     # row scales need rows_per_image
    assert lib.ldit_scale_residual_rows_bwd(None, p16, None, None, 1, p16, None, 8, 128, None) == -1
    assert lib.ldit_resample_taps_bwd(p16, None, 1, 14, 14, 768, 2.0, None) == -1
    assert lib.ldit_resample_taps_bwd(p16, p16, 1, 14, 14, 770, 2.0, None) == -2
    assert lib.ldit_batch_sum(p16, p16, 4, 6, None) == -2                                     # R % 4
    assert lib.ldit_layernorm_bwd(p16, p16, p16, None, p16, p16, p16, 4, 100, 1e-12, None) == -2
